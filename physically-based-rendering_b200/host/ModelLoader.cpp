#include "ModelLoader.h"


ModelLoader::ModelLoader() {
	mObjParser = new ObjParser();
}


ModelLoader::~ModelLoader() {
	delete mObjParser;
}


/**
 * The faces of one object as (v0, v1, v2, global face index); the global index is
 * `offset + position in the object` (reference: ModelLoader.cpp:28-41).
 */
void ModelLoader::getFacesOfObject( const object3D& object, vector<cl_uint4>* faces, cl_int offset ) {
	faces->reserve( faces->size() + object.facesV.size() / 3 );
	for( size_t i = 0; i + 2 < object.facesV.size(); i += 3 ) {
		cl_uint4 f = { object.facesV[i], object.facesV[i + 1], object.facesV[i + 2], (cl_uint) ( offset + (cl_int) faces->size() ) };
		faces->push_back( f );
	}
}


/** Same for the normal indices (reference: ModelLoader.cpp:44-57). */
void ModelLoader::getFaceNormalsOfObject( const object3D& object, vector<cl_uint4>* faceNormals, cl_int offset ) {
	faceNormals->reserve( faceNormals->size() + object.facesVN.size() / 3 );
	for( size_t i = 0; i + 2 < object.facesVN.size(); i += 3 ) {
		cl_uint4 fn = { object.facesVN[i], object.facesVN[i + 1], object.facesVN[i + 2], (cl_uint) ( offset + (cl_int) faceNormals->size() ) };
		faceNormals->push_back( fn );
	}
}


ObjParser* ModelLoader::getObjParser() {
	return mObjParser;
}


/**
 * Load 3D model (reference: ModelLoader.cpp:74-88).
 * @param {std::string} filepath Path to the file, without file name.
 * @param {std::string} filename Name of the file.
 */
void ModelLoader::loadModel( string filepath, string filename ) {
	Logger::logInfo( "[ModelLoader] Importing model \"" + filename + "\" ..." );
	Logger::indent( LOG_INDENT );
	mObjParser->load( filepath, filename );
	Logger::indent( 0 );
	Logger::logInfo( "[ModelLoader] ... Done." );
}
