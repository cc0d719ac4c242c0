#include "ModelLoader.h"

using std::string;
using std::vector;


ModelLoader::ModelLoader() {
	mObjParser = new ObjParser();
}


ModelLoader::~ModelLoader() {
	delete mObjParser;
}


namespace {

/* Index triples of one object -> ( i0, i1, i2, offset + running count of `out` ). */
void appendTriples( const vector<cl_uint>& triples, vector<cl_uint4>* out, cl_int offset ) {
	out->reserve( out->size() + triples.size() / 3 );
	for( size_t i = 0; i + 2 < triples.size(); i += 3 ) {
		cl_uint4 entry = { triples[i], triples[i + 1], triples[i + 2], (cl_uint) ( offset + (cl_int) out->size() ) };
		out->push_back( entry );
	}
}

}


/** The faces of one object with their scene-wide face number in .w (reference: ModelLoader.cpp:28-41). */
void ModelLoader::getFacesOfObject( const object3D& object, vector<cl_uint4>* faces, cl_int offset ) {
	appendTriples( object.facesV, faces, offset );
}


/** Same for the normal indices (reference: ModelLoader.cpp:44-57). */
void ModelLoader::getFaceNormalsOfObject( const object3D& object, vector<cl_uint4>* faceNormals, cl_int offset ) {
	appendTriples( object.facesVN, faceNormals, offset );
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
