#include "ObjParser.h"

#include <algorithm>
#include <chrono>
#include <fstream>
#include <stdio.h>

#include "strtools.h"

using strtools::Token;


ObjParser::ObjParser() {
	mMtlParser = new MtlParser();
	mLightParser = new LightParser();
}


ObjParser::~ObjParser() {
	delete mMtlParser;
	delete mLightParser;
}


vector<cl_int> ObjParser::getFacesMtl() { return mFacesMtl; }
vector<cl_uint> ObjParser::getFacesV() { return mFacesV; }
vector<cl_uint> ObjParser::getFacesVN() { return mFacesVN; }
vector<cl_uint> ObjParser::getFacesVT() { return mFacesVT; }
vector<light_t> ObjParser::getLights() { return mLightParser->getLights(); }
vector<material_t> ObjParser::getMaterials() { return mMtlParser->getMaterials(); }
vector<cl_float> ObjParser::getNormals() { return mNormals; }
vector<object3D> ObjParser::getObjects() { return mObjects; }
vector<cl_float> ObjParser::getTextureCoordinates() { return mTextures; }
vector<cl_float> ObjParser::getVertices() { return mVertices; }


void ObjParser::setScene(
	const vector<cl_float>& vertices, const vector<cl_float>& normals,
	const vector<cl_uint>& facesV, const vector<cl_uint>& facesVN, const vector<cl_int>& facesMtl,
	const vector<object3D>& objects, const vector<material_t>& materials, const vector<light_t>& lights
) {
	mVertices = vertices;
	mNormals = normals;
	mFacesV = facesV;
	mFacesVN = facesVN;
	mFacesMtl = facesMtl;
	mObjects = objects;
	mFacesVT.clear();
	mTextures.clear();
	mMtlParser->setMaterials( materials );
	mLightParser->setLights( lights );
}


namespace {

/** Append the three coordinates of a "v x y z" / "vn x y z" line (ObjParser.cpp:314-335). */
inline void pushVec3( const vector<Token>& parts, vector<cl_float>* out ) {
	for( size_t k = 1; k <= 3; k++ ) {
		out->push_back( k < parts.size() ? (cl_float) strtools::toDouble( parts[k] ) : 0.0f );
	}
}

/** One "v", "v/vt", "v//vn" or "v/vt/vn" group of an `f` line (ObjParser.cpp:262-306). */
inline void pushFaceGroup( const Token& group, vector<cl_uint>* v, vector<cl_uint>* vn, vector<cl_uint>* vt ) {
	/* split at every '/' */
	Token e[4];
	size_t ne = 0, count = 1;
	const char* start = group.p;
	for( size_t i = 0; i < group.n; i++ ) {
		if( group.p[i] == '/' ) {
			if( ne < 4 ) { e[ne].p = start; e[ne].n = (size_t) ( group.p + i - start ); ne++; }
			start = group.p + i + 1;
			count++;
		}
	}
	if( ne < 4 ) { e[ne].p = start; e[ne].n = (size_t) ( group.p + group.n - start ); ne++; }

	if( count == 2 ) {
		/* the reference's "v//vn" branch: exactly two fields */
		v->push_back( (cl_uint) strtools::toLong( e[0] ) - 1 );
		vn->push_back( (cl_uint) strtools::toLong( e[1] ) - 1 );
		return;
	}
	v->push_back( (cl_uint) strtools::toLong( e[0] ) - 1 );
	if( count >= 2 ) {
		vt->push_back( (cl_uint) strtools::toLong( e[1] ) - 1 );
	}
	if( count >= 3 ) {
		vn->push_back( (cl_uint) strtools::toLong( e[2] ) - 1 );
	}
}

}


/**
 * Load an OBJ file (reference: ObjParser.cpp:121-221).
 * @param {std::string} filepath Path to the file.
 * @param {std::string} filename Name of the file.
 */
void ObjParser::load( string filepath, string filename ) {
	mObjects.clear();
	mFacesMtl.clear();
	mFacesV.clear();
	mFacesVN.clear();
	mFacesVT.clear();
	mNormals.clear();
	mTextures.clear();
	mVertices.clear();

	filepath.append( filename );

	if( Cfg::get().value<int>( Cfg::RENDER_SHADOWRAYS ) > 0 ) {
		this->loadLights( filepath );
	}
	else {
		mLightParser->setLights( vector<light_t>() );
	}

	this->loadMtl( filepath );
	vector<material_t> materials = mMtlParser->getMaterials();
	vector<string> materialNames;
	cl_int currentMtl = -1;

	for( size_t i = 0; i < materials.size(); i++ ) {
		materialNames.push_back( materials[i].mtlName );
	}

	std::chrono::steady_clock::time_point timerStart = std::chrono::steady_clock::now();

	string text;
	{
		std::ifstream fileIn( filepath.c_str(), std::ios::binary );
		if( fileIn ) {
			fileIn.seekg( 0, std::ios::end );
			std::streamoff len = fileIn.tellg();
			fileIn.seekg( 0, std::ios::beg );
			text.resize( len > 0 ? (size_t) len : 0 );
			if( len > 0 ) { fileIn.read( &text[0], len ); }
		}
		else {
			Logger::logError( "[ObjParser] Could not open file \"" + filepath + "\"." );
		}
	}

	vector<Token> parts;
	vector<cl_uint> lineV, lineVN, lineVT;
	size_t pos = 0;

	while( pos < text.size() ) {
		size_t nl = pos;
		const char* data = text.data();
		const void* found = memchr( data + pos, '\n', text.size() - pos );
		nl = found ? (size_t) ( (const char*) found - data ) : text.size();
		const char* b = data + pos;
		const char* e = data + nl;
		pos = nl + 1;
		strtools::trim( b, e );
		const size_t len = (size_t) ( e - b );
		const char c0 = len > 0 ? b[0] : '\0';
		const char c1 = len > 1 ? b[1] : '\0';
		const char c2 = len > 2 ? b[2] : '\0';

		if( c0 == '#' ) {
			continue;
		}

		if( c0 == 'o' ) {
			strtools::split( parts, b, e, " \t" );
			object3D o;
			o.oName = parts.size() > 1 ? parts[1].str() : "";
			mObjects.push_back( o );
		}
		else if( c0 == 'v' ) {
			if( c1 == ' ' ) {
				strtools::split( parts, b, e, " \t" );
				pushVec3( parts, &mVertices );
			}
			else if( c1 == 'n' && c2 == ' ' ) {
				strtools::split( parts, b, e, " \t" );
				pushVec3( parts, &mNormals );
			}
			else if( c1 == 't' && c2 == ' ' ) {
				strtools::split( parts, b, e, " \t" );
				const cl_float weight = ( parts.size() >= 4 ) ? (cl_float) strtools::toDouble( parts[3] ) : 0.0f;
				mTextures.push_back( parts.size() > 1 ? (cl_float) strtools::toDouble( parts[1] ) : 0.0f );
				mTextures.push_back( parts.size() > 2 ? (cl_float) strtools::toDouble( parts[2] ) : 0.0f );
				mTextures.push_back( weight );
			}
		}
		else if( c0 == 'f' ) {
			if( c1 == ' ' ) {
				strtools::split( parts, b, e, " \t" );
				lineV.clear();
				lineVN.clear();
				lineVT.clear();
				for( size_t i = 1; i < parts.size(); i++ ) {
					pushFaceGroup( parts[i], &lineV, &lineVN, &lineVT );
				}

				mFacesV.insert( mFacesV.end(), lineV.begin(), lineV.end() );
				mFacesVN.insert( mFacesVN.end(), lineVN.begin(), lineVN.end() );
				mFacesVT.insert( mFacesVT.end(), lineVT.begin(), lineVT.end() );
				mFacesMtl.push_back( currentMtl );

				if( mObjects.size() > 0 ) {
					object3D* op = &( mObjects[mObjects.size() - 1] );
					op->facesV.insert( op->facesV.end(), lineV.begin(), lineV.end() );
					op->facesVN.insert( op->facesVN.end(), lineVN.begin(), lineVN.end() );
				}
			}
		}
		else if( len >= 6 && std::search( b, e, "usemtl", "usemtl" + 6 ) != e ) {
			strtools::split( parts, b, e, " \t" );
			const string name = parts.size() > 1 ? parts[1].str() : "";
			vector<string>::iterator it = std::find( materialNames.begin(), materialNames.end(), name );
			currentMtl = ( it != materialNames.end() ) ? (cl_int) ( it - materialNames.begin() ) : -1;
		}
	}

	const double timeDiff = std::chrono::duration<double>( std::chrono::steady_clock::now() - timerStart ).count();
	char msg[256];
	snprintf(
		msg, 256, "[ObjParser] Loaded %lu vertices, %lu normals, and %lu faces in %g s.",
		(unsigned long) ( mVertices.size() / 3 ), (unsigned long) ( mFacesVN.size() / 3 ),
		(unsigned long) ( mFacesV.size() / 3 ), timeDiff
	);
	Logger::logInfo( msg );
}


/**
 * Load the LIGHTS file to the OBJ: same name, extension ".lights" (reference: ObjParser.cpp:228-233).
 */
void ObjParser::loadLights( string file ) {
	size_t extensionIndex = file.rfind( ".obj" );
	if( extensionIndex == string::npos ) { file.append( ".lights" ); }
	else { file.replace( extensionIndex, 4, ".lights" ); }
	mLightParser->load( file );
}


/**
 * Load the MTL file to the OBJ: same name, extension ".mtl" (reference: ObjParser.cpp:240-245).
 */
void ObjParser::loadMtl( string file ) {
	size_t extensionIndex = file.rfind( ".obj" );
	if( extensionIndex == string::npos ) { file.append( ".mtl" ); }
	else { file.replace( extensionIndex, 4, ".mtl" ); }
	mMtlParser->load( file );
}
