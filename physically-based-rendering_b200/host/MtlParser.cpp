#include "MtlParser.h"

using std::string;
using std::vector;

#include <fstream>
#include <sstream>

#include "strtools.h"

using strtools::Token;


/**
 * Get a material with default values (reference: MtlParser.cpp:11-36).
 */
material_t MtlParser::getEmptyMaterial() {
	const cl_float4 white = { 1.0f, 1.0f, 1.0f, 0.0f };
	material_t mtl;
	mtl.mtlName = "";
	mtl.Ka = white;
	mtl.Kd = white;
	mtl.Ks = white;
	mtl.d = 1.0f;
	mtl.Ni = 1.0f;
	mtl.Ns = 100.0f;
	mtl.illum = 2;
	mtl.light = 0;
	mtl.rough = 1.0f;
	mtl.p = 1.0f;
	mtl.nu = 0.0f;
	mtl.nv = 0.0f;
	mtl.Rs = 0.0f;
	mtl.Rd = 1.0f;
	return mtl;
}


vector<material_t> MtlParser::getMaterials() {
	return mMaterials;
}


void MtlParser::setMaterials( const vector<material_t>& materials ) {
	mMaterials = materials;
}


namespace {

/* attribute name -> what to do with it; `need` = minimum number of tokens on the line */
enum Attr { A_D, A_TR, A_ILLUM, A_KA, A_KD, A_KS, A_NI, A_NS, A_LIGHT, A_ROUGH, A_P, A_NU, A_NV, A_RS, A_RD };
struct AttrDef { const char* key; Attr attr; size_t need; };
const AttrDef ATTRS[] = {
	{ "d", A_D, 2 }, { "Tr", A_TR, 2 }, { "illum", A_ILLUM, 2 }, { "Ka", A_KA, 4 }, { "Kd", A_KD, 4 },
	{ "Ks", A_KS, 4 }, { "Ni", A_NI, 2 }, { "Ns", A_NS, 2 }, { "light", A_LIGHT, 2 }, { "rough", A_ROUGH, 2 },
	{ "p", A_P, 2 }, { "nu", A_NU, 2 }, { "nv", A_NV, 2 }, { "Rs", A_RS, 2 }, { "Rd", A_RD, 2 },
};

void setRGB( cl_float4* c, const vector<Token>& parts ) {
	c->x = (cl_float) strtools::toDouble( parts[1] );
	c->y = (cl_float) strtools::toDouble( parts[2] );
	c->z = (cl_float) strtools::toDouble( parts[3] );
}

}


/**
 * Load the materials from the file (reference: MtlParser.cpp:51-236).
 * @param {std::string} file File path and name of the MTL file.
 */
void MtlParser::load( string file ) {
	mMaterials.clear();

	std::ifstream fileIn( file.c_str() );
	if( !fileIn ) {
		Logger::logWarning( "[MtlParser] Could not open file \"" + file + "\". No materials loaded." );
		return;
	}
	std::stringstream ss;
	ss << fileIn.rdbuf();
	const string text = ss.str();

	material_t mtl = getEmptyMaterial();
	int numMtlFound = 0;
	bool isSetTransparency = false;
	vector<Token> parts;

	size_t pos = 0;
	while( pos <= text.size() ) {
		size_t nl = text.find( '\n', pos );
		if( nl == string::npos ) { nl = text.size(); }
		const char* b = text.data() + pos;
		const char* e = text.data() + nl;
		pos = nl + 1;
		strtools::trim( b, e );

		if( e - b < 3 || *b == '#' ) {
			continue;
		}
		strtools::split( parts, b, e, " \t" );

		if( parts[0].equals( "newmtl" ) ) {
			if( parts.size() < 2 ) {
				Logger::logWarning( "[MtlParser] No name for <newmtl>. Ignoring entry." );
				continue;
			}
			if( numMtlFound > 0 ) {
				mMaterials.push_back( mtl );
			}
			numMtlFound++;
			mtl = getEmptyMaterial();
			mtl.mtlName = parts[1].str();
			continue;
		}

		for( size_t a = 0; a < sizeof( ATTRS ) / sizeof( ATTRS[0] ); a++ ) {
			if( !parts[0].equals( ATTRS[a].key ) ) {
				continue;
			}
			if( ATTRS[a].attr == A_TR && isSetTransparency ) {
				break;   /* the reference's `else if( parts[0] == "Tr" && !isSetTransparency )` falls through to nothing */
			}
			if( parts.size() < ATTRS[a].need ) {
				Logger::logWarning( string( "[MtlParser] Not enough parameters for <" ) + ATTRS[a].key + ">. Ignoring attribute." );
				break;
			}
			const double v = strtools::toDouble( parts[1] );
			switch( ATTRS[a].attr ) {
				case A_D: mtl.d = (cl_float) v; isSetTransparency = true; break;
				case A_TR: mtl.d = (cl_float) ( 1.0f - v ); break;
				case A_ILLUM:
					mtl.illum = (cl_char) strtools::toLong( parts[1] );
					if( mtl.illum < 0 || mtl.illum > 10 ) {
						Logger::logWarning( "[MtlParser] Invalid value for <illum>. Has to be between 0 and 10. Ignoring attribute." );
						mtl.illum = 2;
					}
					break;
				case A_KA: setRGB( &mtl.Ka, parts ); break;
				case A_KD: setRGB( &mtl.Kd, parts ); break;
				case A_KS: setRGB( &mtl.Ks, parts ); break;
				case A_NI: mtl.Ni = (cl_float) v; break;
				case A_NS: mtl.Ns = (cl_float) v; break;
				case A_LIGHT: mtl.light = (cl_char) atoi( parts[1].str().c_str() ); break;
				case A_ROUGH: mtl.rough = (cl_float) v; break;
				case A_P: mtl.p = (cl_float) v; break;
				case A_NU: mtl.nu = (cl_float) v; break;
				case A_NV: mtl.nv = (cl_float) v; break;
				case A_RS: mtl.Rs = (cl_float) v; break;
				case A_RD: mtl.Rd = (cl_float) v; break;
			}
			break;
		}
	}

	if( numMtlFound > 0 ) {
		mMaterials.push_back( mtl );
	}

	char msg[64];
	snprintf( msg, 64, "[MtlParser] Loaded %lu material(s).", (unsigned long) mMaterials.size() );
	Logger::logInfo( msg );
}
