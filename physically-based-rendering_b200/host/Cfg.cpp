#include "Cfg.h"

#include <ctype.h>
#include <fstream>
#include <vector>


const char* Cfg::ACCEL_STRUCT = "accel_struct";
const char* Cfg::BVH_MAXFACES = "bvh.max_faces";
const char* Cfg::BVH_SAHFACESLIMIT = "bvh.sah_faces_limit";
const char* Cfg::BVH_SKIPAHEAD = "bvh.skip_ahead";
const char* Cfg::BVH_SKIPAHEAD_CMP = "bvh.skip_ahead_compare";
const char* Cfg::CAM_CENTER_X = "camera.center.x";
const char* Cfg::CAM_CENTER_Y = "camera.center.y";
const char* Cfg::CAM_CENTER_Z = "camera.center.z";
const char* Cfg::CAM_EYE_X = "camera.eye.x";
const char* Cfg::CAM_EYE_Y = "camera.eye.y";
const char* Cfg::CAM_EYE_Z = "camera.eye.z";
const char* Cfg::CAM_LENSE_APERTURE = "camera.thin_lense.aperture";
const char* Cfg::CAM_LENSE_FOCALLENGTH = "camera.thin_lense.focal_length";
const char* Cfg::CAM_SPEED = "camera.speed";
const char* Cfg::IMPORT_PATH = "import_path";
const char* Cfg::INFO_KERNELTIMES = "info.kernel_times";
const char* Cfg::LOG_LEVEL = "logging.level";
const char* Cfg::OPENCL_BUILDOPTIONS = "opencl.build_options";
const char* Cfg::OPENCL_CHECKERRORS = "opencl.check_errors";
const char* Cfg::OPENCL_LOCALGROUPSIZE = "opencl.localgroupsize";
const char* Cfg::OPENCL_PROGRAM = "opencl.program";
const char* Cfg::PERS_FOV = "camera.perspective.fov";
const char* Cfg::PERS_ZFAR = "camera.perspective.zfar";
const char* Cfg::PERS_ZNEAR = "camera.perspective.znear";
const char* Cfg::RENDER_ANTIALIAS = "render.antialiasing";
const char* Cfg::RENDER_BRDF = "render.brdf";
const char* Cfg::RENDER_INTERVAL = "render.interval";
const char* Cfg::RENDER_MAXADDEDDEPTH = "render.max_added_depth";
const char* Cfg::RENDER_MAXDEPTH = "render.max_depth";
const char* Cfg::RENDER_PHONGTESS = "render.phong_tessellation";
const char* Cfg::RENDER_SAMPLES = "render.samples";
const char* Cfg::RENDER_SHADOWRAYS = "render.shadow_rays";
const char* Cfg::SHADER_NAME = "shader.name";
const char* Cfg::SHADER_PATH = "shader.path";
const char* Cfg::WINDOW_HEIGHT = "window.height";
const char* Cfg::WINDOW_WIDTH = "window.width";


namespace {

/** Recursive-descent reader: objects become dotted prefixes, scalars are kept as their text. */
struct JsonReader {
	const std::string& s;
	size_t i;
	std::map<std::string, std::string>& out;

	JsonReader( const std::string& text, std::map<std::string, std::string>& o ) : s( text ), i( 0 ), out( o ) {}

	void fail( const char* what ) {
		throw std::runtime_error( std::string( "[Cfg] JSON: " ) + what + " at offset " + std::to_string( i ) );
	}

	void skip() {
		while( i < s.size() ) {
			if( isspace( (unsigned char) s[i] ) ) { i++; }
			else if( s[i] == '/' && i + 1 < s.size() && s[i + 1] == '/' ) {
				while( i < s.size() && s[i] != '\n' ) { i++; }
			}
			else if( s[i] == '/' && i + 1 < s.size() && s[i + 1] == '*' ) {
				i += 2;
				while( i + 1 < s.size() && !( s[i] == '*' && s[i + 1] == '/' ) ) { i++; }
				i += 2;
			}
			else { break; }
		}
	}

	std::string readString() {
		if( s[i] != '"' ) { fail( "expected string" ); }
		i++;
		std::string r;
		while( i < s.size() && s[i] != '"' ) {
			if( s[i] == '\\' && i + 1 < s.size() ) {
				i++;
				switch( s[i] ) {
					case 'n': r.push_back( '\n' ); break;
					case 't': r.push_back( '\t' ); break;
					case 'r': r.push_back( '\r' ); break;
					case 'b': r.push_back( '\b' ); break;
					case 'f': r.push_back( '\f' ); break;
					default: r.push_back( s[i] ); break;
				}
				i++;
			}
			else { r.push_back( s[i++] ); }
		}
		if( i >= s.size() ) { fail( "unterminated string" ); }
		i++;
		return r;
	}

	void readValue( const std::string& prefix ) {
		skip();
		if( i >= s.size() ) { fail( "unexpected end" ); }
		if( s[i] == '{' ) {
			i++;
			skip();
			if( s[i] == '}' ) { i++; return; }
			while( true ) {
				skip();
				std::string key = readString();
				skip();
				if( s[i] != ':' ) { fail( "expected ':'" ); }
				i++;
				readValue( prefix.empty() ? key : prefix + "." + key );
				skip();
				if( s[i] == ',' ) { i++; continue; }
				if( s[i] == '}' ) { i++; break; }
				fail( "expected ',' or '}'" );
			}
		}
		else if( s[i] == '[' ) {
			/* property_tree stores array elements under empty keys; nothing in config.json uses one */
			i++;
			int idx = 0;
			skip();
			if( s[i] == ']' ) { i++; return; }
			while( true ) {
				readValue( prefix + "." + std::to_string( idx++ ) );
				skip();
				if( s[i] == ',' ) { i++; continue; }
				if( s[i] == ']' ) { i++; break; }
				fail( "expected ',' or ']'" );
			}
		}
		else if( s[i] == '"' ) {
			out[prefix] = readString();
		}
		else {
			size_t b = i;
			while( i < s.size() && s[i] != ',' && s[i] != '}' && s[i] != ']' && !isspace( (unsigned char) s[i] ) ) { i++; }
			if( b == i ) { fail( "expected value" ); }
			out[prefix] = s.substr( b, i - b );
		}
	}
};

/* The reference's shipped config.json (config.json:1-126), minus machine-specific import_path. */
const char* DEFAULT_CONFIG =
	"{ \"camera\": { \"eye\": { \"x\": 0.00, \"y\": 1.00, \"z\": 3.00 },"
	" \"center\": { \"x\": 0.00, \"y\": 0.00, \"z\": 1.00 },"
	" \"perspective\": { \"fov\": 45.0, \"zfar\": 1000.0, \"znear\": 0.1 },"
	" \"thin_lense\": { \"aperture\": 1.8, \"focal_length\": 0.035 }, \"speed\": 0.2 },"
	" \"import_path\": \"resources/models/\","
	" \"info\": { \"kernel_times\": 250.0 },"
	" \"accel_struct\": 0,"
	" \"bvh\": { \"max_faces\": 2, \"sah_faces_limit\": 100000, \"skip_ahead\": true, \"skip_ahead_compare\": 0.7 },"
	" \"logging\": { \"level\": 1 },"
	" \"opencl\": { \"build_options\": \"\", \"check_errors\": true,"
	" \"program\": \"source/opencl/pathtracing.cl\", \"localgroupsize\": 8 },"
	" \"render\": { \"antialiasing\": 0.7, \"brdf\": 1, \"interval\": 33.3, \"max_added_depth\": 5,"
	" \"max_depth\": 3, \"phong_tessellation\": 0.0, \"samples\": 1, \"shadow_rays\": 0 },"
	" \"shader\": { \"name\": \"pathtracing\", \"path\": \"source/shader/\" },"
	" \"window\": { \"height\": 600, \"width\": 800 } }";

}


/**
 * Load the config file (JSON, `//` comments allowed).
 * Reference: Cfg::loadConfigFile (Cfg.cpp:46-48).  Keys already set and absent from the file keep
 * their value, so a partial file overrides the defaults.
 * @param {const char*} filepath File path.
 */
void Cfg::loadConfigFile( const char* filepath ) {
	std::ifstream in( filepath );
	if( !in ) {
		throw std::runtime_error( std::string( "[Cfg] cannot open file " ) + filepath );
	}
	std::stringstream ss;
	ss << in.rdbuf();
	this->loadConfigString( ss.str() );
}


void Cfg::loadConfigString( const std::string& json ) {
	JsonReader r( json, mValues );
	r.readValue( "" );
}


void Cfg::loadDefaults() {
	mValues.clear();
	this->loadConfigString( DEFAULT_CONFIG );
}
