#include "LightParser.h"

using std::string;
using std::vector;

#include <fstream>
#include <sstream>

#include "strtools.h"

using strtools::Token;


/** Reference: LightParser.cpp:11-22. */
light_t LightParser::getEmptyLight() {
	const cl_float4 white = { 1.0f, 1.0f, 1.0f, 0.0f };
	light_t light;
	light.lightName = "";
	light.type = 0;
	light.pos = white;
	light.rgb = white;
	light.radius = 0.0f;
	return light;
}


vector<light_t> LightParser::getLights() {
	return mLights;
}


void LightParser::setLights( const vector<light_t>& lights ) {
	mLights = lights;
}


/**
 * Load the lights from the file (reference: LightParser.cpp:38-128).
 * @param {std::string} file File path and name of the LIGHTS file.
 */
void LightParser::load( string file ) {
	mLights.clear();

	std::ifstream fileIn( file.c_str() );
	if( !fileIn ) {
		Logger::logWarning( "[LightParser] Could not open file \"" + file + "\". No lights loaded." );
		return;
	}
	std::stringstream ss;
	ss << fileIn.rdbuf();
	const string text = ss.str();

	light_t light = getEmptyLight();
	int numLightsFound = 0;
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
		const Token& key = parts[0];

		if( key.equals( "newlight" ) ) {
			if( parts.size() < 2 ) {
				Logger::logWarning( "[LightParser] No name for <newlight>. Ignoring entry." );
				continue;
			}
			if( numLightsFound > 0 ) {
				mLights.push_back( light );
			}
			numLightsFound++;
			light = getEmptyLight();
			light.lightName = parts[1].str();
		}
		else if( key.equals( "type" ) ) {
			if( parts.size() < 2 ) { Logger::logWarning( "[LightParser] Not enough parameters for <type>. Ignoring attribute." ); continue; }
			light.type = (cl_uint) strtools::toLong( parts[1] );
		}
		else if( key.equals( "rgb" ) || key.equals( "pos" ) ) {
			if( parts.size() < 4 ) { Logger::logWarning( "[LightParser] Not enough parameters for <" + key.str() + ">. Ignoring attribute." ); continue; }
			cl_float4* dst = key.equals( "rgb" ) ? &light.rgb : &light.pos;
			dst->x = (cl_float) strtools::toDouble( parts[1] );
			dst->y = (cl_float) strtools::toDouble( parts[2] );
			dst->z = (cl_float) strtools::toDouble( parts[3] );
		}
		else if( key.equals( "radius" ) ) {
			if( parts.size() < 2 ) { Logger::logWarning( "[LightParser] Not enoug parameters for <radius>. Ignoring attribute." ); continue; }
			light.radius = (cl_float) strtools::toDouble( parts[1] );
		}
	}

	if( numLightsFound > 0 ) {
		mLights.push_back( light );
	}
	else {
		Cfg::get().value( Cfg::RENDER_SHADOWRAYS, 0 );
	}

	char msg[64];
	snprintf( msg, 64, "[LightParser] Loaded %lu light(s).", (unsigned long) mLights.size() );
	Logger::logInfo( msg );
}
