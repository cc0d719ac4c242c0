/*
 * Cfg -- the configuration singleton of the reference (source/Cfg.h, source/Cfg.cpp) without Boost.
 *
 * Same surface: Cfg::get(), loadConfigFile(path), value<T>(key), value(key, newValue), and the 36
 * dotted key constants (Cfg.cpp:4-39, kept verbatim).  The reference reads config.json with
 * boost::property_tree::read_json, which tolerates `//` comments in this file; the reader here is a
 * small JSON parser with `//` and `/ * * /` comments that flattens objects into dotted keys, which is
 * what property_tree's get<T>("a.b.c") resolves.
 */
#ifndef CFG_H
#define CFG_H

#include <map>
#include <sstream>
#include <stdexcept>
#include <string>

class Cfg {

	public:
		static Cfg& get() {
			static Cfg instance;
			return instance;
		}
		void loadConfigFile( const char* filepath );
		void loadConfigString( const std::string& json );
		/** Reset to the defaults of the reference's shipped config.json. */
		void loadDefaults();
		bool has( const char* key ) const { return mValues.count( key ) > 0; }

		template<typename T> T value( const char* key ) {
			std::map<std::string, std::string>::const_iterator it = mValues.find( key );
			if( it == mValues.end() ) {
				throw std::runtime_error( std::string( "[Cfg] No such node (" ) + key + ")" );
			}
			return convert<T>( it->second, key );
		}
		template<typename T> void value( const char* key, T newValue ) {
			std::ostringstream os;
			os.precision( 9 );
			os << newValue;
			mValues[key] = os.str();
		}

		/* the dotted keys of config.json (Cfg.cpp:4-39) */
		static const char *ACCEL_STRUCT, *IMPORT_PATH, *INFO_KERNELTIMES, *LOG_LEVEL;
		static const char *BVH_MAXFACES, *BVH_SAHFACESLIMIT, *BVH_SKIPAHEAD, *BVH_SKIPAHEAD_CMP;
		static const char *CAM_EYE_X, *CAM_EYE_Y, *CAM_EYE_Z, *CAM_CENTER_X, *CAM_CENTER_Y, *CAM_CENTER_Z;
		static const char *CAM_LENSE_APERTURE, *CAM_LENSE_FOCALLENGTH, *CAM_SPEED, *PERS_FOV, *PERS_ZFAR, *PERS_ZNEAR;
		static const char *OPENCL_BUILDOPTIONS, *OPENCL_CHECKERRORS, *OPENCL_LOCALGROUPSIZE, *OPENCL_PROGRAM;
		static const char *RENDER_ANTIALIAS, *RENDER_BRDF, *RENDER_INTERVAL, *RENDER_MAXADDEDDEPTH, *RENDER_MAXDEPTH;
		static const char *RENDER_PHONGTESS, *RENDER_SAMPLES, *RENDER_SHADOWRAYS;
		static const char *SHADER_NAME, *SHADER_PATH, *WINDOW_HEIGHT, *WINDOW_WIDTH;

	private:
		Cfg() { loadDefaults(); }
		Cfg( Cfg const& );
		void operator=( Cfg const& );

		template<typename T> static T convert( const std::string& s, const char* key );

		std::map<std::string, std::string> mValues;

};

/* property_tree's stream translator: numbers via operator>>, bool accepts true/false/1/0. */
template<typename T> T Cfg::convert( const std::string& s, const char* key ) {
	std::istringstream is( s );
	T v;
	is >> v;
	if( is.fail() ) {
		throw std::runtime_error( std::string( "[Cfg] conversion of data to type failed (" ) + key + " = \"" + s + "\")" );
	}
	return v;
}
template<> inline std::string Cfg::convert<std::string>( const std::string& s, const char* ) { return s; }
template<> inline bool Cfg::convert<bool>( const std::string& s, const char* key ) {
	if( s == "true" || s == "1" ) { return true; }
	if( s == "false" || s == "0" ) { return false; }
	throw std::runtime_error( std::string( "[Cfg] conversion of data to type failed (" ) + key + ")" );
}
/* operator>> on (unsigned) char types would read one character; property_tree reads shorts etc. as numbers */
template<> inline short Cfg::convert<short>( const std::string& s, const char* key ) { return (short) convert<int>( s, key ); }

#endif
