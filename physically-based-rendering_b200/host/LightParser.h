/*
 * LightParser -- reference: source/LightParser.{h,cpp}.  Reads the sibling <name>.lights file
 * (newlight / type / pos / rgb / radius).  When the file holds no light the reference switches
 * shadow rays off in the configuration (LightParser.cpp:119-121); kept.
 */
#ifndef LIGHTPARSER_H
#define LIGHTPARSER_H

#include <string>
#include <vector>

#include "cl_types.h"
#include "Logger.h"

using std::string;
using std::vector;

// Light types:
// 1: Point light
// 2: Orb light
struct light_t {
	string lightName;
	cl_uint type;
	cl_float4 pos;
	cl_float4 rgb;
	cl_float radius;
};


class LightParser {

	public:
		vector<light_t> getLights();
		void load( string file );
		/** Additive: install lights that were not read from a file (synthetic scenes). */
		void setLights( const vector<light_t>& lights );
		static light_t getEmptyLight();

	private:
		vector<light_t> mLights;

};

#endif
