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

struct light_t {
	std::string lightName;
	cl_uint type;                /* 1 point light, 2 orb light */
	cl_float4 pos, rgb;
	cl_float radius;
};

class LightParser {
	public:
		void load( std::string file );
		std::vector<light_t> getLights();
		static light_t getEmptyLight();
		/** Additive: install lights that were not read from a file (synthetic scenes). */
		void setLights( const std::vector<light_t>& lights );

	private:
		std::vector<light_t> mLights;
};

#endif
