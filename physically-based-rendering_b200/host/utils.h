/* utils -- reference: source/utils.h (formatBytes, loadFileAsString). */
#ifndef UTILS_H
#define UTILS_H

#include <fstream>
#include <sstream>
#include <string>

namespace utils {

	/** Format a value of bytes into more readable units (utils.h:18-35). */
	inline void formatBytes( size_t bytes, float* bytesFloat, std::string* unit ) {
		static const char* units[] = { "bytes", "KB", "MB", "GB" };
		int u = 0;
		float v = (float) bytes;
		while( v >= 1024.0f && u < 3 ) {
			v /= 1024.0f;
			u++;
		}
		*bytesFloat = v;
		*unit = units[u];
	}

	/** Read the contents of a file as string (utils.h:42-55). */
	inline std::string loadFileAsString( const char* filename ) {
		std::ifstream fileIn( filename );
		std::stringstream ss;
		ss << fileIn.rdbuf();
		return ss.str();
	}

}

#endif
