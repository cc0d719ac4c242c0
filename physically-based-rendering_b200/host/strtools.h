/*
 * strtools -- the two Boost string algorithms the reference's parsers rely on, restated for
 * in-place use on a character buffer:
 *     boost::algorithm::trim                          (ObjParser.cpp:149, MtlParser.cpp:70, LightParser.cpp:56)
 *     boost::split( parts, line, is_any_of(" \t") )   (token_compress_off: every delimiter ends a token,
 *                                                       so runs of blanks yield EMPTY tokens)
 * Tokens are (pointer, length) views into the line; nothing is allocated per line.
 */
#ifndef PBR_HOST_STRTOOLS_H
#define PBR_HOST_STRTOOLS_H

#include <ctype.h>
#include <stdlib.h>
#include <string.h>
#include <string>
#include <vector>

namespace strtools {

struct Token {
	const char* p;
	size_t n;
	std::string str() const { return std::string( p, n ); }
	bool equals( const char* s ) const { return strlen( s ) == n && memcmp( p, s, n ) == 0; }
};

/** Trim [b, e) on both sides (isspace in the "C" locale). */
inline void trim( const char*& b, const char*& e ) {
	while( b < e && isspace( (unsigned char) *b ) ) { b++; }
	while( e > b && isspace( (unsigned char) e[-1] ) ) { e--; }
}

/** Split [b, e) at every character contained in `delims`; always yields at least one token. */
inline void split( std::vector<Token>& out, const char* b, const char* e, const char* delims ) {
	out.clear();
	const char* start = b;
	for( const char* c = b; c < e; c++ ) {
		if( strchr( delims, *c ) != NULL ) {
			Token t = { start, (size_t) ( c - start ) };
			out.push_back( t );
			start = c + 1;
		}
	}
	Token t = { start, (size_t) ( e - start ) };
	out.push_back( t );
}

/** atof() of a token (the token is not NUL terminated in the buffer). */
inline double toDouble( const Token& t ) {
	char buf[64];
	if( t.n == 0 ) { return 0.0; }
	if( t.n < sizeof( buf ) ) {
		memcpy( buf, t.p, t.n );
		buf[t.n] = 0;
		return atof( buf );
	}
	return atof( t.str().c_str() );
}

/** atol() of a token. */
inline long toLong( const Token& t ) {
	char buf[32];
	if( t.n == 0 ) { return 0; }
	if( t.n < sizeof( buf ) ) {
		memcpy( buf, t.p, t.n );
		buf[t.n] = 0;
		return atol( buf );
	}
	return atol( t.str().c_str() );
}

}

#endif
