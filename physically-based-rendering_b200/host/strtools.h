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

#include <charconv>
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
	const char d0 = delims[0];
	const char d1 = delims[1] ? delims[1] : d0;
	const bool two = ( delims[0] == '\0' || delims[1] == '\0' || delims[2] == '\0' );   /* at most two delimiters: no strchr */
	for( const char* c = b; c < e; c++ ) {
		if( two ? ( *c == d0 || *c == d1 ) : ( strchr( delims, *c ) != NULL ) ) {
			Token t = { start, (size_t) ( c - start ) };
			out.push_back( t );
			start = c + 1;
		}
	}
	Token t = { start, (size_t) ( e - start ) };
	out.push_back( t );
}

/** atof() of a token (the token is not NUL terminated in the buffer).  A token that is one plain decimal number
 *  from its first to its last character goes through std::from_chars -- correctly rounded like strtod, so the
 *  same double, about ten times faster; everything else (leading blanks or '+', trailing garbage, hex, inf / nan)
 *  keeps atof's own rules. */
inline double toDouble( const Token& t ) {
	char buf[64];
	if( t.n == 0 ) { return 0.0; }
	{
		const char c = t.p[0];
		if( ( c >= '0' && c <= '9' ) || ( ( c == '-' || c == '.' ) && t.n > 1 && t.p[1] != 'x' && t.p[1] != 'X' ) ) {
			bool plain = !( c == '0' && t.n > 1 && ( t.p[1] == 'x' || t.p[1] == 'X' ) );
			if( plain && c == '-' && t.n > 2 && t.p[1] == '0' && ( t.p[2] == 'x' || t.p[2] == 'X' ) ) { plain = false; }
			if( plain ) {
				double d = 0.0;
				const std::from_chars_result r = std::from_chars( t.p, t.p + t.n, d, std::chars_format::general );
				if( r.ec == std::errc() && r.ptr == t.p + t.n ) { return d; }
			}
		}
	}
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
