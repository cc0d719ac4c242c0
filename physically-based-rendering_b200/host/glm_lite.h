/*
 * glm_lite.h -- the subset of GLM the reference's host code touches (glm::vec3 with +,-,*,/ and
 * operator[], glm::min/max/dot/cross/normalize), with GLM's exact formulas so that float results
 * are the same:  min(x,y) = (y < x) ? y : x;  max(x,y) = (x < y) ? y : x;
 * dot = x*x' + y*y' + z*z' (left to right);  normalize(v) = v * (1 / sqrt(dot(v,v))).
 */
#ifndef PBR_HOST_GLM_LITE_H
#define PBR_HOST_GLM_LITE_H

#include <math.h>

namespace glm {

struct vec3 {
	float x, y, z;
	vec3() : x( 0.0f ), y( 0.0f ), z( 0.0f ) {}      /* GLM 0.9.x zero-initialises */
	vec3( float a, float b, float c ) : x( a ), y( b ), z( c ) {}
	float& operator[]( int i ) { return ( &x )[i]; }
	const float& operator[]( int i ) const { return ( &x )[i]; }
};

struct vec2 {
	float x, y;
	vec2() : x( 0.0f ), y( 0.0f ) {}
};

inline vec3 operator+( const vec3& a, const vec3& b ) { return vec3( a.x + b.x, a.y + b.y, a.z + b.z ); }
inline vec3 operator-( const vec3& a, const vec3& b ) { return vec3( a.x - b.x, a.y - b.y, a.z - b.z ); }
inline vec3 operator*( const vec3& a, float s ) { return vec3( a.x * s, a.y * s, a.z * s ); }
inline vec3 operator*( float s, const vec3& a ) { return vec3( s * a.x, s * a.y, s * a.z ); }
inline vec3 operator/( const vec3& a, float s ) { return vec3( a.x / s, a.y / s, a.z / s ); }

inline float min( float x, float y ) { return ( y < x ) ? y : x; }
inline float max( float x, float y ) { return ( x < y ) ? y : x; }
inline vec3 min( const vec3& a, const vec3& b ) { return vec3( min( a.x, b.x ), min( a.y, b.y ), min( a.z, b.z ) ); }
inline vec3 max( const vec3& a, const vec3& b ) { return vec3( max( a.x, b.x ), max( a.y, b.y ), max( a.z, b.z ) ); }
inline float dot( const vec3& a, const vec3& b ) { return a.x * b.x + a.y * b.y + a.z * b.z; }
inline vec3 cross( const vec3& x, const vec3& y ) {
	return vec3( x.y * y.z - y.y * x.z, x.z * y.x - y.z * x.x, x.x * y.y - y.x * x.y );
}
inline vec3 normalize( const vec3& v ) { return v * ( 1.0f / sqrtf( dot( v, v ) ) ); }

}

#endif
