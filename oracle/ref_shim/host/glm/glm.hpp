/*
 * glm/glm.hpp stand-in: the part of GLM the reference's BVH builder and MathHelp use -- glm::vec3 with
 * component access and arithmetic, dot, cross, normalize, min, max, abs -- computing what GLM computes:
 * plain binary32 operations, dot = (x*x' + y*y') + z*z', normalize(v) = v * (1 / sqrt(dot(v, v))).
 * TEST INFRASTRUCTURE (oracle/build_ref_host.py).
 */
#ifndef PBR_REF_GLM_HPP
#define PBR_REF_GLM_HPP

#include <cmath>

namespace glm {

struct vec2 {
	float x, y;
	vec2() : x(0.0f), y(0.0f) {}
	vec2(float a, float b) : x(a), y(b) {}
	float& operator[](int i) { return (&x)[i]; }
	const float& operator[](int i) const { return (&x)[i]; }
};

struct vec3 {
	float x, y, z;
	vec3() : x(0.0f), y(0.0f), z(0.0f) {}
	explicit vec3(float s) : x(s), y(s), z(s) {}
	vec3(float a, float b, float c) : x(a), y(b), z(c) {}
	float& operator[](int i) { return (&x)[i]; }
	const float& operator[](int i) const { return (&x)[i]; }
	vec3& operator+=(const vec3& o) { x += o.x; y += o.y; z += o.z; return *this; }
	vec3& operator-=(const vec3& o) { x -= o.x; y -= o.y; z -= o.z; return *this; }
	vec3& operator*=(float s) { x *= s; y *= s; z *= s; return *this; }
	vec3& operator/=(float s) { x /= s; y /= s; z /= s; return *this; }
};

inline vec3 operator+(const vec3& a, const vec3& b) { return vec3(a.x + b.x, a.y + b.y, a.z + b.z); }
inline vec3 operator-(const vec3& a, const vec3& b) { return vec3(a.x - b.x, a.y - b.y, a.z - b.z); }
inline vec3 operator*(const vec3& a, const vec3& b) { return vec3(a.x * b.x, a.y * b.y, a.z * b.z); }
inline vec3 operator/(const vec3& a, const vec3& b) { return vec3(a.x / b.x, a.y / b.y, a.z / b.z); }
inline vec3 operator+(const vec3& a, float s) { return vec3(a.x + s, a.y + s, a.z + s); }
inline vec3 operator-(const vec3& a, float s) { return vec3(a.x - s, a.y - s, a.z - s); }
inline vec3 operator*(const vec3& a, float s) { return vec3(a.x * s, a.y * s, a.z * s); }
inline vec3 operator/(const vec3& a, float s) { return vec3(a.x / s, a.y / s, a.z / s); }
inline vec3 operator+(float s, const vec3& a) { return vec3(s + a.x, s + a.y, s + a.z); }
inline vec3 operator-(float s, const vec3& a) { return vec3(s - a.x, s - a.y, s - a.z); }
inline vec3 operator*(float s, const vec3& a) { return vec3(s * a.x, s * a.y, s * a.z); }
inline vec3 operator/(float s, const vec3& a) { return vec3(s / a.x, s / a.y, s / a.z); }
inline vec3 operator-(const vec3& a) { return vec3(-a.x, -a.y, -a.z); }
inline bool operator==(const vec3& a, const vec3& b) { return a.x == b.x && a.y == b.y && a.z == b.z; }
inline bool operator!=(const vec3& a, const vec3& b) { return !(a == b); }

inline float min(float a, float b) { return (b < a) ? b : a; }
inline float max(float a, float b) { return (a < b) ? b : a; }
inline vec3 min(const vec3& a, const vec3& b) { return vec3(min(a.x, b.x), min(a.y, b.y), min(a.z, b.z)); }
inline vec3 max(const vec3& a, const vec3& b) { return vec3(max(a.x, b.x), max(a.y, b.y), max(a.z, b.z)); }
inline float abs(float a) { return ::fabsf(a); }
inline vec3 abs(const vec3& a) { return vec3(::fabsf(a.x), ::fabsf(a.y), ::fabsf(a.z)); }

inline float dot(const vec3& a, const vec3& b) { return (a.x * b.x + a.y * b.y) + a.z * b.z; }
inline vec3 cross(const vec3& a, const vec3& b) {
	return vec3(a.y * b.z - b.y * a.z, a.z * b.x - b.z * a.x, a.x * b.y - b.x * a.y);
}
inline float length(const vec3& a) { return ::sqrtf(dot(a, a)); }
inline vec3 normalize(const vec3& a) { return a * (1.0f / ::sqrtf(dot(a, a))); }

} /* namespace glm */

#endif
