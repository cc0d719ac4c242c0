/*
 * MathHelp -- reference: source/MathHelp.{h,cpp}; only the helpers that feed the hot path
 * (SURVEY.md section 2, row 5): degToRad, getAABB (both overloads), getSurfaceArea, triCalcAABB with
 * its Phong-tessellation growth (triThicknessAndSidedrop, phongTessellate, projectOnPlane),
 * longestAxis and radToDeg.
 */
#ifndef MATHHELP_H
#define MATHHELP_H

#include <vector>

#include "cl_types.h"
#include "glm_lite.h"
#include "Cfg.h"
#include "accelstructures/AccelStructure.h"

using std::vector;

#define MH_PI 3.14159265359                 /* a double: degToRad( float ) is evaluated in binary64 */
#define FLOAT4_TO_VEC3( v ) ( glm::vec3( ( v ).x, ( v ).y, ( v ).z ) )

class MathHelp {
	public:
		/* --- angles */
		static cl_float degToRad( cl_float deg );
		static cl_float radToDeg( cl_float rad );

		/* --- boxes */
		static void getAABB( const vector<cl_float4>& vertices, glm::vec3* bbMin, glm::vec3* bbMax );
		static void getAABB( const vector<glm::vec3>& bbMins, const vector<glm::vec3>& bbMaxs, glm::vec3* bbMin, glm::vec3* bbMax );
		static cl_float getSurfaceArea( const glm::vec3& bbMin, const glm::vec3& bbMax );
		static short longestAxis( glm::vec3 bbMin, glm::vec3 bbMax );

		/* --- a triangle's box; with render.phong_tessellation > 0 grown by the bulge of the curved patch */
		static void triCalcAABB( Tri* tri, const vector<cl_float4>* vertices, const vector<cl_float4>* normals );
		static void triCalcAABB( Tri* tri, const vector<cl_float4>* vertices, const vector<cl_float4>* normals, float phongTessAlpha );
		static void triThicknessAndSidedrop(
			const float alpha,
			const glm::vec3 p1, const glm::vec3 p2, const glm::vec3 p3, const glm::vec3 n1, const glm::vec3 n2, const glm::vec3 n3,
			float* thickness, glm::vec3* sidedropMin, glm::vec3* sidedropMax
		);
		static glm::vec3 phongTessellate(
			const glm::vec3 p1, const glm::vec3 p2, const glm::vec3 p3, const glm::vec3 n1, const glm::vec3 n2, const glm::vec3 n3,
			const float alpha, const float u, const float v
		);
		static glm::vec3 projectOnPlane( glm::vec3 q, glm::vec3 p, glm::vec3 n );
};

#endif
