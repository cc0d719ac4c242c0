/* AccelStructure -- reference: source/accelstructures/AccelStructure.{h,cpp}: the base of BVH, and the Tri record. */
#ifndef ACCELSTRUCT_H
#define ACCELSTRUCT_H

#include <vector>

#include "../cl_types.h"
#include "../glm_lite.h"

#define ACCELSTRUCT_BVH 0      /* the only value config.json's accel_struct knows */

using std::vector;

/* A triangle on its way through the builder: vertex indices (w = face number of the scene), normal indices
 * (w = the same face in facesVN) and its box, grown for Phong tessellation when that is on. */
struct Tri {
	cl_uint4 face, normals;
	glm::vec3 bbMin, bbMax;
};

class AccelStructure {
	public:
		virtual ~AccelStructure() {}

		/** Line-list geometry of the structure for an overlay (reference: BVH::visualize). */
		virtual void visualize( vector<cl_float>* vertices, vector<cl_uint>* indices ) = 0;

		/** xyz triples -> float4 with w = 0 (reference: AccelStructure.cpp:11-25). */
		static vector<cl_float4> packFloatAsFloat4( const vector<cl_float>* vertices ) {
			const size_t count = vertices->size() / 3;
			vector<cl_float4> packed( count );
			for( size_t i = 0; i < count; i++ ) {
				const cl_float* xyz = &( *vertices )[3 * i];
				cl_float4 v = { xyz[0], xyz[1], xyz[2], 0.0f };
				packed[i] = v;
			}
			return packed;
		}
};

#endif
