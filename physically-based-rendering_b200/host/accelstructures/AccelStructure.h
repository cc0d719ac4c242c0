/* AccelStructure -- reference: source/accelstructures/AccelStructure.{h,cpp}. */
#ifndef ACCELSTRUCT_H
#define ACCELSTRUCT_H

#define ACCELSTRUCT_BVH 0

#include <vector>

#include "../cl_types.h"
#include "../glm_lite.h"

using std::vector;


struct Tri {
	cl_uint4 face;      // w: global face index
	cl_uint4 normals;   // w: global index into facesVN
	glm::vec3 bbMin;
	glm::vec3 bbMax;
};


class AccelStructure {

	public:
		virtual ~AccelStructure() {}
		/** Pack xyz triples as float4 with w = 0 (reference: AccelStructure.cpp:11-25). */
		static vector<cl_float4> packFloatAsFloat4( const vector<cl_float>* vertices ) {
			vector<cl_float4> vertices4( vertices->size() / 3 );
			for( size_t i = 0; i < vertices4.size(); i++ ) {
				cl_float4 v = { ( *vertices )[3 * i], ( *vertices )[3 * i + 1], ( *vertices )[3 * i + 2], 0.0f };
				vertices4[i] = v;
			}
			return vertices4;
		}
		/** Line-list geometry of the structure for an overlay (reference: BVH::visualize). */
		virtual void visualize( vector<cl_float>* vertices, vector<cl_uint>* indices ) = 0;

};

#endif
