/*
 * bvh_oracle.cpp -- CPU ORACLE (test infrastructure, NOT product code).
 *
 * Literal restatement of the reference's host-side acceleration-structure path:
 *     source/accelstructures/BVH.cpp      (build, ordering, skip-ahead)
 *     source/MathHelp.cpp                 (AABBs, surface area, Phong-tessellation growth)
 *     source/ModelLoader.cpp:28-57        (per-object face lists)
 *     source/PathTracer.cpp:238-347       (flattening into bvhNode_cl[] + leaf-ordered faces)
 * It deliberately keeps the reference's structure (recursive, vectors copied by value, one
 * std::sort per axis per node) so that it is easy to audit against the source; the product's
 * builder (host/BVH.cpp) is an allocation-free design that must produce the same arrays.
 *
 * PARITY STATUS: PINNED against outputs of the reference itself.  oracle/build_ref_host.py compiles the
 * reference's own BVH.cpp / MathHelp.cpp / ModelLoader.cpp / parsers from /root/reference/source against
 * stand-ins for the Boost / GLM / OpenCL headers they include (oracle/ref_shim/host/) into
 * oracle/_ref/libref_host.so; tests/test_oracle_vs_reference_host.py requires this restatement -- and the
 * product's builder -- to produce the same flattened node array, leaf-ordered faces and counts, bit for
 * bit, for the bundled models under 7 builder configurations and for generated scenes (SAH sweep,
 * mean split, 16 objects with tied centres, Phong-tessellation boxes).  The flattening step itself
 * (PathTracer.cpp:238-347, in a file that needs Qt + OpenCL) is restated on both sides.
 * Soft pin kept: suzanne.obj has 1082 faces (pathtracing.cl:75).
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's baseline legs may load this library.
 */
#include <algorithm>
#include <float.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <vector>

#include "../include/pbr_types.h"

namespace {

struct V3 {
	float v[3];
	float& operator[](int i) { return v[i]; }
	float operator[](int i) const { return v[i]; }
};
inline V3 mk(float x, float y, float z) { V3 r; r.v[0] = x; r.v[1] = y; r.v[2] = z; return r; }
inline V3 operator+(V3 a, V3 b) { return mk(a[0] + b[0], a[1] + b[1], a[2] + b[2]); }
inline V3 operator-(V3 a, V3 b) { return mk(a[0] - b[0], a[1] - b[1], a[2] - b[2]); }
inline V3 operator*(V3 a, float s) { return mk(a[0] * s, a[1] * s, a[2] * s); }
inline V3 operator*(float s, V3 a) { return mk(s * a[0], s * a[1], s * a[2]); }
inline V3 operator/(V3 a, float s) { return mk(a[0] / s, a[1] / s, a[2] / s); }
/* glm::min(x, y) = (y < x) ? y : x;  glm::max(x, y) = (x < y) ? y : x */
inline float gmin(float x, float y) { return (y < x) ? y : x; }
inline float gmax(float x, float y) { return (x < y) ? y : x; }
inline V3 vmin(V3 a, V3 b) { return mk(gmin(a[0], b[0]), gmin(a[1], b[1]), gmin(a[2], b[2])); }
inline V3 vmax(V3 a, V3 b) { return mk(gmax(a[0], b[0]), gmax(a[1], b[1]), gmax(a[2], b[2])); }
inline float gdot(V3 a, V3 b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }
inline V3 gcross(V3 x, V3 y) {
	return mk(x[1] * y[2] - y[1] * x[2], x[2] * y[0] - y[2] * x[0], x[0] * y[1] - y[0] * x[1]);
}
inline V3 gnormalize(V3 a) { return a * (1.0f / sqrtf(gdot(a, a))); }

/* accelstructures/AccelStructure.h:13-18 */
struct Tri {
	pbr_uint4 face;
	pbr_uint4 normals;
	V3 bbMin;
	V3 bbMax;
};

/* accelstructures/BVH.h:16-27 */
struct BVHNode {
	BVHNode* leftChild;
	BVHNode* rightChild;
	BVHNode* parent;
	std::vector<Tri> faces;
	V3 bbMin;
	V3 bbMax;
	uint32_t id;
	uint32_t depth;
	uint32_t numSkipsToHere;
	bool skipNextLeft;
};

struct Config {
	uint32_t maxFaces;
	uint32_t sahFacesLimit;
	bool skipAhead;
	float skipAheadCmp;
	float phongTess;
};

/* ---------------------------------------------------------------- MathHelp.cpp */

/* MathHelp.cpp:95-101 */
float getSurfaceArea(V3 bbMin, V3 bbMax) {
	float xy = fabsf(bbMax[0] - bbMin[0]) * fabsf(bbMax[1] - bbMin[1]);
	float zy = fabsf(bbMax[2] - bbMin[2]) * fabsf(bbMax[1] - bbMin[1]);
	float xz = fabsf(bbMax[0] - bbMin[0]) * fabsf(bbMax[2] - bbMin[2]);
	return 2.0f * (xy + zy + xz);
}

/* MathHelp.cpp:20-36 */
void getAABB_vertices(const pbr_float4* vertices, size_t n, V3* bbMin, V3* bbMax) {
	*bbMin = mk(vertices[0].x, vertices[0].y, vertices[0].z);
	*bbMax = mk(vertices[0].x, vertices[0].y, vertices[0].z);
	for (size_t i = 1; i < n; i++) {
		pbr_float4 v = vertices[i];
		(*bbMin)[0] = ((*bbMin)[0] < v.x) ? (*bbMin)[0] : v.x;
		(*bbMin)[1] = ((*bbMin)[1] < v.y) ? (*bbMin)[1] : v.y;
		(*bbMin)[2] = ((*bbMin)[2] < v.z) ? (*bbMin)[2] : v.z;
		(*bbMax)[0] = ((*bbMax)[0] > v.x) ? (*bbMax)[0] : v.x;
		(*bbMax)[1] = ((*bbMax)[1] > v.y) ? (*bbMax)[1] : v.y;
		(*bbMax)[2] = ((*bbMax)[2] > v.z) ? (*bbMax)[2] : v.z;
	}
}

/* MathHelp.cpp:46-65.  The reference indexes element 0 of an empty vector when a side is
 * empty; here an empty list yields a zero box (only reachable on the logged error path). */
void getAABB_boxes(const std::vector<V3>& bbMins, const std::vector<V3>& bbMaxs, V3* bbMin, V3* bbMax) {
	if (bbMins.empty()) { *bbMin = mk(0, 0, 0); *bbMax = mk(0, 0, 0); return; }
	*bbMin = bbMins[0];
	*bbMax = bbMaxs[0];
	for (size_t i = 1; i < bbMins.size(); i++) {
		(*bbMin)[0] = gmin(bbMins[i][0], (*bbMin)[0]);
		(*bbMin)[1] = gmin(bbMins[i][1], (*bbMin)[1]);
		(*bbMin)[2] = gmin(bbMins[i][2], (*bbMin)[2]);
		(*bbMax)[0] = gmax(bbMaxs[i][0], (*bbMax)[0]);
		(*bbMax)[1] = gmax(bbMaxs[i][1], (*bbMax)[1]);
		(*bbMax)[2] = gmax(bbMaxs[i][2], (*bbMax)[2]);
	}
}

/* MathHelp.cpp:227-229 */
V3 projectOnPlane(V3 q, V3 p, V3 n) { return q - gdot(q - p, n) * n; }

/* MathHelp.cpp:211-224 */
V3 phongTessellate(V3 p1, V3 p2, V3 p3, V3 n1, V3 n2, V3 n3, float alpha, float u, float v) {
	float w = 1.0f - u - v;
	V3 pBary = p1 * u + p2 * v + p3 * w;
	V3 pTessellated =
		u * projectOnPlane(pBary, p1, n1) +
		v * projectOnPlane(pBary, p2, n2) +
		w * projectOnPlane(pBary, p3, n3);
	return (1.0f - alpha) * pBary + alpha * pTessellated;
}

/* MathHelp.cpp:324-378 */
void triThicknessAndSidedrop(
	float alpha, V3 p1, V3 p2, V3 p3, V3 n1, V3 n2, V3 n3,
	float* thickness, V3* sidedropMin, V3* sidedropMax
) {
	V3 e12 = p2 - p1;
	V3 e13 = p3 - p1;
	V3 e23 = p3 - p2;
	V3 e31 = p1 - p3;
	V3 c12 = alpha * (gdot(n2, e12) * n2 - gdot(n1, e12) * n1);
	V3 c23 = alpha * (gdot(n3, e23) * n3 - gdot(n2, e23) * n2);
	V3 c31 = alpha * (gdot(n1, e31) * n1 - gdot(n3, e31) * n3);
	V3 ng = gnormalize(gcross(e12, e13));

	float k_tmp = gdot(ng, c12 - c23 - c31);
	float k = 1.0f / (4.0f * gdot(ng, c23) * gdot(ng, c31) - k_tmp * k_tmp);

	float u = k * (
		2.0f * gdot(ng, c23) * gdot(ng, c31 + e31) +
		gdot(ng, c23 - e23) * gdot(ng, c12 - c23 - c31)
	);
	float v = k * (
		2.0f * gdot(ng, c31) * gdot(ng, c23 - e23) +
		gdot(ng, c31 + e31) * gdot(ng, c12 - c23 - c31)
	);

	u = (u < 0.0f || u > 1.0f) ? 0.0f : u;
	v = (v < 0.0f || v > 1.0f) ? 0.0f : v;

	V3 pt = phongTessellate(p1, p2, p3, n1, n2, n3, alpha, u, v);
	*thickness = gdot(ng, pt - p1);

	const float uv[9][2] = {
		{0.0f, 0.5f}, {0.5f, 0.0f}, {0.5f, 0.5f}, {0.25f, 0.75f}, {0.75f, 0.25f},
		{0.25f, 0.0f}, {0.75f, 0.0f}, {0.0f, 0.25f}, {0.0f, 0.75f}
	};
	V3 ptsd[9];
	for (int i = 0; i < 9; i++) ptsd[i] = phongTessellate(p1, p2, p3, n1, n2, n3, alpha, uv[i][0], uv[i][1]);

	*sidedropMin = ptsd[0];
	*sidedropMax = ptsd[0];
	for (int i = 1; i < 9; i++) {
		*sidedropMin = vmin(*sidedropMin, ptsd[i]);
		*sidedropMax = vmax(*sidedropMax, ptsd[i]);
	}
}

/* MathHelp.cpp:250-310 */
void triCalcAABB(const Config& cfg, Tri* tri, const std::vector<pbr_float4>* vertices, const std::vector<pbr_float4>* normals) {
	pbr_float4 v[3] = { (*vertices)[tri->face.x], (*vertices)[tri->face.y], (*vertices)[tri->face.z] };

	V3 bbMin, bbMax;
	getAABB_vertices(v, 3, &bbMin, &bbMax);
	tri->bbMin = bbMin;
	tri->bbMax = bbMax;

	if (cfg.phongTess <= 0.0f) {
		return;
	}

	V3 p1 = mk(v[0].x, v[0].y, v[0].z);
	V3 p2 = mk(v[1].x, v[1].y, v[1].z);
	V3 p3 = mk(v[2].x, v[2].y, v[2].z);

	pbr_float4 fn1 = (*normals)[tri->normals.x];
	pbr_float4 fn2 = (*normals)[tri->normals.y];
	pbr_float4 fn3 = (*normals)[tri->normals.z];

	V3 n1 = mk(fn1.x, fn1.y, fn1.z);
	V3 n2 = mk(fn2.x, fn2.y, fn2.z);
	V3 n3 = mk(fn3.x, fn3.y, fn3.z);

	V3 test = (n1 - n2) + (n2 - n3);
	if (fabsf(test[0]) <= 0.000001f && fabsf(test[1]) <= 0.000001f && fabsf(test[2]) <= 0.000001f) {
		return;
	}

	float thickness;
	V3 sidedropMin, sidedropMax;
	triThicknessAndSidedrop(cfg.phongTess, p1, p2, p3, n1, n2, n3, &thickness, &sidedropMin, &sidedropMax);

	V3 e12 = p2 - p1;
	V3 e13 = p3 - p1;
	V3 ng = gnormalize(gcross(e12, e13));

	V3 p1thick = p1 + thickness * ng;
	V3 p2thick = p2 + thickness * ng;
	V3 p3thick = p3 + thickness * ng;

	tri->bbMin = vmin(vmin(tri->bbMin, p1thick), vmin(p2thick, p3thick));
	tri->bbMax = vmax(vmax(tri->bbMax, p1thick), vmax(p2thick, p3thick));
	tri->bbMin = vmin(tri->bbMin, sidedropMin);
	tri->bbMax = vmax(tri->bbMax, sidedropMax);
}

/* ---------------------------------------------------------------------- BVH.cpp */

/* BVH.cpp:9-35 */
struct sortFacesCmp {
	uint32_t axis;
	explicit sortFacesCmp(const uint32_t axis) { this->axis = axis; }
	bool operator()(const Tri a, const Tri b) {
		float cenA = (a.bbMin[this->axis] + a.bbMax[this->axis]) * 0.5f;
		float cenB = (b.bbMin[this->axis] + b.bbMax[this->axis]) * 0.5f;
		return cenA < cenB;
	}
};

struct object3D {
	std::vector<uint32_t> facesV;
	std::vector<uint32_t> facesVN;
};

class BVH {
public:
	Config cfg;
	std::vector<BVHNode*> mContainerNodes;
	std::vector<BVHNode*> mLeafNodes;
	std::vector<BVHNode*> mNodes;
	std::vector<BVHNode*> mAllocated;
	BVHNode* mRoot;
	uint32_t mMaxFaces;
	uint32_t mDepthReached;
	uint32_t mSkipped;

	/* flattened output (PathTracer.cpp:238-347) */
	std::vector<pbr_bvh_node> flatNodes;
	std::vector<pbr_uint4> flatFacesV;
	std::vector<pbr_uint4> flatFacesN;

	~BVH() { for (BVHNode* n : mAllocated) delete n; }

	BVHNode* newNode() {
		BVHNode* node = new BVHNode();
		mAllocated.push_back(node);
		node->leftChild = NULL;
		node->rightChild = NULL;
		node->parent = NULL;
		node->depth = 0;
		node->id = 0;
		node->skipNextLeft = false;
		node->numSkipsToHere = 0;
		return node;
	}

	/* BVH.cpp:50-64 */
	void build(const std::vector<object3D>& sceneObjects, const std::vector<float>& vertices, const std::vector<float>& normals) {
		mDepthReached = 0;
		mSkipped = 0;
		mMaxFaces = (uint32_t) fmax((int) cfg.maxFaces, 1);   /* BVH.cpp:759-763 */

		std::vector<BVHNode*> subTrees = buildTreesFromObjects(&sceneObjects, &vertices, &normals);
		mRoot = makeContainerNode(subTrees, true);
		groupTreesToNodes(subTrees, mRoot, mDepthReached);
		combineNodes((uint32_t) subTrees.size());
	}

	/* ModelLoader.cpp:28-41 / :44-57 */
	static void getFacesOfObject(const std::vector<uint32_t>& objFaces, std::vector<pbr_uint4>* faces, int32_t offset) {
		for (uint32_t i = 0; i < objFaces.size(); i += 3) {
			pbr_uint4 f = { objFaces[i + 0], objFaces[i + 1], objFaces[i + 2], (uint32_t) (offset + (int32_t) faces->size()) };
			faces->push_back(f);
		}
	}

	/* BVH.cpp:735-750 */
	static std::vector<pbr_float4> packFloatAsFloat4(const std::vector<float>* vertices) {
		std::vector<pbr_float4> vertices4;
		for (uint32_t i = 0; i + 2 < vertices->size(); i += 3) {
			pbr_float4 v = { (*vertices)[i + 0], (*vertices)[i + 1], (*vertices)[i + 2], 0.0f };
			vertices4.push_back(v);
		}
		return vertices4;
	}

	/* BVH.cpp:363-380 */
	std::vector<Tri> facesToTriStructs(
		const std::vector<pbr_uint4>* facesThisObj, const std::vector<pbr_uint4>* faceNormalsThisObj,
		const std::vector<pbr_float4>* vertices4, const std::vector<pbr_float4>* normals4
	) {
		std::vector<Tri> triFaces;
		for (uint32_t j = 0; j < facesThisObj->size(); j++) {
			Tri tri;
			tri.face = (*facesThisObj)[j];
			if (j < faceNormalsThisObj->size()) tri.normals = (*faceNormalsThisObj)[j];
			else { pbr_uint4 z = { 0, 0, 0, 0 }; tri.normals = z; }
			triCalcAABB(cfg, &tri, vertices4, normals4);
			triFaces.push_back(tri);
		}
		return triFaces;
	}

	/* BVH.cpp:203-245 */
	std::vector<BVHNode*> buildTreesFromObjects(
		const std::vector<object3D>* sceneObjects, const std::vector<float>* vertices, const std::vector<float>* normals
	) {
		std::vector<BVHNode*> subTrees;
		uint32_t offset = 0;
		uint32_t offsetN = 0;

		std::vector<pbr_float4> vertices4 = packFloatAsFloat4(vertices);
		std::vector<pbr_float4> normals4 = packFloatAsFloat4(normals);

		for (uint32_t i = 0; i < sceneObjects->size(); i++) {
			std::vector<pbr_uint4> facesThisObj;
			getFacesOfObject((*sceneObjects)[i].facesV, &facesThisObj, (int32_t) offset);
			offset += (uint32_t) facesThisObj.size();

			std::vector<pbr_uint4> faceNormalsThisObj;
			getFacesOfObject((*sceneObjects)[i].facesVN, &faceNormalsThisObj, (int32_t) offsetN);
			offsetN += (uint32_t) faceNormalsThisObj.size();

			std::vector<Tri> triFaces = facesToTriStructs(&facesThisObj, &faceNormalsThisObj, &vertices4, &normals4);

			makeNode(triFaces, true);   /* only used for rootSA, which buildTree never reads */
			BVHNode* st = buildTree(triFaces, 1);
			subTrees.push_back(st);
		}
		return subTrees;
	}

	/* BVH.cpp:133-193 */
	BVHNode* buildTree(std::vector<Tri> faces, uint32_t depth) {
		BVHNode* containerNode = makeNode(faces, false);

		containerNode->depth = depth;
		mDepthReached = (depth > mDepthReached) ? depth : mDepthReached;

		if (faces.size() <= mMaxFaces) {
			containerNode->faces = faces;
			return containerNode;
		}

		std::vector<Tri> leftFaces, rightFaces;

		if (faces.size() <= cfg.sahFacesLimit) {
			buildWithSAH(faces, &leftFaces, &rightFaces);
		}
		else {
			buildWithMeanSplit(faces, &leftFaces, &rightFaces);
		}

		if (leftFaces.size() == 0 || rightFaces.size() == 0) {
			containerNode->faces = faces;
			return containerNode;
		}

		containerNode->leftChild = buildTree(leftFaces, depth + 1);
		containerNode->rightChild = buildTree(rightFaces, depth + 1);

		return containerNode;
	}

	/* BVH.cpp:255-272 */
	void buildWithMeanSplit(const std::vector<Tri> faces, std::vector<Tri>* leftFaces, std::vector<Tri>* rightFaces) {
		float bestSAH = FLT_MAX;
		for (uint32_t axis = 0; axis <= 2; axis++) {
			std::vector<Tri> leftFacesTmp, rightFacesTmp;
			float splitPos = getMean(faces, axis);
			float sah = splitFaces(faces, splitPos, axis, &leftFacesTmp, &rightFacesTmp);

			if (sah < bestSAH) {
				bestSAH = sah;
				*leftFaces = leftFacesTmp;
				*rightFaces = rightFacesTmp;
			}
		}
	}

	/* BVH.cpp:283-294 */
	float buildWithSAH(std::vector<Tri> faces, std::vector<Tri>* leftFaces, std::vector<Tri>* rightFaces) {
		float bestSAH = FLT_MAX;
		for (uint32_t axis = 0; axis <= 2; axis++) {
			splitBySAH(&bestSAH, axis, faces, leftFaces, rightFaces);
		}
		return bestSAH;
	}

	/* BVH.cpp:318-352 */
	void combineNodes(const uint32_t numSubTrees) {
		if (numSubTrees > 1) {
			mNodes.push_back(mRoot);
		}
		mNodes.insert(mNodes.end(), mContainerNodes.begin(), mContainerNodes.end());

		for (uint32_t i = 0; i < mNodes.size(); i++) {
			if (mNodes[i]->faces.size() > 0) {
				mLeafNodes.push_back(mNodes[i]);
			}
			else {
				mNodes[i]->leftChild->parent = mNodes[i];
				mNodes[i]->rightChild->parent = mNodes[i];

				float leftSA = getSurfaceArea(mNodes[i]->leftChild->bbMin, mNodes[i]->leftChild->bbMax);
				float rightSA = getSurfaceArea(mNodes[i]->rightChild->bbMin, mNodes[i]->rightChild->bbMax);

				if (rightSA > leftSA) {
					BVHNode* tmp = mNodes[i]->leftChild;
					mNodes[i]->leftChild = mNodes[i]->rightChild;
					mNodes[i]->rightChild = tmp;
				}
			}
		}

		orderNodesByTraversal();

		if (cfg.skipAhead) {
			skipAheadOfNodes();
		}
	}

	/* BVH.cpp:410-420 */
	float getMean(const std::vector<Tri> faces, const uint32_t axis) {
		float sum = 0.0f;
		for (uint32_t i = 0; i < faces.size(); i++) {
			Tri tri = faces[i];
			V3 center = 0.5f * (tri.bbMin + tri.bbMax);
			sum += center[axis];
		}
		return sum / faces.size();
	}

	/* BVH.cpp:429-438 -- note: half extent, not centre (reference quirk) */
	float getMeanOfNodes(const std::vector<BVHNode*> nodes, const uint32_t axis) {
		float sum = 0.0f;
		for (uint32_t i = 0; i < nodes.size(); i++) {
			V3 center = (nodes[i]->bbMax - nodes[i]->bbMin) * 0.5f;
			sum += center[axis];
		}
		return sum / nodes.size();
	}

	/* BVH.cpp:471-491 */
	void groupTreesToNodes(std::vector<BVHNode*> nodes, BVHNode* parent, uint32_t depth) {
		if (nodes.size() == 1) {
			return;
		}

		parent->depth = depth;
		mDepthReached = (depth > mDepthReached) ? depth : mDepthReached;

		uint32_t axis = longestAxis(parent);
		std::vector<BVHNode*> leftGroup, rightGroup;
		float mean = getMeanOfNodes(nodes, axis);
		splitNodes(nodes, mean, axis, &leftGroup, &rightGroup);

		BVHNode* leftNode = makeContainerNode(leftGroup, false);
		parent->leftChild = leftNode;
		groupTreesToNodes(leftGroup, parent->leftChild, depth + 1);

		BVHNode* rightNode = makeContainerNode(rightGroup, false);
		parent->rightChild = rightNode;
		groupTreesToNodes(rightGroup, parent->rightChild, depth + 1);
	}

	/* BVH.cpp:502-553 */
	void growAABBsForSAH(
		const std::vector<Tri>* faces,
		std::vector<V3>* leftMin, std::vector<V3>* leftMax, std::vector<V3>* rightMin, std::vector<V3>* rightMax,
		std::vector<float>* leftSA, std::vector<float>* rightSA
	) {
		V3 bbMin = mk(0, 0, 0), bbMax = mk(0, 0, 0);
		const int numFaces = (int) faces->size();

		for (int i = 0; i < numFaces - 1; i++) {
			const Tri& f = (*faces)[i];
			if (i == 0) {
				bbMin = f.bbMin;
				bbMax = f.bbMax;
			}
			else {
				bbMin = vmin(bbMin, f.bbMin);
				bbMax = vmax(bbMax, f.bbMax);
			}
			(*leftMin)[i] = bbMin;
			(*leftMax)[i] = bbMax;
			(*leftSA)[i] = getSurfaceArea(bbMin, bbMax);
		}

		for (int i = numFaces - 2; i >= 0; i--) {
			const Tri& f = (*faces)[i + 1];
			if (i == numFaces - 2) {
				bbMin = f.bbMin;
				bbMax = f.bbMax;
			}
			else {
				bbMin = vmin(bbMin, f.bbMin);
				bbMax = vmax(bbMax, f.bbMax);
			}
			(*rightMin)[i] = bbMin;
			(*rightMax)[i] = bbMax;
			(*rightSA)[i] = getSurfaceArea(bbMin, bbMax);
		}
	}

	/* BVH.cpp:585-594 */
	uint32_t longestAxis(const BVHNode* node) {
		V3 sides = node->bbMax - node->bbMin;
		if (sides[0] > sides[1]) {
			return (sides[0] > sides[2]) ? 0 : 2;
		}
		else {
			return (sides[1] > sides[2]) ? 1 : 2;
		}
	}

	/* BVH.cpp:602-628 */
	BVHNode* makeContainerNode(const std::vector<BVHNode*> subTrees, const bool isRoot) {
		if (subTrees.size() == 1) {
			return subTrees[0];
		}

		BVHNode* node = newNode();
		node->bbMin = subTrees[0]->bbMin;
		node->bbMax = subTrees[0]->bbMax;

		for (uint32_t i = 1; i < subTrees.size(); i++) {
			node->bbMin = vmin(node->bbMin, subTrees[i]->bbMin);
			node->bbMax = vmax(node->bbMax, subTrees[i]->bbMax);
		}

		if (!isRoot) {
			mContainerNodes.push_back(node);
		}
		return node;
	}

	/* BVH.cpp:637-664 */
	BVHNode* makeNode(const std::vector<Tri>& tris, const bool ignore) {
		BVHNode* node = newNode();

		std::vector<V3> bbMins, bbMaxs;
		for (uint32_t i = 0; i < tris.size(); i++) {
			bbMins.push_back(tris[i].bbMin);
			bbMaxs.push_back(tris[i].bbMax);
		}

		V3 bbMin, bbMax;
		getAABB_boxes(bbMins, bbMaxs, &bbMin, &bbMax);
		node->bbMin = bbMin;
		node->bbMax = bbMax;

		if (!ignore) {
			mContainerNodes.push_back(node);
		}
		return node;
	}

	/* BVH.cpp:671-729 */
	void orderNodesByTraversal() {
		std::vector<BVHNode*> nodesOrdered;
		BVHNode* node = mNodes[0];

		while (true) {
			nodesOrdered.push_back(node);

			if (nodesOrdered.size() >= mNodes.size()) {
				break;   /* moved before the parent dereference: a one-node tree has no parent */
			}

			if (node->leftChild != NULL) {
				node = node->leftChild;
			}
			else {
				if (node->parent->leftChild == node) {
					node = node->parent->rightChild;
				}
				else if (node->parent->parent != NULL) {
					BVHNode* dummyParent = node->parent;

					while (dummyParent->parent->rightChild == dummyParent) {
						dummyParent = dummyParent->parent;
						if (dummyParent->parent == NULL) {
							break;
						}
					}

					if (dummyParent->parent != NULL) {
						node = dummyParent->parent->rightChild;
					}
				}
			}
		}

		for (uint32_t i = 0; i < mNodes.size(); i++) {
			BVHNode* n = nodesOrdered[i];
			n->id = i;
			mNodes[i] = n;
		}
	}

	/* BVH.cpp:770-795 */
	void skipAheadOfNodes() {
		float cmp = cfg.skipAheadCmp;
		uint32_t skippedLeft = 0;

		for (uint32_t i = 0; i < mNodes.size(); i++) {
			BVHNode* node = mNodes[i];
			node->numSkipsToHere = skippedLeft;

			if (node->leftChild != NULL && node->leftChild->leftChild != NULL) {
				BVHNode* left = node->leftChild;

				float saNode = getSurfaceArea(node->bbMin, node->bbMax);
				float saLeft = getSurfaceArea(left->bbMin, left->bbMax);

				if (saLeft / saNode >= cmp) {
					node->skipNextLeft = true;
					skippedLeft++;
				}
			}
		}
		mSkipped = skippedLeft;
	}

	/* BVH.cpp:807-851 */
	void splitBySAH(
		float* bestSAH, const uint32_t axis, std::vector<Tri> faces,
		std::vector<Tri>* leftFaces, std::vector<Tri>* rightFaces
	) {
		std::sort(faces.begin(), faces.end(), sortFacesCmp(axis));
		const uint32_t numFaces = (uint32_t) faces.size();

		std::vector<float> leftSA(numFaces - 1);
		std::vector<float> rightSA(numFaces - 1);
		std::vector<V3> leftMin(numFaces - 1), leftMax(numFaces - 1), rightMin(numFaces - 1), rightMax(numFaces - 1);

		growAABBsForSAH(&faces, &leftMin, &leftMax, &rightMin, &rightMax, &leftSA, &rightSA);

		int splitAfter = -1;
		float newSAH;

		for (uint32_t i = 0; i < numFaces - 1; i++) {
			float numFacesLeft = (float) (i + 1);
			float numFacesRight = (float) (numFaces - i - 1);

			newSAH = leftSA[i] * numFacesLeft + rightSA[i] * numFacesRight;

			if (newSAH < *bestSAH) {
				*bestSAH = newSAH;
				splitAfter = (int) i + 1;
			}
		}

		if (splitAfter >= 0) {
			leftFaces->clear();
			rightFaces->clear();
			leftFaces->insert(leftFaces->begin(), faces.begin(), faces.begin() + splitAfter);
			rightFaces->insert(rightFaces->begin(), faces.begin() + splitAfter, faces.end());
		}
	}

	/* BVH.cpp:862-935 */
	float splitFaces(
		const std::vector<Tri> faces, const float pos, const uint32_t axis,
		std::vector<Tri>* leftFaces, std::vector<Tri>* rightFaces
	) {
		float sah = FLT_MAX;
		std::vector<V3> bbMinsL, bbMinsR, bbMaxsL, bbMaxsR;

		leftFaces->clear();
		rightFaces->clear();

		for (uint32_t i = 0; i < faces.size(); i++) {
			Tri tri = faces[i];
			V3 cen = (tri.bbMin + tri.bbMax) * 0.5f;

			if (cen[axis] <= pos) {
				leftFaces->push_back(tri);
				bbMinsL.push_back(tri.bbMin);
				bbMaxsL.push_back(tri.bbMax);
			}
			else {
				rightFaces->push_back(tri);
				bbMinsR.push_back(tri.bbMin);
				bbMaxsR.push_back(tri.bbMax);
			}
		}

		if (leftFaces->size() == 0 || rightFaces->size() == 0) {
			bbMinsL.clear();
			bbMaxsL.clear();
			bbMinsR.clear();
			bbMaxsR.clear();
			leftFaces->clear();
			rightFaces->clear();

			for (uint32_t i = 0; i < faces.size(); i++) {
				Tri tri = faces[i];
				if (i < faces.size() / 2) {
					leftFaces->push_back(tri);
					bbMinsL.push_back(tri.bbMin);
					bbMaxsL.push_back(tri.bbMax);
				}
				else {
					rightFaces->push_back(tri);
					bbMinsR.push_back(tri.bbMin);
					bbMaxsR.push_back(tri.bbMax);
				}
			}
		}

		/* The reference computes the left box from the left lists and then reuses the
		 * DEFAULT-constructed bbMinR/bbMaxR (never filled: BVH.cpp:913-916) for the right
		 * surface area.  glm::vec3 default-constructs to zero in the GLM 0.9.x this code was
		 * written against, so rightSA is 0 and the SAH reduces to leftSA * |left|. */
		V3 bbMinL, bbMaxL;
		V3 bbMinR = mk(0, 0, 0), bbMaxR = mk(0, 0, 0);
		getAABB_boxes(bbMinsL, bbMaxsL, &bbMinL, &bbMaxL);
		float leftSA = getSurfaceArea(bbMinL, bbMaxL);
		float rightSA = getSurfaceArea(bbMinR, bbMaxR);

		sah = leftSA * leftFaces->size() + rightSA * rightFaces->size();

		if (leftFaces->size() == 0 || rightFaces->size() == 0) {
			sah = FLT_MAX;
		}
		return sah;
	}

	/* BVH.cpp:946-987 */
	void splitNodes(
		const std::vector<BVHNode*> nodes, const float pos, const uint32_t axis,
		std::vector<BVHNode*>* leftGroup, std::vector<BVHNode*>* rightGroup
	) {
		for (uint32_t i = 0; i < nodes.size(); i++) {
			BVHNode* node = nodes[i];
			V3 center = (node->bbMax - node->bbMin) / 2.0f;
			if (center[axis] < pos) {
				leftGroup->push_back(node);
			}
			else {
				rightGroup->push_back(node);
			}
		}

		if (leftGroup->size() == 0 || rightGroup->size() == 0) {
			leftGroup->clear();
			rightGroup->clear();
			for (uint32_t i = 0; i < nodes.size(); i++) {
				if (i < nodes.size() / 2) {
					leftGroup->push_back(nodes[i]);
				}
				else {
					rightGroup->push_back(nodes[i]);
				}
			}
		}
	}

	/* PathTracer.cpp:238-347 */
	void flatten(const std::vector<uint32_t>& faces, const std::vector<uint32_t>& facesVN, const std::vector<int32_t>& facesMtl) {
		std::vector<BVHNode*>& bvhNodes = mNodes;
		flatNodes.clear();
		flatFacesV.clear();
		flatFacesN.clear();

		bool skipNext = false;

		for (uint32_t i = 0; i < bvhNodes.size(); i++) {
			BVHNode* node = bvhNodes[i];

			if (skipNext) {
				skipNext = node->skipNextLeft;
				continue;
			}

			pbr_bvh_node sn;
			sn.bbMin.x = node->bbMin[0]; sn.bbMin.y = node->bbMin[1]; sn.bbMin.z = node->bbMin[2]; sn.bbMin.w = 0.0f;
			sn.bbMax.x = node->bbMax[0]; sn.bbMax.y = node->bbMax[1]; sn.bbMax.z = node->bbMax[2]; sn.bbMax.w = 0.0f;

			std::vector<Tri>& facesVec = node->faces;
			uint32_t fvecLen = (uint32_t) facesVec.size();
			sn.bbMin.w = (fvecLen > 0) ? (float) flatFacesV.size() + 0 : -1.0f;
			sn.bbMax.w = (fvecLen > 1) ? (float) flatFacesV.size() + 1 : -1.0f;

			if (fvecLen == 0 && node->skipNextLeft) {
				skipNext = true;
			}

			if (node->parent != NULL && fvecLen == 0) {
				bool isLeftNode = (node->parent->leftChild == node);

				if (!isLeftNode) {
					if (node->parent->parent != NULL) {
						BVHNode* dummyParent = node->parent;

						while (dummyParent->parent->rightChild == dummyParent) {
							dummyParent = dummyParent->parent;
							if (dummyParent->parent == NULL) {
								break;
							}
						}

						if (dummyParent->parent != NULL) {
							BVHNode* tgt = dummyParent->parent->rightChild;
							sn.bbMax.w = (float) (tgt->id - tgt->numSkipsToHere);
						}
					}
				}
				else {
					BVHNode* tgt = node->parent->rightChild;
					sn.bbMax.w = (float) (tgt->id - tgt->numSkipsToHere);
				}
			}

			flatNodes.push_back(sn);

			for (uint32_t j = 0; j < fvecLen; j++) {
				Tri tri = facesVec[j];
				pbr_uint4 fv;
				pbr_uint4 fn;

				fv.x = faces[tri.face.w * 3];
				fv.y = faces[tri.face.w * 3 + 1];
				fv.z = faces[tri.face.w * 3 + 2];
				fv.w = (uint32_t) facesMtl[tri.face.w];

				const size_t ni = (size_t) tri.normals.w * 3;
				fn.x = ni + 2 < facesVN.size() ? facesVN[ni] : 0;
				fn.y = ni + 2 < facesVN.size() ? facesVN[ni + 1] : 0;
				fn.z = ni + 2 < facesVN.size() ? facesVN[ni + 2] : 0;
				fn.w = 0;

				flatFacesV.push_back(fv);
				flatFacesN.push_back(fn);
			}
		}
	}
};

} /* namespace */

extern "C" {

/* Build + flatten.  Objects are given as concatenated per-object index lists (3 indices per
 * face) with per-object face counts, exactly what ObjParser::getObjects() holds.
 * `faces`, `facesVN`, `facesMtl` are the parser's global lists (getFacesV/VN/Mtl). */
void* oracle_bvh_build(
	const float* vertices, int64_t numVertexFloats,
	const float* normals, int64_t numNormalFloats,
	const uint32_t* objFacesV, const uint32_t* objFacesVN, const uint32_t* objFaceCounts,
	const uint32_t* objNormalFaceCounts, int32_t numObjects,
	const uint32_t* faces, int64_t numFaceIdx,
	const uint32_t* facesVN, int64_t numFaceVNIdx,
	const int32_t* facesMtl, int64_t numFacesMtl,
	uint32_t maxFaces, uint32_t sahFacesLimit, int32_t skipAhead, float skipAheadCmp, float phongTess
) {
	BVH* bvh = new BVH();
	bvh->cfg.maxFaces = maxFaces;
	bvh->cfg.sahFacesLimit = sahFacesLimit;
	bvh->cfg.skipAhead = skipAhead != 0;
	bvh->cfg.skipAheadCmp = skipAheadCmp;
	bvh->cfg.phongTess = phongTess;

	std::vector<float> v(vertices, vertices + numVertexFloats);
	std::vector<float> n(normals, normals + numNormalFloats);
	std::vector<object3D> objects((size_t) numObjects);
	size_t offV = 0, offN = 0;
	for (int32_t i = 0; i < numObjects; i++) {
		objects[i].facesV.assign(objFacesV + offV, objFacesV + offV + (size_t) objFaceCounts[i] * 3);
		offV += (size_t) objFaceCounts[i] * 3;
		objects[i].facesVN.assign(objFacesVN + offN, objFacesVN + offN + (size_t) objNormalFaceCounts[i] * 3);
		offN += (size_t) objNormalFaceCounts[i] * 3;
	}

	bvh->build(objects, v, n);

	std::vector<uint32_t> f(faces, faces + numFaceIdx);
	std::vector<uint32_t> fvn(facesVN, facesVN + numFaceVNIdx);
	std::vector<int32_t> fm(facesMtl, facesMtl + numFacesMtl);
	bvh->flatten(f, fvn, fm);
	return bvh;
}

/* info[0..5] = all nodes (before skip-ahead removal), leaves, depth reached, skipped left
 * children, emitted (flattened) nodes, flattened faces. */
void oracle_bvh_info(void* h, int64_t* info) {
	BVH* bvh = (BVH*) h;
	info[0] = (int64_t) bvh->mNodes.size();
	info[1] = (int64_t) bvh->mLeafNodes.size();
	info[2] = (int64_t) bvh->mDepthReached;
	info[3] = (int64_t) bvh->mSkipped;
	info[4] = (int64_t) bvh->flatNodes.size();
	info[5] = (int64_t) bvh->flatFacesV.size();
}

void oracle_bvh_get(void* h, pbr_bvh_node* nodes, pbr_uint4* facesV, pbr_uint4* facesN) {
	BVH* bvh = (BVH*) h;
	if (nodes) memcpy(nodes, bvh->flatNodes.data(), bvh->flatNodes.size() * sizeof(pbr_bvh_node));
	if (facesV) memcpy(facesV, bvh->flatFacesV.data(), bvh->flatFacesV.size() * sizeof(pbr_uint4));
	if (facesN) memcpy(facesN, bvh->flatFacesN.data(), bvh->flatFacesN.size() * sizeof(pbr_uint4));
}

void oracle_bvh_free(void* h) { delete (BVH*) h; }

} /* extern "C" */
