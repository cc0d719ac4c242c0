/*
 * BVH -- reference: source/accelstructures/BVH.{h,cpp}.
 *
 * Produces the SAME tree, node order, skip-ahead marks and leaf face order as the reference builder
 * (that order is the layout contract the kernel relies on), but is built
 * differently: faces are never copied -- every (sub)tree works on a range of an index array; the
 * three per-axis sorts of the SAH sweep run on compact (centre, index) pairs with the same
 * comparison function and the same std::sort, so ties fall exactly as in the reference; bounding
 * boxes for the sweep are two running prefix/suffix scans; independent subtrees are built on worker
 * threads.  The reference passes vector<Tri> by value through the recursion and allocates a
 * vector<vector<glm::vec3>> per split candidate (BVH.cpp:133-193, 502-553, 807-851), which is what
 * makes it take minutes at 10 M triangles.
 *
 * Same public surface: BVH(objects, vertices, normals), getNodes / getRoot / getLeafNodes /
 * getContainerNodes / getDepth / visualize, and struct BVHNode with the reference's fields.
 * `BVHNode::faces` is a read-only view (size(), operator[], iteration, conversion to vector<Tri>)
 * into storage owned by the BVH instead of a vector per node.
 */
#ifndef BVH_H
#define BVH_H

#include <vector>

#include "AccelStructure.h"
#include "../Cfg.h"
#include "../Logger.h"
#include "../MathHelp.h"
#include "../ModelLoader.h"

using std::vector;


/** Read-only list of the faces of a leaf; stands in for the reference's `vector<Tri> faces`. */
struct FaceList {
	const Tri* first;
	cl_uint count;
	FaceList() : first( NULL ), count( 0 ) {}
	size_t size() const { return count; }
	const Tri& operator[]( size_t i ) const { return first[i]; }
	const Tri* begin() const { return first; }
	const Tri* end() const { return first + count; }
	operator vector<Tri>() const { return vector<Tri>( first, first + count ); }
};


struct BVHNode {
	BVHNode* leftChild;
	BVHNode* rightChild;
	BVHNode* parent;
	FaceList faces;
	glm::vec3 bbMin;
	glm::vec3 bbMax;
	cl_uint id;
	cl_uint depth;
	cl_uint numSkipsToHere;
	bool skipNextLeft;
};


class BVH : public AccelStructure {

	public:
		BVH();
		BVH(
			const vector<object3D>& sceneObjects,
			const vector<cl_float>& vertices,
			const vector<cl_float>& normals
		);
		~BVH();
		vector<BVHNode*> getContainerNodes();
		cl_uint getDepth();
		vector<BVHNode*> getLeafNodes();
		vector<BVHNode*> getNodes();
		BVHNode* getRoot();
		virtual void visualize( vector<cl_float>* vertices, vector<cl_uint>* indices );

		/** Additive: no-copy access and build statistics. */
		const vector<BVHNode*>& nodes() const { return mNodes; }
		cl_uint getNumSkipped() const { return mSkipped; }
		double getBuildSeconds() const { return mBuildSeconds; }

	protected:
		struct BuildCtx;
		BVHNode* buildTree( BuildCtx* ctx, cl_uint lo, cl_uint hi, cl_uint depth, int spawnBudget );
		void buildTreesFromObjects(
			const vector<object3D>* sceneObjects, const vector<cl_float>* vertices, const vector<cl_float>* normals,
			vector<BVHNode*>* subTrees
		);
		void collectInCreationOrder( BVHNode* node );
		void combineNodes( const cl_uint numSubTrees );
		cl_float getMeanOfNodes( const vector<BVHNode*>& nodes, const cl_uint axis );
		void groupTreesToNodes( vector<BVHNode*> nodes, BVHNode* parent, cl_uint depth );
		void logStats();
		cl_uint longestAxis( const BVHNode* node );
		BVHNode* makeContainerNode( const vector<BVHNode*>& subTrees, const bool isRoot );
		BVHNode* newNode();
		void orderNodesByTraversal();
		cl_uint setMaxFaces( const int value );
		void skipAheadOfNodes();
		void splitNodes(
			const vector<BVHNode*>& nodes, const cl_float midpoint, const cl_uint axis,
			vector<BVHNode*>* leftGroup, vector<BVHNode*>* rightGroup
		);

		vector<BVHNode*> mContainerNodes;
		vector<BVHNode*> mLeafNodes;
		vector<BVHNode*> mNodes;
		BVHNode* mRoot;

		vector<Tri> mLeafTris;      // faces of all leaves, leaf after leaf (what FaceList points into)
		vector< vector<BVHNode>* > mArenas;

		cl_uint mMaxFaces;
		cl_uint mSahFacesLimit;
		cl_uint mDepthReached;
		cl_uint mSkipped;
		double mBuildSeconds;

};

#endif
