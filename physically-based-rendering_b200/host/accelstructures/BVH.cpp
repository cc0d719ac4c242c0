#include "BVH.h"

#include <algorithm>
#include <atomic>
#include <chrono>
#include <float.h>
#include <mutex>
#include <stdio.h>
#include <thread>

using std::vector;


namespace {

/** (centre on the sort axis, position in the tri array) -- 8 bytes, sorted instead of 56-byte Tris. */
struct Key {
	cl_float cen;
	cl_uint idx;
};

/** The reference's comparator (BVH.cpp:9-35): order by AABB centre on one axis, nothing else. */
struct KeyLess {
	bool operator()( const Key& a, const Key& b ) const { return a.cen < b.cen; }
};

/** Scratch buffers of one worker; grown on demand, reused down the recursion. */
struct Scratch {
	vector<Key> keys;
	vector<cl_uint> best;
	vector<cl_float> leftSA, rightSA;
	vector<cl_uint> tmpL, tmpR, bestL, bestR;
};

/** Node storage in fixed-size chunks so that pointers stay valid while other threads allocate. */
const size_t ARENA_CHUNK = 1 << 16;

}


/** Everything the recursion needs about one scene object. */
struct BVH::BuildCtx {
	const Tri* tris;            // all Tris of the scene, object after object
	cl_uint* order;             // index array being partitioned in place (same length as tris)
	std::mutex* arenaMutex;
	std::atomic<int>* activeWorkers;
	int maxWorkers;
	std::atomic<cl_uint>* depthReached;
};


BVH::BVH() {
	mRoot = NULL;
	mMaxFaces = 2;
	mSahFacesLimit = 0;
	mDepthReached = 0;
	mSkipped = 0;
	mBuildSeconds = 0.0;
}


/**
 * Build a BVH tree for each object in the scene and combine them into one big tree
 * (reference: BVH.cpp:50-64).
 */
BVH::BVH(
	const vector<object3D>& sceneObjects,
	const vector<cl_float>& vertices,
	const vector<cl_float>& normals
) {
	const std::chrono::steady_clock::time_point timerStart = std::chrono::steady_clock::now();
	mRoot = NULL;
	mDepthReached = 0;
	mSkipped = 0;
	this->setMaxFaces( Cfg::get().value<cl_uint>( Cfg::BVH_MAXFACES ) );
	mSahFacesLimit = Cfg::get().value<cl_uint>( Cfg::BVH_SAHFACESLIMIT );

	vector<BVHNode*> subTrees;
	this->buildTreesFromObjects( &sceneObjects, &vertices, &normals, &subTrees );
	if( subTrees.empty() ) {
		Logger::logError( "[BVH] No objects with faces in the scene." );
		mBuildSeconds = 0.0;
		return;
	}
	mRoot = this->makeContainerNode( subTrees, true );
	this->groupTreesToNodes( subTrees, mRoot, mDepthReached );
	this->combineNodes( (cl_uint) subTrees.size() );
	mBuildSeconds = std::chrono::duration<double>( std::chrono::steady_clock::now() - timerStart ).count();
	this->logStats();
}


BVH::~BVH() {
	for( size_t i = 0; i < mArenas.size(); i++ ) {
		delete mArenas[i];
	}
}


BVHNode* BVH::newNode() {
	if( mArenas.empty() || mArenas.back()->size() == ARENA_CHUNK ) {
		vector<BVHNode>* chunk = new vector<BVHNode>();
		chunk->reserve( ARENA_CHUNK );
		mArenas.push_back( chunk );
	}
	mArenas.back()->push_back( BVHNode() );
	BVHNode* node = &mArenas.back()->back();
	node->leftChild = NULL;
	node->rightChild = NULL;
	node->parent = NULL;
	node->id = 0;
	node->depth = 0;
	node->numSkipsToHere = 0;
	node->skipNextLeft = false;
	return node;
}


/**
 * One tree per scene object (reference: BVH.cpp:203-245 with facesToTriStructs :363-380).
 * The global face index stored in Tri::face.w is `faces of all previous objects + position`
 * (ModelLoader.cpp:28-41).
 */
void BVH::buildTreesFromObjects(
	const vector<object3D>* sceneObjects, const vector<cl_float>* vertices, const vector<cl_float>* normals,
	vector<BVHNode*>* subTrees
) {
	const vector<cl_float4> vertices4 = this->packFloatAsFloat4( vertices );
	const vector<cl_float4> normals4 = this->packFloatAsFloat4( normals );
	const float phongTess = Cfg::get().value<float>( Cfg::RENDER_PHONGTESS );

	size_t totalFaces = 0;
	for( size_t i = 0; i < sceneObjects->size(); i++ ) {
		totalFaces += ( *sceneObjects )[i].facesV.size() / 3;
	}

	vector<Tri> tris( totalFaces );
	vector<cl_uint> objBegin( sceneObjects->size() + 1, 0 );
	{
		cl_uint offset = 0, offsetN = 0;
		for( size_t i = 0; i < sceneObjects->size(); i++ ) {
			const object3D& o = ( *sceneObjects )[i];
			const cl_uint nf = (cl_uint) ( o.facesV.size() / 3 );
			const cl_uint nfn = (cl_uint) ( o.facesVN.size() / 3 );
			objBegin[i] = offset;
			for( cl_uint j = 0; j < nf; j++ ) {
				Tri& tri = tris[offset + j];
				cl_uint4 f = { o.facesV[3 * j], o.facesV[3 * j + 1], o.facesV[3 * j + 2], offset + j };
				tri.face = f;
				if( j < nfn ) {
					cl_uint4 fn = { o.facesVN[3 * j], o.facesVN[3 * j + 1], o.facesVN[3 * j + 2], offsetN + j };
					tri.normals = fn;
				}
				else {
					cl_uint4 fn = { 0, 0, 0, 0 };
					tri.normals = fn;
				}
			}
			offset += nf;
			offsetN += nfn;
		}
		objBegin[sceneObjects->size()] = offset;
	}

	/* per-face boxes, in parallel */
	const unsigned hw = std::max( 1u, std::thread::hardware_concurrency() );
	const int maxWorkers = (int) std::min( hw, 64u );
	{
		vector<std::thread> pool;
		const size_t chunk = ( totalFaces + (size_t) maxWorkers - 1 ) / (size_t) maxWorkers;
		for( int w = 0; w < maxWorkers; w++ ) {
			const size_t b = (size_t) w * chunk, e = std::min( totalFaces, b + chunk );
			if( b >= e ) { break; }
			pool.emplace_back( [&, b, e]() {
				for( size_t k = b; k < e; k++ ) {
					MathHelp::triCalcAABB( &tris[k], &vertices4, &normals4, phongTess );
				}
			} );
		}
		for( size_t w = 0; w < pool.size(); w++ ) { pool[w].join(); }
	}

	vector<cl_uint> order( totalFaces );
	for( size_t k = 0; k < totalFaces; k++ ) { order[k] = (cl_uint) k; }

	std::mutex arenaMutex;
	std::atomic<int> activeWorkers( 1 );
	std::atomic<cl_uint> depthReached( 0 );
	BuildCtx ctx = { tris.data(), order.data(), &arenaMutex, &activeWorkers, maxWorkers, &depthReached };

	char msg[256];
	for( size_t i = 0; i < sceneObjects->size(); i++ ) {
		const cl_uint lo = objBegin[i], hi = objBegin[i + 1];
		snprintf(
			msg, 256, "[BVH] Building tree %lu/%lu: \"%s\". %u faces.",
			(unsigned long) ( i + 1 ), (unsigned long) sceneObjects->size(), ( *sceneObjects )[i].oName.c_str(), hi - lo
		);
		Logger::logInfo( msg );
		if( hi == lo ) {
			Logger::logWarning( "[BVH] No faces in node." );
			continue;
		}
		subTrees->push_back( this->buildTree( &ctx, lo, hi, 1, 8 ) );
	}
	mDepthReached = depthReached.load();

	/* leaf faces -> one contiguous array, in creation order of the leaves */
	for( size_t i = 0; i < subTrees->size(); i++ ) {
		this->collectInCreationOrder( ( *subTrees )[i] );
	}
	mLeafTris.reserve( totalFaces );
	for( size_t i = 0; i < mContainerNodes.size(); i++ ) {
		BVHNode* n = mContainerNodes[i];
		if( n->leftChild == NULL ) {
			/* during the build a leaf keeps its range of `order` in (id, numSkipsToHere) */
			const cl_uint lo = n->id, cnt = n->numSkipsToHere;
			for( cl_uint k = 0; k < cnt; k++ ) { mLeafTris.push_back( tris[order[lo + k]] ); }
		}
	}
	size_t cursor = 0;
	for( size_t i = 0; i < mContainerNodes.size(); i++ ) {
		BVHNode* n = mContainerNodes[i];
		if( n->leftChild == NULL ) {
			n->faces.first = &mLeafTris[cursor];
			n->faces.count = n->numSkipsToHere;
			cursor += n->numSkipsToHere;
			n->id = 0;
			n->numSkipsToHere = 0;
		}
	}
}


/** Pre-order walk of the tree as built = the order in which the reference creates nodes. */
void BVH::collectInCreationOrder( BVHNode* root ) {
	vector<BVHNode*> stack;
	stack.push_back( root );
	while( !stack.empty() ) {
		BVHNode* n = stack.back();
		stack.pop_back();
		mContainerNodes.push_back( n );
		if( n->leftChild != NULL ) {
			stack.push_back( n->rightChild );
			stack.push_back( n->leftChild );
		}
	}
}


/**
 * Build the (sub)tree over order[lo, hi) (reference: buildTree BVH.cpp:133-193, buildWithSAH /
 * splitBySAH :283-294, 807-851, growAABBsForSAH :502-553, buildWithMeanSplit / getMean /
 * splitFaces :255-272, 410-420, 862-935).
 */
BVHNode* BVH::buildTree( BuildCtx* ctx, cl_uint lo, cl_uint hi, cl_uint depth, int spawnBudget ) {
	static thread_local Scratch S;
	const Tri* tris = ctx->tris;
	cl_uint* order = ctx->order;
	const cl_uint n = hi - lo;

	BVHNode* node;
	{
		std::lock_guard<std::mutex> lock( *ctx->arenaMutex );
		node = this->newNode();
	}

	/* makeNode: box over the faces in their current order (BVH.cpp:637-664) */
	glm::vec3 bbMin = tris[order[lo]].bbMin, bbMax = tris[order[lo]].bbMax;
	for( cl_uint k = lo + 1; k < hi; k++ ) {
		bbMin = glm::min( tris[order[k]].bbMin, bbMin );
		bbMax = glm::max( tris[order[k]].bbMax, bbMax );
	}
	node->bbMin = bbMin;
	node->bbMax = bbMax;
	node->depth = depth;
	{
		cl_uint seen = ctx->depthReached->load();
		while( depth > seen && !ctx->depthReached->compare_exchange_weak( seen, depth ) ) {}
	}

	if( n <= mMaxFaces ) {
		node->id = lo;
		node->numSkipsToHere = n;
		return node;
	}

	cl_uint numLeft = 0;

	if( n <= mSahFacesLimit ) {
		/* SAH sweep on all three axes; bestSAH carries over from axis to axis, strict `<` */
		cl_float bestSAH = FLT_MAX;
		bool found = false;
		if( S.keys.size() < n ) { S.keys.resize( n ); }
		if( S.best.size() < n ) { S.best.resize( n ); }
		if( S.leftSA.size() < n ) { S.leftSA.resize( n ); S.rightSA.resize( n ); }
		Key* keys = S.keys.data();

		for( int axis = 0; axis <= 2; axis++ ) {
			for( cl_uint k = 0; k < n; k++ ) {
				const Tri& t = tris[order[lo + k]];
				keys[k].cen = ( t.bbMin[axis] + t.bbMax[axis] ) * 0.5f;
				keys[k].idx = order[lo + k];
			}
			std::sort( keys, keys + n, KeyLess() );

			/* boxes grown from the left / from the right, surface area after each step */
			glm::vec3 gMin = tris[keys[0].idx].bbMin, gMax = tris[keys[0].idx].bbMax;
			S.leftSA[0] = MathHelp::getSurfaceArea( gMin, gMax );
			for( cl_uint k = 1; k + 1 < n; k++ ) {
				gMin = glm::min( gMin, tris[keys[k].idx].bbMin );
				gMax = glm::max( gMax, tris[keys[k].idx].bbMax );
				S.leftSA[k] = MathHelp::getSurfaceArea( gMin, gMax );
			}
			gMin = tris[keys[n - 1].idx].bbMin;
			gMax = tris[keys[n - 1].idx].bbMax;
			S.rightSA[n - 2] = MathHelp::getSurfaceArea( gMin, gMax );
			for( cl_uint k = n - 2; k-- > 0; ) {
				gMin = glm::min( gMin, tris[keys[k + 1].idx].bbMin );
				gMax = glm::max( gMax, tris[keys[k + 1].idx].bbMax );
				S.rightSA[k] = MathHelp::getSurfaceArea( gMin, gMax );
			}

			int splitAfter = -1;
			for( cl_uint k = 0; k + 1 < n; k++ ) {
				const cl_float numFacesLeft = (cl_float) ( k + 1 );
				const cl_float numFacesRight = (cl_float) ( n - k - 1 );
				const cl_float newSAH = S.leftSA[k] * numFacesLeft + S.rightSA[k] * numFacesRight;
				if( newSAH < bestSAH ) {
					bestSAH = newSAH;
					splitAfter = (int) k + 1;
				}
			}
			if( splitAfter >= 0 ) {
				found = true;
				numLeft = (cl_uint) splitAfter;
				for( cl_uint k = 0; k < n; k++ ) { S.best[k] = keys[k].idx; }
			}
		}
		if( found ) {
			for( cl_uint k = 0; k < n; k++ ) { order[lo + k] = S.best[k]; }
		}
		else {
			numLeft = 0;
		}
	}
	else {
		/* too many faces for SAH: split at the mean centre of each axis, keep the cheapest */
		char msg[256];
		snprintf( msg, 256, "[BVH] Too many faces in node for SAH. Splitting by mean position. (%u faces)", n );
		Logger::logDebug( msg );

		cl_float bestSAH = FLT_MAX;
		bool found = false;
		S.bestL.clear();
		S.bestR.clear();

		for( int axis = 0; axis <= 2; axis++ ) {
			cl_float sum = 0.0f;
			for( cl_uint k = lo; k < hi; k++ ) {
				const Tri& t = tris[order[k]];
				sum += 0.5f * ( t.bbMin[axis] + t.bbMax[axis] );
			}
			const cl_float splitPos = sum / n;

			S.tmpL.clear();
			S.tmpR.clear();
			for( cl_uint k = lo; k < hi; k++ ) {
				const Tri& t = tris[order[k]];
				const cl_float cen = ( t.bbMin[axis] + t.bbMax[axis] ) * 0.5f;
				if( cen <= splitPos ) { S.tmpL.push_back( order[k] ); }
				else { S.tmpR.push_back( order[k] ); }
			}
			if( S.tmpL.empty() || S.tmpR.empty() ) {
				Logger::logDebugVerbose( "[BVH] Dividing faces by center left one side empty. Just doing it 50:50 now." );
				S.tmpL.clear();
				S.tmpR.clear();
				for( cl_uint k = 0; k < n; k++ ) {
					if( k < n / 2 ) { S.tmpL.push_back( order[lo + k] ); }
					else { S.tmpR.push_back( order[lo + k] ); }
				}
			}

			/* The reference measures only the LEFT box: bbMinR / bbMaxR are default-constructed
			 * (zero) and never filled (BVH.cpp:913-916), so rightSA = 0. */
			cl_float sah = FLT_MAX;
			if( !S.tmpL.empty() && !S.tmpR.empty() ) {
				glm::vec3 lMin = tris[S.tmpL[0]].bbMin, lMax = tris[S.tmpL[0]].bbMax;
				for( size_t k = 1; k < S.tmpL.size(); k++ ) {
					lMin = glm::min( tris[S.tmpL[k]].bbMin, lMin );
					lMax = glm::max( tris[S.tmpL[k]].bbMax, lMax );
				}
				const cl_float leftSA = MathHelp::getSurfaceArea( lMin, lMax );
				const cl_float rightSA = MathHelp::getSurfaceArea( glm::vec3(), glm::vec3() );
				sah = leftSA * S.tmpL.size() + rightSA * S.tmpR.size();
			}
			else {
				Logger::logError( "[BVH] Dividing faces 50:50 left one side empty." );
			}

			if( sah < bestSAH ) {
				bestSAH = sah;
				found = true;
				S.bestL.swap( S.tmpL );
				S.bestR.swap( S.tmpR );
			}
		}
		if( found ) {
			numLeft = (cl_uint) S.bestL.size();
			std::copy( S.bestL.begin(), S.bestL.end(), order + lo );
			std::copy( S.bestR.begin(), S.bestR.end(), order + lo + numLeft );
		}
	}

	if( numLeft == 0 || numLeft == n ) {
		if( n > mMaxFaces ) {
			Logger::logWarning( "[BVH] More faces than can be traversed in node." );
		}
		node->id = lo;
		node->numSkipsToHere = n;
		return node;
	}

	/* children: independent, so the left one may run on another thread */
	const cl_uint mid = lo + numLeft;
	bool spawned = false;
	std::thread worker;
	BVHNode* left = NULL;
	if( spawnBudget > 0 && n >= 20000 ) {
		int active = ctx->activeWorkers->load();
		while( active < ctx->maxWorkers ) {
			if( ctx->activeWorkers->compare_exchange_weak( active, active + 1 ) ) {
				spawned = true;
				break;
			}
		}
	}
	if( spawned ) {
		worker = std::thread( [&, this]() {
			left = this->buildTree( ctx, lo, mid, depth + 1, spawnBudget - 1 );
			ctx->activeWorkers->fetch_sub( 1 );
		} );
	}
	else {
		left = this->buildTree( ctx, lo, mid, depth + 1, spawnBudget - 1 );
	}
	BVHNode* right = this->buildTree( ctx, mid, hi, depth + 1, spawnBudget - 1 );
	if( spawned ) {
		worker.join();
	}
	node->leftChild = left;
	node->rightChild = right;
	return node;
}


/**
 * Combine the container nodes, leaf nodes and the root node into one list, set parents, put the
 * child with the bigger surface area on the left, order for traversal, mark skip-ahead
 * (reference: BVH.cpp:318-352).
 */
void BVH::combineNodes( const cl_uint numSubTrees ) {
	if( numSubTrees > 1 ) {
		mNodes.push_back( mRoot );
	}
	mNodes.insert( mNodes.end(), mContainerNodes.begin(), mContainerNodes.end() );

	for( size_t i = 0; i < mNodes.size(); i++ ) {
		BVHNode* node = mNodes[i];
		if( node->faces.size() > 0 ) {
			mLeafNodes.push_back( node );
		}
		else {
			node->leftChild->parent = node;
			node->rightChild->parent = node;

			const cl_float leftSA = MathHelp::getSurfaceArea( node->leftChild->bbMin, node->leftChild->bbMax );
			const cl_float rightSA = MathHelp::getSurfaceArea( node->rightChild->bbMin, node->rightChild->bbMax );

			if( rightSA > leftSA ) {
				std::swap( node->leftChild, node->rightChild );
			}
		}
	}

	this->orderNodesByTraversal();

	if( Cfg::get().value<bool>( Cfg::BVH_SKIPAHEAD ) ) {
		this->skipAheadOfNodes();
	}
}


vector<BVHNode*> BVH::getContainerNodes() { return mContainerNodes; }
cl_uint BVH::getDepth() { return mDepthReached; }
vector<BVHNode*> BVH::getLeafNodes() { return mLeafNodes; }
vector<BVHNode*> BVH::getNodes() { return mNodes; }
BVHNode* BVH::getRoot() { return mRoot; }


/** Mean of the HALF EXTENTS of the nodes -- not of their centres (reference quirk, BVH.cpp:429-438). */
cl_float BVH::getMeanOfNodes( const vector<BVHNode*>& nodes, const cl_uint axis ) {
	cl_float sum = 0.0f;
	for( size_t i = 0; i < nodes.size(); i++ ) {
		const glm::vec3 center = ( nodes[i]->bbMax - nodes[i]->bbMin ) * 0.5f;
		sum += center[axis];
	}
	return sum / nodes.size();
}


/** Group the object trees into a top-level tree (reference: BVH.cpp:471-491). */
void BVH::groupTreesToNodes( vector<BVHNode*> nodes, BVHNode* parent, cl_uint depth ) {
	if( nodes.size() == 1 ) {
		return;
	}

	parent->depth = depth;
	mDepthReached = ( depth > mDepthReached ) ? depth : mDepthReached;

	const cl_uint axis = this->longestAxis( parent );
	vector<BVHNode*> leftGroup, rightGroup;
	const cl_float mean = this->getMeanOfNodes( nodes, axis );
	this->splitNodes( nodes, mean, axis, &leftGroup, &rightGroup );

	BVHNode* leftNode = this->makeContainerNode( leftGroup, false );
	parent->leftChild = leftNode;
	this->groupTreesToNodes( leftGroup, parent->leftChild, depth + 1 );

	BVHNode* rightNode = this->makeContainerNode( rightGroup, false );
	parent->rightChild = rightNode;
	this->groupTreesToNodes( rightGroup, parent->rightChild, depth + 1 );
}


void BVH::logStats() {
	char msg[512];
	snprintf(
		msg, 512, "[BVH] Generated in %.2f s. Contains %lu nodes (%lu leaves). Max faces of %u. Max depth of %u.",
		mBuildSeconds, (unsigned long) mNodes.size(), (unsigned long) mLeafNodes.size(), mMaxFaces, mDepthReached
	);
	Logger::logInfo( msg );
}


/** Reference: BVH.cpp:585-594. */
cl_uint BVH::longestAxis( const BVHNode* node ) {
	const glm::vec3 sides = node->bbMax - node->bbMin;
	if( sides[0] > sides[1] ) {
		return ( sides[0] > sides[2] ) ? 0 : 2;
	}
	return ( sides[1] > sides[2] ) ? 1 : 2;
}


/** Container over several (sub)trees; a single tree is its own container (reference: BVH.cpp:602-628). */
BVHNode* BVH::makeContainerNode( const vector<BVHNode*>& subTrees, const bool isRoot ) {
	if( subTrees.size() == 1 ) {
		return subTrees[0];
	}

	BVHNode* node = this->newNode();
	node->bbMin = subTrees[0]->bbMin;
	node->bbMax = subTrees[0]->bbMax;
	for( size_t i = 1; i < subTrees.size(); i++ ) {
		node->bbMin = glm::min( node->bbMin, subTrees[i]->bbMin );
		node->bbMax = glm::max( node->bbMax, subTrees[i]->bbMax );
	}

	if( !isRoot ) {
		mContainerNodes.push_back( node );
	}
	return node;
}


/**
 * Order all nodes for left-first, stackless traversal and assign ids (reference: BVH.cpp:671-729).
 * This is a pre-order walk; "next after a right child" is the right sibling of the closest ancestor
 * that is a left child.
 */
void BVH::orderNodesByTraversal() {
	const std::chrono::steady_clock::time_point timerStart = std::chrono::steady_clock::now();

	vector<BVHNode*> nodesOrdered;
	nodesOrdered.reserve( mNodes.size() );
	BVHNode* node = mNodes[0];

	while( true ) {
		nodesOrdered.push_back( node );
		if( nodesOrdered.size() >= mNodes.size() ) {
			break;
		}

		if( node->leftChild != NULL ) {
			node = node->leftChild;
			continue;
		}
		/* leaf: climb while we are a right child, then step to the right sibling */
		BVHNode* climb = node;
		while( climb->parent != NULL && climb->parent->rightChild == climb ) {
			climb = climb->parent;
		}
		if( climb->parent == NULL ) {
			break;   /* rightmost leaf reached before all nodes were seen (cannot happen in a proper tree) */
		}
		node = climb->parent->rightChild;
	}

	for( size_t i = 0; i < nodesOrdered.size(); i++ ) {
		nodesOrdered[i]->id = (cl_uint) i;
		mNodes[i] = nodesOrdered[i];
	}

	char msg[128];
	snprintf(
		msg, 128, "[BVH] Ordered nodes for traversal in %g ms.",
		std::chrono::duration<double, std::milli>( std::chrono::steady_clock::now() - timerStart ).count()
	);
	Logger::logInfo( msg );
}


/** Reference: BVH.cpp:759-763. */
cl_uint BVH::setMaxFaces( const int value ) {
	mMaxFaces = (cl_uint) fmax( value, 1 );
	return mMaxFaces;
}


/**
 * Mark inner left children whose box is nearly as big as their parent's: the flattening leaves
 * them out and a ray tests their children unconditionally (reference: BVH.cpp:770-795).
 */
void BVH::skipAheadOfNodes() {
	const cl_float cmp = Cfg::get().value<cl_float>( Cfg::BVH_SKIPAHEAD_CMP );
	cl_uint skippedLeft = 0;

	for( size_t i = 0; i < mNodes.size(); i++ ) {
		BVHNode* node = mNodes[i];
		node->numSkipsToHere = skippedLeft;

		if( node->leftChild != NULL && node->leftChild->leftChild != NULL ) {
			const BVHNode* left = node->leftChild;
			const cl_float saNode = MathHelp::getSurfaceArea( node->bbMin, node->bbMax );
			const cl_float saLeft = MathHelp::getSurfaceArea( left->bbMin, left->bbMax );

			if( saLeft / saNode >= cmp ) {
				node->skipNextLeft = true;
				skippedLeft++;
			}
		}
	}
	mSkipped = skippedLeft;

	char msg[128];
	snprintf( msg, 128, "[BVH] Marked %u left child nodes as skippable.", skippedLeft );
	Logger::logInfo( msg );
}


/** Split object trees into two groups at `pos` (half-extent quirk kept; reference: BVH.cpp:946-987). */
void BVH::splitNodes(
	const vector<BVHNode*>& nodes, const cl_float pos, const cl_uint axis,
	vector<BVHNode*>* leftGroup, vector<BVHNode*>* rightGroup
) {
	for( size_t i = 0; i < nodes.size(); i++ ) {
		const glm::vec3 center = ( nodes[i]->bbMax - nodes[i]->bbMin ) / 2.0f;
		if( center[axis] < pos ) { leftGroup->push_back( nodes[i] ); }
		else { rightGroup->push_back( nodes[i] ); }
	}

	if( leftGroup->size() == 0 || rightGroup->size() == 0 ) {
		Logger::logDebugVerbose( "[BVH] Dividing nodes by the given position left one side empty. Just doing it 50:50 now." );
		leftGroup->clear();
		rightGroup->clear();
		for( size_t i = 0; i < nodes.size(); i++ ) {
			if( i < nodes.size() / 2 ) { leftGroup->push_back( nodes[i] ); }
			else { rightGroup->push_back( nodes[i] ); }
		}
	}

	if( leftGroup->size() == 0 || rightGroup->size() == 0 ) {
		Logger::logError( "[BVH] Dividing nodes 50:50 left one side empty." );
	}
}


/**
 * Line list (12 edges per box) of all nodes for an overlay (reference: BVH.cpp:995-1055).
 */
void BVH::visualize( vector<cl_float>* vertices, vector<cl_uint>* indices ) {
	static const cl_uint EDGES[24] = { 0,1, 1,3, 3,2, 2,0, 4,5, 5,7, 7,6, 6,4, 0,4, 1,5, 2,6, 3,7 };
	for( size_t i = 0; i < mNodes.size(); i++ ) {
		const BVHNode* n = mNodes[i];
		const cl_uint base = (cl_uint) ( vertices->size() / 3 );
		for( int c = 0; c < 8; c++ ) {
			vertices->push_back( ( c & 1 ) ? n->bbMax.x : n->bbMin.x );
			vertices->push_back( ( c & 2 ) ? n->bbMax.y : n->bbMin.y );
			vertices->push_back( ( c & 4 ) ? n->bbMax.z : n->bbMin.z );
		}
		for( int e = 0; e < 24; e++ ) {
			indices->push_back( base + EDGES[e] );
		}
	}
}
