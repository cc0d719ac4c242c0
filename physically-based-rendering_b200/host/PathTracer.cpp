#include "PathTracer.h"

#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#include "qt/GLWidget.h"

using std::string;
using std::vector;


namespace {

double msSince( const std::chrono::steady_clock::time_point& t0 ) {
	return std::chrono::duration<double, std::milli>( std::chrono::steady_clock::now() - t0 ).count();
}

void logBuffer( const char* what, double ms, size_t bytes ) {
	float bytesFloat;
	string unit;
	utils::formatBytes( bytes, &bytesFloat, &unit );
	char msg[160];
	snprintf( msg, 160, "[PathTracer] Created %s buffer in %g ms -- %.2f %s.", what, ms, bytesFloat, unit.c_str() );
	Logger::logInfo( msg );
}

}


/** Reference: PathTracer.cpp:11-29. */
PathTracer::PathTracer( GLWidget* parent ) {
	srand( (unsigned) time( 0 ) );

	mWidth = Cfg::get().value<cl_uint>( Cfg::WINDOW_WIDTH );
	mHeight = Cfg::get().value<cl_uint>( Cfg::WINDOW_HEIGHT );

	mGLWidget = parent;
	mCL = NULL;
	mCamera = NULL;
	mTextureOut = NULL;
	mTextureDebugHost = NULL;
	mKernelPathTracing = 0;
	mBufBVH = mBufFacesV = mBufFacesN = mBufVertices = mBufNormals = mBufMaterials = 0;
	mBufTextureIn = mBufTextureOut = mBufTextureDebug = mBufLights = 0;

	mFOV = Cfg::get().value<cl_float>( Cfg::PERS_FOV );
	mPxDim = 0.0f;
	mSampleCount = 0;
	mDeterministicSeeds = false;
	mSeedStride = 1;
	mSeedOffset = 0;
	mFrameTimeMs = 0;
	mRenderAhead = 0;
	mHaveOutput = false;
	mRank = 0;
	mWorld = 1;
	mSharding = SHARD_SPP;
	for( int i = 0; i < NUM_DISPLAYS; i++ ) { mBufTextureDisplay[i] = 0; }
	mCombines = 0;
	mTimeSinceStart = std::chrono::steady_clock::now();

	memset( &mStructCam, 0, sizeof( mStructCam ) );
	mStructCam.focusPoint.x = -1;
	mStructCam.focusPoint.y = -1;
	mStructCam.lense.x = Cfg::get().value<cl_float>( Cfg::CAM_LENSE_FOCALLENGTH );
	mStructCam.lense.y = Cfg::get().value<cl_float>( Cfg::CAM_LENSE_APERTURE );
}


PathTracer::~PathTracer() {
	delete mCL;   /* pinned host buffers belong to the context and go with it */
}


/**
 * Seed, mixing weight and camera for this frame, then launch (reference: PathTracer.cpp:43-52).
 * The launch is left running; callers synchronise through a read-back or finish().
 */
void PathTracer::clPathTracing( cl_float timeSinceStart ) {
	cl_float pixelWeight = mSampleCount / (cl_float) ( mSampleCount + 1 );

	mCL->setKernelArg( mKernelPathTracing, 0, sizeof( cl_float ), &timeSinceStart );
	mCL->setKernelArg( mKernelPathTracing, 1, sizeof( cl_float ), &pixelWeight );
	mCL->setKernelArg( mKernelPathTracing, 3, sizeof( camera_cl ), &mStructCam );

	mCL->execute( mKernelPathTracing );
}


/**
 * One frame on the device.  The previous output becomes this frame's input by swapping the two
 * image handles (the reference reads imageOut back and uploads it as imageIn, PathTracer.cpp:61-66).
 */
void PathTracer::launchFrame() {
	this->updateEyeBuffer();
	if( mHaveOutput ) {
		this->advanceImages();
	}
	this->clPathTracing( this->nextSeed() );
	mSampleCount++;
	mHaveOutput = true;
	this->combineFrame();
}


/**
 * The previous output becomes the next input; the next output is the image behind it in the ring.  Two images (a swap)
 * unless frames are traced ahead: a frame that has not been delivered yet keeps its image until it has been copied out.
 */
void PathTracer::advanceImages() {
	size_t at = 0;
	for( size_t i = 0; i < mImages.size(); i++ ) { if( mImages[i] == mBufTextureOut ) { at = i; } }
	mBufTextureIn = mBufTextureOut;
	mBufTextureOut = mImages[( at + 1 ) % mImages.size()];
	mCL->setKernelArg( mKernelPathTracing, 11, sizeof( cl_mem ), &mBufTextureIn );
	mCL->setKernelArg( mKernelPathTracing, 12, sizeof( cl_mem ), &mBufTextureOut );
}


/**
 * The frame just launched is completed across the ranks: one collective, enqueued behind it on the communicator's
 * stream (the next frame is traced meanwhile; the library orders the buffers, include/pbr_b200.h).
 */
void PathTracer::combineFrame() {
	if( mWorld <= 1 || mSharding == SHARD_NONE ) { return; }
	if( mSharding == SHARD_SPP ) {
		mCL->frameCombine( mBufTextureOut, PBR_COMBINE_SPP, mBufTextureDisplay[mCombines % NUM_DISPLAYS] );
	}
	else {
		mCL->frameCombine( mBufTextureOut, PBR_COMBINE_ROWS, 0 );
	}
	mCombines++;
}


/** The image a caller receives: the accumulation buffer, or with SHARD_SPP the mean over ranks of the last frame. */
cl_mem PathTracer::deliveredImage() const {
	if( mWorld > 1 && mSharding == SHARD_SPP && mCombines > 0 ) { return mBufTextureDisplay[( mCombines - 1 ) % NUM_DISPLAYS]; }
	return mBufTextureOut;
}


bool PathTracer::setRanks( int rank, int world, const void* ncclId, int sharding ) {
	this->dropFrameAhead();
	if( world <= 1 ) { mWorld = 1; mRank = 0; return true; }
	if( mCL == NULL ) {
		Logger::logError( "[PathTracer] setRanks: load a model first (the device context is created with it)." );
		return false;
	}
	if( mWorld > 1 ) {
		Logger::logError( "[PathTracer] setRanks: this renderer already has its ranks (setSharding changes the sharding)." );
		return false;
	}
	if( !mCL->commInit( ncclId, rank, world ) ) { return false; }
	mRank = rank; mWorld = world;
	for( int i = 0; i < NUM_DISPLAYS; i++ ) { mBufTextureDisplay[i] = mCL->createImage2DWriteOnly( mWidth, mHeight ); }
	return this->setSharding( sharding );
}


bool PathTracer::setSharding( int sharding ) {
	this->dropFrameAhead();
	if( mWorld <= 1 ) { return true; }
	mCL->commFence();
	mCL->setTileStripes( 0, 1, 0 );
	mCL->setTile( -1, -1 );
	this->setSeedSchedule( 1, 0 );
	mCombines = 0;
	this->resetSampleCount();
	if( sharding == SHARD_SPP ) {
		this->setSeedSchedule( (cl_uint) mWorld, (cl_uint) mRank );
	}
	else if( sharding == SHARD_STRIPES ) {
		int stripe = 0;
		if( mHeight % (cl_uint) mWorld == 0 ) {
			const int local = (int) mHeight / mWorld;
			for( int s = std::min( 8, local ); s > 0; s-- ) { if( local % s == 0 ) { stripe = s; break; } }
		}
		if( stripe == 0 ) {
			Logger::logError( "[PathTracer] window.height is not a multiple of the number of ranks; use SHARD_ROWS." );
			return false;
		}
		mCL->setTileStripes( stripe, mWorld, mRank );
	}
	else if( sharding == SHARD_ROWS ) {
		int32_t y0 = 0, y1 = 0;
		pbr_tile_rows( (int32_t) mHeight, mRank, mWorld, &y0, &y1 );
		mCL->setTile( y0, y1 );
	}
	else if( sharding != SHARD_NONE ) {
		return false;
	}
	mSharding = sharding;
	return true;
}


/**
 * Generate the path traced image (reference: PathTracer.cpp:59-71).
 * @param  {std::vector<cl_float>*} textureDebug Receives the debug image; may be NULL (additive).
 * @return {std::vector<cl_float>}               The accumulated frame, RGBA float, row 0 = bottom.
 */
vector<cl_float> PathTracer::generateImage( vector<cl_float>* textureDebug ) {
	const size_t n = (size_t) mWidth * mHeight * 4;
	cl_float* dbg = NULL;
	if( textureDebug != NULL ) {
		if( mTextureDebugHost == NULL ) {
			mTextureDebugHost = (cl_float*) mCL->allocHost( n * sizeof( cl_float ) );
		}
		dbg = mTextureDebugHost;
	}
	this->generateImageInto( mTextureOut, dbg );
	if( textureDebug != NULL ) {
		textureDebug->assign( dbg, dbg + n );
	}
	return vector<cl_float>( mTextureOut, mTextureOut + n );
}


void PathTracer::generateImageInto( cl_float* target, cl_float* targetDebug ) {
	if( targetDebug != NULL ) {
		this->dropFrameAhead();          /* a frame traced ahead has no debug image */
	}
	if( mAhead.empty() ) {
		mCL->setDebugImage( targetDebug != NULL );
		this->launchAhead();
	}
	const AheadFrame f = mAhead.front();  /* the frame this call delivers */

	if( mRenderAhead > 0 && targetDebug == NULL ) {
		/* copy this frame out on the copy stream while the next ones are traced: they read the image that is being
		 * copied (or one traced after it) and write images further down the ring (with ranks: the copy waits for this
		 * frame's collective, and the next frame's mean lands in the other display image) */
		mCL->readImageOutputBegin( f.delivered, mWidth, mHeight, target );
		mCL->setDebugImage( false );
		while( mAhead.size() < 1 + (size_t) this->aheadDepth() ) { this->launchAhead(); }
		mCL->readImageOutputEnd();
		mAhead.pop_front();
		return;
	}
	mAhead.pop_front();
	mCL->readImageOutput( f.delivered, mWidth, mHeight, target );
	if( targetDebug != NULL ) {
		mCL->readImageOutput( mBufTextureDebug, mWidth, mHeight, targetDebug );
	}
}


/** launchFrame() for a frame that is delivered later: remember what has to be undone if it is not wanted after all. */
void PathTracer::launchAhead() {
	AheadFrame f;
	f.in = mBufTextureIn; f.out = mBufTextureOut; f.sampleCount = mSampleCount; f.combines = mCombines; f.haveOutput = mHaveOutput;
	this->launchFrame();
	f.delivered = this->deliveredImage();
	mAhead.push_back( f );
}


/** How many frames are traced beyond the one being delivered: setRenderAhead's (with ranks: each of them keeps one of
 *  the NUM_DISPLAYS display images until it has been delivered). */
int PathTracer::aheadDepth() const {
	return ( mWorld > 1 && mSharding != SHARD_NONE ) ? std::min( mRenderAhead, (int) NUM_DISPLAYS - 1 ) : mRenderAhead;
}


void PathTracer::setRenderAhead( int depth ) {
	/* frames already traced ahead stay valid: the next calls return them */
	mRenderAhead = std::max( 0, std::min( depth, 3 ) );
	while( mCL != NULL && mImages.size() < 2 + (size_t) mRenderAhead ) {
		mImages.push_back( mCL->createImage2DWriteOnly( mWidth, mHeight ) );
	}
}


/**
 * Frames traced ahead are not wanted after all -- all of them, or all but the first `keep`: undo launchFrame's
 * bookkeeping, so that the last frame kept (or delivered) is the current output again.  (The device work is simply wasted;
 * stream order keeps it harmless.)
 */
void PathTracer::dropFrameAhead( size_t keep ) {
	if( mAhead.size() <= keep ) { mAhead.resize( std::min( keep, mAhead.size() ) ); return; }
	const AheadFrame f = mAhead[keep];      /* the state before the first frame that goes */
	mAhead.resize( keep );
	mBufTextureIn = f.in;
	mBufTextureOut = f.out;
	mCL->setKernelArg( mKernelPathTracing, 11, sizeof( cl_mem ), &mBufTextureIn );
	mCL->setKernelArg( mKernelPathTracing, 12, sizeof( cl_mem ), &mBufTextureOut );
	mSampleCount = f.sampleCount;
	mCombines = f.combines;
	mHaveOutput = f.haveOutput;
}


/**
 * `frames` more frames without a read-back in between, handed to the device as one batch: the seeds and
 * mixing weights of all of them are fixed up front (wall-clock seeds are spaced by 1/30 s, the step of the
 * deterministic schedule, instead of by the time the frames happen to take).
 */
void PathTracer::renderFrames( cl_uint frames ) {
	/* frames traced ahead are the first of these; more of them than asked for are dropped */
	const size_t keep = std::min( (size_t) frames, mAhead.size() );
	this->dropFrameAhead( keep );
	mAhead.clear();
	frames -= (cl_uint) keep;
	if( frames == 0 ) { return; }
	mCL->setDebugImage( false );
	/* with ranks, every frame of the batch is completed across the ranks by the library (progressive display) */
	const bool combine = mWorld > 1 && mSharding != SHARD_NONE;
	mCL->setBatchCombine( combine ? ( mSharding == SHARD_SPP ? PBR_COMBINE_SPP : PBR_COMBINE_ROWS ) : -1,
		mBufTextureDisplay, NUM_DISPLAYS, (int) ( mCombines % NUM_DISPLAYS ) );
	if( combine ) { mCombines += frames; }
	this->updateEyeBuffer();
	if( mHaveOutput ) {
		this->advanceImages();
	}
	vector<cl_float> seeds( frames ), weights( frames );
	const cl_float now = this->getTimeSinceStart();
	for( cl_uint i = 0; i < frames; i++ ) {
		seeds[i] = ( mDeterministicSeeds || mFrameTimeMs > 0 ) ? this->nextSeed() : now + 0.0333f * (cl_float) i;
		weights[i] = mSampleCount / (cl_float) ( mSampleCount + 1 );
		mSampleCount++;
	}
	mCL->setKernelArg( mKernelPathTracing, 3, sizeof( camera_cl ), &mStructCam );
	mCL->executeBatch( mKernelPathTracing, frames, &seeds[0], &weights[0] );
	mHaveOutput = true;
}


void PathTracer::readImage( cl_float* target, cl_float* targetDebug ) {
	this->dropFrameAhead();
	mCL->readImageOutput( this->deliveredImage(), mWidth, mHeight, target );
	if( targetDebug != NULL ) {
		mCL->readImageOutput( mBufTextureDebug, mWidth, mHeight, targetDebug );
	}
}


void PathTracer::writeImage( const cl_float* source, cl_uint sampleCount ) {
	this->dropFrameAhead();
	mCL->updateImageReadOnly( mBufTextureOut, mWidth, mHeight, (cl_float*) source );
	mSampleCount = sampleCount;
	mHaveOutput = true;
}


void PathTracer::setTileRows( int y0, int y1 ) {
	this->dropFrameAhead();
	mCL->setTile( y0, y1 );
}


void PathTracer::setTileStripes( int stripeRows, int world, int rank ) {
	this->dropFrameAhead();
	mCL->setTileStripes( stripeRows, world, rank );
}


double PathTracer::getLastKernelMs() {
	map<cl_kernel, double> times = mCL->getKernelTimes();
	return times.count( mKernelPathTracing ) ? times[mKernelPathTracing] : 0.0;
}


/** Seconds since construction (reference: PathTracer.cpp:78-82). */
cl_float PathTracer::getTimeSinceStart() {
	const long long ms = std::chrono::duration_cast<std::chrono::milliseconds>(
		std::chrono::steady_clock::now() - mTimeSinceStart ).count();
	return ms * 0.001f;
}


cl_float PathTracer::nextSeed() {
	if( mFrameTimeMs > 0 ) {
		const long long ms = (long long) mFrameTimeMs * ( (long long) mSampleCount * mSeedStride + mSeedOffset + 1 );
		return ms * 0.001f;
	}
	if( mDeterministicSeeds ) {
		return 0.0333f * (cl_float) ( mSampleCount * mSeedStride + mSeedOffset + 1 );
	}
	return this->getTimeSinceStart();
}


/**
 * Kernel arguments that do not change per frame (reference: PathTracer.cpp:88-125).
 * pxDim = aspect * 2 * tan(fov/2) / width, the tangent evaluated in double.
 */
void PathTracer::initKernelArgs() {
	cl_float aspect = (cl_float) mWidth / (cl_float) mHeight;
	cl_float f = aspect * 2.0f * tan( (double) ( MathHelp::degToRad( mFOV ) / 2.0f ) );
	cl_float pxDim = f / (cl_float) mWidth;
	mPxDim = pxDim;

	char msg[128];
	snprintf( msg, 128, "[PathTracer] Aspect ratio: %g. Pixel size: %g", aspect, pxDim );
	Logger::logDebugVerbose( msg );

	cl_uint i = 2;   // 0: timeSinceStart, 1: pixelWeight
	mCL->setKernelArg( mKernelPathTracing, i++, sizeof( cl_float ), &pxDim );
	mCL->setKernelArg( mKernelPathTracing, i++, sizeof( camera_cl ), &mStructCam );

	if( Cfg::get().value<int>( Cfg::ACCEL_STRUCT ) != ACCELSTRUCT_BVH ) {
		Logger::logError( "[PathTracer] Unknown acceleration structure." );
		exit( EXIT_FAILURE );
	}
	cl_mem* handles[] = {
		&mBufBVH, &mBufFacesV, &mBufFacesN, &mBufVertices, &mBufNormals, &mBufMaterials, &mBufLights,
		&mBufTextureIn, &mBufTextureOut, &mBufTextureDebug
	};
	for( size_t h = 0; h < sizeof( handles ) / sizeof( handles[0] ); h++ ) {
		mCL->setKernelArg( mKernelPathTracing, i++, sizeof( cl_mem ), handles[h] );
	}
}


/**
 * Upload the scene and prepare the kernel (reference: PathTracer.cpp:136-230).
 */
void PathTracer::initOpenCLBuffers(
	const vector<cl_float>& vertices, const vector<cl_uint>& faces, const vector<cl_float>& normals,
	ModelLoader* ml, AccelStructure* accelStruc
) {
	if( mCL != NULL ) {
		delete mCL;
	}
	mCL = new CL();
	mTextureOut = NULL;
	mTextureDebugHost = NULL;
	mLights.clear();

	Logger::logInfo( "[PathTracer] Initializing OpenCL buffers ..." );
	Logger::indent( LOG_INDENT );

	std::chrono::steady_clock::time_point t0 = std::chrono::steady_clock::now();
	size_t bytes = this->initOpenCLBuffers_Faces( ml, vertices, faces, normals );
	logBuffer( "faces", msSince( t0 ), bytes );

	t0 = std::chrono::steady_clock::now();
	if( Cfg::get().value<short>( Cfg::ACCEL_STRUCT ) == ACCELSTRUCT_BVH ) {
		bytes = this->initOpenCLBuffers_BVH( (BVH*) accelStruc, ml, faces );
	}
	logBuffer( "BVH", msSince( t0 ), bytes );

	t0 = std::chrono::steady_clock::now();
	bytes = this->initOpenCLBuffers_Materials( ml );
	logBuffer( "material", msSince( t0 ), bytes );

	t0 = std::chrono::steady_clock::now();
	bytes = this->initOpenCLBuffers_Lights( ml );
	logBuffer( "light", msSince( t0 ), bytes );

	char msg[32];
	snprintf( msg, 32, "%lu", (unsigned long) mLights.size() );
	mCL->setReplacement( string( "#NUM_LIGHTS#" ), string( msg ) );

	t0 = std::chrono::steady_clock::now();
	bytes = this->initOpenCLBuffers_Textures();
	logBuffer( "texture", msSince( t0 ), bytes );

	Logger::indent( 0 );
	Logger::logInfo( "[PathTracer] ... Done." );

	mCL->loadProgram( Cfg::get().value<string>( Cfg::OPENCL_PROGRAM ) );
	mKernelPathTracing = mCL->createKernel( "pathTracing" );
	if( mGLWidget != NULL ) {
		mGLWidget->createKernelWindow( mCL );
	}

	this->initKernelArgs();
	mSampleCount = 0;
	mHaveOutput = false;
	mAhead.clear();
	mImages.clear();
	mImages.push_back( mBufTextureIn );
	mImages.push_back( mBufTextureOut );
}


/**
 * Flatten the BVH into bvhNode_cl[] and reorder the faces into leaf order
 * (reference: PathTracer.cpp:238-347).
 *
 * For every node in traversal order that is not dropped by skip-ahead:
 *   leaf  : bbMin.w = index of its first face in the new face array, bbMax.w = index of the second
 *           face or -1; the leaf's faces are appended to facesV / facesN.  Only the first two faces
 *           of an over-full leaf are referenced (reference behaviour).
 *   inner : bbMin.w = -1; bbMax.w = where to continue when the box is missed: the right sibling for
 *           a left child, for a right child the right sibling of the closest ancestor that is a left
 *           child, -1 on the right spine.  Ids are shifted by the number of dropped nodes before the
 *           target.
 */
void PathTracer::flattenBVH(
	const BVH* bvh, const ObjParser* op, const vector<cl_uint>& faces,
	vector<bvhNode_cl>* flatNodes, vector<cl_uint4>* flatFacesV, vector<cl_uint4>* flatFacesN
) {
	const vector<BVHNode*>& bvhNodes = bvh->nodes();
	const vector<cl_uint>& facesVN = op->facesVN();
	const vector<cl_int>& facesMtl = op->facesMtl();
	vector<bvhNode_cl>& mFlatNodes = *flatNodes;
	vector<cl_uint4>& mFlatFacesV = *flatFacesV;
	vector<cl_uint4>& mFlatFacesN = *flatFacesN;

	mFlatNodes.clear();
	mFlatFacesV.clear();
	mFlatFacesN.clear();
	mFlatNodes.reserve( bvhNodes.size() );
	mFlatFacesV.reserve( faces.size() / 3 );
	mFlatFacesN.reserve( faces.size() / 3 );

	bool skipNext = false;

	for( size_t i = 0; i < bvhNodes.size(); i++ ) {
		const BVHNode* node = bvhNodes[i];

		if( skipNext ) {
			skipNext = node->skipNextLeft;
			continue;
		}

		const cl_uint numFaces = (cl_uint) node->faces.size();
		bvhNode_cl sn;
		sn.bbMin.x = node->bbMin[0]; sn.bbMin.y = node->bbMin[1]; sn.bbMin.z = node->bbMin[2];
		sn.bbMax.x = node->bbMax[0]; sn.bbMax.y = node->bbMax[1]; sn.bbMax.z = node->bbMax[2];
		sn.bbMin.w = ( numFaces > 0 ) ? (cl_float) mFlatFacesV.size() + 0 : -1.0f;
		sn.bbMax.w = ( numFaces > 1 ) ? (cl_float) mFlatFacesV.size() + 1 : -1.0f;

		if( numFaces == 0 ) {
			skipNext = node->skipNextLeft;

			if( node->parent != NULL ) {
				/* escape target: climb while we are a right child */
				const BVHNode* climb = node;
				while( climb->parent != NULL && climb->parent->rightChild == climb ) {
					climb = climb->parent;
				}
				if( climb->parent != NULL ) {
					const BVHNode* target = climb->parent->rightChild;
					sn.bbMax.w = (cl_float) ( target->id - target->numSkipsToHere );
				}
			}
		}

		mFlatNodes.push_back( sn );

		for( cl_uint j = 0; j < numFaces; j++ ) {
			const Tri& tri = node->faces[j];
			cl_uint4 fv, fn;
			fv.x = faces[tri.face.w * 3];
			fv.y = faces[tri.face.w * 3 + 1];
			fv.z = faces[tri.face.w * 3 + 2];
			fv.w = (cl_uint) facesMtl[tri.face.w];

			const size_t ni = (size_t) tri.normals.w * 3;
			const bool hasNormals = ni + 2 < facesVN.size();
			fn.x = hasNormals ? facesVN[ni] : 0;
			fn.y = hasNormals ? facesVN[ni + 1] : 0;
			fn.z = hasNormals ? facesVN[ni + 2] : 0;
			fn.w = 0;

			mFlatFacesV.push_back( fv );
			mFlatFacesN.push_back( fn );
		}
	}
}


/** Flatten (see flattenBVH) and upload (reference: PathTracer.cpp:238-347). */
size_t PathTracer::initOpenCLBuffers_BVH( BVH* bvh, ModelLoader* ml, const vector<cl_uint>& faces ) {
	PathTracer::flattenBVH( bvh, ml->getObjParser(), faces, &mFlatNodes, &mFlatFacesV, &mFlatFacesN );

	const size_t bytesBVH = sizeof( bvhNode_cl ) * mFlatNodes.size();
	mBufBVH = mCL->createBufferFromPtr( mFlatNodes.data(), bytesBVH );

	char msg[32];
	snprintf( msg, 32, "%lu", (unsigned long) mFlatNodes.size() );
	mCL->setReplacement( string( "#BVH_NUM_NODES#" ), string( msg ) );

	const size_t bytesFV = sizeof( cl_uint4 ) * mFlatFacesV.size();
	mBufFacesV = mCL->createBufferFromPtr( mFlatFacesV.data(), bytesFV );

	const size_t bytesFN = sizeof( cl_uint4 ) * mFlatFacesN.size();
	mBufFacesN = mCL->createBufferFromPtr( mFlatFacesN.data(), bytesFN );

	return bytesBVH + bytesFV + bytesFN;
}


/** Vertices and normals as float4 (reference: PathTracer.cpp:357-380). */
size_t PathTracer::initOpenCLBuffers_Faces(
	ModelLoader* ml, const vector<cl_float>& vertices, const vector<cl_uint>& faces, const vector<cl_float>& normals
) {
	(void) ml;
	(void) faces;
	vector<cl_float4> vertices4 = AccelStructure::packFloatAsFloat4( &vertices );
	vector<cl_float4> normals4 = AccelStructure::packFloatAsFloat4( &normals );

	const size_t bytesV = sizeof( cl_float4 ) * vertices4.size();
	const size_t bytesN = sizeof( cl_float4 ) * normals4.size();

	mBufVertices = mCL->createBufferFromPtr( vertices4.data(), bytesV );
	mBufNormals = mCL->createBufferFromPtr( normals4.data(), bytesN );

	return bytesV + bytesN;
}


/**
 * Lights (reference: PathTracer.cpp:387-428).  Without lights one zeroed dummy is uploaded so that
 * the kernel argument is valid; NUM_LIGHTS stays 0.
 */
size_t PathTracer::initOpenCLBuffers_Lights( ModelLoader* ml ) {
	vector<light_t> lights = ml->getObjParser()->getLights();

	for( size_t i = 0; i < lights.size(); i++ ) {
		light_cl light;
		memset( &light, 0, sizeof( light ) );
		light.pos = lights[i].pos;
		light.rgb = lights[i].rgb;
		light.data.x = lights[i].type;
		if( light.data.x == 2 ) {
			light.data.y = lights[i].radius;   // orb
		}
		mLights.push_back( light );
	}

	vector<light_cl> upload = mLights;
	if( upload.empty() ) {
		light_cl dummy;
		memset( &dummy, 0, sizeof( dummy ) );
		upload.push_back( dummy );
	}

	const size_t bytes = sizeof( light_cl ) * upload.size();
	mBufLights = mCL->createBufferFromPtr( upload.data(), bytes );
	return bytes;
}


size_t PathTracer::initOpenCLBuffers_Materials( ModelLoader* ml ) {
	return this->initOpenCLBuffers_MaterialsRGB( ml->getObjParser()->getMaterials() );
}


/**
 * Materials in the layout of the selected BRDF; the Kd of a material named "sky_light" becomes the
 * SKY_LIGHT define, written with "%f" like the reference (reference: PathTracer.cpp:435-519).
 */
size_t PathTracer::initOpenCLBuffers_MaterialsRGB( const vector<material_t>& materials ) {
	const int brdf = Cfg::get().value<int>( Cfg::RENDER_BRDF );
	size_t bytesMTL = 0;
	string skyLight = "(float4)( 1.0f, 1.0f, 1.0f, 0.0f )";

	for( size_t i = 0; i < materials.size(); i++ ) {
		if( materials[i].mtlName == "sky_light" ) {
			const cl_float4 Kd = materials[i].Kd;
			char msg[128];
			snprintf( msg, 128, "(float4)( %f, %f, %f, 0.0f )", Kd.x, Kd.y, Kd.z );
			skyLight = msg;
		}
	}

	if( brdf == 0 ) {
		vector<material_schlick_rgb> materialsCL( materials.size() );
		for( size_t i = 0; i < materials.size(); i++ ) {
			material_schlick_rgb& mtl = materialsCL[i];
			mtl.data.x = materials[i].d;
			mtl.data.y = materials[i].Ni;
			mtl.data.z = materials[i].p;
			mtl.data.w = materials[i].rough;
			mtl.rgbDiff = materials[i].Kd;
			mtl.rgbSpec = materials[i].Ks;
		}
		bytesMTL = sizeof( material_schlick_rgb ) * materialsCL.size();
		mBufMaterials = mCL->createBufferFromPtr( materialsCL.data(), bytesMTL );
	}
	else if( brdf == 1 ) {
		vector<material_shirley_ashikhmin_rgb> materialsCL( materials.size() );
		for( size_t i = 0; i < materials.size(); i++ ) {
			material_shirley_ashikhmin_rgb& mtl = materialsCL[i];
			memset( mtl.data, 0, sizeof( mtl.data ) );
			mtl.data[0] = materials[i].d;
			mtl.data[1] = materials[i].Ni;
			mtl.data[2] = materials[i].nu;
			mtl.data[3] = materials[i].nv;
			mtl.data[4] = materials[i].Rs;
			mtl.data[5] = materials[i].Rd;
			mtl.rgbDiff = materials[i].Kd;
			mtl.rgbSpec = materials[i].Ks;
		}
		bytesMTL = sizeof( material_shirley_ashikhmin_rgb ) * materialsCL.size();
		mBufMaterials = mCL->createBufferFromPtr( materialsCL.data(), bytesMTL );
	}
	else {
		Logger::logError( "[PathTracer] Unknown BRDF selected." );
		exit( EXIT_FAILURE );
	}

	mCL->setReplacement( string( "#SKY_LIGHT#" ), skyLight );
	return bytesMTL;
}


/** The three RGBA-float images (reference: PathTracer.cpp:525-533); the host copy is pinned. */
size_t PathTracer::initOpenCLBuffers_Textures() {
	const size_t n = (size_t) mWidth * mHeight * 4;
	mTextureOut = (cl_float*) mCL->allocHost( n * sizeof( cl_float ) );
	if( mTextureOut == NULL ) {
		Logger::logError( "[PathTracer] Could not allocate the host image." );
		exit( EXIT_FAILURE );
	}
	memset( mTextureOut, 0, n * sizeof( cl_float ) );

	mBufTextureIn = mCL->createImage2DReadOnly( mWidth, mHeight, mTextureOut );
	mBufTextureOut = mCL->createImage2DWriteOnly( mWidth, mHeight );
	mBufTextureDebug = mCL->createImage2DWriteOnly( mWidth, mHeight );

	return sizeof( cl_float ) * n * 3;
}


void PathTracer::moveSun( const int key ) {
	(void) key;   /* the body is commented out in the reference (PathTracer.cpp:540-570) */
	this->resetSampleCount();
}


/** Reset the sample counter: the next frame ignores the history (reference: PathTracer.cpp:576-578). */
void PathTracer::resetSampleCount() {
	this->dropFrameAhead();
	mSampleCount = 0;
}


void PathTracer::setCamera( Camera* camera ) {
	mCamera = camera;
}


/** Camera focus at a pixel; negative = no depth of field (reference: PathTracer.cpp:596-602). */
void PathTracer::setFocus( int x, int y ) {
	mStructCam.focusPoint.x = x;
	mStructCam.focusPoint.y = y;

	this->resetSampleCount();
	if( mGLWidget != NULL ) {
		mGLWidget->resetRenderTime();
	}
}


void PathTracer::setFOV( cl_float fov ) {
	mFOV = fov;
}


void PathTracer::setWidthAndHeight( cl_uint width, cl_uint height ) {
	mWidth = width;
	mHeight = height;
}


/** eye, w = view direction, u = right, v = up (reference: PathTracer.cpp:628-652). */
void PathTracer::updateEyeBuffer() {
	const glm::vec3 c = mCamera->getAdjustedCenter_glmVec3();
	const glm::vec3 eye = mCamera->getEye_glmVec3();
	const glm::vec3 up = mCamera->getUp_glmVec3();

	const glm::vec3 w = glm::normalize( c - eye );
	const glm::vec3 u = glm::normalize( glm::cross( w, up ) );
	const glm::vec3 v = glm::normalize( glm::cross( u, w ) );

	const glm::vec3* src[4] = { &eye, &w, &u, &v };
	cl_float3* dst[4] = { &mStructCam.eye, &mStructCam.w, &mStructCam.u, &mStructCam.v };
	for( int i = 0; i < 4; i++ ) {
		dst[i]->x = src[i]->x;
		dst[i]->y = src[i]->y;
		dst[i]->z = src[i]->z;
	}
}
