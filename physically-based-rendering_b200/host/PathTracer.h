/*
 * PathTracer -- the reference's render driver (source/PathTracer.{h,cpp}) on top of the CL shim.
 *
 * Same public API (PathTracer.h:83-95): generateImage, initOpenCLBuffers, moveSun, resetSampleCount,
 * setCamera, setFocus, setFOV, setWidthAndHeight; same host-side structs (camera_cl, light_cl,
 * material_*, bvhNode_cl, PathTracer.h:25-73 -- here typedefs of the C-ABI records); same buffer
 * packing and the same flattening of the BVH into bvhNode_cl[] + leaf-ordered faces
 * (PathTracer.cpp:238-347), which is the layout contract of the kernel.
 *
 * What differs, deliberately:
 *   - generateImage keeps the accumulated frame on the device: imageIn/imageOut are swapped from
 *     frame to frame instead of reading the frame back and uploading it again (PathTracer.cpp:61-66);
 *     the host copy is refreshed by one read-back per frame into pinned memory.  The debug image is
 *     only read back when the caller passes a vector for it.
 *   - the seed is wall-clock seconds as in the reference (PathTracer.cpp:78-82) unless a deterministic
 *     schedule seed_k = 0.0333f * (k + 1) is selected (setDeterministicSeeds), which is what makes
 *     frames reproducible and testable.
 *   - additive: renderFrames(n) (n frames, no read-back until asked), readImage(), tile rows for
 *     multi-GPU sharding, counters.
 */
#ifndef PATH_TRACER_H
#define PATH_TRACER_H

#include <chrono>
#include <deque>
#include <string>
#include <vector>

#include "Camera.h"
#include "CL.h"
#include "Cfg.h"
#include "MtlParser.h"
#include "accelstructures/BVH.h"

using std::vector;

typedef pbr_camera camera_cl;
typedef pbr_light light_cl;
typedef pbr_material_schlick material_schlick_rgb;
typedef pbr_material_sa material_shirley_ashikhmin_rgb;
typedef pbr_bvh_node bvhNode_cl;

struct face_cl {
	cl_uint4 vertices; // w: material
	cl_uint4 normals;
};


class Camera;
class GLWidget;


class PathTracer {

	public:
		PathTracer( GLWidget* parent );
		~PathTracer();
		vector<cl_float> generateImage( vector<cl_float>* textureDebug );
		void initOpenCLBuffers(
			const vector<cl_float>& vertices, const vector<cl_uint>& faces, const vector<cl_float>& normals,
			ModelLoader* ml, AccelStructure* bvh
		);
		void moveSun( const int key );
		void resetSampleCount();
		void setCamera( Camera* camera );
		void setFocus( int x, int y );
		void setFOV( cl_float fov );
		void setWidthAndHeight( cl_uint width, cl_uint height );

		/* ---- additive ---- */
		/** Use seed_k = 0.0333f * (k + 1) for frame k instead of wall-clock seconds. */
		void setDeterministicSeeds( bool enabled ) { this->dropFrameAhead(); mDeterministicSeeds = enabled; }
		/** Deterministic schedule for sample-sharded multi-GPU runs: frame k of this renderer uses the
		 *  global index k * stride + offset, i.e. seed = 0.0333f * (k * stride + offset + 1). */
		void setSeedSchedule( cl_uint stride, cl_uint offset ) { this->dropFrameAhead(); mSeedStride = stride; mSeedOffset = offset; }
		/** Simulated clock: frame with global index g is "rendered g * ms + ms milliseconds after start", so
		 *  its seed is what the reference computes from its wall clock, ( ms * ( g + 1 ) ) * 0.001f
		 *  (reference: PathTracer.cpp:78-82), but reproducible.  0 = off. */
		void setFrameTimeMs( cl_uint ms ) { this->dropFrameAhead(); mFrameTimeMs = ms; }
		/** Render `frames` frames back to back without reading anything back. */
		void renderFrames( cl_uint frames );
		/** One frame, result read into `target` (W*H*4 floats; pinned memory recommended). */
		void generateImageInto( cl_float* target, cl_float* targetDebug );
		/** Read the current accumulated frame (and optionally the debug image) from the device. */
		void readImage( cl_float* target, cl_float* targetDebug );
		/** Replace the accumulated frame on the device (resume from a checkpoint). */
		void writeImage( const cl_float* source, cl_uint sampleCount );
		void setTileRows( int y0, int y1 );
		/** Interleaved rows for load balance: stripes of `stripeRows` rows, stripe index % world == rank. */
		void setTileStripes( int stripeRows, int world, int rank );
		/** Multi-GPU, one process per GPU: this renderer is rank `rank` of `world`; `ncclId` are the 128
		 *  bytes of CL::commUniqueId() of rank 0.  From then on every frame ends with ONE collective inside the library:
		 *    SHARD_SPP      every rank renders whole frames with its own seeds (seed schedule stride = world, offset =
		 *                   rank); the frame delivered by generateImage / readImage is the mean over ranks;
		 *    SHARD_ROWS     rank r renders the row block pbr_tile_rows gives it, the others are gathered in;
		 *    SHARD_STRIPES  the same with interleaved stripes of rows (load balance).
		 *  Replaces the reference's single-device assumption (CL.cpp:355, 470, 521). */
		enum { SHARD_SPP = 0, SHARD_ROWS = 1, SHARD_STRIPES = 2, SHARD_NONE = 3 };   /* NONE: whole frames, no collective */
		bool setRanks( int rank, int world, const void* ncclId, int sharding );
		/** Another sharding on the communicator setRanks created (the accumulation restarts). */
		bool setSharding( int sharding );
		/** The render stream waits -- on the device -- for every collective enqueued so far. */
		void commFence() { if( mWorld > 1 ) { mCL->commFence(); } }
		int getWorld() const { return mWorld; }
		/** -1 automatic (default), 0 the reference's visiting order, 1 the ordered walk (pbr_set_traversal). */
		void setTraversal( int mode ) { mCL->setTraversal( mode ); }
		cl_uint getSampleCount() const { return mSampleCount - (cl_uint) mAhead.size(); }
		/** Render ahead: generateImage() starts tracing the next `depth` (0..3) frames before it waits for the copy of this
		 *  one, so the read-back -- and the tail of every frame's launches -- is hidden behind the following frames
		 *  (each frame traced ahead keeps an image of its own until it has been delivered).  Anything that changes what
		 *  the next frame should look like (camera, focus, sample count, tile, seeds) discards the frames traced ahead.
		 *  0 by default: then a frame is only ever traced inside the call that returns it, as upstream. */
		void setRenderAhead( int depth );
		cl_uint getWidth() const { return mWidth; }
		cl_uint getHeight() const { return mHeight; }
		CL* getCL() { return mCL; }
		cl_mem getImageHandle() const { return mBufTextureOut; }
		double getLastKernelMs();
		/** The flattening of initOpenCLBuffers_BVH without the upload (no device needed). */
		static void flattenBVH(
			const BVH* bvh, const ObjParser* op, const vector<cl_uint>& faces,
			vector<bvhNode_cl>* flatNodes, vector<cl_uint4>* flatFacesV, vector<cl_uint4>* flatFacesN
		);
		/** The flattened scene as uploaded (for tests and the explicit-ray entry points). */
		const vector<bvhNode_cl>& getFlatNodes() const { return mFlatNodes; }
		const vector<cl_uint4>& getFlatFacesV() const { return mFlatFacesV; }
		const vector<cl_uint4>& getFlatFacesN() const { return mFlatFacesN; }
		cl_mem getBufBVH() const { return mBufBVH; }
		cl_mem getBufFacesV() const { return mBufFacesV; }
		cl_mem getBufVertices() const { return mBufVertices; }
		cl_mem getBufLights() const { return mBufLights; }
		cl_uint getNumLights() const { return (cl_uint) mLights.size(); }
		/** The camera as the next frame will see it (additive accessor). */
		const camera_cl& getCameraStruct() { this->updateEyeBuffer(); return mStructCam; }
		cl_float getPxDim() const { return mPxDim; }

	protected:
		void clPathTracing( cl_float timeSinceStart );
		cl_float getTimeSinceStart();
		cl_float nextSeed();
		void initKernelArgs();
		size_t initOpenCLBuffers_BVH( BVH* bvh, ModelLoader* ml, const vector<cl_uint>& faces );
		size_t initOpenCLBuffers_Faces(
			ModelLoader* ml,
			const vector<cl_float>& vertices, const vector<cl_uint>& faces, const vector<cl_float>& normals
		);
		size_t initOpenCLBuffers_Lights( ModelLoader* ml );
		size_t initOpenCLBuffers_Materials( ModelLoader* ml );
		size_t initOpenCLBuffers_MaterialsRGB( const vector<material_t>& materials );
		size_t initOpenCLBuffers_Textures();
		void updateEyeBuffer();
		void launchFrame();

	private:
		cl_uint mHeight;
		cl_uint mWidth;
		cl_float mFOV;
		cl_float mPxDim;
		cl_uint mSampleCount;
		bool mDeterministicSeeds;
		cl_uint mSeedStride, mSeedOffset;
		cl_uint mFrameTimeMs;
		int mRenderAhead;
		struct AheadFrame {           // a frame launched but not delivered yet, and the state before it was launched
			cl_mem in, out, delivered;
			cl_uint sampleCount, combines;
			bool haveOutput;
		};
		std::deque<AheadFrame> mAhead;
		std::vector<cl_mem> mImages;  // the accumulation images: In / Out walk around this ring
		void advanceImages();
		void launchAhead();
		int aheadDepth() const;
		void dropFrameAhead( size_t keep = 0 );
		bool mHaveOutput;             // imageOut holds a frame that the next launch must read as imageIn
		int mRank, mWorld, mSharding;
		enum { NUM_DISPLAYS = 4 };
		cl_mem mBufTextureDisplay[NUM_DISPLAYS]; // SHARD_SPP: the mean over ranks of frame k lands in [k % NUM_DISPLAYS]
		cl_uint mCombines;
		void combineFrame();
		cl_mem deliveredImage() const;

		cl_float* mTextureOut;        // pinned host copy of the frame, W*H*4
		cl_float* mTextureDebugHost;  // pinned, allocated on first use

		cl_kernel mKernelPathTracing;

		cl_mem mBufBVH;
		cl_mem mBufFacesV;
		cl_mem mBufFacesN;
		cl_mem mBufVertices;
		cl_mem mBufNormals;
		cl_mem mBufMaterials;

		camera_cl mStructCam;
		cl_mem mBufTextureIn;
		cl_mem mBufTextureOut;
		cl_mem mBufTextureDebug;

		vector<light_cl> mLights;
		cl_mem mBufLights;

		vector<bvhNode_cl> mFlatNodes;
		vector<cl_uint4> mFlatFacesV;
		vector<cl_uint4> mFlatFacesN;

		GLWidget* mGLWidget;
		Camera* mCamera;
		CL* mCL;
		std::chrono::steady_clock::time_point mTimeSinceStart;

};

#endif
