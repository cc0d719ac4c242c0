/*
 * GLWidget -- headless stand-in for the reference's Qt/OpenGL widget (source/qt/GLWidget.{h,cpp}).
 *
 * The reference's GLWidget owns the PathTracer and the Camera, loads a model (loadModel,
 * qt/GLWidget.cpp:339-387), asks the PathTracer for one image per timer tick (paintGL, :504-517) and
 * blits it with OpenGL.  Everything that touches Qt or OpenGL is display-only and out of scope
 * (SURVEY.md section 2, rows 10-11); what remains is the owner object PathTracer and Camera call back
 * into: cameraUpdate(), createKernelWindow(), resetRenderTime().  paintGL() becomes renderFrame().
 */
#ifndef GLWIDGET_H
#define GLWIDGET_H

#include <string>
#include <vector>

#include "../cl_types.h"
#include "../Camera.h"
#include "../CL.h"
#include "../ModelLoader.h"
#include "../PathTracer.h"
#include "../accelstructures/BVH.h"

using std::string;
using std::vector;


class GLWidget {

	public:
		GLWidget();
		~GLWidget();
		void cameraUpdate();
		void createKernelWindow( CL* cl );
		void loadModel( string filepath, string filename );
		/** Additive: a scene that is already in memory (synthetic generators); takes ownership of `ml`. */
		void loadModel( ModelLoader* ml );
		void resetRenderTime();
		/** paintGL without the painting: one PathTracer::generateImage (qt/GLWidget.cpp:504-517). */
		const vector<cl_float>& renderFrame( bool withDebugImage );
		void toggleViewDebug() { mViewDebug = !mViewDebug; }

		Camera* getCamera() { return mCamera; }
		PathTracer* getPathTracer() { return mPathTracer; }
		const vector<cl_float>& getTextureOut() const { return mTextureOut; }
		const vector<cl_float>& getTextureDebug() const { return mTextureDebug; }
		double getBvhBuildSeconds() const { return mBvhBuildSeconds; }
		cl_uint getBvhNumNodes() const { return mBvhNumNodes; }
		cl_uint getBvhNumSkipped() const { return mBvhNumSkipped; }
		cl_uint getBvhDepth() const { return mBvhDepth; }
		bool isReady() const { return mDoRendering; }

	private:
		void deleteOldModel();

		bool mDoRendering;
		bool mViewDebug;
		Camera* mCamera;
		PathTracer* mPathTracer;
		CL* mKernelWindowCL;

		vector<cl_uint> mFaces;
		vector<cl_float> mNormals;
		vector<cl_float> mVertices;
		vector<cl_float> mTextureOut;
		vector<cl_float> mTextureDebug;

		double mBvhBuildSeconds;
		cl_uint mBvhNumNodes, mBvhNumSkipped, mBvhDepth;

};

#endif
