#include "GLWidget.h"


/** Reference: qt/GLWidget.cpp:12-34 (without QTimer / GL state). */
GLWidget::GLWidget() {
	mDoRendering = false;
	mViewDebug = false;
	mKernelWindowCL = NULL;
	mBvhBuildSeconds = 0.0;
	mBvhNumNodes = mBvhNumSkipped = mBvhDepth = 0;
	mPathTracer = new PathTracer( this );
	mCamera = new Camera( this );
	mPathTracer->setCamera( mCamera );
}


GLWidget::~GLWidget() {
	delete mPathTracer;
	delete mCamera;
}


/** The camera changed: restart the accumulation (reference: qt/GLWidget.cpp:80-84). */
void GLWidget::cameraUpdate() {
	if( mPathTracer != NULL ) {
		mPathTracer->resetSampleCount();
	}
}


/** The reference opens the "Info -> Kernels" window on this CL object (qt/GLWidget.cpp:90-99). */
void GLWidget::createKernelWindow( CL* cl ) {
	mKernelWindowCL = cl;
}


void GLWidget::deleteOldModel() {
	mDoRendering = false;
	mFaces.clear();
	mNormals.clear();
	mVertices.clear();
}


/** Load 3D model and get ready to render it (reference: qt/GLWidget.cpp:339-387). */
void GLWidget::loadModel( string filepath, string filename ) {
	ModelLoader* ml = new ModelLoader();
	ml->loadModel( filepath, filename );
	this->loadModel( ml );
}


void GLWidget::loadModel( ModelLoader* ml ) {
	this->deleteOldModel();

	ObjParser* op = ml->getObjParser();
	mFaces = op->facesV();
	mNormals = op->normals();
	mVertices = op->vertices();

	AccelStructure* accelStruct = NULL;
	if( Cfg::get().value<short>( Cfg::ACCEL_STRUCT ) == ACCELSTRUCT_BVH ) {
		BVH* bvh = new BVH( op->objects(), mVertices, mNormals );
		mBvhBuildSeconds = bvh->getBuildSeconds();
		mBvhNumNodes = (cl_uint) bvh->nodes().size();
		mBvhNumSkipped = bvh->getNumSkipped();
		mBvhDepth = bvh->getDepth();
		accelStruct = bvh;
	}
	else {
		Logger::logError( "[GLWidget] Unknown acceleration structure." );
		exit( EXIT_FAILURE );
	}

	mPathTracer->initOpenCLBuffers( mVertices, mFaces, mNormals, ml, accelStruct );

	delete ml;
	delete accelStruct;

	mDoRendering = true;
}


void GLWidget::resetRenderTime() {}


const vector<cl_float>& GLWidget::renderFrame( bool withDebugImage ) {
	if( !mDoRendering || mVertices.size() <= 0 ) {
		return mTextureOut;
	}
	mTextureOut = mPathTracer->generateImage( ( withDebugImage || mViewDebug ) ? &mTextureDebug : NULL );
	return mTextureOut;
}
