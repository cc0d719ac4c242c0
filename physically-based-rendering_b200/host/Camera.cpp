#include "Camera.h"

using std::string;
using std::vector;

#include "qt/GLWidget.h"


Camera::Camera( GLWidget* parent ) {
	mParent = parent;
	mCameraSpeed = Cfg::get().value<float>( Cfg::CAM_SPEED );
	this->cameraReset();
}


/* One step of `speed` along the view axes (reference: Camera.cpp:20-75).  rot.x turns around the
 * vertical axis, rot.y tilts; the forward vector is (sin x cos y, -sin y, -cos x cos y).
 *
 * The reference calls unqualified sin( float ) / cos( float ) / fabs() with only <cmath> included: with the
 * toolchain it was written for these are the C functions, i.e. the expressions are evaluated in binary64 and
 * rounded when stored.  Spelled out here with explicit doubles, so that the result does not depend on which
 * overloads a standard library happens to put into the global namespace (checked against the reference's own
 * Camera.cpp in tests/test_gpu_host.py::test_product_equals_reference_renderer). */
void Camera::stepAlongView( double sign ) {
	const double rx = MathHelp::degToRad( mCamera.rot.x ), ry = MathHelp::degToRad( mCamera.rot.y );
	const double fx = sin( rx ) * cos( ry ) * mCameraSpeed;
	const double fy = sin( ry ) * mCameraSpeed;
	const double fz = cos( rx ) * cos( ry ) * mCameraSpeed;
	/* forward = ( +fx, -fy, -fz ); a - b and a + ( -b ) are the same binary64 sum, rounded once when stored */
	mCamera.eye.x = (float) ( (double) mCamera.eye.x + sign * fx );
	mCamera.eye.y = (float) ( (double) mCamera.eye.y - sign * fy );
	mCamera.eye.z = (float) ( (double) mCamera.eye.z - sign * fz );
	this->updateParent();
}


void Camera::stepSideways( double sign ) {
	const double rx = MathHelp::degToRad( mCamera.rot.x );
	mCamera.eye.x = (float) ( (double) mCamera.eye.x + sign * ( cos( rx ) * mCameraSpeed ) );
	mCamera.eye.z = (float) ( (double) mCamera.eye.z + sign * ( sin( rx ) * mCameraSpeed ) );
	this->updateParent();
}


void Camera::stepVertically( float sign ) {
	mCamera.eye.y += sign * mCameraSpeed;
	this->updateParent();
}


void Camera::cameraMoveForward() { this->stepAlongView( 1.0 ); }
void Camera::cameraMoveBackward() { this->stepAlongView( -1.0 ); }
void Camera::cameraMoveRight() { this->stepSideways( 1.0 ); }
void Camera::cameraMoveLeft() { this->stepSideways( -1.0 ); }
void Camera::cameraMoveUp() { this->stepVertically( 1.0f ); }
void Camera::cameraMoveDown() { this->stepVertically( -1.0f ); }


/** Eye and center from the config; the center is used as a normalised direction (Camera.cpp:80-95). */
void Camera::cameraReset() {
	mCamera.eye.x = Cfg::get().value<float>( Cfg::CAM_EYE_X );
	mCamera.eye.y = Cfg::get().value<float>( Cfg::CAM_EYE_Y );
	mCamera.eye.z = Cfg::get().value<float>( Cfg::CAM_EYE_Z );
	mCamera.up = glm::vec3( 0.0f, 1.0f, 0.0f );
	mCamera.rot.x = 0.0f;
	mCamera.rot.y = 0.0f;
	this->updateCameraRot( 0, 0 );
	mCamera.center.x = Cfg::get().value<float>( Cfg::CAM_CENTER_X );
	mCamera.center.y = Cfg::get().value<float>( Cfg::CAM_CENTER_Y );
	mCamera.center.z = Cfg::get().value<float>( Cfg::CAM_CENTER_Z );
	mCamera.center = glm::normalize( mCamera.center );
}


/** The point the camera looks at: eye + (c.x, -c.y, -c.z) (Camera.cpp:101-107). */
glm::vec3 Camera::getAdjustedCenter_glmVec3() {
	return glm::vec3(
		mCamera.eye.x + mCamera.center.x,
		mCamera.eye.y - mCamera.center.y,
		mCamera.eye.z - mCamera.center.z
	);
}


glm::vec3 Camera::getCenter_glmVec3() { return mCamera.center; }
glm::vec3 Camera::getEye_glmVec3() { return mCamera.eye; }
glm::vec3 Camera::getUp_glmVec3() { return mCamera.up; }
float Camera::getRotX() { return mCamera.rot.x; }
float Camera::getRotY() { return mCamera.rot.y; }
float Camera::getSpeed() { return mCameraSpeed; }
void Camera::setSpeed( float speed ) { mCameraSpeed = speed; }


vector<float> Camera::getEye() {
	vector<float> eye;
	eye.push_back( mCamera.eye.x );
	eye.push_back( mCamera.eye.y );
	eye.push_back( mCamera.eye.z );
	return eye;
}


void Camera::setEye( float x, float y, float z ) {
	mCamera.eye = glm::vec3( x, y, z );
	this->updateParent();
}


/**
 * Update the viewing direction from a mouse movement in pixels = degrees (Camera.cpp:192-241).
 */
void Camera::updateCameraRot( int moveX, int moveY ) {
	mCamera.rot.x -= moveX;
	mCamera.rot.y -= moveY;

	if( mCamera.rot.x >= 360.0f ) { mCamera.rot.x = 0.0f; }
	else if( mCamera.rot.x < 0.0f ) { mCamera.rot.x = 360.0f; }

	if( mCamera.rot.y > 90.0f ) { mCamera.rot.y = 90.0f; }
	else if( mCamera.rot.y < -90.0f ) { mCamera.rot.y = -90.0f; }

	const double rx = MathHelp::degToRad( mCamera.rot.x ), ry = MathHelp::degToRad( mCamera.rot.y );

	mCamera.center.x = sin( rx ) - fabs( sin( ry ) ) * sin( rx );
	mCamera.center.y = sin( ry );
	mCamera.center.z = cos( rx ) - fabs( sin( ry ) ) * cos( rx );

	const bool lookUp = ( mCamera.center.y == 1.0f ), lookDown = ( mCamera.center.y == -1.0f );
	mCamera.up.x = lookUp ? sin( rx ) : ( lookDown ? -sin( rx ) : 0.0f );
	mCamera.up.y = ( lookUp || lookDown ) ? 0.0f : 1.0f;
	mCamera.up.z = lookUp ? -cos( rx ) : ( lookDown ? cos( rx ) : 0.0f );

	this->updateParent();
}


void Camera::updateParent() {
	if( mParent != NULL ) {
		mParent->cameraUpdate();
	}
}
