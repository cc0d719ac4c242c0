/*
 * Camera -- reference: source/Camera.{h,cpp}.  Same state (eye, center, up, rot), same movement and
 * rotation arithmetic, same getters; every change notifies the owner, which resets the sample count
 * (Camera.cpp:248-251 -> GLWidget::cameraUpdate, qt/GLWidget.cpp:80-84).
 */
#ifndef CAMERA_H
#define CAMERA_H

#include <cmath>
#include <vector>

#include "Cfg.h"
#include "MathHelp.h"
#include "glm_lite.h"

class GLWidget;

/* eye position; viewing direction as a unit offset (`center`, see getAdjustedCenter_glmVec3); rot in degrees */
struct camera_t {
	glm::vec3 eye, center, up, right;
	glm::vec2 rot;
};

class Camera {
	public:
		explicit Camera( GLWidget* parent );

		/* --- one step of getSpeed() along the view axes; cameraReset() re-reads camera.* from the config */
		void cameraMoveForward();   void cameraMoveBackward();
		void cameraMoveLeft();      void cameraMoveRight();
		void cameraMoveUp();        void cameraMoveDown();
		void cameraReset();
		void updateCameraRot( int moveX, int moveY );   /* mouse deltas in pixels = degrees */
		void setEye( float x, float y, float z );       /* additive: headless drivers have no key events */

		/* --- speed of the steps above */
		float getSpeed();
		void setSpeed( float speed );

		/* --- what PathTracer::updateEyeBuffer and the overlay read */
		glm::vec3 getEye_glmVec3();
		std::vector<float> getEye();
		glm::vec3 getCenter_glmVec3();
		glm::vec3 getAdjustedCenter_glmVec3();          /* ( eye.x + c.x, eye.y - c.y, eye.z - c.z ) */
		glm::vec3 getUp_glmVec3();
		float getRotX();
		float getRotY();

	protected:
		void updateParent();
		void stepAlongView( double sign );
		void stepSideways( double sign );
		void stepVertically( float sign );

	private:
		camera_t mCamera;
		float mCameraSpeed;
		GLWidget* mParent;
};

#endif
