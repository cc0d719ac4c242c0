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

using std::vector;


struct camera_t {
	glm::vec3 eye;
	glm::vec3 center;
	glm::vec3 up;
	glm::vec3 right;
	glm::vec2 rot;
};


class GLWidget;


class Camera {

	public:
		Camera( GLWidget* parent );
		void cameraMoveBackward();
		void cameraMoveDown();
		void cameraMoveForward();
		void cameraMoveLeft();
		void cameraMoveRight();
		void cameraMoveUp();
		void cameraReset();
		glm::vec3 getAdjustedCenter_glmVec3();
		glm::vec3 getCenter_glmVec3();
		vector<float> getEye();
		glm::vec3 getEye_glmVec3();
		float getRotX();
		float getRotY();
		float getSpeed();
		glm::vec3 getUp_glmVec3();
		void setSpeed( float speed );
		void updateCameraRot( int moveX, int moveY );
		/** Additive: place the eye directly (headless drivers have no key events). */
		void setEye( float x, float y, float z );

	protected:
		void updateParent();

	private:
		GLWidget* mParent;
		float mCameraSpeed;
		camera_t mCamera;

};

#endif
