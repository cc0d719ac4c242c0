/* ModelLoader -- reference: source/ModelLoader.{h,cpp}: owns the ObjParser; per-object face lists for the BVH. */
#ifndef MODELLOADER_H
#define MODELLOADER_H

#include "ObjParser.h"        /* object3D, cl types, std::string / std::vector */
#include "utils.h"

class ModelLoader {
	ObjParser* mObjParser;

	public:
		ModelLoader();
		~ModelLoader();

		/** Parse <filepath><filename> (an .obj) with its sibling .mtl / .lights. */
		void loadModel( std::string filepath, std::string filename );
		ObjParser* getObjParser();

		/* ( a, b, c, offset + running index ) per triangle of one object (ModelLoader.cpp:28-57) */
		static void getFacesOfObject(
			const object3D& object, std::vector<cl_uint4>* faces, cl_int offset );
		static void getFaceNormalsOfObject(
			const object3D& object, std::vector<cl_uint4>* faceNormals, cl_int offset );
};

#endif
