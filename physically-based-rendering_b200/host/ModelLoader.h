/* ModelLoader -- reference: source/ModelLoader.{h,cpp}. */
#ifndef MODELLOADER_H
#define MODELLOADER_H

#include <string>
#include <vector>

#include "cl_types.h"
#include "ObjParser.h"
#include "utils.h"

using std::string;
using std::vector;


class ModelLoader {

	public:
		ModelLoader();
		~ModelLoader();
		ObjParser* getObjParser();
		void loadModel( string filepath, string filename );

		static void getFaceNormalsOfObject( const object3D& object, vector<cl_uint4>* faceNormals, cl_int offset );
		static void getFacesOfObject( const object3D& object, vector<cl_uint4>* faces, cl_int offset );

	private:
		ObjParser* mObjParser;

};

#endif
