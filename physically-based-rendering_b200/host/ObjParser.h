/*
 * ObjParser -- reference: source/ObjParser.{h,cpp}.  Wavefront OBJ reader with the reference's
 * class surface (load + the get* accessors, ObjParser.h:32-44) and observable behaviour:
 *   - `mtllib` is ignored; the sibling <name>.mtl is loaded, and <name>.lights only when
 *     render.shadow_rays > 0 (ObjParser.cpp:133-137, 228-245)
 *   - triangles only; indices are unsigned, so negative (relative) OBJ indices do not work (:264-273)
 *   - is_any_of("//") is the single-character set {'/'}: "v//vn" and "v/vt/vn" read as intended,
 *     but "v/vt" is taken for "v//vn" (:266-280)
 *   - separators are single blanks/tabs; runs are not compressed (:261, 311)
 *   - faces before the first `o` belong to no object (:192-198)
 * The whole file is read into one buffer and tokenised in place (no per-line allocations), which is
 * what makes the 1M / 10M triangle configurations loadable in seconds.
 */
#ifndef OBJPARSER_H
#define OBJPARSER_H

#include <string>
#include <vector>

#include "cl_types.h"
#include "Logger.h"
#include "LightParser.h"
#include "MtlParser.h"

using std::string;
using std::vector;

/* one `o` block: its triangles as flat index triples (vertex indices, normal indices) */
struct object3D {
	string oName;
	vector<cl_uint> facesV, facesVN;
};

class ObjParser {
	public:
		ObjParser();
		~ObjParser();
		void load( string filepath, string filename );

		/* --- the reference's accessors: copies, as upstream */
		vector<cl_float> getVertices();
		vector<cl_float> getNormals();
		vector<cl_float> getTextureCoordinates();
		vector<cl_uint> getFacesV();
		vector<cl_uint> getFacesVN();
		vector<cl_uint> getFacesVT();
		vector<cl_int> getFacesMtl();
		vector<object3D> getObjects();
		vector<material_t> getMaterials();
		vector<light_t> getLights();

		/* --- additive: the same arrays by reference (large scenes) */
		const vector<cl_float>& vertices() const { return mVertices; }
		const vector<cl_float>& normals() const { return mNormals; }
		const vector<cl_uint>& facesV() const { return mFacesV; }
		const vector<cl_uint>& facesVN() const { return mFacesVN; }
		const vector<cl_int>& facesMtl() const { return mFacesMtl; }
		const vector<object3D>& objects() const { return mObjects; }

		/* --- additive: install an already parsed scene (synthetic generators, tests) */
		void setScene(
			const vector<cl_float>& vertices, const vector<cl_float>& normals,
			const vector<cl_uint>& facesV, const vector<cl_uint>& facesVN, const vector<cl_int>& facesMtl,
			const vector<object3D>& objects, const vector<material_t>& materials, const vector<light_t>& lights
		);

	protected:
		void loadMtl( string file );
		void loadLights( string file );

	private:
		vector<cl_float> mVertices, mNormals, mTextures;
		vector<cl_uint> mFacesV, mFacesVN, mFacesVT;
		vector<cl_int> mFacesMtl;
		vector<object3D> mObjects;
		MtlParser* mMtlParser;
		LightParser* mLightParser;
};

#endif
