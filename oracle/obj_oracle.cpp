/*
 * obj_oracle.cpp -- CPU ORACLE (test infrastructure, NOT product code).
 *
 * Literal restatement of the reference's scene loading:
 *     source/ObjParser.cpp:121-348   (OBJ: v, vn, vt, f, o, usemtl; triangles only)
 *     source/MtlParser.cpp:11-236    (sibling .mtl with the custom keys rough p nu nv Rs Rd light)
 *     source/LightParser.cpp:11-128  (sibling .lights: newlight type pos rgb radius)
 * including its observable quirks: `mtllib` is ignored and the sibling
 * <name>.mtl / <name>.lights are used; .lights is read only when render.shadow_rays > 0;
 * is_any_of("//") is the one-character set {'/'} so "v/vt" is read as "v//vn"; indices are
 * unsigned, so negative OBJ indices never work; separators are single characters, runs of
 * blanks are NOT compressed; `Tr` is ignored once any `d` was seen in the file.
 *
 * boost::algorithm::trim / boost::split(.., is_any_of(" \t")) are restated below.
 *
 * PARITY STATUS: PINNED against the reference's own ObjParser / MtlParser / LightParser compiled from
 * /root/reference/source (oracle/build_ref_host.py -> oracle/_ref/libref_host.so): every parsed array,
 * material and light identical for the bundled models, the quirks fixture and generated scenes
 * (tests/test_oracle_vs_reference_host.py).
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's baseline legs may load this library.
 */
#include <ctype.h>
#include <fstream>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <string>
#include <vector>
#include <algorithm>

namespace {

using std::string;
using std::vector;

/* boost::algorithm::trim (classic locale isspace) */
void trim(string& s) {
	size_t b = 0, e = s.size();
	while (b < e && isspace((unsigned char) s[b])) b++;
	while (e > b && isspace((unsigned char) s[e - 1])) e--;
	s = s.substr(b, e - b);
}

/* boost::split with token_compress_off: every delimiter ends a (possibly empty) token. */
void split(vector<string>& parts, const string& s, const char* anyOf) {
	parts.clear();
	string cur;
	for (size_t i = 0; i < s.size(); i++) {
		if (strchr(anyOf, s[i]) != NULL) {
			parts.push_back(cur);
			cur.clear();
		}
		else {
			cur.push_back(s[i]);
		}
	}
	parts.push_back(cur);
}

struct f4 { float x, y, z, w; };

/* MtlParser.h:43-63 */
struct material_t {
	string mtlName;
	f4 Ka, Kd, Ks;
	float d, Ni, Ns;
	signed char illum, light;
	float rough, p;
	float nu, nv, Rs, Rd;
};

/* LightParser.h:20-26 */
struct light_t {
	string lightName;
	uint32_t type;
	f4 pos, rgb;
	float radius;
};

/* ObjParser.h:23-27 */
struct object3D {
	string oName;
	vector<uint32_t> facesV;
	vector<uint32_t> facesVN;
};

/* MtlParser.cpp:11-36 */
material_t getEmptyMaterial() {
	f4 white = { 1.0f, 1.0f, 1.0f, 0.0f };
	material_t mtl;
	mtl.mtlName = "";
	mtl.Ka = white;
	mtl.Kd = white;
	mtl.Ks = white;
	mtl.Ns = 100.0f;
	mtl.Ni = 1.0f;
	mtl.d = 1.0f;
	mtl.illum = 2;
	mtl.light = 0;
	mtl.rough = 1.0f;
	mtl.p = 1.0f;
	mtl.nu = 0.0f;
	mtl.nv = 0.0f;
	mtl.Rs = 0.0f;
	mtl.Rd = 1.0f;
	return mtl;
}

/* MtlParser.cpp:51-236 */
void loadMtl(const string& file, vector<material_t>& mMaterials) {
	mMaterials.clear();

	std::ifstream fileIn(file.c_str());
	material_t mtl = getEmptyMaterial();
	int numMtlFound = 0;
	bool isSetTransparency = false;

	if (!fileIn) {
		return;
	}

	while (fileIn.good()) {
		string line;
		getline(fileIn, line);
		trim(line);

		if (line.length() < 3 || line[0] == '#') {
			continue;
		}

		vector<string> parts;
		split(parts, line, " \t");

		if (parts[0] == "newmtl") {
			if (parts.size() < 2) continue;
			if (numMtlFound > 0) {
				mMaterials.push_back(mtl);
			}
			numMtlFound++;
			mtl = getEmptyMaterial();
			mtl.mtlName = parts[1];
		}
		else if (parts[0] == "d") {
			if (parts.size() < 2) continue;
			mtl.d = atof(parts[1].c_str());
			isSetTransparency = true;
		}
		else if (parts[0] == "Tr" && !isSetTransparency) {
			if (parts.size() < 2) continue;
			mtl.d = 1.0f - atof(parts[1].c_str());
		}
		else if (parts[0] == "illum") {
			if (parts.size() < 2) continue;
			mtl.illum = (signed char) atol(parts[1].c_str());
			if (mtl.illum < 0 || mtl.illum > 10) {
				mtl.illum = 2;
				continue;
			}
		}
		else if (parts[0] == "Ka") {
			if (parts.size() < 4) continue;
			mtl.Ka.x = atof(parts[1].c_str());
			mtl.Ka.y = atof(parts[2].c_str());
			mtl.Ka.z = atof(parts[3].c_str());
		}
		else if (parts[0] == "Kd") {
			if (parts.size() < 4) continue;
			mtl.Kd.x = atof(parts[1].c_str());
			mtl.Kd.y = atof(parts[2].c_str());
			mtl.Kd.z = atof(parts[3].c_str());
		}
		else if (parts[0] == "Ks") {
			if (parts.size() < 4) continue;
			mtl.Ks.x = atof(parts[1].c_str());
			mtl.Ks.y = atof(parts[2].c_str());
			mtl.Ks.z = atof(parts[3].c_str());
		}
		else if (parts[0] == "Ni") {
			if (parts.size() < 2) continue;
			mtl.Ni = atof(parts[1].c_str());
		}
		else if (parts[0] == "Ns") {
			if (parts.size() < 2) continue;
			mtl.Ns = atof(parts[1].c_str());
		}
		else if (parts[0] == "light") {
			if (parts.size() < 2) continue;
			mtl.light = (signed char) atoi(parts[1].c_str());
		}
		else if (parts[0] == "rough") {
			if (parts.size() < 2) continue;
			mtl.rough = atof(parts[1].c_str());
		}
		else if (parts[0] == "p") {
			if (parts.size() < 2) continue;
			mtl.p = atof(parts[1].c_str());
		}
		else if (parts[0] == "nu") {
			if (parts.size() < 2) continue;
			mtl.nu = atof(parts[1].c_str());
		}
		else if (parts[0] == "nv") {
			if (parts.size() < 2) continue;
			mtl.nv = atof(parts[1].c_str());
		}
		else if (parts[0] == "Rs") {
			if (parts.size() < 2) continue;
			mtl.Rs = atof(parts[1].c_str());
		}
		else if (parts[0] == "Rd") {
			if (parts.size() < 2) continue;
			mtl.Rd = atof(parts[1].c_str());
		}
	}

	if (numMtlFound > 0) {
		mMaterials.push_back(mtl);
	}
}

/* LightParser.cpp:11-22 */
light_t getEmptyLight() {
	f4 white = { 1.0f, 1.0f, 1.0f, 0.0f };
	light_t light;
	light.lightName = "";
	light.pos = white;
	light.radius = 0.0f;
	light.rgb = white;
	light.type = 0;
	return light;
}

/* LightParser.cpp:38-128.  Returns false when the reference would force render.shadow_rays = 0
 * (file opened but no `newlight` found, LightParser.cpp:119-121). */
bool loadLights(const string& file, vector<light_t>& mLights) {
	mLights.clear();

	std::ifstream fileIn(file.c_str());
	light_t light = getEmptyLight();
	int numLightsFound = 0;

	if (!fileIn) {
		return true;
	}

	while (fileIn.good()) {
		string line;
		getline(fileIn, line);
		trim(line);

		if (line.length() < 3 || line[0] == '#') {
			continue;
		}

		vector<string> parts;
		split(parts, line, " \t");

		if (parts[0] == "newlight") {
			if (parts.size() < 2) continue;
			if (numLightsFound > 0) {
				mLights.push_back(light);
			}
			numLightsFound++;
			light = getEmptyLight();
			light.lightName = parts[1];
		}
		else if (parts[0] == "type") {
			if (parts.size() < 2) continue;
			light.type = (uint32_t) atol(parts[1].c_str());
		}
		else if (parts[0] == "rgb") {
			if (parts.size() < 4) continue;
			light.rgb.x = atof(parts[1].c_str());
			light.rgb.y = atof(parts[2].c_str());
			light.rgb.z = atof(parts[3].c_str());
		}
		else if (parts[0] == "pos") {
			if (parts.size() < 4) continue;
			light.pos.x = atof(parts[1].c_str());
			light.pos.y = atof(parts[2].c_str());
			light.pos.z = atof(parts[3].c_str());
		}
		else if (parts[0] == "radius") {
			if (parts.size() < 2) continue;
			light.radius = atof(parts[1].c_str());
		}
	}

	if (numLightsFound > 0) {
		mLights.push_back(light);
		return true;
	}
	return false;
}

struct ObjParser {
	vector<object3D> mObjects;
	vector<int32_t> mFacesMtl;
	vector<uint32_t> mFacesV, mFacesVN, mFacesVT;
	vector<float> mNormals, mTextures, mVertices;
	vector<material_t> mMaterials;
	vector<light_t> mLights;
	bool shadowRaysForcedOff;

	/* ObjParser.cpp:262-306 */
	static void parseFace(const string& line, vector<uint32_t>* facesV, vector<uint32_t>* facesVN, vector<uint32_t>* facesVT) {
		vector<string> parts;
		split(parts, line, " \t");

		for (size_t i = 1; i < parts.size(); i++) {
			uint32_t a;
			vector<string> e0, e1;
			split(e0, parts[i], "/");
			split(e1, parts[i], "//");   /* is_any_of("//") == {'/'} */

			if (e1.size() == 2) {
				a = (uint32_t) atol(e1[0].c_str());
				facesV->push_back(a - 1);
				a = (uint32_t) atol(e1[1].c_str());
				facesVN->push_back(a - 1);
			}
			else {
				a = (uint32_t) atol(e0[0].c_str());
				facesV->push_back(a - 1);

				if (e0.size() >= 2) {
					a = (uint32_t) atol(e0[1].c_str());
					facesVT->push_back(a - 1);
				}
				if (e0.size() >= 3) {
					a = (uint32_t) atol(e0[2].c_str());
					facesVN->push_back(a - 1);
				}
			}
		}
	}

	/* ObjParser.cpp:314-335 */
	static void parseVec3(const string& line, vector<float>* out) {
		vector<string> parts;
		split(parts, line, " \t");
		for (int k = 1; k <= 3; k++) {
			out->push_back((size_t) k < parts.size() ? (float) atof(parts[k].c_str()) : 0.0f);
		}
	}

	/* ObjParser.cpp:343-348 */
	static void parseVertexTexture(const string& line, vector<float>* texCoords) {
		vector<string> parts;
		split(parts, line, " \t");
		float weight = (parts.size() >= 4) ? atof(parts[3].c_str()) : 0.0f;
		texCoords->push_back(parts.size() > 1 ? (float) atof(parts[1].c_str()) : 0.0f);
		texCoords->push_back(parts.size() > 2 ? (float) atof(parts[2].c_str()) : 0.0f);
		texCoords->push_back(weight);
	}

	static string sibling(string file, const char* ext) {
		size_t extensionIndex = file.rfind(".obj");
		if (extensionIndex == string::npos) return file + ext;
		file.replace(extensionIndex, 4, ext);
		return file;
	}

	/* ObjParser.cpp:121-221 */
	void load(const string& file, int shadowRays) {
		shadowRaysForcedOff = false;
		std::ifstream fileIn(file.c_str());

		if (shadowRays > 0) {
			if (!loadLights(sibling(file, ".lights"), mLights)) shadowRaysForcedOff = true;
		}

		loadMtl(sibling(file, ".mtl"), mMaterials);
		vector<string> materialNames;
		int32_t currentMtl = -1;

		for (size_t i = 0; i < mMaterials.size(); i++) {
			materialNames.push_back(mMaterials[i].mtlName);
		}

		while (fileIn.good()) {
			string line;
			getline(fileIn, line);
			trim(line);

			if (line[0] == '#') {
				continue;
			}

			if (line[0] == 'o') {
				object3D o;
				vector<string> parts;
				split(parts, line, " \t");
				o.oName = parts.size() > 1 ? parts[1] : "";
				mObjects.push_back(o);
			}
			else if (line[0] == 'v') {
				if (line[1] == ' ') {
					parseVec3(line, &mVertices);
				}
				else if (line[1] == 'n' && line[2] == ' ') {
					parseVec3(line, &mNormals);
				}
				else if (line[1] == 't' && line[2] == ' ') {
					parseVertexTexture(line, &mTextures);
				}
			}
			else if (line[0] == 'f') {
				if (line[1] == ' ') {
					vector<uint32_t> lineFacesV, lineFacesVN, lineFacesVT;
					parseFace(line, &lineFacesV, &lineFacesVN, &lineFacesVT);

					mFacesV.insert(mFacesV.end(), lineFacesV.begin(), lineFacesV.end());
					mFacesVN.insert(mFacesVN.end(), lineFacesVN.begin(), lineFacesVN.end());
					mFacesVT.insert(mFacesVT.end(), lineFacesVT.begin(), lineFacesVT.end());

					mFacesMtl.push_back(currentMtl);

					if (mObjects.size() > 0) {
						object3D* op = &(mObjects[mObjects.size() - 1]);
						op->facesV.insert(op->facesV.end(), lineFacesV.begin(), lineFacesV.end());
						op->facesVN.insert(op->facesVN.end(), lineFacesVN.begin(), lineFacesVN.end());
					}
				}
			}
			else if (line.find("usemtl") != string::npos) {
				vector<string> parts;
				split(parts, line, " \t");
				const string name = parts.size() > 1 ? parts[1] : "";
				vector<string>::iterator it = std::find(materialNames.begin(), materialNames.end(), name);
				currentMtl = (it != materialNames.end()) ? (int32_t) (it - materialNames.begin()) : -1;
			}
		}
	}
};

} /* namespace */

extern "C" {

void* oracle_obj_load(const char* file, int32_t shadowRays) {
	ObjParser* p = new ObjParser();
	p->load(file, shadowRays);
	return p;
}

void oracle_obj_free(void* h) { delete (ObjParser*) h; }

/* what: 0 vertices(f32) 1 normals(f32) 2 facesV(u32) 3 facesVN(u32) 4 facesMtl(i32)
 *       5 per-object face counts (u32) 6 per-object facesV, concatenated (u32)
 *       7 per-object facesVN, concatenated (u32) 8 per-object normal-face counts (u32)
 *       9 materials (f32 x 24 each: Ka4 Kd4 Ks4 d Ni Ns illum light rough p nu nv Rs Rd pad)
 *      10 lights (f32 x 10 each: type pos4 rgb4 radius)  11 shadow-rays-forced-off flag (count)
 *      12 facesVT(u32) 13 texture coords (f32)
 * Returns the element count; copies when dst != NULL. */
int64_t oracle_obj_get(void* h, int32_t what, void* dst) {
	ObjParser* p = (ObjParser*) h;
#define COPY_VEC(v) do { if (dst) memcpy(dst, (v).data(), (v).size() * sizeof((v)[0])); return (int64_t) (v).size(); } while (0)
	switch (what) {
		case 0: COPY_VEC(p->mVertices);
		case 1: COPY_VEC(p->mNormals);
		case 2: COPY_VEC(p->mFacesV);
		case 3: COPY_VEC(p->mFacesVN);
		case 4: COPY_VEC(p->mFacesMtl);
		case 12: COPY_VEC(p->mFacesVT);
		case 13: COPY_VEC(p->mTextures);
		case 5: case 8: {
			if (dst) for (size_t i = 0; i < p->mObjects.size(); i++)
				((uint32_t*) dst)[i] = (uint32_t) ((what == 5 ? p->mObjects[i].facesV.size() : p->mObjects[i].facesVN.size()) / 3);
			return (int64_t) p->mObjects.size();
		}
		case 6: case 7: {
			int64_t n = 0;
			for (size_t i = 0; i < p->mObjects.size(); i++) {
				const std::vector<uint32_t>& v = (what == 6) ? p->mObjects[i].facesV : p->mObjects[i].facesVN;
				if (dst) memcpy((uint32_t*) dst + n, v.data(), v.size() * 4);
				n += (int64_t) v.size();
			}
			return n;
		}
		case 9: {
			if (dst) for (size_t i = 0; i < p->mMaterials.size(); i++) {
				const material_t& m = p->mMaterials[i];
				float* o = (float*) dst + i * 24;
				o[0] = m.Ka.x; o[1] = m.Ka.y; o[2] = m.Ka.z; o[3] = m.Ka.w;
				o[4] = m.Kd.x; o[5] = m.Kd.y; o[6] = m.Kd.z; o[7] = m.Kd.w;
				o[8] = m.Ks.x; o[9] = m.Ks.y; o[10] = m.Ks.z; o[11] = m.Ks.w;
				o[12] = m.d; o[13] = m.Ni; o[14] = m.Ns; o[15] = (float) m.illum; o[16] = (float) m.light;
				o[17] = m.rough; o[18] = m.p; o[19] = m.nu; o[20] = m.nv; o[21] = m.Rs; o[22] = m.Rd; o[23] = 0.0f;
			}
			return (int64_t) p->mMaterials.size();
		}
		case 10: {
			if (dst) for (size_t i = 0; i < p->mLights.size(); i++) {
				const light_t& l = p->mLights[i];
				float* o = (float*) dst + i * 10;
				o[0] = (float) l.type;
				o[1] = l.pos.x; o[2] = l.pos.y; o[3] = l.pos.z; o[4] = l.pos.w;
				o[5] = l.rgb.x; o[6] = l.rgb.y; o[7] = l.rgb.z; o[8] = l.rgb.w;
				o[9] = l.radius;
			}
			return (int64_t) p->mLights.size();
		}
		case 11: return p->shadowRaysForcedOff ? 1 : 0;
	}
#undef COPY_VEC
	return -1;
}

/* kind: 0 object, 1 material, 2 light */
const char* oracle_obj_name(void* h, int32_t kind, int32_t idx) {
	ObjParser* p = (ObjParser*) h;
	if (kind == 0 && idx >= 0 && (size_t) idx < p->mObjects.size()) return p->mObjects[idx].oName.c_str();
	if (kind == 1 && idx >= 0 && (size_t) idx < p->mMaterials.size()) return p->mMaterials[idx].mtlName.c_str();
	if (kind == 2 && idx >= 0 && (size_t) idx < p->mLights.size()) return p->mLights[idx].lightName.c_str();
	return "";
}

} /* extern "C" */
