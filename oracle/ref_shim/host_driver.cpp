/*
 * host_driver.cpp -- C entry points around the REFERENCE's own host classes (ModelLoader / ObjParser /
 * MtlParser / LightParser / BVH / MathHelp / Cfg / Logger), compiled from the sources where they lie under
 * /root/reference/source by oracle/build_ref_host.py into oracle/_ref/libref_host.so.  TEST INFRASTRUCTURE:
 * the yardstick for oracle/obj_oracle.cpp, oracle/bvh_oracle.cpp and, through them, for the product's host
 * library.
 *
 * The sequence is the reference's (qt/GLWidget.cpp:339-355): ModelLoader::loadModel, then
 * BVH( objects, vertices, normals ).  PathTracer.cpp itself cannot be compiled here (Qt, OpenCL), so the last
 * step -- PathTracer::initOpenCLBuffers_BVH, which turns the node list into bvhNode_cl[] and the leaf-ordered
 * face arrays (PathTracer.cpp:238-347) -- is restated below on the reference's BVHNode objects.
 *
 * `what` codes of refhost_obj_get are those of oracle_obj_get (oracle/obj_oracle.cpp).
 */
#include <stdint.h>
#include <string.h>
#include <string>
#include <vector>

#include "Cfg.h"
#include "ModelLoader.h"
#include "accelstructures/BVH.h"
#include "Camera.h"
#include "PathTracer.h"

/* qt/GLWidget.h stand-in (ref_shim/host/cl_types.h); GLWidget.cpp:78-82 */
void GLWidget::cameraUpdate() {
	if (mPathTracer) mPathTracer->resetSampleCount();
}

namespace {

struct Loaded {
	ModelLoader* ml;
	BVH* bvh;
	std::vector<float> nodes;           /* 8 floats per emitted node */
	std::vector<uint32_t> facesV, facesN;   /* 4 per face, leaf order */
	int64_t info[6];                    /* allNodes, leaves, depth, skipped, emitted, faces */
	std::vector<object3D> objects;
	std::vector<material_t> materials;
	std::vector<light_t> lights;
	Loaded() : ml(NULL), bvh(NULL) { memset(info, 0, sizeof(info)); }
};

/* PathTracer::initOpenCLBuffers_BVH (PathTracer.cpp:238-347) */
void flatten(Loaded* L) {
	std::vector<BVHNode*> bvhNodes = L->bvh->getNodes();
	ObjParser* op = L->ml->getObjParser();
	std::vector<cl_uint> faces = op->getFacesV();
	std::vector<cl_uint> facesVN = op->getFacesVN();
	std::vector<cl_int> facesMtl = op->getFacesMtl();
	bool skipNext = false;
	int64_t skipped = 0;
	size_t numFaces = 0;

	for (cl_uint i = 0; i < bvhNodes.size(); i++) {
		BVHNode* node = bvhNodes[i];
		if (skipNext) {
			skipNext = node->skipNextLeft;
			skipped++;
			continue;
		}
		float sn[8] = { node->bbMin[0], node->bbMin[1], node->bbMin[2], 0.0f, node->bbMax[0], node->bbMax[1], node->bbMax[2], 0.0f };
		const cl_uint fvecLen = (cl_uint) node->faces.size();
		sn[3] = (fvecLen > 0) ? (cl_float) numFaces + 0 : -1.0f;
		sn[7] = (fvecLen > 1) ? (cl_float) numFaces + 1 : -1.0f;
		if (fvecLen == 0 && node->skipNextLeft) skipNext = true;

		if (node->parent != NULL && fvecLen == 0) {
			const bool isLeftNode = (node->parent->leftChild == node);
			if (!isLeftNode) {
				if (node->parent->parent != NULL) {
					BVHNode* p = node->parent;
					while (p->parent->rightChild == p) {
						p = p->parent;
						if (p->parent == NULL) break;
					}
					if (p->parent != NULL) {
						sn[7] = p->parent->rightChild->id - p->parent->rightChild->numSkipsToHere;
					}
				}
			}
			else {
				sn[7] = node->parent->rightChild->id - node->parent->rightChild->numSkipsToHere;
			}
		}
		L->nodes.insert(L->nodes.end(), sn, sn + 8);

		for (cl_uint j = 0; j < fvecLen; j++) {
			const Tri& tri = node->faces[j];
			L->facesV.push_back(faces[tri.face.w * 3]);
			L->facesV.push_back(faces[tri.face.w * 3 + 1]);
			L->facesV.push_back(faces[tri.face.w * 3 + 2]);
			L->facesV.push_back((uint32_t) facesMtl[tri.face.w]);
			L->facesN.push_back(facesVN[tri.normals.w * 3]);
			L->facesN.push_back(facesVN[tri.normals.w * 3 + 1]);
			L->facesN.push_back(facesVN[tri.normals.w * 3 + 2]);
			L->facesN.push_back(0u);
			numFaces++;
		}
	}
	L->info[0] = (int64_t) bvhNodes.size();
	L->info[1] = (int64_t) L->bvh->getLeafNodes().size();
	L->info[2] = (int64_t) L->bvh->getDepth();
	L->info[3] = skipped;
	L->info[4] = (int64_t) (L->nodes.size() / 8);
	L->info[5] = (int64_t) numFaces;
}

template <typename T>
int64_t copyOut(const std::vector<T>& v, void* dst) {
	if (dst && !v.empty()) memcpy(dst, &v[0], v.size() * sizeof(T));
	return (int64_t) v.size();
}

} /* namespace */

extern "C" {

/* configJson: a config.json as the reference reads it at start-up (main.cpp -> Cfg::loadConfigFile). */
void* refhost_load(const char* dir, const char* file, const char* configJson, int buildBvh) {
	Cfg::get().loadConfigFile(configJson);
	Loaded* L = new Loaded();
	L->ml = new ModelLoader();
	L->ml->loadModel(std::string(dir), std::string(file));
	ObjParser* op = L->ml->getObjParser();
	L->objects = op->getObjects();
	L->materials = op->getMaterials();
	L->lights = op->getLights();
	if (buildBvh) {
		L->bvh = new BVH(op->getObjects(), op->getVertices(), op->getNormals());
		flatten(L);
	}
	return L;
}

int64_t refhost_obj_get(void* h, int32_t what, void* dst) {
	Loaded* L = (Loaded*) h;
	ObjParser* op = L->ml->getObjParser();
	switch (what) {
		case 0: return copyOut(op->getVertices(), dst);
		case 1: return copyOut(op->getNormals(), dst);
		case 2: return copyOut(op->getFacesV(), dst);
		case 3: return copyOut(op->getFacesVN(), dst);
		case 4: return copyOut(op->getFacesMtl(), dst);
		case 5: case 8: {
			if (dst) for (size_t i = 0; i < L->objects.size(); i++)
				((uint32_t*) dst)[i] = (uint32_t) ((what == 5 ? L->objects[i].facesV.size() : L->objects[i].facesVN.size()) / 3);
			return (int64_t) L->objects.size();
		}
		case 6: case 7: {
			int64_t n = 0;
			for (size_t i = 0; i < L->objects.size(); i++) {
				const std::vector<cl_uint>& v = (what == 6) ? L->objects[i].facesV : L->objects[i].facesVN;
				if (dst && !v.empty()) memcpy((uint32_t*) dst + n, &v[0], v.size() * 4);
				n += (int64_t) v.size();
			}
			return n;
		}
		case 9: {
			if (dst) for (size_t i = 0; i < L->materials.size(); i++) {
				const material_t& m = L->materials[i];
				float* o = (float*) dst + i * 24;
				o[0] = m.Ka.x; o[1] = m.Ka.y; o[2] = m.Ka.z; o[3] = m.Ka.w;
				o[4] = m.Kd.x; o[5] = m.Kd.y; o[6] = m.Kd.z; o[7] = m.Kd.w;
				o[8] = m.Ks.x; o[9] = m.Ks.y; o[10] = m.Ks.z; o[11] = m.Ks.w;
				o[12] = m.d; o[13] = m.Ni; o[14] = m.Ns; o[15] = (float) m.illum; o[16] = (float) m.light;
				o[17] = m.rough; o[18] = m.p; o[19] = m.nu; o[20] = m.nv; o[21] = m.Rs; o[22] = m.Rd; o[23] = 0.0f;
			}
			return (int64_t) L->materials.size();
		}
		case 10: {
			if (dst) for (size_t i = 0; i < L->lights.size(); i++) {
				const light_t& l = L->lights[i];
				float* o = (float*) dst + i * 10;
				o[0] = (float) l.type;
				o[1] = l.pos.x; o[2] = l.pos.y; o[3] = l.pos.z; o[4] = l.pos.w;
				o[5] = l.rgb.x; o[6] = l.rgb.y; o[7] = l.rgb.z; o[8] = l.rgb.w;
				o[9] = l.radius;
			}
			return (int64_t) L->lights.size();
		}
		case 12: return copyOut(op->getFacesVT(), dst);
		case 13: return copyOut(op->getTextureCoordinates(), dst);
	}
	return -1;
}

/* kind: 0 object, 1 material, 2 light */
const char* refhost_obj_name(void* h, int32_t kind, int32_t idx) {
	Loaded* L = (Loaded*) h;
	if (kind == 0 && idx >= 0 && (size_t) idx < L->objects.size()) return L->objects[idx].oName.c_str();
	if (kind == 1 && idx >= 0 && (size_t) idx < L->materials.size()) return L->materials[idx].mtlName.c_str();
	if (kind == 2 && idx >= 0 && (size_t) idx < L->lights.size()) return L->lights[idx].lightName.c_str();
	return "";
}

void refhost_bvh_info(void* h, int64_t info[6]) { memcpy(info, ((Loaded*) h)->info, 6 * sizeof(int64_t)); }

void refhost_bvh_get(void* h, float* nodes, uint32_t* facesV, uint32_t* facesN) {
	Loaded* L = (Loaded*) h;
	copyOut(L->nodes, nodes);
	copyOut(L->facesV, facesV);
	copyOut(L->facesN, facesN);
}

/* ---- the reference's renderer core: GLWidget's constructor + loadModel + paintGL sequence
 * (qt/GLWidget.cpp:28-32, 339-387, 504-517) without the widget ---- */

struct Renderer {
	GLWidget widget;
	PathTracer* pt;
	Camera* camera;
	std::vector<cl_float> image, debug;
};

void* refhost_renderer_create(const char* configJson, const char* dir, const char* file) {
	Cfg::get().loadConfigFile(configJson);
	boost::posix_time::clockOverrideUs() = 0;          /* PathTracer's constructor notes the start time */
	Renderer* r = new Renderer();
	r->pt = new PathTracer(&r->widget);
	r->camera = new Camera(&r->widget);
	r->widget.mPathTracer = r->pt;
	r->pt->setCamera(r->camera);
	/* GLWidget::resizeGL -> setWidthAndHeight( window size ) */
	r->pt->setWidthAndHeight(Cfg::get().value<cl_uint>(Cfg::WINDOW_WIDTH), Cfg::get().value<cl_uint>(Cfg::WINDOW_HEIGHT));

	ModelLoader* ml = new ModelLoader();
	ml->loadModel(std::string(dir), std::string(file));
	ObjParser* op = ml->getObjParser();
	std::vector<cl_uint> faces = op->getFacesV();
	std::vector<cl_float> normals = op->getNormals();
	std::vector<cl_float> vertices = op->getVertices();
	AccelStructure* accel = new BVH(op->getObjects(), vertices, normals);
	r->pt->initOpenCLBuffers(vertices, faces, normals, ml, accel);
	delete ml;
	delete accel;
	return r;
}

/* One PathTracer::generateImage() at time `ms` after start: the seed is ms * 0.001f (PathTracer.cpp:78-82). */
int refhost_renderer_generate(void* h, long long ms, float* image, float* debug) {
	Renderer* r = (Renderer*) h;
	boost::posix_time::clockOverrideUs() = ms * 1000;
	/* generateImage reads the debug image straight into the caller's storage (GLWidget keeps it sized) */
	r->debug.resize((size_t) Cfg::get().value<cl_uint>(Cfg::WINDOW_WIDTH) * Cfg::get().value<cl_uint>(Cfg::WINDOW_HEIGHT) * 4);
	r->image = r->pt->generateImage(&r->debug);
	if (image && !r->image.empty()) memcpy(image, &r->image[0], r->image.size() * 4);
	if (debug && !r->debug.empty()) memcpy(debug, &r->debug[0], r->debug.size() * 4);
	return (int) r->image.size();
}

/* what: 0 setFocus(a, b); 1 resetSampleCount; 2 updateCameraRot(a, b); 3.. cameraMove{Forward,Backward,Left,Right,Up,Down};
 * 9 cameraReset; 10 setSpeed(a / 1000) */
void refhost_renderer_command(void* h, int what, int a, int b) {
	Renderer* r = (Renderer*) h;
	switch (what) {
		case 0: r->pt->setFocus(a, b); break;
		case 1: r->pt->resetSampleCount(); break;
		case 2: r->camera->updateCameraRot(a, b); break;
		case 3: r->camera->cameraMoveForward(); break;
		case 4: r->camera->cameraMoveBackward(); break;
		case 5: r->camera->cameraMoveLeft(); break;
		case 6: r->camera->cameraMoveRight(); break;
		case 7: r->camera->cameraMoveUp(); break;
		case 8: r->camera->cameraMoveDown(); break;
		case 9: r->camera->cameraReset(); break;
		case 10: r->camera->setSpeed(a / 1000.0f); break;
	}
}

void refhost_renderer_free(void* h) {
	Renderer* r = (Renderer*) h;
	delete r->pt;
	delete r->camera;
	delete r;
}

void refhost_free(void* h) {
	Loaded* L = (Loaded*) h;
	delete L->bvh;
	delete L->ml;
	delete L;
}

} /* extern "C" */
