/*
 * wide_walk_model.cpp -- TEST INFRASTRUCTURE: a CPU model of the ordered wide-BVH walk (csrc/pt_wide.cuh) next to
 * a literal copy of the reference-order walk (traverse, source/opencl/pt_bvh.cl:82-123), built as a shared library
 * by tests/test_wide_walk.py.
 *
 * What it pins, without a GPU:
 *   - the product's wide-BVH builder (csrc/wide_bvh.h, the very header pbr_capi.cu includes) on real trees;
 *   - the exactness argument of the ordered walk: for every ray, (t bits, face, leaf) of the ordered walk -- nearest
 *     child first, pruning with a margin, "ambiguous" rays re-walked in reference order -- equal those of the
 *     reference-order walk;
 *   - statistics: wide-node visits, triangle tests, stack depth, how many rays fall back.
 *
 * Arithmetic: include/pbr_pinned_math.h, -ffp-contract=off (the same contract as oracle/ and csrc/).
 */
#include <math.h>
#include <stdint.h>
#include <string.h>

#include <algorithm>
#include <vector>

#include "../include/pbr_pinned_math.h"
#include "../physically-based-rendering_b200/csrc/wide_bvh.h"

namespace {

struct V3 { float x, y, z; };
inline V3 sub(V3 a, V3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
inline V3 cross(V3 a, V3 b) { return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x}; }
inline float dot(V3 a, V3 b) { return (a.x * b.x + a.y * b.y) + a.z * b.z; }

const float EPS5 = 0.00001f;

struct Scene {
	const float* nodes; int numNodes;
	const uint32_t* facesV; int numFaces;
	const float* vertices;          /* float4 per vertex */
};

/* intersectBox (pt_intersect.cl:11-25) */
bool intersectBox(V3 o, V3 inv, const float* lo, const float* hi, float& tNear, float& tFar) {
	const float t1x = (lo[0] - o.x) * inv.x, t1y = (lo[1] - o.y) * inv.y, t1z = (lo[2] - o.z) * inv.z;
	const float t2x = (hi[0] - o.x) * inv.x, t2y = (hi[1] - o.y) * inv.y, t2z = (hi[2] - o.z) * inv.z;
	const float tMinX = fminf(t1x, t2x), tMinY = fminf(t1y, t2y), tMinZ = fminf(t1z, t2z);
	const float tMaxX = fmaxf(t1x, t2x), tMaxY = fmaxf(t1y, t2y), tMaxZ = fmaxf(t1z, t2z);
	tNear = fmaxf(fmaxf(tMinX, tMinY), tMinZ);
	tFar = fminf(fminf(tMaxX, tMaxY), fminf(tMaxZ, INFINITY));
	return tNear <= tFar;
}

/* flatTriAndRayIntersect (pt_intersect.cl:92-129): t of the face against a ray whose current ray.t is `rt`;
 * INFINITY when rejected */
float faceT(const Scene& S, int face, V3 o, V3 d, float tNear, float rt) {
	const uint32_t* fv = S.facesV + 4 * (size_t) face;
	const float* pa = S.vertices + 4 * (size_t) fv[0];
	const float* pb = S.vertices + 4 * (size_t) fv[1];
	const float* pc = S.vertices + 4 * (size_t) fv[2];
	const V3 a = {pa[0], pa[1], pa[2]}, b = {pb[0], pb[1], pb[2]}, c = {pc[0], pc[1], pc[2]};
	const float f = fmaxf(0.0f, tNear - 0.001f);
	const V3 co = {fmaf(d.x, f, o.x), fmaf(d.y, f, o.y), fmaf(d.z, f, o.z)};
	const V3 e1 = sub(b, a), e2 = sub(c, a), tv = sub(co, a);
	const V3 pv = cross(d, e2), qv = cross(tv, e1);
	const float invDet = pm::rcp(dot(e1, pv));
	float t = dot(e2, qv) * invDet;
	if (t >= rt || t < EPS5) return INFINITY;
	const float u = dot(tv, pv) * invDet, v = dot(d, qv) * invDet;
	if (u + v > 1.0f || fminf(u, v) < 0.0f) return INFINITY;
	return t + f;
}

struct Hit { float t; int face; int leaf; uint32_t nodes, tris; };

/* traverse (pt_bvh.cl:82-123), no lights */
Hit strictWalk(const Scene& S, V3 o, V3 d, float rt0) {
	Hit h = {rt0, 0, -1, 0, 0};
	const V3 inv = {pm::rcp(d.x), pm::rcp(d.y), pm::rcp(d.z)};
	int index = 1;
	if (S.numNodes < 2) return h;
	do {
		h.nodes++;
		const float* lo = S.nodes + 8 * (size_t) index;
		const float* hi = lo + 4;
		const int cur = index;
		index = (lo[3] <= -1.0f) ? (int) hi[3] : cur + 1;
		float tNear, tFar;
		const bool hit = intersectBox(o, inv, lo, hi, tNear, tFar) && tFar > EPS5 && h.t > tNear;
		if (!hit) continue;
		index = cur + 1;
		if (lo[3] >= 0.0f) {
			const int f0 = (int) lo[3];
			float t = faceT(S, f0, o, d, tNear, h.t);
			h.tris++;
			if (h.t > t) { h.t = t; h.face = f0; h.leaf = cur; }
			if (hi[3] != -1.0f) {
				const int f1 = (int) hi[3];
				t = faceT(S, f1, o, d, tNear, h.t);
				h.tris++;
				if (h.t > t) { h.t = t; h.face = f1; h.leaf = cur; }
			}
		}
	} while (index > 0 && index < S.numNodes);
	return h;
}

struct FastStats { uint64_t wideVisits = 0, triTests = 0, fallbacks = 0, overflow = 0, insaneWinners = 0; int maxStack = 0; };

inline float pruneLimit(float rt) { return rt + (0.0022f + 2e-6f * rt); }

/* The ordered walk as the device runs it (csrc/pt_wide.cuh), one ray at a time: a work item is an inner node or a
 * leaf; visiting an inner node tests its four child boxes, keeps the nearest child that is still in reach as the next
 * item and pushes the others (leaves included) with their tNear; a leaf item tests its faces; an exhausted item pops
 * the stack, skipping entries that have fallen out of reach. */
Hit fastWalk(const Scene& S, const wbvh::Result& W, V3 o, V3 d, float rt0, int stackCap, FastStats& st) {
	const V3 inv = {pm::rcp(d.x), pm::rcp(d.y), pm::rcp(d.z)};
	float rt = rt0, bestTn = -INFINITY, t2 = rt0;
	int face = 0, leaf = -1;
	uint32_t nn = 0, nt = 0;
	std::vector<std::pair<int, float>> stack;
	bool overflow = false;
	int item = 0;                       /* >= 0 inner node, < 0 leaf ref */
	float itemTn = 0.0f;
	bool have = !W.nodes.empty();
	float lim = pruneLimit(rt);
	auto pop = [&]() {
		have = false;
		while (!stack.empty()) {
			const std::pair<int, float> e = stack.back();
			stack.pop_back();
			if (e.second <= lim) { item = e.first; itemTn = e.second; have = true; break; }
		}
	};
	while (have && !overflow) {
		if (item >= 0) {
			nn++;
			const wbvh::Node& N = W.nodes[(size_t) item];
			/* the children still in reach, sorted by entry distance (ties: child order): the nearest is visited next,
			 * the others are pushed farthest first */
			int refs[4], n = 0;
			float tns[4];
			for (int j = 0; j < 4; j++) {
				const wbvh::Child& C = N.c[j];
				float tNear, tFar;
				const float lo[3] = {C.lo(0), C.lo(1), C.lo(2)}, hi[3] = {C.hi(0), C.hi(1), C.hi(2)};
				const bool hit = intersectBox(o, inv, lo, hi, tNear, tFar) && tFar > EPS5 && C.ref != wbvh::REF_EMPTY;
				if (!(hit && tNear <= lim && tNear < INFINITY)) continue;
				int k = n++;
				while (k > 0 && tns[k - 1] > tNear) { tns[k] = tns[k - 1]; refs[k] = refs[k - 1]; k--; }
				tns[k] = tNear; refs[k] = C.ref;
			}
			int curRef = wbvh::REF_EMPTY;
			float curTn = INFINITY;
			if (n > 0) { curRef = refs[0]; curTn = tns[0]; }
			for (int k = n - 1; k >= 1; k--) {
				if ((int) stack.size() >= stackCap) { overflow = true; break; }
				stack.push_back({refs[k], tns[k]});
			}
			st.maxStack = std::max(st.maxStack, (int) stack.size());
			if (curRef != wbvh::REF_EMPTY) { item = curRef; itemTn = curTn; }
			else pop();
		}
		else {
			const int f0 = (int) ((uint32_t) item & wbvh::REF_FACE_MASK);
			float tl = faceT(S, f0, o, d, itemTn, INFINITY);
			int fl = f0;
			nt++;
			if ((uint32_t) item & wbvh::REF_TWO) {
				const float t1 = faceT(S, f0 + 1, o, d, itemTn, INFINITY);
				nt++;
				if (t1 < tl) { tl = t1; fl = f0 + 1; }
			}
			if (tl < INFINITY) {
				if (tl < rt || (tl == rt && leaf >= 0 && fl < face)) {
					t2 = fminf(t2, rt);
					rt = tl; face = fl; leaf = W.faceLeaf[(size_t) fl]; bestTn = itemTn;
				}
				else t2 = fminf(t2, tl);
				lim = pruneLimit(rt);
			}
			pop();
		}
	}
	st.wideVisits += nn; st.triTests += nt;
	const bool insane = rt < bestTn;
	if (insane) st.insaneWinners++;
	if (overflow) st.overflow++;
	if (overflow || (insane && t2 <= bestTn)) {
		st.fallbacks++;
		Hit h = strictWalk(S, o, d, rt0);
		h.nodes += nn; h.tris += nt;
		return h;
	}
	return {rt, face, leaf, nn, nt};
}


/* EXPERIMENT (scripts/quant_nodes_stats.py; DESIGN.md §6 "where the next factor is"): the ordered walk on child boxes
 * quantised to `bits` bits per plane relative to the union of the node's children (64-byte nodes), conservative (lo
 * rounded down, hi up; power-of-two grid steps when pow2).  The exact leaf box would move next to the leaf's faces: a
 * leaf item first repeats the box test on the exact bits (its tNear feeds the triangle test, pt_intersect.cl:96), so
 * the candidates -- and with them the result -- are those of the exact walk; only the visits change. */
struct QuantStats { uint64_t inner = 0, leaves = 0, leavesRejected = 0, tris = 0, fallbacks = 0, mismatches = 0; };

void quantBox(const wbvh::Node& N, int j, int bits, bool pow2, float* lo, float* hi) {
	const float steps = (float) ((1 << bits) - 1);
	for (int a = 0; a < 3; a++) {
		float flo = INFINITY, fhi = -INFINITY;
		for (int k = 0; k < 4; k++) {
			if (N.c[k].ref == wbvh::REF_EMPTY) continue;
			flo = fminf(flo, N.c[k].lo(a)); fhi = fmaxf(fhi, N.c[k].hi(a));
		}
		double step = ((double) fhi - (double) flo) / steps;
		if (pow2 && step > 0.0) step = exp2(ceil(log2(step)));
		if (!(step > 0.0)) { lo[a] = N.c[j].lo(a); hi[a] = N.c[j].hi(a); continue; }
		double ql = floor(((double) N.c[j].lo(a) - flo) / step), qh = ceil(((double) N.c[j].hi(a) - flo) / step);
		float l = (float) (flo + ql * step), h = (float) (flo + qh * step);
		while (l > N.c[j].lo(a)) l = nextafterf(l, -INFINITY);
		while (h < N.c[j].hi(a)) h = nextafterf(h, INFINITY);
		lo[a] = l; hi[a] = h;
	}
}

Hit quantWalk(const Scene& S, const wbvh::Result& W, V3 o, V3 d, float rt0, int bits, bool pow2, QuantStats& st) {
	const V3 inv = {pm::rcp(d.x), pm::rcp(d.y), pm::rcp(d.z)};
	float rt = rt0, bestTn = -INFINITY, t2 = rt0;
	int face = 0, leaf = -1;
	std::vector<std::pair<int, float>> stack;
	int item = 0;
	bool have = !W.nodes.empty();
	float lim = pruneLimit(rt);
	auto pop = [&]() {
		have = false;
		while (!stack.empty()) {
			const std::pair<int, float> e = stack.back();
			stack.pop_back();
			if (e.second <= lim) { item = e.first; have = true; break; }
		}
	};
	while (have) {
		if (item >= 0) {
			st.inner++;
			const wbvh::Node& N = W.nodes[(size_t) item];
			int refs[4], n = 0;
			float tns[4];
			for (int j = 0; j < 4; j++) {
				if (N.c[j].ref == wbvh::REF_EMPTY) continue;
				float lo[3], hi[3], tNear, tFar;
				if (bits > 0) quantBox(N, j, bits, pow2, lo, hi);
				else for (int a = 0; a < 3; a++) { lo[a] = N.c[j].lo(a); hi[a] = N.c[j].hi(a); }
				const bool hit = intersectBox(o, inv, lo, hi, tNear, tFar) && tFar > EPS5;
				if (!(hit && tNear <= lim && tNear < INFINITY)) continue;
				int k = n++;
				while (k > 0 && tns[k - 1] > tNear) { tns[k] = tns[k - 1]; refs[k] = refs[k - 1]; k--; }
				tns[k] = tNear; refs[k] = N.c[j].ref;
			}
			for (int k = n - 1; k >= 1; k--) stack.push_back({refs[k], tns[k]});
			if (n > 0) item = refs[0];
			else pop();
		}
		else {
			st.leaves++;
			const int f0 = (int) ((uint32_t) item & wbvh::REF_FACE_MASK);
			/* the exact leaf box, as the leaf record would carry it */
			const float* lo = S.nodes + 8 * (size_t) W.faceLeaf[(size_t) f0];
			float tNear, tFar;
			const bool hit = intersectBox(o, inv, lo, lo + 4, tNear, tFar) && tFar > EPS5 && tNear <= lim && tNear < INFINITY;
			if (!hit) { st.leavesRejected++; pop(); continue; }
			float tl = faceT(S, f0, o, d, tNear, INFINITY);
			int fl = f0;
			st.tris++;
			if ((uint32_t) item & wbvh::REF_TWO) {
				const float t1 = faceT(S, f0 + 1, o, d, tNear, INFINITY);
				st.tris++;
				if (t1 < tl) { tl = t1; fl = f0 + 1; }
			}
			if (tl < INFINITY) {
				if (tl < rt || (tl == rt && leaf >= 0 && fl < face)) {
					t2 = fminf(t2, rt);
					rt = tl; face = fl; leaf = W.faceLeaf[(size_t) fl]; bestTn = tNear;
				}
				else t2 = fminf(t2, tl);
				lim = pruneLimit(rt);
			}
			pop();
		}
	}
	if (rt < bestTn && t2 <= bestTn) { st.fallbacks++; return strictWalk(S, o, d, rt0); }
	return {rt, face, leaf, 0, 0};
}

} /* namespace */

extern "C" {

/* rays: n x 8 floats (origin.xyz, -, dir.xyz, t0).  out: n x 4 int32 (t bits, face, leaf, -) for both walks.
 * stats: wide nodes, wide depth, top count, leaf refs, inner refs, sum strict nodes, sum strict tris, sum wide visits,
 *        sum fast tris, fallbacks, overflows, insane winners, max stack, mismatches.
 * Returns 0, or 1 when the builder refused the tree (why -> msg). */
int wide_model_run(const float* nodes, int numNodes, const uint32_t* facesV, int numFaces, const float* vertices,
                   const float* rays, long long n, int stackCap, int topBudget,
                   int32_t* outStrict, int32_t* outFast, long long* stats, char* msg, int msgLen) {
	const wbvh::Result W = wbvh::build(nodes, numNodes, numFaces, topBudget);
	if (!W.ok) {
		if (msg && msgLen > 0) { strncpy(msg, W.why.c_str(), (size_t) msgLen - 1); msg[msgLen - 1] = 0; }
		return 1;
	}
	const Scene S = {nodes, numNodes, facesV, numFaces, vertices};
	FastStats st;
	long long sn = 0, stt = 0, mism = 0;
	for (long long i = 0; i < n; i++) {
		const float* r = rays + 8 * i;
		const V3 o = {r[0], r[1], r[2]}, d = {r[4], r[5], r[6]};
		const Hit a = strictWalk(S, o, d, r[7]);
		const Hit b = fastWalk(S, W, o, d, r[7], stackCap, st);
		sn += a.nodes; stt += a.tris;
		int32_t ta, tb;
		memcpy(&ta, &a.t, 4); memcpy(&tb, &b.t, 4);
		outStrict[4 * i] = ta; outStrict[4 * i + 1] = a.face; outStrict[4 * i + 2] = a.leaf; outStrict[4 * i + 3] = (int32_t) a.nodes;
		outFast[4 * i] = tb; outFast[4 * i + 1] = b.face; outFast[4 * i + 2] = b.leaf; outFast[4 * i + 3] = (int32_t) b.nodes;
		if (ta != tb || a.face != b.face || a.leaf != b.leaf) mism++;
	}
	stats[0] = (long long) W.nodes.size(); stats[1] = W.depth; stats[2] = W.topCount; stats[3] = W.leafRefs; stats[4] = W.innerRefs;
	stats[5] = sn; stats[6] = stt; stats[7] = (long long) st.wideVisits; stats[8] = (long long) st.triTests;
	stats[9] = (long long) st.fallbacks; stats[10] = (long long) st.overflow; stats[11] = (long long) st.insaneWinners;
	stats[12] = st.maxStack; stats[13] = mism;
	return 0;
}

/* stats: inner visits, leaf visits, leaf visits the exact box rejects, triangle tests, re-walks, mismatches against the
 * reference-order walk.  bits = 0: the exact boxes (the walk as shipped, leaf box test repeated at the leaf). */
int wide_model_quant(const float* nodes, int numNodes, const uint32_t* facesV, int numFaces, const float* vertices,
                     const float* rays, long long n, int topBudget, int bits, int pow2, long long* stats) {
	const wbvh::Result W = wbvh::build(nodes, numNodes, numFaces, topBudget);
	if (!W.ok) return 1;
	const Scene S = {nodes, numNodes, facesV, numFaces, vertices};
	QuantStats st;
	for (long long i = 0; i < n; i++) {
		const float* r = rays + 8 * i;
		const V3 o = {r[0], r[1], r[2]}, d = {r[4], r[5], r[6]};
		const Hit a = strictWalk(S, o, d, r[7]);
		const Hit b = quantWalk(S, W, o, d, r[7], bits, pow2 != 0, st);
		if (memcmp(&a.t, &b.t, 4) != 0 || a.face != b.face || a.leaf != b.leaf) st.mismatches++;
	}
	stats[0] = (long long) st.inner; stats[1] = (long long) st.leaves; stats[2] = (long long) st.leavesRejected;
	stats[3] = (long long) st.tris; stats[4] = (long long) st.fallbacks; stats[5] = (long long) st.mismatches;
	return 0;
}

}
