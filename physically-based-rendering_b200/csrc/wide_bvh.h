/*
 * wide_bvh.h -- host-side construction of the 4-wide BVH the ordered ("fast") traversal walks.
 *
 * Input is the reference's flattened node array exactly as PathTracer::initOpenCLBuffers_BVH uploads it
 * (source/PathTracer.cpp:238-347: pre-order, left child at i + 1, bbMin.w = first face or -1, bbMax.w = second
 * face / miss link, float-encoded).  That array is a tree in disguise: the subtree of an inner node i is
 * [i, link(i)), its children are i + 1, then the node behind each child's subtree -- with skip-ahead
 * (PathTracer.cpp:253-256, 271-273) a node can have three or more children.  This file recovers that tree,
 * checks the properties the ordered walk relies on, and collapses it into nodes of up to four children:
 *
 *     Node  = 4 x Child, 128 bytes, one cache line
 *     Child = box as (min.x, max.x, min.y, max.y, min.z, max.z) -- binary32, copied bit for bit from the reference
 *             node; min / max of an axis side by side so that one packed FADD2 / FMUL2 handles both slab planes --, ref, aux
 *             ref  >= 0          inner child: index of its Node
 *             ref  bit 31 set    leaf: bits 0..29 first face, bit 30 = the leaf has a second face (first + 1)
 *             ref  == REF_EMPTY  unused slot (box is NaN: never hit)
 *             aux  leaf: index of the leaf in the REFERENCE array; inner: unused
 *     faceLeaf[f] = index of the reference leaf that tests face f (the `leaf` of pbr_hit), -1 for a face no leaf tests
 *
 * A leaf's box is the reference leaf's box bit for bit, because flatTriAndRayIntersect starts from a point
 * derived from the leaf box's tNear (pt_intersect.cl:96-97): the t of a hit depends on it.  Inner boxes only
 * have to be conservative.
 *
 * build() REFUSES (ok = false, `why` says it) when the array is not what the reference's host code produces in a way
 * that matters to the ordered walk -- links that do not nest, a child box that sticks out of its parent's (the
 * reference would then skip a leaf whose own box is hit; the ordered walk could not know), the -2 flag of
 * traverseShadows, a second face that is not first + 1.  The caller then keeps the reference-order walk.
 *
 * Plain C++, no CUDA: included by pbr_capi.cu (host side) and by scripts/wide_proto.cpp.
 */
#pragma once

#include <math.h>
#include <stdint.h>
#include <string.h>

#include <algorithm>
#include <string>
#include <vector>

namespace wbvh {

struct Child {
	float b[6];                   /* min.x, max.x, min.y, max.y, min.z, max.z */
	int32_t ref;
	int32_t aux;
	float lo(int a) const { return b[2 * a]; }
	float hi(int a) const { return b[2 * a + 1]; }
};

struct Node {
	Child c[4];
};

static_assert(sizeof(Child) == 32 && sizeof(Node) == 128, "wide node layout");

enum : uint32_t { REF_LEAF = 0x80000000u, REF_TWO = 0x40000000u, REF_FACE_MASK = 0x3fffffffu };
static const int32_t REF_EMPTY = 0x7fffffff;

struct Result {
	bool ok = false;
	std::string why;
	std::vector<Node> nodes;      /* node 0 is the root; nodes [0, topCount) are the top of the tree in breadth-first order */
	std::vector<int32_t> faceLeaf;
	int topCount = 0;
	int depth = 0;                /* levels of wide nodes */
	long long leafRefs = 0;       /* leaf children */
	long long innerRefs = 0;      /* inner children */
};

namespace detail {

struct Tmp {                      /* a wide node before the final numbering */
	int kind[4];                  /* 0 empty, 1 leaf (src = reference node), 2 inner (child = Tmp index) */
	int src[4];                   /* reference node the box comes from, or -1: box in `box` */
	int child[4];
	float box[4][6];
	int n = 0;
};

struct Builder {
	const float* src;
	int numNodes, numFaces;
	std::vector<int> end;         /* subtree of i = [i, end[i]) */
	std::vector<unsigned char> kind;   /* 0 inner, 1 leaf, 2 neither (box only: the walk just steps over it) */
	std::vector<Tmp> tmp;
	std::string why;

	const float* lo(int i) const { return src + 8 * (size_t) i; }
	const float* hi(int i) const { return src + 8 * (size_t) i + 4; }

	float area(int i) const {
		const float ex = fmaxf(hi(i)[0] - lo(i)[0], 0.0f), ey = fmaxf(hi(i)[1] - lo(i)[1], 0.0f), ez = fmaxf(hi(i)[2] - lo(i)[2], 0.0f);
		return ex * ey + ey * ez + ez * ex;
	}

	bool fail(const std::string& m) { why = m; return false; }

	/* children of reference node i that can matter to a walk: leaves, and inner nodes that have children */
	void childrenOf(int i, int iEnd, std::vector<int>& out) const {
		for (int c = i + 1; c < iEnd; c = end[(size_t) c]) {
			if (kind[(size_t) c] == 1 || (kind[(size_t) c] == 0 && end[(size_t) c] > c + 1)) out.push_back(c);
		}
	}

	bool parse() {
		const int n = numNodes;
		end.assign((size_t) (n > 0 ? n : 1), 0);
		kind.assign((size_t) (n > 0 ? n : 1), 2);
		std::vector<int> openEnd, openOwner;
		openEnd.push_back(n);
		openOwner.push_back(-1);
		for (int i = 1; i < n; i++) {
			while (openEnd.size() > 1 && i == openEnd.back()) { openEnd.pop_back(); openOwner.pop_back(); }
			if (i > openEnd.back()) return fail("node links do not nest");
			const float lw = lo(i)[3], hw = hi(i)[3];
			for (int a = 0; a < 3; a++) {
				if (!(lo(i)[a] == lo(i)[a]) || !(hi(i)[a] == hi(i)[a])) return fail("NaN in a node box");
			}
			const int owner = openOwner.back();
			if (owner >= 1) {
				for (int a = 0; a < 3; a++) {
					if (!(lo(i)[a] >= lo(owner)[a]) || !(hi(i)[a] <= hi(owner)[a])) return fail("a child box is not inside its parent's box");
				}
			}
			if (lw <= -1.0f) {
				if (lw == -2.0f) return fail("traverseShadows' skip flag (-2) is set");
				kind[(size_t) i] = 0;
				const long long link = (long long) hw;
				int e = n;
				if (hw == hw && link > 0 && link < (long long) n) {
					if (link <= (long long) i) return fail("a miss link points backwards");
					e = (int) link;
				}
				if (e > openEnd.back()) return fail("node links do not nest");
				end[(size_t) i] = e;
				if (e > i + 1) { openEnd.push_back(e); openOwner.push_back(i); }
			}
			else if (lw >= 0.0f) {
				kind[(size_t) i] = 1;
				end[(size_t) i] = i + 1;
				const long long f0 = (long long) lw;
				if (f0 < 0 || f0 >= (long long) numFaces || f0 > (long long) REF_FACE_MASK) return fail("a leaf's face index is out of range");
				if (hw != -1.0f) {
					const long long f1 = (long long) hw;
					if (!(hw == hw) || f1 != f0 + 1 || f1 >= (long long) numFaces) return fail("a leaf's second face is not first + 1");
				}
			}
			else {
				kind[(size_t) i] = 2;
				end[(size_t) i] = i + 1;
			}
		}
		return true;
	}

	/* One wide node from a list of reference nodes (siblings, or the children pulled up from below). */
	int makeWide(std::vector<int> list, int level, int& depth) {
		depth = std::max(depth, level + 1);
		const int me = (int) tmp.size();
		tmp.emplace_back();
		for (int k = 0; k < 4; k++) { tmp[(size_t) me].kind[k] = 0; tmp[(size_t) me].src[k] = -1; tmp[(size_t) me].child[k] = -1; }
		if (list.size() > 4) {
			/* more than four siblings (nested skip-ahead): split the run into four parts, each part with more than one
			 * member becomes an artificial node whose box is the union of its members */
			const size_t k = list.size();
			size_t at = 0;
			for (int part = 0; part < 4; part++) {
				const size_t cnt = k / 4 + ((size_t) part < k % 4 ? 1 : 0);
				std::vector<int> sub(list.begin() + (long) at, list.begin() + (long) (at + cnt));
				at += cnt;
				if (sub.size() == 1) { addRef(me, sub[0], level, depth); continue; }
				float b[6] = {INFINITY, INFINITY, INFINITY, -INFINITY, -INFINITY, -INFINITY};
				for (int r : sub) for (int a = 0; a < 3; a++) { b[a] = fminf(b[a], lo(r)[a]); b[3 + a] = fmaxf(b[3 + a], hi(r)[a]); }
				const int child = makeWide(sub, level + 1, depth);
				Tmp& T = tmp[(size_t) me];
				T.kind[T.n] = 2; T.src[T.n] = -1; T.child[T.n] = child;
				memcpy(T.box[T.n], b, sizeof(b));
				T.n++;
			}
			return me;
		}
		/* pull grandchildren up while there is room: always open the inner node with the largest surface area */
		while (list.size() < 4) {
			int best = -1;
			float bestArea = -1.0f;
			std::vector<int> kids;
			for (size_t k = 0; k < list.size(); k++) {
				const int c = list[k];
				if (kind[(size_t) c] != 0) continue;
				kids.clear();
				childrenOf(c, end[(size_t) c], kids);
				if (kids.empty() || list.size() - 1 + kids.size() > 4) continue;
				const float a = area(c);
				if (a > bestArea) { bestArea = a; best = (int) k; }
			}
			if (best < 0) break;
			kids.clear();
			childrenOf(list[(size_t) best], end[(size_t) list[(size_t) best]], kids);
			list.erase(list.begin() + best);
			list.insert(list.end(), kids.begin(), kids.end());
		}
		for (int r : list) addRef(me, r, level, depth);
		return me;
	}

	void addRef(int me, int r, int level, int& depth) {
		int child = -1;
		if (kind[(size_t) r] == 0) {
			std::vector<int> kids;
			childrenOf(r, end[(size_t) r], kids);
			child = makeWide(kids, level + 1, depth);
		}
		Tmp& T = tmp[(size_t) me];                 /* (after the recursion: tmp may have been reallocated) */
		T.kind[T.n] = kind[(size_t) r] == 1 ? 1 : 2;
		T.src[T.n] = r;
		T.child[T.n] = child;
		T.n++;
	}
};

} /* namespace detail */

/* topBudget: how many nodes of the top of the tree are numbered first, breadth-first (the traversal kernel keeps
 * nodes [0, topCount) in shared memory). */
inline Result build(const float* src, int numNodes, int numFaces, int topBudget) {
	Result R;
	detail::Builder B;
	B.src = src; B.numNodes = numNodes; B.numFaces = numFaces;
	if (!B.parse()) { R.why = B.why; return R; }

	std::vector<int> top;
	B.childrenOf(0, numNodes, top);                /* node 0 is never tested (pt_bvh.cl:84): the walk starts with its children */
	B.tmp.reserve((size_t) numNodes / 2 + 16);
	int depth = 0;
	B.makeWide(top, 0, depth);
	R.depth = depth;

	/* final numbering: breadth-first for the first topBudget nodes, depth-first below them */
	const size_t n = B.tmp.size();
	std::vector<int> order;
	order.reserve(n);
	std::vector<int> newIndex(n, -1);
	std::vector<int> frontier;
	frontier.push_back(0);
	size_t head = 0;
	while (head < frontier.size() && (int) order.size() < topBudget) {
		const int t = frontier[head++];
		newIndex[(size_t) t] = (int) order.size();
		order.push_back(t);
		for (int k = 0; k < B.tmp[(size_t) t].n; k++) if (B.tmp[(size_t) t].kind[k] == 2) frontier.push_back(B.tmp[(size_t) t].child[k]);
	}
	R.topCount = (int) order.size();
	std::vector<int> stack;
	for (size_t f = frontier.size(); f > head; f--) stack.push_back(frontier[f - 1]);
	while (!stack.empty()) {
		const int t = stack.back();
		stack.pop_back();
		newIndex[(size_t) t] = (int) order.size();
		order.push_back(t);
		for (int k = B.tmp[(size_t) t].n - 1; k >= 0; k--) if (B.tmp[(size_t) t].kind[k] == 2) stack.push_back(B.tmp[(size_t) t].child[k]);
	}

	const float nan = nanf("");
	R.faceLeaf.assign((size_t) (numFaces > 0 ? numFaces : 1), -1);
	R.nodes.resize(order.size());
	for (size_t j = 0; j < order.size(); j++) {
		const detail::Tmp& T = B.tmp[(size_t) order[j]];
		Node& N = R.nodes[j];
		for (int k = 0; k < 4; k++) {
			Child& C = N.c[k];
			if (k >= T.n || T.kind[k] == 0) {
				for (int a = 0; a < 6; a++) C.b[a] = nan;
				C.ref = REF_EMPTY;
				C.aux = 0;
				continue;
			}
			const float* blo = T.src[k] >= 0 ? B.lo(T.src[k]) : T.box[k];
			const float* bhi = T.src[k] >= 0 ? B.hi(T.src[k]) : T.box[k] + 3;
			for (int a = 0; a < 3; a++) { C.b[2 * a] = blo[a]; C.b[2 * a + 1] = bhi[a]; }
			if (T.kind[k] == 1) {
				const int r = T.src[k];
				const uint32_t f0 = (uint32_t) (long long) B.lo(r)[3];
				const bool two = B.hi(r)[3] != -1.0f;
				C.ref = (int32_t) (REF_LEAF | (two ? REF_TWO : 0u) | f0);
				C.aux = r;
				R.faceLeaf[(size_t) f0] = r;
				if (two) R.faceLeaf[(size_t) f0 + 1] = r;
				R.leafRefs++;
			}
			else {
				C.ref = newIndex[(size_t) T.child[k]];
				C.aux = 0;
				R.innerRefs++;
			}
		}
	}
	R.ok = true;
	return R;
}

} /* namespace wbvh */
