"""CPU prototype of the node permutation planned in DESIGN.md section 6 (hot nodes dense, explicit links), with
the proof that it does not change a walk: rays are stepped through the reference-order array and through the
permuted one side by side, and at every step both must stand on the same node (mapped back through the index
table).  The device repack kernel of the next round is a port of build_permuted().

Encoding of a permuted record (8 x 32 bit, still one 256-bit load):
    inner:  lo.xyz box min | lo.w = hit link    hi.xyz box max | hi.w = miss link
    leaf:   lo.xyz box min | lo.w = first face  hi.xyz box max | hi.w = successor | LEAF | (TWO if a second face, = first + 1)
    link 0 = walk finished (node 0, the root, is never visited: pt_bvh.cl:84); position 1 stays the first node visited.

    python scripts/node_permutation_proto.py [tris]"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import pbr_b200  # noqa: E402,F401
from pbr_b200 import host, scenes  # noqa: E402
import helpers as Hh  # noqa: E402

LEAF, TWO, LINK = np.int32(-2 ** 31), np.int32(1 << 30), np.int32((1 << 30) - 1)


def build_permuted(nodes):
    """nodes: [N, 8] float32 in the uploaded layout (bbMin.w = first face or -1, bbMax.w = second face / miss link / -1).
    Returns (box [N, 6] f32, lo_w [N] i32, hi_w [N] i32, orig [N] i32), all in permuted order."""
    N = nodes.shape[0]
    inner = nodes[:, 3] <= -1.0
    ext = np.maximum(nodes[:, 4:7] - nodes[:, 0:3], 0)
    area = ext[:, 0] * ext[:, 1] + ext[:, 1] * ext[:, 2] + ext[:, 0] * ext[:, 2]
    key = -area.astype(np.float64)
    key[0], key[1] = -np.inf, -1e300            # positions 0 and 1 stay where they are
    orig = np.argsort(key, kind="stable").astype(np.int32)       # permuted position -> reference index
    pos = np.empty(N, np.int32)
    pos[orig] = np.arange(N, dtype=np.int32)                     # reference index -> permuted position

    def link(target):                                            # reference index (or -1 / N = stop) -> permuted link
        t = np.asarray(target, np.int64)
        ok = (t > 0) & (t < N)
        return np.where(ok, pos[np.clip(t, 0, N - 1)], 0).astype(np.int32)

    idx = np.arange(N, dtype=np.int64)
    hit = link(idx + 1)
    miss = link(nodes[:, 7].astype(np.int64))
    lo_w = np.where(inner, hit, nodes[:, 3].astype(np.int32))
    two = ~inner & (nodes[:, 7] >= 0)
    assert np.all(nodes[two, 7].astype(np.int64) == nodes[two, 3].astype(np.int64) + 1)   # PathTracer.cpp:267-268
    hi_w = np.where(inner, miss, hit | LEAF | np.where(two, TWO, np.int32(0))).astype(np.int32)
    box = np.concatenate([nodes[:, 0:3], nodes[:, 4:7]], axis=1)
    return box[orig], lo_w[orig], hi_w[orig], orig


def box_hit(box, o, inv):
    """intersectBox (pt_intersect.cl:11-25) && tFar > 1e-5, without the t pruning: the link structure is what is tested."""
    t1 = (box[:, 0:3] - o) * inv
    t2 = (box[:, 3:6] - o) * inv
    tmin, tmax = np.fmin(t1, t2), np.fmax(t1, t2)
    near = np.maximum(np.maximum(tmin[:, 0], tmin[:, 1]), tmin[:, 2])
    far = np.minimum(np.minimum(tmax[:, 0], tmax[:, 1]), tmax[:, 2])
    return (near <= far) & (far > np.float32(1e-5))


def main():
    tris = int(sys.argv[1]) if len(sys.argv) > 1 else 50000
    nodes = np.ascontiguousarray(host.Scene.from_arrays(scenes.soup(tris, seed=12345)).build_flat()["nodes"], np.float32).reshape(-1, 8)
    N = nodes.shape[0]
    box, lo_w, hi_w, orig = build_permuted(nodes)
    rays = Hh.random_rays(4000, 5, -1.0, 1.0)
    o = rays[:, 0:3].astype(np.float32)
    with np.errstate(divide="ignore"):
        inv = (np.float32(1.0) / rays[:, 4:7]).astype(np.float32)
    ref_box = np.concatenate([nodes[:, 0:3], nodes[:, 4:7]], axis=1)
    ref_inner = nodes[:, 3] <= -1.0
    a = np.ones(len(rays), np.int64)             # reference-order index
    b = np.ones(len(rays), np.int64)             # permuted index
    steps = visits = leaf_visits = 0
    while True:
        live = (a > 0) & (a < N)
        assert np.array_equal(live, b != 0), "one walk ended, the other did not"
        if not live.any():
            break
        assert np.array_equal(orig[b[live]], a[live]), "the walks stand on different nodes at step %d" % steps
        ai, bi = a[live], b[live]
        # reference order (pt_bvh.cl:88-121)
        h = box_hit(ref_box[ai], o[live], inv[live])
        nxt = np.where(ref_inner[ai], nodes[ai, 7].astype(np.int64), ai + 1)
        a[live] = np.where(h, ai + 1, nxt)
        # permuted
        hb = box_hit(box[bi], o[live], inv[live])
        assert np.array_equal(h, hb)
        leaf = hi_w[bi] < 0
        succ = (hi_w[bi] & LINK).astype(np.int64)
        b[live] = np.where(hb & ~leaf, lo_w[bi].astype(np.int64), succ)
        # a leaf's faces come out of the record unchanged
        lf = leaf & hb
        f0_ref = nodes[ai[lf], 3].astype(np.int64)
        f1_ref = nodes[ai[lf], 7].astype(np.int64)
        f1 = np.where((hi_w[bi[lf]] & TWO) != 0, lo_w[bi[lf]].astype(np.int64) + 1, -1)
        assert np.array_equal(lo_w[bi[lf]].astype(np.int64), f0_ref) and np.array_equal(f1, f1_ref)
        steps += 1
        visits += int(live.sum())
        leaf_visits += int(lf.sum())
    hot = np.sort(orig[:8192])
    print("%d nodes, %d rays, %d steps, %d visits (%d leaves with their box hit): the permuted walk visits the same nodes "
          "in the same order and hands out the same faces" % (N, len(rays), steps, visits, leaf_visits))
    print("the 8192 records now at the front came from positions up to %d of %d (median %d)" % (hot[-1], N, int(np.median(hot))))


if __name__ == "__main__":
    main()
