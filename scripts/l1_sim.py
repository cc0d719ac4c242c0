"""CPU estimate for the node-permutation experiment: an LRU cache of 128-byte lines the size of
an SM's L1 in front of the node and triangle fetches of as many rays as an SM has in flight, stepped round-robin.
The cache is modelled the way the hardware fills it -- a miss brings in the one 32-byte sector that was asked for, the
other three sectors of the line stay invalid -- and, for comparison, as if a miss filled the whole line.  Layouts: the
reference's pre-order against the array permuted by surface area (scripts/node_permutation_proto.py).
    python scripts/l1_sim.py [tris]"""
import ctypes as C
import os
import sys
from collections import OrderedDict

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, os.path.join(ROOT, "scripts"))
import pbr_b200  # noqa: E402,F401
from pbr_b200 import host, scenes  # noqa: E402
from oracle import oracle as O  # noqa: E402
from oracle import scene as S  # noqa: E402
import helpers as Hh  # noqa: E402
from node_permutation_proto import build_permuted  # noqa: E402

tris = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
IN_FLIGHT = 9 * 128            # rays an SM holds: 9 blocks of 128 lanes
CAP = 1024
sc = scenes.soup(tris, seed=12345)
flat = host.Scene.from_arrays(sc).build_flat()
nodes = np.ascontiguousarray(flat["nodes"], np.float32).reshape(-1, 8)
N = nodes.shape[0]
facesV = np.ascontiguousarray(flat["facesV"], np.uint32)
facesN = np.ascontiguousarray(flat["facesN"], np.uint32)
v4 = S.pack_float4(sc["vertices"])
n4 = np.zeros((1, 4), np.float32)
_, _, _, orig = build_permuted(nodes)
pos = np.empty(N, np.int64)
pos[orig] = np.arange(N)


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


class P:
    camera = S.camera(eye=(0.0, 0.0, 3.5), center=(0.0, 0.0, 1.0))


L = O.lib()
D = O.make_defines(img_width=1920, img_height=1080, bvh_num_nodes=N)


def traces(rays):
    rays = np.ascontiguousarray(rays, np.float32)
    tr = np.zeros((len(rays), CAP), np.int32)
    ln = np.zeros(len(rays), np.int32)
    L.oracle_visit_trace(_p(D), _p(nodes), _p(facesV), _p(facesN), _p(v4), _p(n4), _p(rays), C.c_int64(len(rays)), C.c_int32(CAP),
                         _p(tr), _p(ln))
    return tr, np.minimum(ln, CAP)


TRI_TAG, TRIB_TAG = 1 << 40, 1 << 41


def hit_rate(tr, ln, index_of, lines, with_tris=False, sectored=True, neighbour=0.0):
    """Rays enter in order, IN_FLIGHT at a time, one node step per ray per round; a finished ray is replaced by the next."""
    cache = OrderedDict()
    rng = np.random.default_rng(0)
    helped = None
    hits = total = 0
    nxt = min(IN_FLIGHT, len(ln))
    slot_ray = list(range(nxt))
    slot_step = [0] * nxt
    live = nxt
    while live:
        for s in range(len(slot_ray)):
            r = slot_ray[s]
            if r < 0:
                continue
            e = int(tr[r, slot_step[s]])
            if e >= 0:
                k = int(index_of[e])
                touched = ((k >> 2, 1 << (k & 3)),)                 # four 32-byte nodes (sectors) per 128-byte line
                total += 1
                if neighbour > 0.0 and (k & 3) != 3 and rng.random() < neighbour:
                    helped = (k >> 2, 1 << ((k & 3) + 1))           # a resting lane asks for the next sector of the line
            elif not with_tris:
                touched = ()
            else:
                f = -e - 1                                          # 32-byte record + 4-byte word of a face: two more lines
                touched = ((TRI_TAG + (f >> 2), 1 << (f & 3)), (TRIB_TAG + (f >> 5), 1 << ((f >> 3) & 3)))
            if e >= 0 and helped is not None:
                touched = touched + (helped + (True,),)
                helped = None
            for entry in touched:
                line, sector = entry[0], entry[1]
                silent = len(entry) > 2
                valid = cache.get(line)
                if valid is not None:
                    cache.move_to_end(line)
                    if sectored and not (valid & sector):
                        cache[line] = valid | sector                # tag hit, sector miss: only this sector is fetched
                    elif not silent:
                        hits += e >= 0
                else:
                    cache[line] = sector if sectored else 15
                    if len(cache) > lines:
                        cache.popitem(last=False)
            slot_step[s] += 1
            if slot_step[s] >= ln[r]:
                if nxt < len(ln):
                    slot_ray[s], slot_step[s] = nxt, 0
                    nxt += 1
                else:
                    slot_ray[s] = -1
                    live -= 1
    return hits / total


ident = np.arange(N, dtype=np.int64)
# a 1080p frame hands an SM 8x4-pixel blocks; 4608 rays = 144 consecutive blocks of the 320x180 grid used here
sets = (("primary", Hh.primary_rays(P, 320, 180)[:4 * IN_FLIGHT]), ("random", Hh.random_rays(4 * IN_FLIGHT, 1, -1.0, 1.0)))
for name, rays in sets:
    tr, ln = traces(rays)
    for kb in ([] if os.environ.get('L1_SIM_SKIP_CACHE') else [224]):
        lines = kb * 1024 // 128
        a = hit_rate(tr, ln, ident, lines, with_tris=True)
        b = hit_rate(tr, ln, pos, lines, with_tris=True)
        c = hit_rate(tr, ln, ident, lines, with_tris=True, sectored=False)
        d = hit_rate(tr, ln, ident, lines, with_tris=False)
        h1 = hit_rate(tr, ln, ident, lines, with_tris=True, neighbour=0.45)
        h2 = hit_rate(tr, ln, ident, lines, with_tris=True, neighbour=0.7)
        print("%-8s L1 %3d KB: node fetches hitting (triangle records go through the same cache), sector fills: pre-order %.1f %%, "
              "by surface area %.1f %%, pre-order with the triangle records kept out of L1 %.1f %%, pre-order with the neighbour sector "
              "fetched along on 45 %% / 70 %% of the fetches %.1f %% / %.1f %%;  whole-line fills, pre-order: %.1f %%" % (
                  name, kb, 100 * a, 100 * b, 100 * d, 100 * h1, 100 * h2, 100 * c), flush=True)

# What a private line buffer per ray would catch (no cache at all): visits that stay in the 128-byte line of the
# previous visit of the same ray, or in one of the last two lines; the same for 64-byte pairs.
for name, rays in sets:
    tr, ln = traces(rays)
    same = two = pair = total = 0
    for r in range(len(ln)):
        nodes_only = tr[r, :ln[r]][tr[r, :ln[r]] >= 0]
        li = nodes_only >> 2
        total += len(li) - 1
        s1 = li[1:] == li[:-1]
        same += int(s1.sum())
        two += int((s1[1:] | (li[2:] == li[:-2])).sum()) + int(s1[:1].sum())
        pi = nodes_only >> 1
        pair += int((pi[1:] == pi[:-1]).sum())
    print("%-8s private buffer per ray: next node in the same 128-byte line %.1f %% of the visits, in one of the last two lines "
          "%.1f %%, in the same 64-byte pair %.1f %%" % (name, 100 * same / total, 100 * two / total, 100 * pair / total), flush=True)
