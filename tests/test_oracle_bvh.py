"""CPU: the BVH / flatten restatement (oracle/bvh_oracle.cpp) against structural invariants and the
reference's soft known answers."""
import numpy as np
import pytest

import helpers as Hh


def _walk_all_hit(nodes):
    """Order in which a ray that hits every box visits nodes: 1, 2, 3, ... (pt_bvh.cl:115)."""
    n = nodes.shape[0]
    return list(range(1, n))


def _check_flat(nodes, facesV, num_faces):
    n = nodes.shape[0]
    lo_w, hi_w = nodes[:, 3], nodes[:, 7]
    inner = lo_w <= -1.0
    leaf = ~inner
    # leaves reference consecutive faces, in order, each face exactly once
    f0 = lo_w[leaf].astype(np.int64)
    f1 = hi_w[leaf].astype(np.int64)
    cnt = np.where(f1 >= 0, 2, 1)
    assert np.all(f1[f1 >= 0] == f0[f1 >= 0] + 1)
    assert f0[0] == 0
    assert np.all(f0[1:] == f0[:-1] + cnt[:-1])
    assert f0[-1] + cnt[-1] == num_faces == facesV.shape[0]
    # miss links of inner nodes: -1 or a strictly larger index inside the array
    links = hi_w[inner].astype(np.int64)
    idx = np.nonzero(inner)[0]
    assert np.all((links == -1) | ((links > idx) & (links < n)))
    # the all-miss walk terminates and only moves forward
    i, steps = 1, 0
    while 0 < i < n:
        i = int(hi_w[i]) if inner[i] else i + 1
        steps += 1
        assert steps <= n
    return int(leaf.sum())


@pytest.mark.parametrize("name,faces", [("suzanne.obj", 1082), ("pillars.obj", 56)])
def test_bundled_scene_counts(oracle, name, faces):
    s = oracle.load_obj(Hh.model_path(name), 0)
    assert len(s["facesV"]) // 3 == faces          # pathtracing.cl:75: 1082 faces in the test model
    b = oracle.build_bvh(s)
    assert b["info"]["faces"] == faces
    assert b["info"]["emitted"] == b["info"]["allNodes"] - b["info"]["skipped"]
    leaves = _check_flat(b["nodes"], b["facesV"], faces)
    assert leaves == b["info"]["leaves"]


def test_suzanne_parse_statistics(oracle):
    """SURVEY.md 8c soft pins: 603 v / 549 vn / 1082 f / 10 objects; one orb light."""
    s = oracle.load_obj(Hh.model_path("suzanne.obj"), 1)
    assert (len(s["vertices"]) // 3, len(s["normals"]) // 3, len(s["objFaceCounts"])) == (603, 549, 10)
    assert len(s["materials"]) == 13 and "sky_light" in s["materialNames"]
    assert s["lights"].shape == (1, 10)
    li = s["lights"][0]
    assert li[0] == 2 and np.allclose(li[1:4], (0.1, 1.3, 1.2)) and np.isclose(li[9], 0.1)
    # lights are only read when render.shadow_rays > 0 (ObjParser.cpp:133-135)
    assert oracle.load_obj(Hh.model_path("suzanne.obj"), 0)["lights"].shape[0] == 0


def test_leaf_boxes_contain_their_faces(oracle, oscene):
    s = oracle.load_obj(Hh.model_path("suzanne.obj"), 0)
    b = oracle.build_bvh(s)
    v = oscene.pack_float4(s["vertices"])
    nodes, fv = b["nodes"], b["facesV"]
    for i in np.nonzero(nodes[:, 3] >= 0)[0]:
        for f in (int(nodes[i, 3]), int(nodes[i, 7])):
            if f < 0:
                continue
            tri = v[fv[f, :3], :3]
            assert np.all(tri.min(0) >= nodes[i, 0:3]) and np.all(tri.max(0) <= nodes[i, 4:7])


@pytest.mark.parametrize("skip_ahead", [True, False])
@pytest.mark.parametrize("max_faces", [1, 2])
def test_soup_invariants(oracle, skip_ahead, max_faces):
    import pbr_b200
    s = pbr_b200.scenes.soup(3000, seed=7)
    b = oracle.build_bvh(s, max_faces=max_faces, skip_ahead=skip_ahead, sah_faces_limit=1000)
    _check_flat(b["nodes"], b["facesV"], 3000)
    if not skip_ahead:
        assert b["info"]["skipped"] == 0
    # material index is carried in facesV.w; each original face appears once
    assert np.all(b["facesV"][:, 3] == 0)
    key = np.sort(b["facesV"][:, 0] // 3)
    assert np.array_equal(key, np.arange(3000))


def test_appendix_b_example(oracle):
    """R(A(a1,a2), B(C(c1,c2), b2)) without skip-ahead gives miss links
    A->B, B->-1, C->b2.  Built from two objects of 4 and 6 well separated triangles."""
    # Construct geometry whose SAH tree has that shape is brittle; instead check the generic rule on
    # a real tree: an inner LEFT child links to its right sibling = the first node after its subtree.
    s = oracle.load_obj(Hh.model_path("pillars.obj"), 0)
    b = oracle.build_bvh(s, skip_ahead=False)
    nodes = b["nodes"]
    n = nodes.shape[0]
    inner = nodes[:, 3] <= -1.0

    def subtree_end(i):
        # pre-order: a leaf ends at i+1; an inner node ends where its right child's subtree ends
        if not inner[i]:
            return i + 1
        left_end = subtree_end(i + 1)
        return subtree_end(left_end)
    for i in range(1, n):
        if inner[i]:
            end = subtree_end(i)
            link = int(nodes[i, 7])
            assert link == (end if end < n else -1)
