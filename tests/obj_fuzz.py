"""Random, plausible OBJ text for the parser comparison: every number spelling, blank and tab runs, CR line ends,
all four face styles, objects, usemtl in odd places, unknown statements.  Lines on which the reference indexes past
its token vector (a bare `o`, `usemtl`, `v 1 2`) are not generated: undefined upstream."""
import numpy as np


def num(rng):
    k = rng.integers(0, 9)
    x = rng.uniform(-3, 3)
    return ["%.6f" % x, "%.3f" % x, "%g" % x, "%e" % x, "%d" % int(x), "%.9g" % x, "+%.2f" % abs(x), ".%d" % rng.integers(0, 999), "%de-%d" % (int(x * 10), rng.integers(0, 5))][k]

def sep(rng):
    return [" ", " ", " ", "  ", "\t", " \t"][rng.integers(0, 6)]

def gen(rng, nlines):
    out = []
    nv = nn = nt = 0
    mtls = ["red", "glass", "nope", "sky_light"]
    for _ in range(nlines):
        k = rng.integers(0, 100)
        if k < 30:
            parts = ["v"] + [num(rng) for _ in range(3 + (rng.integers(0, 8) == 0))]; nv += 1
            line = parts[0] + " " + sep(rng).join(parts[1:]) if rng.integers(0, 4) else " ".join(parts)
        elif k < 45:
            line = "vn " + " ".join(num(rng) for _ in range(3)); nn += 1
        elif k < 52:
            line = "vt " + " ".join(num(rng) for _ in range(2 + rng.integers(0, 2))); nt += 1
        elif k < 80 and nv > 3 and nn > 3:
            style = rng.integers(0, 5)
            groups = []
            for _ in range(3 + (rng.integers(0, 10) == 0)):
                v, n, t = rng.integers(1, nv + 1), rng.integers(1, nn + 1), rng.integers(1, max(nt, 1) + 1)
                groups.append(["%d" % v, "%d//%d" % (v, n), "%d/%d/%d" % (v, t, n), "%d/%d" % (v, t), "%d//%d" % (v, n)][style])
            line = "f " + (sep(rng) if rng.integers(0, 6) == 0 else " ").join(groups)
        elif k < 84:
            line = "o obj%d" % rng.integers(0, 99) + ("" if rng.integers(0, 3) else " extra")
        elif k < 90:
            line = ["usemtl ", "  usemtl ", "g usemtl "][rng.integers(0, 3)] + mtls[rng.integers(0, 4)]
        elif k < 93:
            line = "# " + "comment usemtl red"
        elif k < 95:
            line = ""
        elif k < 97:
            line = ["s off", "g group", "mtllib other.mtl", "off 1 2", "vp 1 2 3", "fx 1 2 3", "vnx 1 2 3"][rng.integers(0, 7)]
        else:
            line = "   " + "v " + " ".join(num(rng) for _ in range(3)) + "   "; nv += 1
        if rng.integers(0, 12) == 0:
            line += "\r"
        if rng.integers(0, 15) == 0:
            line = "\t" + line + " "
        out.append(line)
    return "\n".join(out) + ("\n" if rng.integers(0, 2) else "")


def gen_mtl(rng, nlines):
    """Random .mtl text: every keyword of MtlParser.cpp:79-226, too few / too many values, unknown keywords."""
    keys3, keys1 = ["Ka", "Kd", "Ks"], ["d", "Tr", "illum", "Ni", "Ns", "light", "rough", "p", "nu", "nv", "Rs", "Rd"]
    out = []
    for _ in range(nlines):
        k = rng.integers(0, 100)
        if k < 12:
            line = "newmtl" + ("" if rng.integers(0, 8) == 0 else " " + ["red", "glass", "sky_light", "m%d" % rng.integers(0, 9)][rng.integers(0, 4)])
        elif k < 45:
            line = keys3[rng.integers(0, 3)] + sep(rng) + sep(rng).join(num(rng) for _ in range(int(rng.integers(1, 5))))
        elif k < 88:
            key = keys1[rng.integers(0, len(keys1))]
            vals = [num(rng) for _ in range(int(rng.integers(0, 3)))]
            if key in ("illum", "light") and vals and rng.integers(0, 2):
                vals[0] = "%d" % rng.integers(-2, 14)
            line = key + ("" if not vals else sep(rng) + sep(rng).join(vals))
        elif k < 92:
            line = "# Kd 1 1 1"
        elif k < 95:
            line = ["", "Kd", "d", "map_Kd tex.png", "Ke 1 1 1", "kd 1 1 1"][rng.integers(0, 6)]
        else:
            line = "  \t" + "Kd " + " ".join(num(rng) for _ in range(3)) + " "
        if rng.integers(0, 12) == 0:
            line += "\r"
        out.append(line)
    return "\n".join(out) + ("\n" if rng.integers(0, 2) else "")


def gen_lights(rng, nlines):
    """Random .lights text (LightParser.cpp:65-113)."""
    out = []
    for _ in range(nlines):
        k = rng.integers(0, 100)
        if k < 20:
            line = "newlight" + ("" if rng.integers(0, 8) == 0 else " l%d" % rng.integers(0, 9))
        elif k < 40:
            line = "type " + ("%d" % rng.integers(0, 4) if rng.integers(0, 4) else num(rng))
        elif k < 60:
            line = "pos" + sep(rng) + sep(rng).join(num(rng) for _ in range(int(rng.integers(2, 5))))
        elif k < 78:
            line = "rgb " + " ".join(num(rng) for _ in range(int(rng.integers(2, 5))))
        elif k < 92:
            line = "radius" + ("" if rng.integers(0, 6) == 0 else " " + num(rng))
        else:
            line = ["# radius 3", "", "ab", "power 3", "  radius 0.25  "][rng.integers(0, 5)]
        if rng.integers(0, 12) == 0:
            line += "\r"
        out.append(line)
    return "\n".join(out) + ("\n" if rng.integers(0, 2) else "")


def gen_bvh_scene(rng):
    """Random small scene for the builder comparison: 1-4 objects of 1-200 triangles, coordinates optionally snapped to
    a grid (tied centres, flat boxes), some degenerate triangles, and a random builder configuration.
    Returns (obj text, BVH keyword arguments, number of faces)."""
    nobj = int(rng.integers(1, 5)); lines = ["vn 0 0 1"]; nv = 0
    snap = [None, 0.5, 0.25, 0.1][rng.integers(0, 4)]
    for o in range(nobj):
        lines += ["o obj%d" % o, "usemtl m"]
        nf = int(rng.integers(1, [4, 30, 200][rng.integers(0, 3)]))
        for f in range(nf):
            tri = rng.uniform(-1, 1, 3) + rng.uniform(-0.3, 0.3, (3, 3))
            if snap: tri = np.round(tri / snap) * snap
            if rng.integers(0, 15) == 0: tri[2] = tri[1]
            for v in tri: lines.append("v %.6f %.6f %.6f" % tuple(v))
            lines.append("f %d//1 %d//1 %d//1" % (nv + 1, nv + 2, nv + 3)); nv += 3
    kw = dict(max_faces=int(rng.integers(1, 3)), skip_ahead=bool(rng.integers(0, 2)), skip_ahead_compare=float(np.round(rng.uniform(0.05, 1.0), 3)),
              sah_faces_limit=int([100000, 50, 8, 3][rng.integers(0, 4)]))
    if rng.integers(0, 4) == 0: kw["phong_tess"] = float(np.round(rng.uniform(0.1, 1.0), 2))
    return "\n".join(lines) + "\n", kw, nv // 3
