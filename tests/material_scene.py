"""A small scene whose .mtl walks through the BRDF parameter space the bundled models leave out: partial
transparency (the `d <= rand()` coin and refract()'s Fresnel coin), anisotropic Shirley-Ashikhmin lobes
(nu != nv, all four quadrant branches), Schlick with rough in (0, 1) and isotropy p < 1 (the phi += pi/2
branch), a near-mirror, Rs / Rd mixes, plus an orb light for shadow rays and light hits.  Written on the fly
(deterministic text), parsed by the reference's own parsers, the oracle's and the product's."""
import os

_BOX = [(0, 1, 2), (0, 2, 3), (4, 6, 5), (4, 7, 6), (0, 4, 5), (0, 5, 1), (1, 5, 6), (1, 6, 2), (2, 6, 7), (2, 7, 3), (3, 7, 4), (3, 4, 0)]

MATERIALS = """
newmtl floor
Kd 0.7 0.7 0.65
Ks 0.2 0.2 0.2
d 1.0
Ni 1.0
rough 1.0
p 1.0
nu 0.0
nv 0.0
Rs 0.05
Rd 1.0

newmtl wall
Kd 0.35 0.5 0.8
Ks 0.5 0.5 0.5
d 1.0
Ni 1.0
rough 0.6
p 0.3
nu 10.0
nv 1000.0
Rs 0.4
Rd 0.6

newmtl frosted
Kd 0.9 0.95 0.9
Ks 0.9 0.9 0.9
d 0.5
Ni 1.33
rough 0.4
p 0.7
nu 40.0
nv 40.0
Rs 0.3
Rd 0.7

newmtl brushed
Kd 0.8 0.6 0.3
Ks 1.0 0.9 0.7
d 1.0
Ni 1.0
rough 0.15
p 0.1
nu 2000.0
nv 5.0
Rs 0.8
Rd 0.2

newmtl mirrorish
Kd 0.95 0.95 0.95
Ks 1.0 1.0 1.0
d 1.0
Ni 1.0
rough 0.02
p 1.0
nu 30000.0
nv 30000.0
Rs 0.95
Rd 0.05

newmtl glass
Kd 1.0 1.0 1.0
Ks 1.0 1.0 1.0
d 0.1
Ni 1.5
rough 0.0
p 1.0
nu 100000.0
nv 100000.0
Rs 0.9
Rd 0.1

newmtl sky_light
Kd 0.9 0.85 0.8
"""

LIGHTS = """newlight key
type 2
pos 0.4 1.6 1.4
rgb 1.0 0.95 0.9
radius 0.2
"""


def write(directory, name="materials", transparent=True):
    """Write <name>.obj / .mtl / .lights into `directory`; returns the .obj path.  transparent=False makes the two
    see-through materials opaque: under BRDF 1 the reference turns most transparent paths into NaN pixels, which
    compare equal but say little."""
    verts, normals, out = [], [], []

    def box(x0, y0, z0, x1, y1, z1):
        base = len(verts)
        verts.extend([(x0, y0, z0), (x1, y0, z0), (x1, y1, z0), (x0, y1, z0), (x0, y0, z1), (x1, y0, z1), (x1, y1, z1), (x0, y1, z1)])
        return [(base + a + 1, base + b + 1, base + c + 1) for a, b, c in _BOX]

    normals.extend([(0, 1, 0), (0, 0, 1), (1, 0, 0), (0.6, 0.8, 0.0), (0.0, 0.6, 0.8)])
    objects = [
        ("floor", "floor", box(-2.0, -0.1, -2.0, 2.0, 0.0, 3.0)),
        ("back", "wall", box(-2.0, 0.0, -2.1, 2.0, 2.2, -2.0)),
        ("side", "wall", box(-2.1, 0.0, -2.0, -2.0, 2.2, 3.0)),
        ("frosted", "frosted", box(-1.2, 0.0, -0.6, -0.5, 0.9, 0.1)),
        ("brushed", "brushed", box(-0.3, 0.0, -1.0, 0.4, 1.3, -0.3)),
        ("mirror", "mirrorish", box(0.7, 0.0, -0.7, 1.5, 0.7, 0.1)),
        ("glass", "glass", box(-0.2, 0.0, 0.4, 0.5, 0.6, 1.0)),
    ]
    out.append("# written by tests/material_scene.py")
    for v in verts:
        out.append("v %.6f %.6f %.6f" % v)
    for n in normals:
        out.append("vn %.6f %.6f %.6f" % n)
    k = 0
    for oname, mtl, faces in objects:
        out.append("o %s" % oname)
        out.append("usemtl %s" % mtl)
        for a, b, c in faces:
            # vertex normals that differ within a face on some objects: exercises the Phong-tessellation branch
            n0, n1, n2 = (1 + k % 5, 1 + (k + 1) % 5, 1 + (k + 2) % 5) if oname in ("brushed", "mirror") else (1 + k % 5,) * 3
            out.append("f %d//%d %d//%d %d//%d" % (a, n0, b, n1, c, n2))
            k += 1
    os.makedirs(directory, exist_ok=True)
    base = os.path.join(directory, name)
    with open(base + ".obj", "w") as f:
        f.write("\n".join(out) + "\n")
    mtl = MATERIALS.lstrip("\n")
    if not transparent:
        mtl = mtl.replace("d 0.5\nNi 1.33", "d 1.0\nNi 1.33").replace("d 0.1\nNi 1.5", "d 1.0\nNi 1.5")
    with open(base + ".mtl", "w") as f:
        f.write(mtl)
    with open(base + ".lights", "w") as f:
        f.write(LIGHTS)
    return base + ".obj"
