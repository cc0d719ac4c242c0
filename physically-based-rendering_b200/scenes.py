"""Synthetic scenes for the BASELINE.json configurations.

Every generator is a pure function of its arguments: positions come from a counter-based integer
hash (lowbias32), not from a stateful RNG, so the same scene can be regenerated bit-for-bit anywhere.
A scene is a dict of flat numpy arrays in the shape the reference's ObjParser produces
(source/ObjParser.h:32-44): vertices / normals (float32, 3 per element), facesV / facesVN (uint32,
3 per face, zero-based), facesMtl (int32 per face), per-object face lists, materials, lights.
`write_obj` emits the same scene as OBJ + MTL (+ .lights) for the reference's own loader path.
"""
import os

import numpy as np

f32 = np.float32


def lowbias32(x):
    """Chris Wellons' lowbias32 integer hash on uint32 arrays."""
    x = np.asarray(x, np.uint64) & 0xffffffff
    x ^= x >> 16
    x = (x * 0x7feb352d) & 0xffffffff
    x ^= x >> 15
    x = (x * 0x846ca68b) & 0xffffffff
    x ^= x >> 16
    return x.astype(np.uint32)


def uniform(counter, stream, seed):
    """U[0,1) float32 (24-bit mantissa) for integer counters; `stream` separates coordinates."""
    c = np.asarray(counter, np.uint64)
    h = lowbias32((c * 16 + stream) ^ (np.uint64(seed) * np.uint64(0x9e3779b9) & np.uint64(0xffffffff)))
    h = lowbias32(h.astype(np.uint64) + np.uint64(seed))
    return ((h >> 8).astype(f32) * f32(1.0 / 16777216.0)).astype(f32)


def default_material(name="mat", **kw):
    """Row of the 24-float material table (see oracle/obj_oracle.cpp: oracle_obj_get what=9),
    initialised like MtlParser::getEmptyMaterial (MtlParser.cpp:11-36)."""
    m = np.zeros(24, f32)
    m[0:4] = (1, 1, 1, 0)      # Ka
    m[4:8] = (1, 1, 1, 0)      # Kd
    m[8:12] = (1, 1, 1, 0)     # Ks
    m[12], m[13], m[14], m[15], m[16] = 1.0, 1.0, 100.0, 2.0, 0.0   # d Ni Ns illum light
    m[17], m[18] = 1.0, 1.0    # rough p
    m[19], m[20], m[21], m[22] = 0.0, 0.0, 0.0, 1.0                  # nu nv Rs Rd
    idx = {"Kd": slice(4, 7), "Ks": slice(8, 11), "d": 12, "Ni": 13, "rough": 17, "p": 18,
           "nu": 19, "nv": 20, "Rs": 21, "Rd": 22, "light": 16}
    for k, v in kw.items():
        m[idx[k]] = v
    return name, m


def _finish(vertices, faces, face_mtl, object_counts, materials, normals=None, facesVN=None, lights=None,
            object_names=None):
    faces = np.ascontiguousarray(faces, np.uint32).reshape(-1)
    nf = faces.size // 3
    normals = np.zeros(0, f32) if normals is None else np.ascontiguousarray(normals, f32).reshape(-1)
    facesVN = np.zeros(0, np.uint32) if facesVN is None else np.ascontiguousarray(facesVN, np.uint32).reshape(-1)
    counts = np.asarray(object_counts, np.uint32)
    assert counts.sum() == nf
    has_n = facesVN.size == faces.size
    return {
        "vertices": np.ascontiguousarray(vertices, f32).reshape(-1),
        "normals": normals,
        "facesV": faces,
        "facesVN": facesVN,
        "facesMtl": np.ascontiguousarray(face_mtl, np.int32),
        "objFaceCounts": counts,
        "objFacesV": faces.copy(),
        "objFacesVN": facesVN.copy(),
        "objNormalFaceCounts": counts.copy() if has_n else np.zeros_like(counts),
        "materials": np.stack([m for _, m in materials]).astype(f32) if materials else np.zeros((0, 24), f32),
        "materialNames": [n for n, _ in materials],
        "lights": np.zeros((0, 10), f32) if lights is None else np.asarray(lights, f32).reshape(-1, 10),
        "lightNames": ["light%d" % i for i in range(0 if lights is None else len(lights))],
        "objectNames": object_names or ["object%d" % i for i in range(len(counts))],
        "shadowRaysForcedOff": False,
    }


def soup(num_triangles=1_000_000, seed=12345, extent=1.0, edge=0.02, kd=0.8):
    """BASELINE config 2: random triangle soup.  Centre ~ U[-extent, extent]^3, two edge vectors
    ~ U[-edge, edge]^3; one object, one diffuse material, white sky."""
    i = np.arange(num_triangles, dtype=np.uint64)
    u = [uniform(i, k, seed) for k in range(9)]
    two = f32(2.0)
    c = np.stack([(u[k] * two - f32(1.0)) * f32(extent) for k in range(3)], 1)
    e1 = np.stack([(u[k] * two - f32(1.0)) * f32(edge) for k in range(3, 6)], 1)
    e2 = np.stack([(u[k] * two - f32(1.0)) * f32(edge) for k in range(6, 9)], 1)
    v = np.empty((num_triangles, 3, 3), f32)
    v[:, 0] = c
    v[:, 1] = c + e1
    v[:, 2] = c + e2
    faces = np.arange(num_triangles * 3, dtype=np.uint32)
    mats = [default_material("soup", Kd=(kd, kd, kd))]
    return _finish(v.reshape(-1), faces, np.zeros(num_triangles, np.int32), [num_triangles], mats,
                   object_names=["soup"])


def _value_noise(x, z, seed, octaves=4):
    """Fixed-seed value noise on a float32 lattice (bilinear, smoothstep), summed over octaves."""
    h = np.zeros_like(x, dtype=f32)
    amp, freq = f32(1.0), f32(1.0)
    for o in range(octaves):
        fx, fz = x * freq, z * freq
        ix, iz = np.floor(fx), np.floor(fz)
        tx, tz = (fx - ix).astype(f32), (fz - iz).astype(f32)
        tx = tx * tx * (f32(3.0) - f32(2.0) * tx)
        tz = tz * tz * (f32(3.0) - f32(2.0) * tz)
        ixi, izi = ix.astype(np.int64) & 0xffff, iz.astype(np.int64) & 0xffff

        def lat(a, b):
            return uniform(((a & 0xffff) << 16 | (b & 0xffff)).astype(np.uint64), o, seed)
        v00, v10 = lat(ixi, izi), lat(ixi + 1, izi)
        v01, v11 = lat(ixi, izi + 1), lat(ixi + 1, izi + 1)
        top = v00 + (v10 - v00) * tx
        bot = v01 + (v11 - v01) * tx
        h += amp * (top + (bot - top) * tz)
        amp *= f32(0.5)
        freq *= f32(2.0)
    return h


def displaced_grid(cells_x=2237, cells_z=2236, patches=8, seed=777, size=2.0, height=0.25):
    """BASELINE config 4: height-field grid, 2 triangles per cell (2237 x 2236 -> 10 003 864),
    split into patches x patches `o` objects so that the per-object tree path is exercised."""
    nx, nz = cells_x + 1, cells_z + 1
    gx = (np.arange(nx, dtype=f32) / f32(cells_x) - f32(0.5)) * f32(size)
    gz = (np.arange(nz, dtype=f32) / f32(cells_z) - f32(0.5)) * f32(size)
    X, Z = np.meshgrid(gx, gz, indexing="xy")           # [nz, nx]
    Y = (_value_noise(X * f32(4.0), Z * f32(4.0), seed) * f32(height)).astype(f32)
    vertices = np.stack([X, Y, Z], -1).reshape(-1, 3)

    cx, cz = np.meshgrid(np.arange(cells_x), np.arange(cells_z), indexing="xy")   # [cells_z, cells_x]
    v00 = (cz * nx + cx).astype(np.uint32)
    v10, v01, v11 = v00 + 1, v00 + nx, v00 + nx + 1
    tri = np.stack([np.stack([v00, v01, v10], -1), np.stack([v10, v01, v11], -1)], 2)   # [cz, cx, 2, 3]
    # order faces patch by patch (each patch is one `o` object)
    px = np.minimum(cx * patches // cells_x, patches - 1)
    pz = np.minimum(cz * patches // cells_z, patches - 1)
    pid = (pz * patches + px).reshape(-1)
    order = np.argsort(pid, kind="stable")
    faces = tri.reshape(-1, 2, 3)[order].reshape(-1, 3)
    counts = np.bincount(pid, minlength=patches * patches).astype(np.uint32) * 2
    mats = [default_material("ground", Kd=(0.7, 0.7, 0.7))]
    return _finish(vertices.reshape(-1), faces.reshape(-1), np.zeros(faces.shape[0], np.int32), counts, mats,
                   object_names=["patch%02d" % i for i in range(patches * patches)])


class _Mesh:
    """Accumulates indexed triangles with per-vertex normals, object by object."""

    def __init__(self):
        self.v, self.n, self.f, self.m, self.counts, self.names = [], [], [], [], [], []
        self.nv = 0

    def add(self, name, verts, normals, faces, mtl):
        verts = np.asarray(verts, f32).reshape(-1, 3)
        normals = np.asarray(normals, f32).reshape(-1, 3)
        faces = np.asarray(faces, np.uint32).reshape(-1, 3)
        assert len(verts) == len(normals)
        self.v.append(verts)
        self.n.append(normals)
        self.f.append(faces + np.uint32(self.nv))
        mtl = np.asarray(mtl, np.int32)
        self.m.append(np.broadcast_to(mtl, (len(faces),)).astype(np.int32))
        self.counts.append(len(faces))
        self.names.append(name)
        self.nv += len(verts)

    def grid(self, name, origin, du, dv, nu, nv, normal, mtl):
        """nu x nv quads spanning origin + s*du + t*dv; mtl may be a function (i, j) -> index."""
        s, t = np.meshgrid(np.arange(nu + 1, dtype=f32) / f32(nu), np.arange(nv + 1, dtype=f32) / f32(nv), indexing="xy")
        P = (np.asarray(origin, f32)[None, None] + s[..., None] * np.asarray(du, f32)[None, None] +
             t[..., None] * np.asarray(dv, f32)[None, None]).astype(f32)
        i, j = np.meshgrid(np.arange(nu), np.arange(nv), indexing="xy")
        a = (j * (nu + 1) + i).reshape(-1)
        faces = np.stack([np.stack([a, a + 1, a + nu + 2], -1), np.stack([a, a + nu + 2, a + nu + 1], -1)], 1).reshape(-1, 3)
        m = mtl if not callable(mtl) else np.repeat(mtl(i.reshape(-1), j.reshape(-1)), 2)
        self.add(name, P.reshape(-1, 3), np.broadcast_to(np.asarray(normal, f32), (P.size // 3, 3)), faces, m)

    def revolve(self, name, centre, radius_fn, y0, y1, segments, rings, mtl):
        """Surface of revolution around the y axis through `centre`: radius_fn(t in [0,1]) -> radius."""
        t = np.arange(rings + 1, dtype=f32) / f32(rings)
        ang = np.arange(segments + 1, dtype=f32) * f32(2.0 * np.pi / segments)
        r = radius_fn(t).astype(f32)
        y = (f32(y0) + t * f32(y1 - y0)).astype(f32)
        ca, sa = np.cos(ang).astype(f32), np.sin(ang).astype(f32)
        P = np.stack([centre[0] + r[:, None] * ca[None], np.broadcast_to(y[:, None], (rings + 1, segments + 1)),
                      centre[2] + r[:, None] * sa[None]], -1).astype(f32)
        dr = np.gradient(r.astype(np.float64), y.astype(np.float64))
        N = np.stack([ca[None] * np.ones_like(dr)[:, None], np.broadcast_to(-dr[:, None], (rings + 1, segments + 1)),
                      sa[None] * np.ones_like(dr)[:, None]], -1)
        N = (N / np.linalg.norm(N, axis=-1, keepdims=True)).astype(f32)
        i, j = np.meshgrid(np.arange(segments), np.arange(rings), indexing="xy")
        a = (j * (segments + 1) + i).reshape(-1)
        faces = np.stack([np.stack([a, a + segments + 2, a + 1], -1), np.stack([a, a + segments + 1, a + segments + 2], -1)], 1)
        self.add(name, P.reshape(-1, 3), N.reshape(-1, 3), faces.reshape(-1, 3), mtl)

    def arch(self, name, a, b, height, tube, segments, rings, mtl):
        """Half torus from point a to point b (same y), rising by `height`."""
        a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
        mid, half = (a + b) / 2, (b - a) / 2
        R = np.linalg.norm(half)
        ex = half / R
        up = np.array([0.0, 1.0, 0.0])
        ez = np.cross(ex, up)
        th = np.arange(segments + 1) * (np.pi / segments)
        ph = np.arange(rings + 1) * (2 * np.pi / rings)
        cth, sth = np.cos(th), np.sin(th) * (height / R)
        centre = mid[None] - cth[:, None] * R * ex[None] + sth[:, None] * R * up[None]
        radial = -cth[:, None] * ex[None] + (np.sin(th))[:, None] * up[None]
        P = centre[:, None] + tube * (np.cos(ph)[None, :, None] * radial[:, None] + np.sin(ph)[None, :, None] * ez[None, None])
        N = np.cos(ph)[None, :, None] * radial[:, None] + np.sin(ph)[None, :, None] * ez[None, None]
        N = N / np.linalg.norm(N, axis=-1, keepdims=True)
        i, j = np.meshgrid(np.arange(rings), np.arange(segments), indexing="xy")
        q = (j * (rings + 1) + i).reshape(-1)
        faces = np.stack([np.stack([q, q + 1, q + rings + 2], -1), np.stack([q, q + rings + 2, q + rings + 1], -1)], 1)
        self.add(name, P.reshape(-1, 3).astype(f32), N.reshape(-1, 3).astype(f32), faces.reshape(-1, 3), mtl)

    def finish(self, materials, lights=None):
        v = np.concatenate(self.v)
        n = np.concatenate(self.n)
        f = np.concatenate(self.f)
        return _finish(v.reshape(-1), f.reshape(-1), np.concatenate(self.m), self.counts, materials,
                       normals=n.reshape(-1), facesVN=f.reshape(-1), lights=lights, object_names=self.names)


def interior(detail=1.0, with_light=True):
    """BASELINE config 3: procedural interior (~260 k triangles at detail = 1): a room with an open
    ceiling, a tiled floor, two rows of fluted columns joined by arches, two spheres; 8 materials, diffuse
    (nu = nv = 0, Rs = 0, Rd = 1 / rough = 1) and glossy-to-specular (nu = nv in {100, 1000}, Rs 0.8 /
    rough in {0, 0.3}).  Vertex normals are written, faces are `v//vn`."""
    d = float(detail)

    def k(x):
        return max(2, int(round(x * d)))
    mats = [
        default_material("floor_light", Kd=(0.75, 0.72, 0.68), rough=1.0),
        default_material("floor_dark", Kd=(0.25, 0.24, 0.22), nu=100.0, nv=100.0, Rs=0.8, Rd=0.6, rough=0.3),
        default_material("wall", Kd=(0.8, 0.78, 0.7), rough=1.0),
        default_material("wall_accent", Kd=(0.55, 0.2, 0.15), rough=1.0),
        default_material("column", Kd=(0.85, 0.85, 0.8), nu=100.0, nv=100.0, Rs=0.3, Rd=0.9, rough=0.3),
        default_material("arch", Kd=(0.7, 0.6, 0.4), nu=1000.0, nv=1000.0, Rs=0.8, Rd=0.5, rough=0.0),
        default_material("mirror", Kd=(0.95, 0.95, 0.95), Ks=(1.0, 1.0, 1.0), nu=1000.0, nv=1000.0, Rs=0.8, Rd=0.1, rough=0.0),
        default_material("glass", Kd=(0.9, 0.95, 1.0), d=0.3, Ni=1.5, nu=1000.0, nv=1000.0, Rs=0.8, Rd=0.2, rough=0.0),
    ]
    M = _Mesh()
    X, Z, Hh = 4.0, 6.0, 3.2
    nt = k(100)
    M.grid("floor", (-X, 0, -Z), (2 * X, 0, 0), (0, 0, 2 * Z), nt, nt, (0, 1, 0), lambda i, j: ((i + j) & 1).astype(np.int32))
    wq = (k(60), k(30))
    M.grid("wall_back", (-X, 0, -Z), (2 * X, 0, 0), (0, Hh, 0), wq[0], wq[1], (0, 0, 1), 2)
    M.grid("wall_front", (X, 0, Z), (-2 * X, 0, 0), (0, Hh, 0), wq[0], wq[1], (0, 0, -1), 2)
    M.grid("wall_left", (-X, 0, Z), (0, 0, -2 * Z), (0, Hh, 0), wq[0], wq[1], (1, 0, 0), lambda i, j: np.where(j < wq[1] // 4, 3, 2).astype(np.int32))
    M.grid("wall_right", (X, 0, -Z), (0, 0, 2 * Z), (0, Hh, 0), wq[0], wq[1], (-1, 0, 0), lambda i, j: np.where(j < wq[1] // 4, 3, 2).astype(np.int32))
    seg, rings = k(96), k(80)

    def flute(t):
        # base, entasis, capital
        r = 0.22 - 0.04 * t + 0.06 * np.exp(-((t - 0.02) / 0.03) ** 2) + 0.07 * np.exp(-((t - 0.98) / 0.03) ** 2)
        return r
    zs = np.linspace(-Z + 1.0, Z - 1.0, 6)
    for side, x in enumerate((-2.2, 2.2)):
        for ci, z in enumerate(zs):
            M.revolve("column_%d_%d" % (side, ci), (x, 0.0, z), flute, 0.0, 2.4, seg, rings, 4)
        for ci in range(len(zs) - 1):
            M.arch("arch_%d_%d" % (side, ci), (x, 2.4, zs[ci]), (x, 2.4, zs[ci + 1]), 0.6, 0.09, k(48), k(24), 5)

    def ball(t):
        return 0.55 * np.sqrt(np.maximum(1e-6, 1.0 - (2 * t - 1) ** 2))
    M.revolve("sphere_mirror", (-0.8, 0.0, -1.5), ball, 0.0, 1.1, k(64), k(48), 6)
    M.revolve("sphere_glass", (0.9, 0.0, 0.6), ball, 0.0, 1.1, k(64), k(48), 7)
    lights = np.array([[1, 0.0, 3.0, 0.0, 0.0, 1.0, 0.95, 0.9, 0.0, 0.0]], f32) if with_light else None
    return M.finish(mats, lights)


def write_obj(scene, path):
    """Write <path>.obj / .mtl (/.lights) in the dialect the reference's parsers read: one `o` per
    object, `usemtl` on material change, `f v v v` or `f v//vn v//vn v//vn`, single blanks."""
    base, _ = os.path.splitext(path)
    v = scene["vertices"].reshape(-1, 3)
    n = scene["normals"].reshape(-1, 3)
    f = scene["facesV"].reshape(-1, 3) + 1
    has_n = scene["facesVN"].size == scene["facesV"].size
    fn = scene["facesVN"].reshape(-1, 3) + 1 if has_n else None
    with open(base + ".obj", "w") as fh:
        fh.write("# generated by physically-based-rendering_b200/scenes.py\n")
        for row in v:
            fh.write("v %.9g %.9g %.9g\n" % tuple(row))
        for row in n:
            fh.write("vn %.9g %.9g %.9g\n" % tuple(row))
        k = 0
        for oi, cnt in enumerate(scene["objFaceCounts"]):
            fh.write("o %s\n" % scene["objectNames"][oi])
            cur = None
            for j in range(k, k + int(cnt)):
                m = int(scene["facesMtl"][j])
                if m != cur and m >= 0:
                    fh.write("usemtl %s\n" % scene["materialNames"][m])
                    cur = m
                if has_n:
                    fh.write("f %d//%d %d//%d %d//%d\n" % (f[j, 0], fn[j, 0], f[j, 1], fn[j, 1], f[j, 2], fn[j, 2]))
                else:
                    fh.write("f %d %d %d\n" % tuple(f[j]))
            k += int(cnt)
    with open(base + ".mtl", "w") as fh:
        for name, m in zip(scene["materialNames"], scene["materials"]):
            fh.write("newmtl %s\nKd %.9g %.9g %.9g\nKs %.9g %.9g %.9g\nd %.9g\nNi %.9g\n" % (
                name, m[4], m[5], m[6], m[8], m[9], m[10], m[12], m[13]))
            fh.write("rough %.9g\np %.9g\nnu %.9g\nnv %.9g\nRs %.9g\nRd %.9g\nlight %d\n\n" % (
                m[17], m[18], m[19], m[20], m[21], m[22], int(m[16])))
    if len(scene["lights"]):
        with open(base + ".lights", "w") as fh:
            for name, li in zip(scene["lightNames"], scene["lights"]):
                fh.write("newlight %s\ntype %d\npos %.9g %.9g %.9g\nrgb %.9g %.9g %.9g\nradius %.9g\n" % (
                    name, int(li[0]), li[1], li[2], li[3], li[5], li[6], li[7], li[9]))
    return base + ".obj"
