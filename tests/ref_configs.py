"""The configurations for which the reference's own kernel is compiled (oracle/build_ref.py) -- one shared
library per configuration, exactly like one OpenCL program per configuration upstream.  Listed in one place
so that `python oracle/build_ref.py` (and __graft_entry__.build()) can prebuild them where /root/reference
exists; on the GPU box only the prebuilt libraries are there."""
import os
import sys

_HERE = os.path.dirname(os.path.abspath(__file__))
_ROOT = os.path.dirname(_HERE)
for p in (_ROOT, _HERE):
    if p not in sys.path:
        sys.path.insert(0, p)

# name -> (scene, Prepared kwargs).  scene: a model in tests/golden/models or ("soup", triangles, seed).
CASES = {
    "smoke": ("suzanne.obj", dict(width=64, height=64, max_depth=4)),          # __graft_entry__.smoke()
    "suzanne_sa": ("suzanne.obj", dict(width=96, height=64, brdf=1, max_depth=4)),
    "suzanne_schlick": ("suzanne.obj", dict(width=96, height=64, brdf=0, max_depth=4)),
    "suzanne_sa_shadow": ("suzanne.obj", dict(width=80, height=64, brdf=1, shadow_rays=1, max_depth=3)),
    "suzanne_schlick_shadow_ms": ("suzanne.obj", dict(width=64, height=48, brdf=0, shadow_rays=1, samples=3, max_depth=3)),
    "suzanne_dof_ms": ("suzanne.obj", dict(width=64, height=64, brdf=1, samples=2, max_depth=3, focus_point=(32, 20))),
    "suzanne_phong_sa": ("suzanne.obj", dict(width=80, height=64, brdf=1, max_depth=3, phong_tessellation=0.7)),
    "suzanne_phong_schlick_shadow": ("suzanne.obj", dict(width=72, height=56, brdf=0, shadow_rays=1, max_depth=3,
                                                         phong_tessellation=0.7)),
    "pillars_sa": ("pillars.obj", dict(width=96, height=64, brdf=1, max_depth=5, max_added_depth=3)),
    "pillars_schlick": ("pillars.obj", dict(width=96, height=64, brdf=0, max_depth=5, max_added_depth=3)),
    # (tests/golden/models/quirks.obj is not here: faces without `usemtl` carry material index (uint)-1, which
    #  the reference kernel reads out of bounds -- undefined upstream, defined in the oracle and on the device)
    "suzanne_sa_close": ("suzanne.obj", dict(width=64, height=48, brdf=1, shadow_rays=1, max_depth=6, max_added_depth=2,
                                             eye=(0.3, 0.8, 1.6), center=(0.2, 0.3, 1.0), antialiasing=0.333)),
    "soup_sa": (("soup", 20000, 5), dict(width=96, height=64, brdf=1, eye=(0.0, 0.0, 3.5))),
    # tests/material_scene.py: partial transparency, anisotropic lobes, rough / isotropy in (0, 1), an orb light
    "materials_sa": ("@materials", dict(width=96, height=64, brdf=1, shadow_rays=1, max_depth=5, max_added_depth=3,
                                        eye=(0.3, 1.1, 3.2), center=(-0.05, 0.25, 1.0))),
    "materials_sa_opaque": ("@materials_opaque", dict(width=96, height=64, brdf=1, shadow_rays=1, max_depth=5, max_added_depth=3,
                                                      eye=(0.3, 1.1, 3.2), center=(-0.05, 0.25, 1.0))),
    "materials_schlick": ("@materials", dict(width=96, height=64, brdf=0, shadow_rays=1, max_depth=5, max_added_depth=3,
                                             eye=(0.3, 1.1, 3.2), center=(-0.05, 0.25, 1.0))),
    "materials_sa_ms_noshadow": ("@materials_opaque", dict(width=64, height=48, brdf=1, samples=3, max_depth=6, max_added_depth=4,
                                                    eye=(-0.8, 0.7, 2.6), center=(0.2, 0.15, 1.0))),
    "materials_schlick_phong": ("@materials", dict(width=64, height=48, brdf=0, max_depth=4, phong_tessellation=0.6,
                                                   eye=(0.3, 1.1, 3.2), center=(-0.05, 0.25, 1.0))),
}

_SCENES = {}


def scene_of(spec):
    from oracle import oracle as O
    import helpers as Hh
    if spec not in _SCENES:
        if isinstance(spec, tuple):
            import pbr_b200
            _SCENES[spec] = pbr_b200.scenes.soup(spec[1], seed=spec[2])
        elif spec in ("@materials", "@materials_opaque"):
            import tempfile
            import material_scene
            with tempfile.TemporaryDirectory() as d:
                _SCENES[spec] = O.load_obj(material_scene.write(d, transparent=(spec == "@materials")), 1)
        else:
            _SCENES[spec] = O.load_obj(Hh.model_path(spec), 1)      # with the .lights; used when shadow_rays = 1
    return _SCENES[spec]


def prepared(name):
    import helpers as Hh
    spec, kw = CASES[name]
    return Hh.Prepared(scene_of(spec), **kw)


def bench_c2_values():
    """The configuration bench.py's reference arm runs (C2: 1M-triangle soup, 1920x1080)."""
    import bench
    return bench.reference_program_values()


def renderer_program_values():
    """The configurations of the end-to-end renderer tests (tests/test_gpu_host.py, test_oracle_vs_reference_host.py):
    the values are whatever the reference's own host code hands to CL::setValues."""
    from oracle import ref_host as RH
    import helpers as Hh
    import test_gpu_host as G
    import test_oracle_vs_reference_host as T
    path = Hh.model_path("suzanne.obj")
    for table, (w, h) in ((G.REF_RENDER, (G.REF_W, G.REF_H)), (T.RENDER_CONFIGS, (88, 56))):
        for name in sorted(table):
            kw = dict(table[name])
            kw.setdefault("max_depth", 4)
            r = RH.Renderer(path, width=w, height=h, **kw)      # builds the kernel library as a side effect
            yield r.values
            r.close()


def static_program_values():
    """Configurations whose program values are known without running the reference's host code."""
    from oracle import ref as R
    for name in CASES:
        yield R.values_from_defines(prepared(name).defines)
    if os.environ.get("PBR_REF_SKIP_BENCH") != "1":
        yield bench_c2_values()
    import test_random_parity as T
    for brdf in (1, 0):
        for seed in range(T.TRIALS):
            yield R.values_from_defines(T._trial(seed, brdf).defines)
    import config_cases                                            # C1 at 512x512, C3 at 1920x1080 (both BRDFs)
    for v in config_cases.program_values():
        yield v


def all_program_values():
    from oracle import ref_host as RH
    for v in static_program_values():
        yield v
    if RH.available():
        for v in renderer_program_values():
            yield v
