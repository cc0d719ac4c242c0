"""Build the REFERENCE's own kernel for the host CPU: oracle/_ref/pt_ref_<key>.so.

TEST INFRASTRUCTURE.  The reference compiles its OpenCL program at run time, once per configuration, from
text it assembles itself: CL::combineParts splices the `#FILE:name:FILE#` includes (CL.cpp:107-127),
CL::setValues substitutes the `#NAME#` placeholders with values printed from config.json and from the scene
(CL.cpp:626-705, PathTracer.cpp:210, 338, 470-515).  This script does the same two steps with the sources where
they lie under /root/reference/source/opencl, and compiles the result as C++ behind oracle/ref_shim/cl_compat.h
instead of handing it to an OpenCL driver.  Two mechanical adaptations of the text, nothing else:

  * OpenCL vector literals `(float4)( a, b, c, d )` are a C cast of a comma expression in C++, so they are
    rewritten to constructor calls `float4( a, b, c, d )`;
  * the kernel writes through `const Scene*` (scene->debugColor, pt_bvh.cl:17, 85), which OpenCL compilers
    let pass and C++ does not: the program text is compiled with `const` defined away.

Everything is written to oracle/_ref/ (git-ignored): no reference source enters the repository.  One shared
library per configuration, like one cl_program per configuration upstream; built on demand and cached.
/root/reference does not exist on the GPU box -- there only the libraries built here are available.

    python oracle/build_ref.py            # prebuild the configurations the tests use
"""
import hashlib
import os
import re
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
OUT = os.path.join(HERE, "_ref")
REF_CL_DIR = "/root/reference/source/opencl"
MAIN = "pathtracing.cl"

INT_KEYS = ["ACCEL_STRUCT", "BRDF", "IMG_HEIGHT", "IMG_WIDTH", "SHADOW_RAYS", "MAX_DEPTH", "MAX_ADDED_DEPTH",
            "PHONGTESS", "SAMPLES"]                      # CL.cpp:641-649, "%u"
FLOAT_KEYS = ["ANTI_ALIASING", "PHONGTESS_ALPHA"]        # CL.cpp:676-677, "%ff"
STRING_KEYS = ["BVH_NUM_NODES", "NUM_LIGHTS", "SKY_LIGHT"]   # CL::setReplacement


def reference_available():
    return os.path.isfile(os.path.join(REF_CL_DIR, MAIN))


def c_format(fmt, v):
    """printf-style formatting as the reference's snprintf does it."""
    return fmt % v


def program_values(img_width, img_height, bvh_num_nodes, num_lights, sky_light=(1.0, 1.0, 1.0), brdf=1, samples=1,
                   max_depth=3, max_added_depth=5, shadow_rays=0, antialiasing=0.7, phong_tessellation=0.0):
    """The placeholder -> text map CL::setValues builds (phongtess = alpha > 0, CL.cpp:660)."""
    v = {
        "ACCEL_STRUCT": "%u" % 0, "BRDF": "%u" % brdf, "IMG_HEIGHT": "%u" % img_height, "IMG_WIDTH": "%u" % img_width,
        "SHADOW_RAYS": "%u" % shadow_rays, "MAX_DEPTH": "%u" % max_depth, "MAX_ADDED_DEPTH": "%u" % max_added_depth,
        "PHONGTESS": "%u" % (1 if phong_tessellation > 0.0 else 0), "SAMPLES": "%u" % samples,
        "ANTI_ALIASING": "%ff" % antialiasing, "PHONGTESS_ALPHA": "%ff" % phong_tessellation,
        "BVH_NUM_NODES": "%u" % bvh_num_nodes, "NUM_LIGHTS": "%u" % num_lights,
        "SKY_LIGHT": "(float4)( %f, %f, %f, 0.0f )" % tuple(float(c) for c in sky_light[:3]),
    }
    return v


def assemble(values):
    """combineParts + setValues."""
    def load(name):
        with open(os.path.join(REF_CL_DIR, name)) as f:
            return f.read()
    text = load(MAIN)
    while True:
        a, b = text.find("#FILE:"), text.find(":FILE#")
        if a < 0 or b < 0:
            break
        text = text[:a] + load(text[a + 6:b]) + text[b + 6:]
    for k, v in values.items():
        text = text.replace("#%s#" % k, v, 1)             # first occurrence only, like string::replace upstream
    return text


def adapt(text):
    return re.sub(r"\(\s*(float[2348]|int[23]|uint4)\s*\)\s*\(", r"\1(", text)


def _shim_digest():
    """The host-side code a library is built with: a change of it invalidates every cached library."""
    h = hashlib.sha1()
    for path in (os.path.join(HERE, "ref_shim", "cl_compat.h"), os.path.join(HERE, "ref_shim", "ref_driver.inc"),
                 os.path.join(ROOT, "include", "pbr_pinned_math.h")):
        with open(path, "rb") as f:
            h.update(f.read())
    return h.hexdigest()


def variant_key(values):
    blob = ";".join("%s=%s" % kv for kv in sorted(values.items()))
    return hashlib.sha1((blob + "|" + _shim_digest()).encode()).hexdigest()[:16], blob


def library_path(values):
    key, _ = variant_key(values)
    return os.path.join(OUT, "pt_ref_%s.so" % key)


def build(values, force=False, verbose=False):
    """Returns the path of the shared library for this configuration, building it if necessary."""
    so = library_path(values)
    if os.path.isfile(so) and not force:
        return so
    if not reference_available():
        raise FileNotFoundError("reference sources not present (%s) and %s not prebuilt" % (REF_CL_DIR, so))
    os.makedirs(OUT, exist_ok=True)
    key, blob = variant_key(values)
    cpp = os.path.join(OUT, "pt_ref_%s.cpp" % key)
    with open(os.path.join(HERE, "ref_shim", "ref_driver.inc")) as f:
        driver = f.read()
    body = adapt(assemble(values))
    with open(cpp, "w") as f:
        f.write('#include "cl_compat.h"\n')
        f.write('#define REF_DEFINES_STRING "%s"\n' % blob.replace("\\", "\\\\").replace('"', '\\"'))
        f.write("namespace clref {\n#define const\n")
        f.write(body)
        f.write("\n#undef const\n} /* namespace clref */\n")
        f.write(driver)
    cmd = ["g++", "-O2", "-std=c++17", "-ffp-contract=off", "-fPIC", "-shared", "-pthread", "-w",
           "-I", os.path.join(ROOT, "include"), "-I", os.path.join(HERE, "ref_shim"), cpp, "-o", so]
    if verbose:
        print(" ".join(cmd), flush=True)
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("reference kernel does not compile for %s:\n%s" % (blob, r.stderr[-6000:]))
    if os.environ.get("PBR_REF_KEEP_SOURCE") != "1":
        os.remove(cpp)                                   # the assembled reference text is not kept around
    return so


def main():
    """Prebuild every configuration the tests and bench.py use; the g++ runs go side by side."""
    from concurrent.futures import ThreadPoolExecutor
    for p in (ROOT, os.path.join(ROOT, "tests")):
        if p not in sys.path:
            sys.path.insert(0, p)
    import ref_configs
    verbose = "-v" in sys.argv
    todo = {}
    for values in ref_configs.static_program_values():
        so = library_path(values)
        if not os.path.isfile(so):
            todo[so] = values                                     # two cases may share one configuration
    todo = list(todo.values())
    with ThreadPoolExecutor(max_workers=max(1, min(8, os.cpu_count() or 1))) as pool:
        list(pool.map(lambda v: build(v, verbose=verbose), todo))
    n = sum(1 for _ in ref_configs.all_program_values())          # the rest (renderer cases) build as they are created
    print("oracle/_ref: %d configuration(s) of the reference kernel built" % n)


if __name__ == "__main__":
    main()
