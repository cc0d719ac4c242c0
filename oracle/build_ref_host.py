"""Build the REFERENCE's own host classes for the tests: oracle/_ref/libref_host.so.

TEST INFRASTRUCTURE.  ObjParser, MtlParser, LightParser, ModelLoader, MathHelp, BVH, Cfg and Logger are
compiled from the sources where they lie under /root/reference/source -- nothing is copied -- against small
stand-ins for the third-party headers they include and this image does not have (oracle/ref_shim/host/:
boost::split / trim / posix_time / property_tree, glm::vec3, GL scalar typedefs, the OpenCL host vector types;
the vendored Khronos cl.hpp is skipped through its include guard).  oracle/ref_shim/host_driver.cpp adds C
entry points and the one step that lives in PathTracer.cpp (Qt + OpenCL, not compilable here): flattening the
BVH into the kernel's arrays.

The library is the yardstick for oracle/obj_oracle.cpp and oracle/bvh_oracle.cpp
(tests/test_oracle_vs_reference_host.py).  /root/reference does not exist on the GPU box; the library built
here travels with the snapshot.
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
OUT = os.path.join(HERE, "_ref")
REF_SRC = "/root/reference/source"
SOURCES = ["Cfg.cpp", "Logger.cpp", "MathHelp.cpp", "ObjParser.cpp", "MtlParser.cpp", "LightParser.cpp",
           "ModelLoader.cpp", "accelstructures/AccelStructure.cpp", "accelstructures/BVH.cpp",
           "Camera.cpp", "PathTracer.cpp"]
SO = os.path.join(OUT, "libref_host.so")


def reference_available():
    return all(os.path.isfile(os.path.join(REF_SRC, s)) for s in SOURCES)


def build(force=False, verbose=False):
    shim = os.path.join(HERE, "ref_shim", "host")
    driver = os.path.join(HERE, "ref_shim", "host_driver.cpp")
    fake_cl = os.path.join(HERE, "ref_shim", "fake_cl.cpp")
    deps = [driver, fake_cl] + [os.path.join(dp, f) for dp, _, fs in os.walk(shim) for f in fs]
    if os.path.isfile(SO) and not force and all(os.path.getmtime(d) <= os.path.getmtime(SO) for d in deps):
        return SO
    if not reference_available():
        if os.path.isfile(SO):
            return SO
        raise FileNotFoundError("reference sources not present (%s) and %s not prebuilt" % (REF_SRC, SO))
    os.makedirs(OUT, exist_ok=True)
    cmd = ["g++", "-O2", "-std=gnu++11", "-ffp-contract=off", "-fPIC", "-shared", "-pthread", "-w", "-Wno-narrowing",
           "-fpermissive", "-include", os.path.join(shim, "cl_types.h"), "-I", shim, "-I", REF_SRC, driver, fake_cl]
    cmd += [os.path.join(REF_SRC, s) for s in SOURCES] + ["-ldl", "-o", SO]
    if verbose:
        print(" ".join(cmd), flush=True)
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("reference host sources do not compile:\n%s" % r.stderr[-8000:])
    return SO


if __name__ == "__main__":
    print(build(force="-f" in sys.argv, verbose="-v" in sys.argv))
