"""CPU: the C-ABI library loads and exports every symbol include/pbr_b200.h declares."""
import ctypes
import os
import re

from conftest import ROOT


def _declared():
    text = open(os.path.join(ROOT, "include", "pbr_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(pbr_[a-z0-9_]+)\s*\(", text)))


def test_header_declares_the_reference_surface():
    names = _declared()
    # one entry point per CL method PathTracer uses
    for need in ("pbr_create", "pbr_destroy", "pbr_buffer_create", "pbr_image_create", "pbr_image_write",
                 "pbr_image_read", "pbr_set_define", "pbr_program_load", "pbr_kernel_get", "pbr_kernel_set_arg",
                 "pbr_kernel_launch", "pbr_finish", "pbr_kernel_time_ms", "pbr_trace"):
        assert need in names


def test_library_exports_every_declared_symbol():
    import pbr_b200
    assert os.path.exists(pbr_b200.capi.LIB_PATH), "libpbr_b200.so not built: run __graft_entry__.build()"
    lib = ctypes.CDLL(pbr_b200.capi.LIB_PATH)
    missing = [n for n in _declared() if not hasattr(lib, n)]
    assert not missing, missing
    assert sorted(pbr_b200.capi.SYMBOLS) == _declared()


def test_host_library_exports_every_declared_symbol():
    """include/pbr_host.h (the flat C view of the host mirror: Cfg, loaders, BVH, PathTracer, Camera) <-> libpbr_host.so."""
    import pbr_b200
    text = open(os.path.join(ROOT, "include", "pbr_host.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    declared = sorted(set(re.findall(r"\b(pbrh_[a-z0-9_]+)\s*\(", text)))
    assert len(declared) >= 40
    assert os.path.exists(pbr_b200.host.LIB_PATH), "libpbr_host.so not built: run __graft_entry__.build()"
    lib = ctypes.CDLL(pbr_b200.host.LIB_PATH)
    missing = [n for n in declared if not hasattr(lib, n)]
    assert not missing, missing


def test_headers_are_plain_c_and_link(tmp_path):
    """The boundary is a C ABI: both public headers compile as strict C99 and a C program links against the
    library and gets an error code -- not a crash, not a CPU path -- from pbr_create when there is no device."""
    import shutil
    import subprocess
    import pbr_b200
    if not shutil.which("gcc"):
        return
    src = tmp_path / "abi.c"
    src.write_text('#include "pbr_b200.h"\n#include "pbr_host.h"\n#include <stdio.h>\n'
                   'int main(void) { pbr_ctx* c = 0; int rc = pbr_create(0, &c);\n'
                   '  printf("%d %d\\n", rc, (int) (sizeof(pbr_camera) + sizeof(pbr_ray) + sizeof(pbr_hit)));\n'
                   '  if (rc == PBR_OK) pbr_destroy(c);\n  return 0; }\n')
    inc = os.path.join(ROOT, "include")
    libdir = os.path.dirname(pbr_b200.capi.LIB_PATH)
    subprocess.check_call(["gcc", "-std=c99", "-pedantic", "-Wall", "-Wextra", "-Werror", "-I", inc, "-fsyntax-only", str(src)])
    exe = str(tmp_path / "abi")
    subprocess.check_call(["gcc", "-std=c99", "-I", inc, str(src), "-o", exe, "-L", libdir, "-lpbr_b200", "-Wl,-rpath," + libdir])
    out = subprocess.check_output([exe]).decode().split()
    assert int(out[1]) == 80 + 32 + 16                      # camera_cl (PathTracer.h:25-32), ray, hit record
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    assert (int(out[0]) == 0) == has_gpu


def test_no_cpu_fallback_without_device():
    """Without a CUDA device pbr_create must fail (the product never routes to the CPU)."""
    import pbr_b200
    lib = pbr_b200.capi.load_library()
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if has_gpu:
        return
    ctx = ctypes.c_void_p()
    rc = lib.pbr_create(0, ctypes.byref(ctx))
    assert rc != 0 and not ctx.value
    try:
        pbr_b200.Device(0)
    except pbr_b200.PbrError:
        pass
    else:
        raise AssertionError("Device() must raise without a GPU")


def test_product_does_not_reference_the_oracle():
    """Nothing under the product package or include/ may import, include or link oracle/."""
    bad = []
    for base in ("physically-based-rendering_b200", "include"):
        for dirpath, _, files in os.walk(os.path.join(ROOT, base)):
            for f in files:
                if f.endswith((".so", ".o", ".pyc")):
                    continue
                text = open(os.path.join(dirpath, f), errors="ignore").read()
                for line in text.splitlines():
                    s = line.strip()
                    if re.search(r'#include\s+".*oracle', s) or re.search(r"^\s*(from|import)\s+oracle\b", s) or "liboracle" in s:
                        bad.append((f, s))
    assert not bad, bad
