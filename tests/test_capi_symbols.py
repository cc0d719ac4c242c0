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
