"""pbr_b200 -- B200-native path-tracing core behind the host API of sebadorn/Physically-based-Rendering.

Layout of this package (only what the hot path needs):
    csrc/      CUDA kernels (sm_100a) and the C ABI  -> libpbr_b200.so   (include/pbr_b200.h)
    host/      C++ mirror of the reference's host classes (PathTracer, CL shim, BVH, parsers, Cfg)
               -> libpbr_host.so + the headless `pbr_headless` driver
    capi.py    ctypes view of the C ABI
    scenes.py  synthetic scene generators for the BASELINE configurations

The directory name contains a hyphen, so import it through the alias module `pbr_b200` at the
repository root (`import pbr_b200`).  Nothing here falls back to the CPU: without the built shared
library or without a CUDA device, constructing a Device raises.
"""
from . import capi, host, multigpu, scenes    # noqa: F401
from .capi import Device, PbrError  # noqa: F401
