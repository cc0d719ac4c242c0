#!/bin/bash
# Build experiment variants of libpbr_b200.so next to it (csrc/var_<name>.so; git-ignored, they travel with gpurun).
#   scripts/build_variants.sh help1=-DPT_HELPER_PREFETCH=1 help2=-DPT_HELPER_PREFETCH=2 trinol1=-DPT_TRI_NO_L1=1 order=-DPT_NODE_ORDER=1
# The in-tree libpbr_b200.so is not touched.  Run here (nvcc cross-compiles), then scripts/ab_variants.sh on the GPU box.
set -e
cd "$(dirname "$0")/../physically-based-rendering_b200/csrc"
FLAGS="-gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -fmad=false -Xcompiler -fPIC"
for spec in "$@"; do
	name="${spec%%=*}"
	defs="${spec#*=}"
	( nvcc $FLAGS $defs -shared -o "var_${name}.so" pbr_capi.cu -lcudart && echo "built var_${name}.so ($defs)" ) &
done
wait
