#!/bin/bash
# On the GPU box: time the in-tree library and every csrc/var_<name>.so on the C2 workload (scripts/ab_frames.py: best and
# median of 5, image and hit checksums -- they must not change), twice in alternation, then run the GPU parity tests
# on each variant.  The in-tree library is restored at the end.
#   scripts/ab_variants.sh help1 help2 trinol1        (writes gpurun_out/ab_variants.log)
cd "$(dirname "$0")/.."
D=physically-based-rendering_b200/csrc
mkdir -p gpurun_out
cp $D/libpbr_b200.so $D/keep.so
trap 'cp $D/keep.so $D/libpbr_b200.so' EXIT
for pass in 1 2; do
	for v in base "$@"; do
		if [ "$v" = base ]; then cp $D/keep.so $D/libpbr_b200.so; else cp $D/var_$v.so $D/libpbr_b200.so; fi
		timeout 60 python scripts/ab_frames.py $v 2>&1 | tail -2 | tee -a gpurun_out/ab_variants.log
	done
done
[ -n "$AB_NO_TESTS" ] && exit 0
for v in "$@"; do
	cp $D/var_$v.so $D/libpbr_b200.so
	echo "== GPU tests with $v" | tee -a gpurun_out/ab_variants.log
	timeout 120 python -m pytest tests -m gpu -x -q 2>&1 | tail -2 | tee -a gpurun_out/ab_variants.log
done
