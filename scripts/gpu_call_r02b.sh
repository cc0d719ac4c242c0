set -x
mkdir -p gpurun_out
( time timeout 1200 python -m pytest tests -m gpu -x -q --durations=8 ) > gpurun_out/r02c_gpu_tests.log 2>&1
tail -25 gpurun_out/r02c_gpu_tests.log
timeout 600 python scripts/ab_traversal.py 1 21 85 2>&1 | tee gpurun_out/r02c_ab_traversal.log
