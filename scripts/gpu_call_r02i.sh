set -x
mkdir -p gpurun_out
T=r02i
( time timeout 1500 python -m pytest tests -m gpu -x -q --durations=5 ) > gpurun_out/${T}_gpu_tests.log 2>&1
tail -8 gpurun_out/${T}_gpu_tests.log
timeout 600 python scripts/ab_inflight.py 2>&1 | tee gpurun_out/${T}_ab_inflight.log
