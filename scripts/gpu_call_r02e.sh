set -x
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py -m gpu -x -q -k "ordered or fullsize" ) > gpurun_out/r02e_gpu_tests.log 2>&1
tail -5 gpurun_out/r02e_gpu_tests.log
timeout 900 python scripts/ab_traversal.py 21,16,4 21,8,4 21,12,4 21,20,4 21,24,4 21,28,4 21,16,2 21,16,8 21,16,12 1,16,4 2>&1 | tee gpurun_out/r02e_ab_traversal.log
