set -x
mkdir -p gpurun_out
T=r02o
( time timeout 1500 python -m pytest tests -m gpu -x -q ) > gpurun_out/${T}_gpu_tests.log 2>&1
tail -4 gpurun_out/${T}_gpu_tests.log
timeout 300 python scripts/run_configs.py --only c1 --out gpurun_out/${T}_configs.json 2>&1 | tail -2
( timeout 600 compute-sanitizer --tool memcheck python scripts/sanitize.py ) > gpurun_out/${T}_memcheck.log 2>&1
tail -6 gpurun_out/${T}_memcheck.log
( timeout 600 compute-sanitizer --tool racecheck python scripts/sanitize.py --no-megakernel ) > gpurun_out/${T}_racecheck.log 2>&1
tail -4 gpurun_out/${T}_racecheck.log
( timeout 400 compute-sanitizer --tool synccheck python scripts/sanitize.py --no-megakernel ) > gpurun_out/${T}_synccheck.log 2>&1
tail -3 gpurun_out/${T}_synccheck.log
