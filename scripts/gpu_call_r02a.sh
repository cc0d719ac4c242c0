set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/r02a_smi.txt
nproc >> gpurun_out/r02a_smi.txt
( time timeout 900 python -m pytest tests -m gpu -x -q --durations=15 ) > gpurun_out/r02a_gpu_tests.log 2>&1
tail -30 gpurun_out/r02a_gpu_tests.log
rm -f gpurun_out/ab_variants.log
AB_NO_TESTS=1 timeout 600 scripts/ab_variants.sh r1base help1 help2 trinol1 order
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02a_launches.csv python scripts/profile_frame.py 4 > gpurun_out/r02a_ncu_launches.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:traverseKernel -s 6 -c 3 -o gpurun_out/r02a_traverse python scripts/profile_frame.py 3 > gpurun_out/r02a_ncu_full.log 2>&1
tail -3 gpurun_out/r02a_ncu_full.log
ls -la gpurun_out | tail -12
