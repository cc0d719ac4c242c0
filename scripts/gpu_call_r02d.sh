set -x
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -x -q --durations=8 ) > gpurun_out/r02d_gpu_tests.log 2>&1
tail -15 gpurun_out/r02d_gpu_tests.log
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02d_launches.csv python scripts/profile_frame.py 4 > gpurun_out/r02d_ncu_launches.log 2>&1
PBR_TRAVERSAL=0 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02d_ref_launches.csv python scripts/profile_frame.py 4 > gpurun_out/r02d_ref_ncu_launches.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:traverseWideKernel -s 3 -c 3 -o gpurun_out/r02d_traverse python scripts/profile_frame.py 3 > gpurun_out/r02d_ncu_full.log 2>&1
tail -3 gpurun_out/r02d_ncu_full.log
PBR_TRAVERSAL=0 timeout 600 ncu --set full --clock-control none --import-source on -k regex:traverseKernel -s 3 -c 3 -o gpurun_out/r02d_ref_traverse python scripts/profile_frame.py 3 > gpurun_out/r02d_ref_ncu_full.log 2>&1
tail -3 gpurun_out/r02d_ref_ncu_full.log
ls -la gpurun_out | tail -8
