set -x
mkdir -p gpurun_out
T=r02f
python -c "import sys; sys.path.insert(0,'.'); import pbr_b200; print(pbr_b200.capi.build_id())" > gpurun_out/${T}_build_id.txt
( time timeout 1500 python -m pytest tests -m gpu -x -q --durations=5 ) > gpurun_out/${T}_gpu_tests.log 2>&1
tail -8 gpurun_out/${T}_gpu_tests.log
timeout 900 python bench.py --steps 6 --warmup 3 > gpurun_out/${T}_bench_n1.json 2> gpurun_out/${T}_bench_n1.err
tail -c 3000 gpurun_out/${T}_bench_n1.json; tail -5 gpurun_out/${T}_bench_n1.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${T}_launches.csv python scripts/profile_frame.py 4 > gpurun_out/${T}_ncu_launches.log 2>&1
PBR_TRAVERSAL=0 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${T}_ref_launches.csv python scripts/profile_frame.py 4 > gpurun_out/${T}_ref_ncu_launches.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:traverseWideKernel -s 3 -c 3 -o gpurun_out/${T}_traverse python scripts/profile_frame.py 3 > gpurun_out/${T}_ncu_full.log 2>&1
PBR_TRAVERSAL=0 timeout 600 ncu --set full --clock-control none --import-source on -k regex:traverseKernel -s 3 -c 3 -o gpurun_out/${T}_ref_traverse python scripts/profile_frame.py 3 > gpurun_out/${T}_ref_ncu_full.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:shadeKernel -s 3 -c 3 -o gpurun_out/${T}_shade python scripts/profile_frame.py 3 > gpurun_out/${T}_shade_ncu_full.log 2>&1
ls -la gpurun_out | tail -12
