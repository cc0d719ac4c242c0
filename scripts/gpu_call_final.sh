# The round's closing GPU call: GPU tests, smoke, the bench line as the driver runs it, and the ncu captures of the very
# build that was benchmarked (build id recorded).  scripts/summarize_ncu.py <tag> turns the captures into profiles/.
set -x
mkdir -p gpurun_out
T=${1:-r02_final}
python -c "import sys; sys.path.insert(0,'.'); import pbr_b200; print(pbr_b200.capi.build_id())" > gpurun_out/${T}_build_id.txt
( time timeout 1500 python -m pytest tests -m gpu -q --durations=5 ) > gpurun_out/${T}_gpu_tests.log 2>&1
tail -8 gpurun_out/${T}_gpu_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 | tee gpurun_out/${T}_smoke.log
timeout 900 python bench.py > gpurun_out/${T}_bench_n1.json 2> gpurun_out/${T}_bench_n1.err
python - <<PY
import json
d=json.load(open('gpurun_out/${T}_bench_n1.json'))
print({k:d[k] for k in ('value','ms_per_step','steps','gpu_launches')}, 'e2e', d['e2e']['value'], d['config']['pipeline'])
print(d['verify']); print(d['reference_walk']); print({k:d['roofline'][k] for k in ('frac','achieved','kernel_share_of_step','traffic','ncu_note')})
PY
tail -3 gpurun_out/${T}_bench_n1.err
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${T}_bench_reference.json 2> gpurun_out/${T}_bench_reference.err; tail -c 600 gpurun_out/${T}_bench_reference.json
export PBR_FRAMES_IN_FLIGHT=1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${T}_launches.csv python scripts/profile_frame.py 4 > gpurun_out/${T}_ncu_launches.log 2>&1
PBR_TRAVERSAL=0 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${T}_ref_launches.csv python scripts/profile_frame.py 4 > gpurun_out/${T}_ref_ncu_launches.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:traverseWideKernel -s 3 -c 3 -o gpurun_out/${T}_traverse python scripts/profile_frame.py 3 > gpurun_out/${T}_ncu_full.log 2>&1
PBR_TRAVERSAL=0 timeout 600 ncu --set full --clock-control none --import-source on -k regex:traverseKernel -s 3 -c 3 -o gpurun_out/${T}_ref_traverse python scripts/profile_frame.py 3 > gpurun_out/${T}_ref_ncu_full.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:shadeKernel -s 3 -c 3 -o gpurun_out/${T}_shade python scripts/profile_frame.py 3 > gpurun_out/${T}_shade_ncu_full.log 2>&1
ls -la gpurun_out | grep ${T} | tail -16
