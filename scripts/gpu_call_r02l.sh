set -x
mkdir -p gpurun_out
T=r02l
( time timeout 1500 python -m pytest tests -m gpu -x -q ) > gpurun_out/${T}_gpu_tests.log 2>&1
tail -4 gpurun_out/${T}_gpu_tests.log
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-reference-walk > gpurun_out/${T}_bench_n1.json 2> gpurun_out/${T}_bench_n1.err
python -c "
import json
d=json.load(open('gpurun_out/${T}_bench_n1.json'))
print({k:d[k] for k in ('value','ms_per_step','steps','gpu_launches')}, 'e2e', d['e2e']['value'], d['e2e']['ms_per_step'])
"
tail -3 gpurun_out/${T}_bench_n1.err
