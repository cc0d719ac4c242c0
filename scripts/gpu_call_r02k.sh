set -x
mkdir -p gpurun_out
T=r02k
( time timeout 1500 python -m pytest tests -m gpu -x -q --durations=5 ) > gpurun_out/${T}_gpu_tests.log 2>&1
tail -8 gpurun_out/${T}_gpu_tests.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/${T}_bench_n1.json 2> gpurun_out/${T}_bench_n1.err
python -c "
import json
d=json.load(open('gpurun_out/${T}_bench_n1.json'))
print({k:d[k] for k in ('value','ms_per_step','steps','gpu_launches')}, 'e2e', d['e2e']['value'], d['config']['pipeline'], d['verify'], d['reference_walk'], d['roofline']['frac'], d['roofline']['kernel_share_of_step'])
"
tail -3 gpurun_out/${T}_bench_n1.err
timeout 300 python scripts/run_configs.py --only c1 --out gpurun_out/${T}_configs.json 2>&1 | tail -5
