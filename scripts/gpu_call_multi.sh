set -x
mkdir -p gpurun_out
T=${2:-r02_multi}
N=${1:-2}
M=tests/golden/models/
H=physically-based-rendering_b200/host/pbr_headless
timeout 60 $H --model $M suzanne.obj --frames 6 --deterministic --set window.width=256 --set window.height=192 --out gpurun_out/${T}_one.pfm
timeout 90 $H --model $M suzanne.obj --frames 6 --deterministic --set window.width=256 --set window.height=192 --ranks $N --shard stripes --out gpurun_out/${T}_stripes.pfm
timeout 90 $H --model $M suzanne.obj --frames 6 --deterministic --set window.width=256 --set window.height=192 --ranks $N --shard rows --out gpurun_out/${T}_rows.pfm
cmp gpurun_out/${T}_one.pfm gpurun_out/${T}_stripes.pfm && echo "STRIPES_IDENTICAL"
cmp gpurun_out/${T}_one.pfm gpurun_out/${T}_rows.pfm && echo "ROWS_IDENTICAL"
# a height the ranks cannot share equally: the gather falls back to grouped broadcasts of unequal row blocks
timeout 60 $H --model $M suzanne.obj --frames 5 --deterministic --set window.width=200 --set window.height=200 --out gpurun_out/${T}_one200.pfm
timeout 90 $H --model $M suzanne.obj --frames 5 --deterministic --set window.width=200 --set window.height=200 --ranks $N --shard rows --out gpurun_out/${T}_rows200.pfm
cmp gpurun_out/${T}_one200.pfm gpurun_out/${T}_rows200.pfm && echo "UNEQUAL_ROWS_IDENTICAL"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 8 --warmup 3 > gpurun_out/${T}_bench_n${N}.json 2> gpurun_out/${T}_bench_n${N}.err
python - <<PY
import json
d=json.load(open('gpurun_out/${T}_bench_n${N}.json'))
print('N=$N', {k:d[k] for k in ('value','ms_per_step')}, 'e2e', d['e2e']['value'], {k:(v.get('speedup'), v.get('ms_per_step'), v.get('bit_identical_pixels')) for k,v in d['strong'].items() if isinstance(v, dict)}, d['strong'].get('error'))
PY
tail -3 gpurun_out/${T}_bench_n${N}.err
