set -x
mkdir -p gpurun_out
T=r02n
N=${1:-8}
for F in 4 8; do
PBR_FRAMES_IN_FLIGHT=$F timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 8 --warmup 3 --no-e2e > gpurun_out/${T}_bench_n${N}_f$F.json 2> gpurun_out/${T}_bench_n${N}_f$F.err
python - <<PY
import json
d=json.load(open('gpurun_out/${T}_bench_n${N}_f$F.json'))
print('F=$F', {k:d[k] for k in ('value','ms_per_step')}, {k:(v.get('speedup'), v.get('ms_per_step'), v.get('bit_identical_pixels')) for k,v in d['strong'].items() if isinstance(v, dict)})
PY
done
