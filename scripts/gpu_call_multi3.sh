set -x
mkdir -p gpurun_out
T=r02p
N=${1:-8}
for P in 0 1 0 1; do
PBR_COMM_PRIORITY=$P timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 8 --warmup 3 --no-e2e --no-strong > gpurun_out/${T}_bench_n${N}_p$P.json 2> gpurun_out/${T}_bench_n${N}_p$P.err
python - <<PY
import json
d=json.load(open('gpurun_out/${T}_bench_n${N}_p$P.json'))
print('PRIO=$P', {k:d[k] for k in ('value','ms_per_step')})
PY
done
