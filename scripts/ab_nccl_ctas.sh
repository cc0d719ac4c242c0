set -x
mkdir -p gpurun_out
N=2
for C in 0 2 4 8 0 4; do
PBR_NCCL_MAX_CTAS=$C timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 8 --warmup 3 --no-e2e --no-strong > gpurun_out/ctas_$C.json 2> gpurun_out/ctas_$C.err
python - <<PY
import json
d=json.load(open('gpurun_out/ctas_$C.json'))
print('CTAS=$C', {k:d[k] for k in ('value','ms_per_step')})
PY
done
