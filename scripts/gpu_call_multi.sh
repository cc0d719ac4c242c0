set -x
mkdir -p gpurun_out
T=r02m
N=${1:-8}
M=tests/golden/models/
H=physically-based-rendering_b200/host/pbr_headless
timeout 60 $H --model $M suzanne.obj --frames 4 --deterministic --set window.width=256 --set window.height=192 --out gpurun_out/${T}_one.pfm
timeout 90 $H --model $M suzanne.obj --frames 4 --deterministic --set window.width=256 --set window.height=192 --ranks $N --shard stripes --out gpurun_out/${T}_stripes.pfm
timeout 90 $H --model $M suzanne.obj --frames 4 --deterministic --set window.width=256 --set window.height=192 --ranks $N --shard rows --out gpurun_out/${T}_rows.pfm
cmp gpurun_out/${T}_one.pfm gpurun_out/${T}_stripes.pfm && echo "STRIPES_IDENTICAL"
cmp gpurun_out/${T}_one.pfm gpurun_out/${T}_rows.pfm && echo "ROWS_IDENTICAL"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 8 --warmup 3 > gpurun_out/${T}_bench_n$N.json 2> gpurun_out/${T}_bench_n$N.err
tail -c 1600 gpurun_out/${T}_bench_n$N.json; tail -4 gpurun_out/${T}_bench_n$N.err
