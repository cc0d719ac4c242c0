set -x
mkdir -p gpurun_out
T=r02g
nvidia-smi -L
# the C++ driver alone, two ranks, rows sharded in stripes: must equal one rank
M=tests/golden/models/
H=physically-based-rendering_b200/host/pbr_headless
timeout 120 $H --model $M suzanne.obj --frames 4 --deterministic --set window.width=256 --set window.height=192 --out gpurun_out/${T}_one.pfm
timeout 120 $H --model $M suzanne.obj --frames 4 --deterministic --set window.width=256 --set window.height=192 --ranks 2 --shard stripes --out gpurun_out/${T}_two_stripes.pfm
timeout 120 $H --model $M suzanne.obj --frames 4 --deterministic --set window.width=256 --set window.height=192 --ranks 2 --shard rows --out gpurun_out/${T}_two_rows.pfm
timeout 120 $H --model $M suzanne.obj --frames 4 --deterministic --set window.width=256 --set window.height=192 --ranks 2 --shard spp --out gpurun_out/${T}_two_spp.pfm
cmp gpurun_out/${T}_one.pfm gpurun_out/${T}_two_stripes.pfm && echo "STRIPES_IDENTICAL"
cmp gpurun_out/${T}_one.pfm gpurun_out/${T}_two_rows.pfm && echo "ROWS_IDENTICAL"
ls -la gpurun_out/${T}_*.pfm
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/${T}_bench_n2.json 2> gpurun_out/${T}_bench_n2.err
tail -c 2500 gpurun_out/${T}_bench_n2.json; tail -8 gpurun_out/${T}_bench_n2.err
