#!/bin/bash
# 2-GPU check: the bench line under torchrun (frames sharded, cfg4 tiles sharded over both ranks)
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
TAG=${1:-f}
N=${2:-2}
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/${TAG}_bench_${N}gpu.json 2> gpurun_out/${TAG}_bench_${N}gpu.err; echo "bench rc=$?"
tail -c 2500 gpurun_out/${TAG}_bench_${N}gpu.json
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29518 bench.py --impl reference --gpus $N --steps 3 --warmup 1 > gpurun_out/${TAG}_ref_${N}gpu.json 2>&1; echo "ref rc=$?"
