#!/bin/bash
# ncu --set full of one kernel of a tools/bench_configs.py config: gpu_call_j.sh TAG CONFIG KERNEL_REGEX [skip]
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
TAG=$1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"$3" -s ${4:-0} -c 1 -o gpurun_out/${TAG}_prof \
    python tools/bench_configs.py $2 > gpurun_out/${TAG}_ncu.log 2>&1; echo "ncu rc=$?"
