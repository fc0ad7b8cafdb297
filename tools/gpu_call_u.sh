#!/bin/bash
# ncu --set full of the row-mask tier-1 encoder kernel (source-level), after the plain probe has run
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
TAG=${1:-u}
timeout 600 python tools/encode_probe.py > gpurun_out/${TAG}_encode_probe.log 2>&1 || exit 1
tail -4 gpurun_out/${TAG}_encode_probe.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_t1_enc -s 11 -c 1 -o gpurun_out/${TAG}_t1enc \
    python tools/encode_probe.py > gpurun_out/${TAG}_ncu.log 2>&1; echo "ncu rc=$?"; tail -3 gpurun_out/${TAG}_ncu.log
