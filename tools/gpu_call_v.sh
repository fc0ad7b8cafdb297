#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
TAG=${1:-v}
timeout 900 python tools/encode_probe.py > gpurun_out/${TAG}_encode_probe.log 2>&1; echo "probe rc=$?"
head -3 gpurun_out/${TAG}_encode_probe.log; tail -3 gpurun_out/${TAG}_encode_probe.log
