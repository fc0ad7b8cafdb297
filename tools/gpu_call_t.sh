#!/bin/bash
# forward path: its tests, the probe (both coders, several contexts at once)
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
TAG=${1:-t}
(time timeout 900 python -m pytest tests/test_gpu_encode.py -m gpu -x -q) > gpurun_out/${TAG}_pytest_enc.log 2>&1; grep -n "passed\|failed" gpurun_out/${TAG}_pytest_enc.log
timeout 900 python tools/encode_probe.py > gpurun_out/${TAG}_encode_probe.log 2>&1; echo "probe rc=$?"
cat gpurun_out/${TAG}_encode_probe.log
if [ -f go-jpeg2000_b200/libj2kgpu_ab.so ]; then
J2K_PROBE_LIB=$PWD/go-jpeg2000_b200/libj2kgpu_ab.so timeout 900 python tools/encode_probe.py > gpurun_out/${TAG}_encode_probe_ab.log 2>&1; echo "probe ab rc=$?"
grep -v "contexts at once" gpurun_out/${TAG}_encode_probe_ab.log
fi
