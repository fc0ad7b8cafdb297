#!/bin/bash
# compute-sanitizer (memcheck, racecheck, synccheck) over a small slice of the GPU tests that covers the round-2 kernels
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
TAG=${1:-k}
SEL='sigprop or htj2k_8bit or garbage or random_blocks or lossy_htj2k or openjpeg_codestream_in or golden or whole_path_lossy or reduce_resolution'
for TOOL in memcheck racecheck synccheck; do
  timeout 1200 compute-sanitizer --tool $TOOL --error-exitcode 7 --print-limit 20 python -m pytest tests/test_gpu_iso.py tests/test_gpu_entropy.py tests/test_gpu_codestream.py -m gpu -x -q -k "$SEL" > gpurun_out/${TAG}_$TOOL.log 2>&1
  echo "$TOOL rc=$?"; grep -E "ERROR SUMMARY|passed|failed|RACECHECK SUMMARY" gpurun_out/${TAG}_$TOOL.log | tail -3
done
