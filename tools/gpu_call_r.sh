#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
TAG=${1:-r}
(timeout 900 python -m pytest tests/test_gpu_path.py tests/test_gpu_iso.py tests/test_gpu_fullsize.py -m gpu -x -q) > gpurun_out/${TAG}_pytest.log 2>&1; tail -2 gpurun_out/${TAG}_pytest.log
for M in 1 0; do
  J2KGPU_NO_COARSE_OVERLAP=$M timeout 600 python bench.py --steps 20 --warmup 5 --no-extra --quick > gpurun_out/${TAG}_bench_$M.json 2> gpurun_out/${TAG}_bench_$M.err
  python - <<PY
import json
d=json.loads(open("gpurun_out/${TAG}_bench_$M.json").read().strip().splitlines()[-1])
print("no_coarse_overlap=$M: value", d["value"], d["ms_per_step"], d["stages_ms"], "sustained", d.get("sustained"), d["guard"]["frames_equal_source"], d["gpu_launches"])
PY
done
