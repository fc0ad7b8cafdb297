#!/bin/bash
# 9-7 path check: DWT/MCT + ISO lossy GPU tests, then the full-size cfg3 / cfg5 side measurements (REF and ISO)
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
TAG=${1:-e}
(timeout 900 python -m pytest tests/test_gpu_dwt_mct.py tests/test_gpu_iso.py tests/test_gpu_path.py tests/test_gpu_fullsize.py -m gpu -x -q) > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${TAG}_pytest.log
tail -3 gpurun_out/${TAG}_pytest.log
timeout 900 python tools/bench_configs.py cfg3 cfg5 iso_cfg5 iso_4k_lossy > gpurun_out/${TAG}_configs.jsonl 2> gpurun_out/${TAG}_configs.err; echo "configs rc=$?"
python - <<PY
import json
for l in open("gpurun_out/${TAG}_configs.jsonl"):
    d=json.loads(l); print(d.get("config"), d.get("frames"), d.get("ms"), d.get("gpixel_s"), {k:v for k,v in d.items() if "exact" in k or "equal" in k or "diff" in k})
PY
if [ -n "$2" ]; then
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"$2" -c 6 -o gpurun_out/${TAG}_prof \
    python tools/bench_configs.py iso_cfg5 cfg5 > gpurun_out/${TAG}_ncu.log 2>&1; echo "ncu rc=$?"
fi
