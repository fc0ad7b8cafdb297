#!/bin/bash
# EBCOT kernel check: tests, then A/B of the lanes-per-block option on the EBCOT configs
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
TAG=${1:-g}
(timeout 900 python -m pytest tests/test_gpu_entropy.py tests/test_gpu_iso.py tests/test_gpu_path.py tests/test_gpu_fullsize.py -m gpu -x -q) > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${TAG}_pytest.log
tail -3 gpurun_out/${TAG}_pytest.log
for G in ${2:-32 16 8 4}; do
  J2KGPU_T1_GROUP=$G timeout 900 python tools/bench_configs.py cfg1 cfg3 iso_4k_ebcot iso_4k_lossy > gpurun_out/${TAG}_configs_g$G.jsonl 2> gpurun_out/${TAG}_configs_g$G.err; echo "G=$G rc=$?"
  python - <<PY
import json
for l in open("gpurun_out/${TAG}_configs_g$G.jsonl"):
    d=json.loads(l); print("  G=$G", d.get("config"), d.get("frames"), d.get("ms"), {k:v for k,v in d.items() if "exact" in k or "diff" in k})
PY
done
