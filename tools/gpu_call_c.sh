#!/bin/bash
# quick check of a kernel change: ISO / entropy / codestream GPU tests + a short bench line
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
TAG=${1:-c}
(timeout 900 python -m pytest tests/test_gpu_iso.py tests/test_gpu_entropy.py tests/test_gpu_codestream.py tests/test_gpu_fullsize.py -m gpu -x -q) > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${TAG}_pytest.log
tail -3 gpurun_out/${TAG}_pytest.log
timeout 600 python bench.py --steps 20 --warmup 5 --no-extra --quick > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench rc=$?"
python - <<PY
import json
d=json.loads(open("gpurun_out/${TAG}_bench.json").read().strip().splitlines()[-1])
print(d["value"], d["ms_per_step"], d["stages_ms"], d["e2e"]["value"], d["guard"])
PY
