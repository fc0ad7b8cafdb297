#!/bin/bash
# kernel change check + ncu capture of the entropy kernels: tests, bench line, then ncu --set full on the HT ISO kernels
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
TAG=${1:-d}
KREGEX=${2:-k_htiso}
(timeout 900 python -m pytest tests/test_gpu_iso.py tests/test_gpu_entropy.py tests/test_gpu_codestream.py tests/test_gpu_fullsize.py tests/test_gpu_path.py -m gpu -x -q) > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${TAG}_pytest.log
tail -3 gpurun_out/${TAG}_pytest.log
timeout 600 python bench.py --steps 20 --warmup 5 --no-extra --quick > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench rc=$?"
python - <<PY
import json
d=json.loads(open("gpurun_out/${TAG}_bench.json").read().strip().splitlines()[-1])
print(d["value"], d["ms_per_step"], d["stages_ms"], d["e2e"]["value"], d["guard"])
PY
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"$KREGEX" -s 6 -c 3 -o gpurun_out/${TAG}_prof \
    python bench.py --steps 2 --warmup 3 --no-extra --quick > gpurun_out/${TAG}_ncu.log 2>&1; echo "ncu rc=$?"
