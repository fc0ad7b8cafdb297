#!/bin/bash
# forward path on the GPU: its tests, then the whole -m gpu suite, smoke, the bench line (with the forward_path side object)
# and a launch list of the forward path alone
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
TAG=${1:-s}
(time timeout 900 python -m pytest tests/test_gpu_encode.py -m gpu -x -q) > gpurun_out/${TAG}_pytest_enc.log 2>&1; tail -3 gpurun_out/${TAG}_pytest_enc.log
if [ "$2" = full ]; then
(time timeout 1500 python -m pytest tests -m gpu -x -q --deselect tests/test_gpu_encode.py) > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${TAG}_pytest.log
tail -4 gpurun_out/${TAG}_pytest.log
fi
timeout 300 python __graft_entry__.py smoke > gpurun_out/${TAG}_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/${TAG}_smoke.log
(time timeout 900 python bench.py --steps 20 --warmup 5) > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench rc=$?"
python - <<PY
import json
d=json.loads(open("gpurun_out/${TAG}_bench.json").read().strip().splitlines()[-1])
print("value", d["value"], d["ms_per_step"], "e2e", d["e2e"]["value"], "launches", d["gpu_launches"])
print("forward_path", json.dumps(d.get("forward_path")))
PY
timeout 600 python tools/encode_probe.py > gpurun_out/${TAG}_encode_probe.log 2>&1 &&
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/${TAG}_encode_launches.csv \
    python tools/encode_probe.py > gpurun_out/${TAG}_ncu_enc.log 2>&1
tail -12 gpurun_out/${TAG}_encode_probe.log
echo done
