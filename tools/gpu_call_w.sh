#!/bin/bash
# validation of the final tree: full -m gpu suite, smoke, bench line + reference arm, forward-path launch list
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
TAG=${1:-w}
(time timeout 1500 python -m pytest tests -m gpu -x -q) > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${TAG}_pytest.log
grep -n "passed\|failed\|rc=" gpurun_out/${TAG}_pytest.log | tail -3
timeout 300 python __graft_entry__.py smoke > gpurun_out/${TAG}_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/${TAG}_smoke.log
(timeout 900 python bench.py --steps 20 --warmup 5) > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench rc=$?"
(timeout 600 python bench.py --impl reference --steps 5 --warmup 1) > gpurun_out/${TAG}_bench_ref.json 2> gpurun_out/${TAG}_bench_ref.err; echo "ref rc=$?"
python - <<PY
import json
d=json.loads([l for l in open("gpurun_out/${TAG}_bench.json") if l.startswith("{")][-1])
print("value", d["value"], d["ms_per_step"], "e2e", d["e2e"]["value"], "launches", d["gpu_launches"], "roofline", d["roofline"]["frac"])
print("forward_path", json.dumps(d.get("forward_path")))
PY
timeout 600 python tools/encode_probe.py > /dev/null 2>&1 &&
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/${TAG}_encode_launches.csv \
    python tools/encode_probe.py > gpurun_out/${TAG}_ncu_enc.log 2>&1
echo done
