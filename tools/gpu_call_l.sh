#!/bin/bash
# packed-RGB transfer: full GPU suite, then e2e of the bench batch with the option off / on / auto
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
TAG=${1:-l}
(timeout 1500 python -m pytest tests -m gpu -x -q) > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${TAG}_pytest.log
tail -3 gpurun_out/${TAG}_pytest.log
nproc
for M in 0 1 auto; do
  if [ "$M" = auto ]; then unset J2KGPU_HOST_ALPHA; else export J2KGPU_HOST_ALPHA=$M; fi
  J2KGPU_DEBUG_PLAN=0 timeout 600 python bench.py --steps 20 --warmup 5 --no-extra --quick > gpurun_out/${TAG}_bench_$M.json 2> gpurun_out/${TAG}_bench_$M.err
  python - <<PY
import json
d=json.loads(open("gpurun_out/${TAG}_bench_$M.json").read().strip().splitlines()[-1])
print("host_alpha=$M: e2e", d["e2e"]["value"], d["e2e"]["ms_per_step"], "prebuilt", d["e2e"]["prebuilt_job"]["ms_per_step"], "device", d["ms_per_step"], d["guard"])
PY
done
