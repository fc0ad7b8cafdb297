#!/bin/bash
# chunk plans of the host-buffer run: e2e of the bench batch under explicit J2KGPU_CHUNKS
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
TAG=${1:-i}
for P in default "1,1,1,1,1,1,1,1,1,1,1,1,1,1,1,1" "1,2,2,2,2,2,2,3" "1,1,2,2,2,2,2,2,2" "2,2,2,2,2,2,2,2" "1,3,4,4,4"; do
  if [ "$P" = default ]; then unset J2KGPU_CHUNKS; else export J2KGPU_CHUNKS=$P; fi
  timeout 600 python bench.py --steps 20 --warmup 5 --no-extra --quick > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
  python - <<PY
import json
d=json.loads(open("gpurun_out/${TAG}_bench.json").read().strip().splitlines()[-1])
print("plan $P: e2e", d["e2e"]["value"], d["e2e"]["ms_per_step"], "prebuilt", d["e2e"]["prebuilt_job"]["ms_per_step"], "device", d["ms_per_step"])
PY
done
