#!/bin/bash
# A/B of library variants built into gpurun_variants/ (each is copied over the product library in this scratch copy only)
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
TAG=${1:-h}
cp go-jpeg2000_b200/libj2kgpu.so /tmp/libj2kgpu_orig.so
for V in $2; do
  cp gpurun_variants/libj2kgpu_$V.so go-jpeg2000_b200/libj2kgpu.so
  for G in $3; do
    J2KGPU_T1_GROUP=$G timeout 600 python tools/bench_configs.py $4 > gpurun_out/${TAG}_$V_g$G.jsonl 2> gpurun_out/${TAG}_$V_g$G.err
    python - <<PY
import json
for l in open("gpurun_out/${TAG}_$V_g$G.jsonl"):
    d=json.loads(l); print("  V=$V G=$G", d.get("config"), d.get("ms",{}).get("entropy"), {k:v for k,v in d.items() if "exact" in k or "diff" in k})
PY
  done
done
cp /tmp/libj2kgpu_orig.so go-jpeg2000_b200/libj2kgpu.so
