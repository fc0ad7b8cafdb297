#!/bin/bash
# round 2, GPU call B: -m gpu suite with the codestream front door, smoke, full bench line + reference arm
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
(time timeout 1500 python -m pytest tests -m gpu -x -q) > gpurun_out/b_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/b_pytest.log
tail -5 gpurun_out/b_pytest.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/b_smoke.log 2>&1; echo "smoke rc=$?"
(time timeout 900 python bench.py --steps 20 --warmup 5) > gpurun_out/b_bench.json 2> gpurun_out/b_bench.err; echo "bench rc=$?"
tail -c 6000 gpurun_out/b_bench.json
(time timeout 600 python bench.py --impl reference --steps 5 --warmup 1) > gpurun_out/b_bench_ref.json 2>&1
