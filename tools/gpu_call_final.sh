#!/bin/bash
# round-2 evidence run: full -m gpu suite, smoke, bench line + reference arm, launch list, ncu --set full of the top kernels,
# the other BASELINE configs at full size
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
TAG=${1:-z}
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/${TAG}_gpu.txt
(time timeout 1500 python -m pytest tests -m gpu -x -q) > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${TAG}_pytest.log
tail -4 gpurun_out/${TAG}_pytest.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/${TAG}_smoke.log 2>&1; echo "smoke rc=$?"
(time timeout 900 python bench.py --steps 20 --warmup 5) > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench rc=$?"
(time timeout 600 python bench.py --impl reference --steps 5 --warmup 1) > gpurun_out/${TAG}_bench_ref.json 2>&1
timeout 900 python tools/bench_configs.py cfg1 cfg3 cfg4 cfg5 iso_cfg1 iso_cfg5 iso_4k_ebcot iso_4k_lossy > gpurun_out/${TAG}_configs.jsonl 2> gpurun_out/${TAG}_configs.err; echo "configs rc=$?"
timeout 600 python bench.py --steps 2 --warmup 3 --no-extra --quick > gpurun_out/${TAG}_plain.log 2>&1 &&
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/${TAG}_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-extra --quick > gpurun_out/${TAG}_ncu1.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_htiso|k_idwt53_wide' -s 10 -c 5 -o gpurun_out/${TAG}_prof \
    python bench.py --steps 2 --warmup 3 --no-extra --quick > gpurun_out/${TAG}_ncu2.log 2>&1
echo done
