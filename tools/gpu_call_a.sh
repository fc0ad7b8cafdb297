#!/bin/bash
# round 2, GPU call A: full -m gpu suite, smoke, bench line, launch list + full ncu capture of the top kernels
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total --format=csv > gpurun_out/a_gpu.txt
(time timeout 1500 python -m pytest tests -m gpu -x -q) > gpurun_out/a_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/a_pytest.log
tail -5 gpurun_out/a_pytest.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/a_smoke.log 2>&1; echo "smoke rc=$?"
(time timeout 900 python bench.py --steps 20 --warmup 5) > gpurun_out/a_bench.json 2> gpurun_out/a_bench.err; echo "bench rc=$?"
tail -c 3000 gpurun_out/a_bench.json
(time timeout 600 python bench.py --impl reference --steps 5 --warmup 1) > gpurun_out/a_bench_ref.json 2>&1
timeout 600 python bench.py --steps 2 --warmup 3 --no-extra --quick > gpurun_out/a_plain.log 2>&1 &&
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/a_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-extra --quick > gpurun_out/a_ncu1.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_htiso|k_idwt53_wide' -s 12 -c 8 -o gpurun_out/a_prof \
    python bench.py --steps 2 --warmup 3 --no-extra --quick > gpurun_out/a_ncu2.log 2>&1
ls -la gpurun_out
