#!/bin/bash
# final regression of round 2 on the code of record: full GPU suite, smoke, bench lines of BASELINE configs 2 / 3 / 4 / 5
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/r2k_tests.log 2>&1; echo "tests rc=$?"; tail -1 gpurun_out/r2k_tests.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2k_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/r2k_smoke.log
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r2k_bench_default.log 2>&1; echo "default rc=$?"; tail -1 gpurun_out/r2k_bench_default.log | cut -c1-200
timeout 600 python bench.py --steps 10 --warmup 3 --config l16 --no-cpu-baseline > gpurun_out/r2k_bench_l16.log 2>&1; echo "l16 rc=$?"; tail -1 gpurun_out/r2k_bench_l16.log | cut -c1-200
timeout 600 python bench.py --steps 10 --warmup 3 --variant orig --no-cpu-baseline > gpurun_out/r2k_bench_orig.log 2>&1; echo "orig rc=$?"; tail -1 gpurun_out/r2k_bench_orig.log | cut -c1-200
timeout 600 python bench.py --steps 10 --warmup 3 --variant kohonen --no-cpu-baseline > gpurun_out/r2k_bench_koh.log 2>&1; echo "koh rc=$?"; tail -1 gpurun_out/r2k_bench_koh.log | cut -c1-200
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2k_bench_ref.log 2>&1; echo "ref rc=$?"; tail -1 gpurun_out/r2k_bench_ref.log | cut -c1-300
