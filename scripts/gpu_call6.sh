#!/bin/bash
python bench.py --steps 10 --warmup 3 --config l16 --no-cpu-baseline > gpurun_out/r2_bench_l16.log 2>&1; echo "l16 rc=$?"
python bench.py --steps 10 --warmup 3 --variant orig --no-cpu-baseline > gpurun_out/r2_bench_orig.log 2>&1; echo "orig rc=$?"
python bench.py --steps 10 --warmup 3 --variant kohonen --no-cpu-baseline > gpurun_out/r2_bench_koh.log 2>&1; echo "koh rc=$?"
python bench.py --steps 20 --warmup 5 > gpurun_out/r2_bench_default.log 2>&1; echo "default rc=$?"
