#!/bin/bash
# bench lines of record for this round's final kernels: default, L/16, original-ViT branch, Kohonen
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r2f_bench_default.log 2>&1; echo "default rc=$?"; tail -1 gpurun_out/r2f_bench_default.log | cut -c1-200
timeout 600 python bench.py --steps 10 --warmup 3 --config l16 --no-cpu-baseline > gpurun_out/r2f_bench_l16.log 2>&1; echo "l16 rc=$?"; tail -1 gpurun_out/r2f_bench_l16.log | cut -c1-200
timeout 600 python bench.py --steps 10 --warmup 3 --variant orig --no-cpu-baseline > gpurun_out/r2f_bench_orig.log 2>&1; echo "orig rc=$?"; tail -1 gpurun_out/r2f_bench_orig.log | cut -c1-200
timeout 600 python bench.py --steps 10 --warmup 3 --variant kohonen --no-cpu-baseline > gpurun_out/r2f_bench_koh.log 2>&1; echo "koh rc=$?"; tail -1 gpurun_out/r2f_bench_koh.log | cut -c1-200
