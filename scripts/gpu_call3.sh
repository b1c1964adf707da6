#!/bin/bash
timeout 400 python -m pytest tests -m gpu -x -q > gpurun_out/r2_t3.log 2>&1; echo "all tests rc=$?"; tail -3 gpurun_out/r2_t3.log
python bench.py --steps 20 --warmup 5 > gpurun_out/r2_bench3.log 2>&1; echo "bench rc=$?"
python bench.py --steps 10 --warmup 3 --config l16 --no-cpu-baseline > gpurun_out/r2_bench_l16.log 2>&1; echo "l16 rc=$?"
