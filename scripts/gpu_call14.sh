#!/bin/bash
# TMEM read-rate probe; residual backward ring sweep at C = 1024; kernel tests + short bench with the new skip-form policy
timeout 60 scripts/probes/tmem_read_bw.bin > gpurun_out/r2i_tmem_read_bw.log 2>&1; cat gpurun_out/r2i_tmem_read_bw.log
timeout 200 python scripts/residual_w_sweep.py 1024 > gpurun_out/r2i_residual_w_1024.log 2>&1; cat gpurun_out/r2i_residual_w_1024.log
timeout 300 python -m pytest tests/test_kernels_gpu.py -x -q -k "residual" > gpurun_out/r2i_res_tests.log 2>&1; echo "residual tests rc=$?"; tail -1 gpurun_out/r2i_res_tests.log
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2i_bench.log 2>&1; echo "bench rc=$?"; tail -1 gpurun_out/r2i_bench.log | cut -c1-200
