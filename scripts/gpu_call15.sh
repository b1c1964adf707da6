#!/bin/bash
# A/B in one process: row-wise kernels bottom up (nvit_set_row_order 1) against top down, graph replay of the whole step
timeout 300 python -m pytest tests/test_kernels_gpu.py -x -q -k "residual" > gpurun_out/r2j_res_tests.log 2>&1; echo "residual tests rc=$?"; tail -1 gpurun_out/r2j_res_tests.log
NVIT_ROW_ORDER_TEST=1 timeout 400 python scripts/step_ab.py --hook nvit_set_row_order --steps 15 --reps 4 --graph-only > gpurun_out/r2j_row_order_ab.log 2>&1; echo "ab rc=$?"; cat gpurun_out/r2j_row_order_ab.log
