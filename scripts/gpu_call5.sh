#!/bin/bash
timeout 300 python -m pytest tests -m gpu -x -q > gpurun_out/r2_t5.log 2>&1; echo "all tests rc=$?"; tail -2 gpurun_out/r2_t5.log
timeout 120 python scripts/profile_step.py > gpurun_out/plain_step.log 2>&1 && timeout 400 ncu --profile-from-start off --metrics gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/launches_r2.csv python scripts/profile_step.py > gpurun_out/ncu_step.log 2>&1; echo "ncu rc=$?"; tail -2 gpurun_out/ncu_step.log
