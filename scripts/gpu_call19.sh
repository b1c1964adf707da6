#!/bin/bash
# code of record: full GPU suite, smoke, launch list of one step, ncu --set full of the attention forward, bench line
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/r2q_tests.log 2>&1; echo "tests rc=$?"; tail -1 gpurun_out/r2q_tests.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2q_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/r2q_smoke.log
M=gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,dram__bytes_read.sum,dram__bytes_write.sum
timeout 120 python scripts/profile_step.py > gpurun_out/r2q_plain_step.log 2>&1 && timeout 500 ncu --profile-from-start off --metrics $M --clock-control none --csv --log-file gpurun_out/launches_r2q.csv python scripts/profile_step.py > gpurun_out/r2q_ncu_step.log 2>&1; echo "ncu list rc=$?"; tail -1 gpurun_out/r2q_ncu_step.log
timeout 300 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:attn_fwd_ws -s 1 -c 1 -o gpurun_out/prof_r2_attn_fwd_ws_v2 python scripts/profile_step.py > gpurun_out/r2q_ncu_fwd.log 2>&1; echo "ncu fwd rc=$?"
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r2q_bench_default.log 2>&1; echo "bench rc=$?"; tail -1 gpurun_out/r2q_bench_default.log | cut -c1-220
