#!/bin/bash
# final evidence of round 2: full GPU suite, smoke, launch list of one step with the final kernels, ncu --set full of the
# persistent attention forward and of the c_fc + gate forward GEMM, bench line of record
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/r2g_tests.log 2>&1; echo "tests rc=$?"; tail -1 gpurun_out/r2g_tests.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2g_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/r2g_smoke.log
M=gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,dram__bytes_read.sum,dram__bytes_write.sum
timeout 120 python scripts/profile_step.py > gpurun_out/r2g_plain_step.log 2>&1 && timeout 500 ncu --profile-from-start off --metrics $M --clock-control none --csv --log-file gpurun_out/launches_r2g.csv python scripts/profile_step.py > gpurun_out/r2g_ncu_step.log 2>&1; echo "ncu list rc=$?"; tail -1 gpurun_out/r2g_ncu_step.log
timeout 300 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:attn_fwd_ws -s 1 -c 1 -o gpurun_out/prof_r2_attn_fwd_ws python scripts/profile_step.py > gpurun_out/r2g_ncu_fwd.log 2>&1; echo "ncu fwd rc=$?"
timeout 300 ncu --profile-from-start off --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:gemm_tcgen05_kernel<256, 0, 0, 1" -s 1 -c 1 -o gpurun_out/prof_r2_swiglu python scripts/profile_step.py > gpurun_out/r2g_ncu_swiglu.log 2>&1; echo "ncu swiglu rc=$?"
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r2g_bench_default.log 2>&1; echo "bench rc=$?"; tail -1 gpurun_out/r2g_bench_default.log | cut -c1-300
ls -la gpurun_out/*.ncu-rep 2>/dev/null | tail -3
