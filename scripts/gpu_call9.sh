#!/bin/bash
# full GPU suite, smoke, bench with the attention backward of this session (variant 3) and the previous default (variant 2)
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2_t9.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/r2_t9.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_smoke9.log 2>&1; echo "smoke rc=$?"
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r2_bench9.log 2>&1; echo "bench rc=$?"; tail -1 gpurun_out/r2_bench9.log | cut -c1-600
NVIT_ATTN_BWD_VARIANT=2 timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2_bench9_v2.log 2>&1; echo "bench v2 rc=$?"; tail -1 gpurun_out/r2_bench9_v2.log | cut -c1-300
