#!/bin/bash
# GPU call of round 2: tests, smoke, attention timing, bench (old / new attention backward), unfused-tail A/B
python -m pytest tests -m gpu -x -q > gpurun_out/r2_t1.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/r2_t1.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_smoke1.log 2>&1; echo "smoke rc=$?"
timeout 120 python scripts/attn_bwd_time.py > gpurun_out/r2_attn_time.log 2>&1; echo "attn time rc=$?"; cat gpurun_out/r2_attn_time.log
python bench.py --steps 10 --warmup 3 > gpurun_out/r2_bench1.log 2>&1; echo "bench rc=$?"
NVIT_ATTN_BWD_VARIANT=2 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench1_v2.log 2>&1; echo "bench v2 rc=$?"
python bench.py --steps 10 --warmup 3 --unfused-tail --no-cpu-baseline --no-e2e > gpurun_out/r2_bench1_unfused.log 2>&1
