#!/bin/bash
timeout 200 python -m pytest tests/test_zz_attention_variants_gpu.py -m gpu -x -q > gpurun_out/r2_t3b.log 2>&1; echo "variants rc=$?"; tail -2 gpurun_out/r2_t3b.log
timeout 90 python scripts/attn_bwd_time.py > gpurun_out/r2_attn_time.log 2>&1; echo "time rc=$?"; cat gpurun_out/r2_attn_time.log
timeout 300 python -m pytest tests -m gpu -x -q > gpurun_out/r2_t3.log 2>&1; echo "all tests rc=$?"; tail -3 gpurun_out/r2_t3.log
python bench.py --steps 10 --warmup 3 > gpurun_out/r2_bench3.log 2>&1; echo "bench rc=$?"
