#!/bin/bash
timeout 150 python -m pytest tests/test_zz_attention_variants_gpu.py -m gpu -x -q > gpurun_out/r2_t2b.log 2>&1; echo "variants rc=$?"; tail -4 gpurun_out/r2_t2b.log
timeout 60 python scripts/attn_bwd_time.py > gpurun_out/r2_attn_time.log 2>&1; echo "time rc=$?"; cat gpurun_out/r2_attn_time.log
timeout 60 python scripts/attn_bwd_phases_v2.py > gpurun_out/r2_attn_phases_v2.log 2>&1; echo "phases rc=$?"; cat gpurun_out/r2_attn_phases_v2.log
timeout 150 python -m pytest tests/test_kohonen_gpu.py -m gpu -x -q > gpurun_out/r2_t2a.log 2>&1; echo "kohonen rc=$?"; tail -3 gpurun_out/r2_t2a.log
