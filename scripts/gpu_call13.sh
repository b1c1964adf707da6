#!/bin/bash
# A/B: TMEM loads of the persistent attention forward fetched a chunk pair ahead; residual backward staged-ring warp counts
timeout 300 python -m pytest tests/test_zz_attention_variants_gpu.py -x -q > gpurun_out/r2h_attn_tests.log 2>&1; echo "attn tests rc=$?"; tail -1 gpurun_out/r2h_attn_tests.log
timeout 200 python scripts/attn_bwd_time.py > gpurun_out/r2h_attn_time.log 2>&1; echo "time rc=$?"; grep -i "fwd" gpurun_out/r2h_attn_time.log
timeout 100 python scripts/attn_fwd_phases.py > gpurun_out/r2h_attn_fwd_phases.log 2>&1; grep -A3 "median CTA" gpurun_out/r2h_attn_fwd_phases.log
timeout 200 python scripts/residual_w_sweep.py > gpurun_out/r2h_residual_w.log 2>&1; echo "sweep rc=$?"; cat gpurun_out/r2h_residual_w.log
