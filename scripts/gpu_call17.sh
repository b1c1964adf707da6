#!/bin/bash
# persistent attention forward with alternating q tiles: parity cases, timing, phase marks
timeout 300 python -m pytest tests/test_zz_attention_variants_gpu.py -x -q > gpurun_out/r2l_attn_tests.log 2>&1; echo "attn tests rc=$?"; tail -1 gpurun_out/r2l_attn_tests.log
timeout 200 python scripts/attn_bwd_time.py > gpurun_out/r2l_attn_time.log 2>&1; echo "time rc=$?"; grep -i "fwd" gpurun_out/r2l_attn_time.log
timeout 100 python scripts/attn_fwd_phases.py > gpurun_out/r2l_attn_fwd_phases.log 2>&1; cat gpurun_out/r2l_attn_fwd_phases.log | tail -14
