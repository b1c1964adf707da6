#!/bin/bash
# variant 3 of the attention backward: parity, timing, phase marks
timeout 900 python -m pytest tests/test_zz_attention_variants_gpu.py -x -q 2>&1 | tail -15
timeout 300 python scripts/attn_bwd_time.py 2>&1 | tee gpurun_out/attn_bwd_time_v3.log
timeout 200 python scripts/attn_bwd_phases_v3.py 2>&1 | tee gpurun_out/attn_bwd_phases_v3.log
