#!/bin/bash
timeout 200 python -m pytest tests/test_zz_attention_variants_gpu.py tests/test_kernels_gpu.py -m gpu -x -q -k "attention" > gpurun_out/r2_t3b.log 2>&1; echo "attention tests rc=$?"; tail -2 gpurun_out/r2_t3b.log
timeout 90 python scripts/attn_bwd_time.py > gpurun_out/r2_attn_time.log 2>&1; cat gpurun_out/r2_attn_time.log
