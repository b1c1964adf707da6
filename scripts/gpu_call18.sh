#!/bin/bash
timeout 100 scripts/probes/tmem_read_bw.bin 2>&1 | tail -6 | tee gpurun_out/r2n_tmem_probe.log
timeout 300 python -m pytest tests/test_zz_attention_variants_gpu.py -x -q > gpurun_out/r2n_attn_tests.log 2>&1; echo "attn tests rc=$?"; tail -1 gpurun_out/r2n_attn_tests.log
timeout 200 python scripts/attn_bwd_time.py > gpurun_out/r2n_attn_time.log 2>&1; echo "time rc=$?"; grep -i "fwd" gpurun_out/r2n_attn_time.log
