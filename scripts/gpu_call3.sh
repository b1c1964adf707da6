#!/bin/bash
timeout 400 python -m pytest tests -m gpu -x -q > gpurun_out/r2_t3.log 2>&1; echo "all tests rc=$?"; tail -3 gpurun_out/r2_t3.log
timeout 60 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/r2_smoke.log
timeout 90 python scripts/attn_bwd_time.py > gpurun_out/r2_attn_time.log 2>&1; grep variant gpurun_out/r2_attn_time.log
