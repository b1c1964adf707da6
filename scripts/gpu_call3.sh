#!/bin/bash
timeout 120 python scripts/attn_bwd_time.py > gpurun_out/r2_attn_time.log 2>&1; cat gpurun_out/r2_attn_time.log | tail -5
