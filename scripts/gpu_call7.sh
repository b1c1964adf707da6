#!/bin/bash
timeout 100 python scripts/one_gemm.py "dgrad+gate" > gpurun_out/plain_gateb.log 2>&1 && timeout 300 ncu --set full --clock-control none --import-source on -k regex:gemm_tcgen05 -s 3 -c 1 -o gpurun_out/prof_r2_gateb2 python scripts/one_gemm.py "dgrad+gate" > gpurun_out/ncu_gateb.log 2>&1; echo "ncu gateb rc=$?"; cat gpurun_out/plain_gateb.log
