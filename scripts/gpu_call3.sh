#!/bin/bash
timeout 300 python -m pytest tests/test_kernels_gpu.py -m gpu -x -q -k "gate or swiglu" > gpurun_out/r2_t3b.log 2>&1; echo "gate tests rc=$?"; tail -2 gpurun_out/r2_t3b.log
timeout 100 python scripts/one_gemm.py "dgrad+gate" 2>&1 | tail -1
timeout 100 python scripts/one_gemm.py "c_fc swiglu" 2>&1 | tail -1
timeout 300 python -m pytest tests/test_model_gpu.py tests/test_kohonen_gpu.py -m gpu -x -q > gpurun_out/r2_t3c.log 2>&1; echo "model tests rc=$?"; tail -2 gpurun_out/r2_t3c.log
