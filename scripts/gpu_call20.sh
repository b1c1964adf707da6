#!/bin/bash
# A/B of the remote mbarrier arrive of the pair-mode GEMM epilogues: release.cluster (old library) against relaxed.cluster
for lib in libnvit_b200_hooks_old.so libnvit_b200_hooks.so libnvit_b200_hooks_old.so libnvit_b200_hooks.so; do
  echo "== $lib"
  NVIT_LIB_PATH=$PWD/nvit_b200/$lib DBG_MODES=0 CG_MODES=0 timeout 200 python scripts/gemm_bench.py 2>&1 | tail -11
done > gpurun_out/r2s_arrive_ab.log 2>&1
cat gpurun_out/r2s_arrive_ab.log
timeout 400 python -m pytest tests/test_kernels_gpu.py -x -q -k "gemm or linear or swiglu or gate" > gpurun_out/r2s_gemm_tests.log 2>&1; echo "gemm tests rc=$?"; tail -1 gpurun_out/r2s_gemm_tests.log
