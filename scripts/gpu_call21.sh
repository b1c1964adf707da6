#!/bin/bash
# same box, alternating libraries: remote arrive of the pair-mode GEMM epilogues release.cluster (old) against relaxed (new, gate backward keeps release)
for lib in libnvit_b200_old.so libnvit_b200.so libnvit_b200_old.so libnvit_b200.so; do
  echo "== $lib"
  NVIT_LIB_PATH=$PWD/nvit_b200/$lib timeout 300 python scripts/step_ab.py --hook nvit_set_pdl --modes 0,0 --reps 1 --graph-only --steps 20 2>&1 | tail -2
done > gpurun_out/r2t_arrive_step_ab.log 2>&1
cat gpurun_out/r2t_arrive_step_ab.log
