#!/bin/bash
# N = 2 data-parallel checks: CUDA-graph capture of the step with the NCCL all-reduce, dp_parity, overlap A/B
run() { timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $1 bench.py --gpus 2 --steps 10 --warmup 3 "${@:3}" > gpurun_out/$2 2>&1; echo "$2 rc=$?"; grep -c '^{' gpurun_out/$2; }
run 29511 r2_dp2_graph.log
run 29512 r2_dp2_overlap_graph.log --overlap
tail -c 600 gpurun_out/r2_dp2_graph.log
