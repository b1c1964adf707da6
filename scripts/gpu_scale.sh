#!/bin/bash
# usage: gpu_scale.sh N tag [bench args...]
N=$1; TAG=$2; shift 2
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 400)) bench.py --gpus $N --steps 20 --warmup 5 "$@" > gpurun_out/r2_scale_${TAG}.log 2>&1; echo "$TAG rc=$?"; grep '^{' gpurun_out/r2_scale_${TAG}.log | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print(d['metric'], d['n_gpus'], round(d['value'],1), round(d['ms_per_step'],2), d['clocks'], d['scaling'], d.get('dp_parity'))"
