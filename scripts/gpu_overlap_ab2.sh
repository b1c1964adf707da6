#!/bin/bash
N=$1
run() { tag=$1; shift; timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 400)) bench.py --gpus $N --steps 20 --warmup 5 --no-cpu-baseline --no-e2e --no-dp-parity "$@" > gpurun_out/r2_ov_${N}_${tag}.log 2>&1; echo "$tag rc=$?"; grep '^{' gpurun_out/r2_ov_${N}_${tag}.log | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print('   ', round(d['value'],1), 'img/s', round(d['ms_per_step'],2), 'ms', d['clocks']['sm_mhz'], 'MHz')"; }
run single_b
run overlap_r16 --overlap --overlap-reserve-sms 16 --overlap-reserve-calls 6
run overlap_r32 --overlap --overlap-reserve-sms 32 --overlap-reserve-calls 8
run overlap_r32b2 --overlap --overlap-reserve-sms 32 --overlap-reserve-calls 8 --bucket-blocks 3
