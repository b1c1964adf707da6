#!/bin/bash
# usage: gpu_overlap_ab.sh N  -- data-parallel overlap A/B: one all-reduce after backward vs overlapped buckets, with / without the SM window
N=$1
run() { tag=$1; shift; timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 400)) bench.py --gpus $N --steps 20 --warmup 5 --no-cpu-baseline --no-e2e "$@" > gpurun_out/r2_ov_${N}_${tag}.log 2>&1; echo "$tag rc=$?"; grep '^{' gpurun_out/r2_ov_${N}_${tag}.log | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print('   ', round(d['value'],1), 'img/s', round(d['ms_per_step'],2), 'ms', d['clocks']['sm_mhz'], 'MHz', (d.get('dp_parity') or {}).get('rel_l2_grad_err'), (d.get('dp_parity') or {}).get('params_bit_identical_across_ranks'))"; }
run single
run overlap --overlap
run overlap_w8 --overlap --overlap-reserve-sms 8 --overlap-reserve-calls 4 --nccl-max-ctas 8
run overlap_w16 --overlap --overlap-reserve-sms 16 --overlap-reserve-calls 6 --nccl-max-ctas 16
run single_ctas8 --nccl-max-ctas 8
