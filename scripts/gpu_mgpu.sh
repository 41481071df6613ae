#!/bin/bash
# N-GPU parity round: the z-slab path against the single-domain oracle, with the optional features
N=${1:-2}
O=gpurun_out
mkdir -p $O
for args in "" "--periodic" "--kerr" "--nonuniform" "--kerr --nonuniform"; do
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
    scripts/mgpu_parity.py $args 2>&1 | grep -E "mgpu parity|Error:" | tail -3
done | tee $O/mgpu${N}_parity.txt
