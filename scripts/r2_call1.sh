#!/bin/bash
# Round 2, GPU call 1 (2 GPUs): parity suite incl. full-size + multi-rank tests, smoke, default bench, 2-GPU bench.
O=gpurun_out; mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $O/r2c1_smi.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -q -x > $O/r2c1_pytest.log 2>&1; echo "pytest rc=$?" >> $O/r2c1_pytest.log
timeout 200 python __graft_entry__.py smoke > $O/r2c1_smoke.log 2>&1; echo "smoke rc=$?" >> $O/r2c1_smoke.log
timeout 900 python bench.py --steps 20 --warmup 5 > $O/r2c1_bench_default.json 2> $O/r2c1_bench_default.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29701 bench.py --gpus 2 --steps 20 --warmup 5 > $O/r2c1_bench_2gpu.json 2> $O/r2c1_bench_2gpu.err
tail -5 $O/r2c1_pytest.log; tail -2 $O/r2c1_smoke.log; cut -c1-1500 $O/r2c1_bench_default.json; echo; tail -3 $O/r2c1_bench_default.err; cut -c1-1200 $O/r2c1_bench_2gpu.json; tail -3 $O/r2c1_bench_2gpu.err
