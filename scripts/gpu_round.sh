#!/bin/bash
# Runs on the GPU box (1 GPU): parity tests, smoke, both bench arms, ncu launch list and
# one full ncu capture of a time step.  Everything lands in gpurun_out/<tag>_*.
TAG=${1:-r01}
O=gpurun_out
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $O/${TAG}_smi.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > $O/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" >> $O/${TAG}_pytest.log
timeout 300 python __graft_entry__.py smoke > $O/${TAG}_smoke.log 2>&1; echo "smoke rc=$?" >> $O/${TAG}_smoke.log
timeout 900 python bench.py > $O/${TAG}_bench_waveguide.json 2> $O/${TAG}_bench_waveguide.err
timeout 600 python bench.py --workload sphere --steps 200 --warmup 10 --no-cpu > $O/${TAG}_bench_sphere.json 2> $O/${TAG}_bench_sphere.err
timeout 600 python bench.py --workload uled --steps 1000 --warmup 10 --no-cpu > $O/${TAG}_bench_uled.json 2> $O/${TAG}_bench_uled.err
timeout 600 python bench.py --impl reference --steps 5 --warmup 3 > $O/${TAG}_bench_reference.json 2> $O/${TAG}_bench_reference.err
# launch list of the default bench command (cold-cache, serialised: shares only)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/${TAG}_launches_waveguide.csv \
  python bench.py --steps 5 --warmup 3 --no-cpu > $O/${TAG}_ncu_launch.log 2>&1
# full capture of one warmed-up time step (6 step kernels) for dram traffic
timeout 900 ncu --set full --clock-control none --import-source on -k regex:step_kernel -s 60 -c 6 -f -o $O/${TAG}_full_waveguide \
  python bench.py --steps 5 --warmup 3 --no-cpu > $O/${TAG}_ncu_full.log 2>&1
# sphere: full capture of one time step as well
timeout 900 ncu --set full --clock-control none --import-source on -k regex:step_kernel -s 60 -c 6 -f -o $O/${TAG}_full_sphere \
  python bench.py --workload sphere --steps 5 --warmup 3 --no-cpu > $O/${TAG}_ncu_full_sphere.log 2>&1
tail -3 $O/${TAG}_pytest.log; cat $O/${TAG}_smoke.log | tail -2; cat $O/${TAG}_bench_waveguide.json | cut -c1-600
