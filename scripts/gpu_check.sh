#!/bin/bash
# Lean GPU round (1 GPU): parity tests, smoke, default bench.  Outputs in gpurun_out/<tag>_*.
TAG=${1:-chk}
O=gpurun_out
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $O/${TAG}_smi.txt 2>&1
timeout 1200 python -m pytest tests -m gpu -q -x --durations=15 ${PYTEST_ARGS} > $O/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" >> $O/${TAG}_pytest.log
timeout 300 python __graft_entry__.py smoke > $O/${TAG}_smoke.log 2>&1; echo "smoke rc=$?" >> $O/${TAG}_smoke.log
timeout 600 python bench.py > $O/${TAG}_bench_waveguide.json 2> $O/${TAG}_bench_waveguide.err
tail -25 $O/${TAG}_pytest.log; tail -2 $O/${TAG}_smoke.log; cut -c1-400 $O/${TAG}_bench_waveguide.json
