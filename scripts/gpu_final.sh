#!/bin/bash
# Final 1-GPU round: parity tests, smoke, bench arms, ncu launch list + one full capture of a time step.
TAG=${1:-fin}
O=gpurun_out; mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $O/${TAG}_smi.txt 2>&1
timeout 900 python -m pytest tests -m gpu -q > $O/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" >> $O/${TAG}_pytest.log
timeout 200 python __graft_entry__.py smoke > $O/${TAG}_smoke.log 2>&1; echo "smoke rc=$?" >> $O/${TAG}_smoke.log
# full capture of one warmed-up time step first: its DRAM bytes per launch feed roofline.traffic of the bench below
timeout 300 ncu --set full --clock-control none --import-source on -k regex:step_kernel -s 70 -c 7 -f -o $O/${TAG}_full_waveguide \
  python bench.py --steps 5 --warmup 3 --no-cpu > $O/${TAG}_ncu_full.log 2>&1
python scripts/ncu_summary.py $O/${TAG}_full_waveguide.ncu-rep $O/${TAG}_ncu_full_waveguide.txt profiles/traffic_waveguide_mode.json > /dev/null 2>&1
cp profiles/traffic_waveguide_mode.json $O/${TAG}_traffic_waveguide_mode.json
timeout 400 python bench.py > $O/${TAG}_bench_waveguide.json 2> $O/${TAG}_bench_waveguide.err
timeout 300 python bench.py --workload sphere --steps 200 --warmup 10 --no-cpu > $O/${TAG}_bench_sphere.json 2> $O/${TAG}_bench_sphere.err
timeout 300 python bench.py --workload uled --steps 1000 --warmup 10 --no-cpu > $O/${TAG}_bench_uled.json 2> $O/${TAG}_bench_uled.err
timeout 300 python bench.py --impl reference --steps 5 --warmup 3 > $O/${TAG}_bench_reference.json 2> $O/${TAG}_bench_reference.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/${TAG}_launches_waveguide.csv \
  python bench.py --steps 5 --warmup 3 --no-cpu > $O/${TAG}_ncu_launch.log 2>&1
tail -3 $O/${TAG}_pytest.log; tail -1 $O/${TAG}_smoke.log; cat $O/${TAG}_traffic_waveguide_mode.json; for w in waveguide sphere uled reference; do cut -c1-330 $O/${TAG}_bench_$w.json; echo; done
