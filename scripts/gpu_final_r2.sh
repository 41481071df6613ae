#!/bin/bash
# Round 2 final 1-GPU evidence: full capture of the dominant kernels first (its DRAM bytes feed roofline.traffic of the
# bench below), parity suite, smoke, both bench arms as the driver runs them, ncu launch list of the same command.
TAG=${1:-r02fin}
O=gpurun_out; mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $O/${TAG}_smi.txt 2>&1
timeout 200 python bench.py --steps 5 --warmup 3 --no-cpu --no-extra > $O/${TAG}_bench_shape.json 2> /dev/null
timeout 300 ncu --set full --clock-control none --import-source on -k regex:pml_tma_kernel -s 20 -c 2 -f -o $O/${TAG}_full_sphere \
  python bench.py --steps 5 --warmup 3 --no-cpu --no-extra > $O/${TAG}_ncu_full.log 2>&1
python scripts/ncu_summary.py $O/${TAG}_full_sphere.ncu-rep $O/${TAG}_ncu_full_sphere.txt profiles/traffic_sphere.json $O/${TAG}_bench_shape.json > /dev/null 2>&1
cp profiles/traffic_sphere.json $O/${TAG}_traffic_sphere.json
timeout 600 python -m pytest tests -m gpu -q --timeout 240 > $O/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" >> $O/${TAG}_pytest.log
timeout 200 python __graft_entry__.py smoke > $O/${TAG}_smoke.log 2>&1; echo "smoke rc=$?" >> $O/${TAG}_smoke.log
timeout 300 python bench.py --impl reference --steps 20 --warmup 5 > $O/${TAG}_bench_reference.json 2> $O/${TAG}_bench_reference.err
timeout 600 python bench.py --steps 20 --warmup 5 > $O/${TAG}_bench_default.json 2> $O/${TAG}_bench_default.err
timeout 300 python bench.py --no-cpu --no-extra > $O/${TAG}_bench_sphere_200.json 2> $O/${TAG}_bench_sphere_200.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/${TAG}_launches_sphere.csv \
  python bench.py --steps 20 --warmup 5 --no-cpu --no-extra > $O/${TAG}_ncu_launch.log 2>&1
tail -3 $O/${TAG}_pytest.log; tail -1 $O/${TAG}_smoke.log; for w in reference default sphere_200; do cut -c1-400 $O/${TAG}_bench_$w.json; echo; done
