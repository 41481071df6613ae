#!/bin/bash
O=gpurun_out; mkdir -p $O
timeout 600 ncu --set full --clock-control none --import-source on -k regex:step_kernel -s 48 -c 6 -f -o $O/r2c13_uled_full \
  python bench.py --workload uled --steps 5 --warmup 3 --no-cpu --no-extra > $O/r2c13_ncu.log 2>&1
tail -2 $O/r2c13_ncu.log; ls -la $O/r2c13_uled_full.ncu-rep
