#!/bin/bash
O=gpurun_out; mkdir -p $O
# ncu full captures: TMA half-step kernels on dipole500 and sphere; LDG kernels on sphere (before/after evidence)
KHR_TMA=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:pml_tma_kernel -s 20 -c 2 -f -o $O/r2c6_tma_dipole500 \
  python bench.py --workload dipole500 --steps 5 --warmup 3 --no-cpu --no-extra > $O/r2c6_ncu_dipole.log 2>&1
KHR_TMA=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:pml_tma_kernel -s 20 -c 2 -f -o $O/r2c6_tma_sphere \
  python bench.py --workload sphere --steps 5 --warmup 3 --no-cpu --no-extra > $O/r2c6_ncu_sphere_tma.log 2>&1
KHR_TMA=0 timeout 600 ncu --set full --clock-control none --import-source on -k regex:step_kernel -s 60 -c 6 -f -o $O/r2c6_ldg_sphere \
  python bench.py --workload sphere --steps 5 --warmup 3 --no-cpu --no-extra > $O/r2c6_ncu_sphere_ldg.log 2>&1
ls -la $O/*.ncu-rep | tail -5
