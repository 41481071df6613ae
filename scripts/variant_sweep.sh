#!/bin/bash
# usage: scripts/variant_sweep.sh  (runs on the GPU box) — bench each library variant
P='import sys,json; d=json.loads(sys.stdin.read().strip().split(chr(10))[-1]); print("%.0f Mcells/s %.4f ms/step e2e %.0f | " % (d["value"], d["ms_per_step"], d["e2e"]["value"]) + " ".join("%s=%.4f" % (k["name"].split("<")[1][:12], k["total_ms"]/max(k["launches"],1)) for k in d["kernels"]))'
for v in default $(ls khronos.jl_b200/lib/variants/ | sed 's/.so//'); do
  if [ $v = default ]; then unset KHRONOS_B200_LIB; else export KHRONOS_B200_LIB=$PWD/khronos.jl_b200/lib/variants/$v.so; fi
  echo "== $v waveguide"; python bench.py --steps 200 --warmup 10 --no-cpu 2>&1 | python -c "$P"
  echo "== $v sphere"; python bench.py --workload sphere --steps 40 --warmup 5 --no-cpu 2>&1 | python -c "$P"
done
