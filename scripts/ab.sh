#!/bin/bash
# usage (GPU box): scripts/ab.sh "<variant list>" "<env settings list, ';'-separated>" -- A/B of library variants
P='import sys,json; d=json.loads(sys.stdin.read().strip().split(chr(10))[-1]); print("%.0f Mcells/s %.4f ms/step (prof %.4f) e2e %.0f | " % (d["value"], d["ms_per_step"], d.get("ms_per_step_serialised_with_kernel_events",0), d["e2e"]["value"]) + " ".join("%s=%.4f" % (k["name"].split("<")[1][:14], k["total_ms"]/max(k["launches"],1)) for k in d["kernels"]))'
VARS=${1:-default}
WL=${2:-"waveguide_mode sphere uled"}
for v in $VARS; do
  if [ $v = default ]; then unset KHRONOS_B200_LIB; else export KHRONOS_B200_LIB=$PWD/khronos.jl_b200/lib/variants/$v.so; fi
  for w in $WL; do
    st=300; [ $w = sphere ] && st=40
    echo "== $v $w $KHR_ENV_NOTE"; python bench.py --workload $w --steps $st --warmup 10 --no-cpu 2>&1 | python -c "$P"
  done
done
