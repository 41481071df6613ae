#!/bin/bash
O=gpurun_out; mkdir -p $O
run() { tag=$1; shift
  env "$@" timeout 200 python bench.py --workload $W --steps 400 --warmup 10 --no-cpu --no-extra > $O/r2c19_${W}_$tag.json 2> $O/r2c19_${W}_$tag.err
  python - <<PY
import json
try:
    j=json.loads(open("$O/r2c19_${W}_$tag.json").read().strip().splitlines()[-1])
    print("$W $tag", round(j["value"]), "e2e", round(j["e2e"]["value"]), "step", round(j["roofline"]["whole_step"]["frac"],3), [(k["name"].replace("step_kernel<f32,","")[:26], k["ctas"], round(k["total_ms"]/max(1,k["launches"]),4)) for k in j["details"]["kernels"] if k["launches"]])
except Exception as e:
    print("$W $tag failed", e); print(open("$O/r2c19_${W}_$tag.err").read()[-400:])
PY
}
V=$PWD/khronos.jl_b200/lib/variants
for W in waveguide_mode uled; do
run ldg KHR_TMA=0
run tma_r112 KHR_TMA=1
run tma_r112_z8 KHR_TMA=1 KHR_ZSEG=8
run tma_noreg KHR_TMA=1 KHRONOS_B200_LIB=$V/libkhr_tma_noreg.so
run tma_r104 KHR_TMA=1 KHRONOS_B200_LIB=$V/libkhr_tma_r104.so
done
for W in sphere dipole500; do
run tma_r112
run tma_noreg KHRONOS_B200_LIB=$V/libkhr_tma_noreg.so
done
