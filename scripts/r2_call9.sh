#!/bin/bash
O=gpurun_out; mkdir -p $O
run() { tag=$1; shift
  env "$@" timeout 300 python bench.py --workload $W --steps 200 --warmup 10 --no-cpu --no-extra > $O/r2c9_${W}_$tag.json 2> $O/r2c9_${W}_$tag.err
  python - <<PY
import json
try:
    j=json.loads(open("$O/r2c9_${W}_$tag.json").read().strip().splitlines()[-1])
    print("$W $tag", round(j["value"]), "e2e", round(j["e2e"]["value"]), "frac", round(j["roofline"]["frac"],3), "step", round(j["roofline"]["whole_step"]["frac"],3), [(k["name"].replace("step_kernel<f32,","")[:30], k["ctas"], round(k["total_ms"]/max(1,k["launches"]),4)) for k in j["details"]["kernels"] if k["launches"]])
except Exception as e:
    print("$W $tag failed", e); print(open("$O/r2c9_${W}_$tag.err").read()[-600:])
PY
}
for W in sphere dipole500; do
for z in 1 2 3 4 6; do run tma_z$z KHR_TMA=1 KHR_ZSEG=$z; done
run tma_z2_s3 KHR_TMA=1 KHR_ZSEG=2 KHR_TMA_STAGES=3
run tma_z2_t0 KHR_TMA=1 KHR_ZSEG=2 KHR_TMA_TAIL=0
done
W=metalens_full
run tma_z2 KHR_TMA=1 KHR_ZSEG=2
run ldg KHR_TMA=0
W=waveguide_mode
run tma_z4 KHR_TMA=1 KHR_ZSEG=4
run ldg_z4 KHR_TMA=0 KHR_ZSEG=4
