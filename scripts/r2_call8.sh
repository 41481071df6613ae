#!/bin/bash
O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_modes.py -q -x -k "TMA or tma" > $O/r2c8_modes.log 2>&1; echo "rc=$?" >> $O/r2c8_modes.log
tail -8 $O/r2c8_modes.log
run() { tag=$1; shift
  env "$@" timeout 300 python bench.py --workload $W --steps 200 --warmup 10 --no-cpu --no-extra > $O/r2c8_${W}_$tag.json 2> $O/r2c8_${W}_$tag.err
  python - <<PY
import json
try:
    j=json.loads(open("$O/r2c8_${W}_$tag.json").read().strip().splitlines()[-1])
    print("$W $tag", round(j["value"]), "e2e", round(j["e2e"]["value"]), "frac", round(j["roofline"]["frac"],3), "step", round(j["roofline"]["whole_step"]["frac"],3), [(k["name"].replace("step_kernel<f32,","")[:30], k["ctas"], round(k["total_ms"]/max(1,k["launches"]),4)) for k in j["details"]["kernels"] if k["launches"]])
except Exception as e:
    print("$W $tag failed", e); print(open("$O/r2c8_${W}_$tag.err").read()[-600:])
PY
}
for W in sphere waveguide_mode; do
run tma KHR_TMA=1
run fuse KHR_TMA=1 KHR_FUSE=1
run fuse_z4 KHR_TMA=1 KHR_FUSE=1 KHR_ZSEG=4
run fuse_z2 KHR_TMA=1 KHR_FUSE=1 KHR_ZSEG=2
run fuse_z2_l2 KHR_TMA=1 KHR_FUSE=1 KHR_ZSEG=2 KHR_FUSE_LAG=2
run fuse_z1 KHR_TMA=1 KHR_FUSE=1 KHR_ZSEG=1
run tma_z2 KHR_TMA=1 KHR_ZSEG=2
done
