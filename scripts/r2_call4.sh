#!/bin/bash
# GPU call 4 (2 GPUs): multi-rank pytest + C ABI client + sweep-mode L2 reuse probe on sphere
O=gpurun_out; mkdir -p $O
timeout 1200 python -m pytest tests/test_gpu_multirank.py tests/test_abi_c.py tests/test_gpu_parity.py tests/test_gpu_post.py -q > $O/r2c4_pytest.log 2>&1; echo "rc=$?" >> $O/r2c4_pytest.log
tail -6 $O/r2c4_pytest.log
run() { tag=$1; shift
  env "$@" timeout 300 python bench.py --workload $W --steps 100 --warmup 10 --no-cpu --no-extra > $O/r2c4_${W}_$tag.json 2> $O/r2c4_${W}_$tag.err
  python - <<PY
import json
try:
    j=json.loads(open("$O/r2c4_${W}_$tag.json").read().strip().splitlines()[-1])
    print("$W $tag", round(j["value"]), "e2e", round(j["e2e"]["value"]), [(k["name"].replace("step_kernel<f32,",""), round(k["total_ms"]/k["launches"],4)) for k in j["details"]["kernels"]])
except Exception as e:
    print("$W $tag failed", e); print(open("$O/r2c4_${W}_$tag.err").read()[-600:])
PY
}
W=sphere
run base KHR_TMA=0
run sweep_l1 KHR_SWEEP=1 KHR_SWEEP_LAG=1
run sweep_l1_z4 KHR_SWEEP=1 KHR_SWEEP_LAG=1 KHR_ZSEG=4
run sweep_l2_z4 KHR_SWEEP=1 KHR_SWEEP_LAG=2 KHR_ZSEG=4
run sweep_l2_z2 KHR_SWEEP=1 KHR_SWEEP_LAG=2 KHR_ZSEG=2
run sweep_l3_z2 KHR_SWEEP=1 KHR_SWEEP_LAG=3 KHR_ZSEG=2
run z4 KHR_ZSEG=4
W=waveguide_mode
run base KHR_TMA=0
run sweep_l2 KHR_SWEEP=1 KHR_SWEEP_LAG=2
run sweep_l2_z4 KHR_SWEEP=1 KHR_SWEEP_LAG=2 KHR_ZSEG=4
