#!/bin/bash
O=gpurun_out; mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -q -x > $O/r2c11_pytest.log 2>&1; echo "rc=$?" >> $O/r2c11_pytest.log
tail -6 $O/r2c11_pytest.log
run() { tag=$1; shift
  env "$@" timeout 300 python bench.py --workload $W --steps 400 --warmup 10 --no-cpu --no-extra > $O/r2c11_${W}_$tag.json 2> $O/r2c11_${W}_$tag.err
  python - <<PY
import json
try:
    j=json.loads(open("$O/r2c11_${W}_$tag.json").read().strip().splitlines()[-1])
    print("$W $tag", round(j["value"]), "e2e", round(j["e2e"]["value"]), "frac", round(j["roofline"]["frac"],3), "step", round(j["roofline"]["whole_step"]["frac"],3), [(k["name"].replace("step_kernel<f32,","")[:30], k["ctas"], round(k["total_ms"]/max(1,k["launches"]),4)) for k in j["details"]["kernels"] if k["launches"]])
except Exception as e:
    print("$W $tag failed", e); print(open("$O/r2c11_${W}_$tag.err").read()[-600:])
PY
}
for W in uled waveguide_mode; do
run base
run m2x2 KHRONOS_B200_LIB=$PWD/khronos.jl_b200/lib/variants/libkhr_m2x2.so
run zf1 KHR_ZSEG_FULL=1
run zf4 KHR_ZSEG_FULL=4
done
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29721 bench.py --gpus 2 --steps 20 --warmup 5 --no-extra > $O/r2c11_bench_2gpu.json 2> $O/r2c11_bench_2gpu.err
python - <<PY
import json
j=json.loads(open("$O/r2c11_bench_2gpu.json").read().strip().splitlines()[-1]); print("2gpu", round(j["value"]), "e2e", round(j["e2e"]["value"]), j["e2e"]["d2h_bytes_per_step"])
PY
