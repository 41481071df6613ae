#!/bin/bash
O=gpurun_out; mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -q -x > $O/r2c14_pytest.log 2>&1; echo "rc=$?" >> $O/r2c14_pytest.log
tail -8 $O/r2c14_pytest.log
run() { tag=$1; shift
  env "$@" timeout 300 python bench.py --workload $W --steps 400 --warmup 10 --no-cpu --no-extra > $O/r2c14_${W}_$tag.json 2> $O/r2c14_${W}_$tag.err
  python - <<PY
import json
try:
    j=json.loads(open("$O/r2c14_${W}_$tag.json").read().strip().splitlines()[-1])
    print("$W $tag", round(j["value"]), "e2e", round(j["e2e"]["value"]), "graph", j["details"]["cuda_graph"], "launches", j["gpu_launches"])
except Exception as e:
    print("$W $tag failed", e); print(open("$O/r2c14_${W}_$tag.err").read()[-600:])
PY
}
for W in waveguide_mode uled sphere; do
run graph
run nograph KHR_GRAPH=0
done
