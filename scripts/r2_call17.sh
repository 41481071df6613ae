#!/bin/bash
O=gpurun_out; mkdir -p $O
run() { tag=$1; shift
  env "$@" timeout 200 python bench.py --workload $W --steps 400 --warmup 10 --no-cpu --no-extra > $O/r2c17_${W}_$tag.json 2> $O/r2c17_${W}_$tag.err
  python - <<PY
import json
try:
    j=json.loads(open("$O/r2c17_${W}_$tag.json").read().strip().splitlines()[-1])
    print("$W $tag", round(j["value"]), "e2e", round(j["e2e"]["value"]))
except Exception as e:
    print("$W $tag failed", e); print(open("$O/r2c17_${W}_$tag.err").read()[-400:])
PY
}
for W in uled waveguide_mode; do
run base
run order1 KHR_ORDER=1
run order2 KHR_ORDER=2
run noprio KHR_STREAM_PRIO=0
run order1_noprio KHR_ORDER=1 KHR_STREAM_PRIO=0
done
