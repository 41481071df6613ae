#!/bin/bash
# post-processing kernels: throughput + one ncu --set full capture (summarised off-box with scripts/ncu_summary.py)
O=gpurun_out; mkdir -p $O
timeout 200 python scripts/post_bench.py | tee $O/post_bench.json
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"near2far_kernel|mode_overlap_kernel|flux_kernel" -c 3 -f -o $O/post_kernels python scripts/post_bench.py 12 8 > $O/ncu_post.log 2>&1
tail -3 $O/ncu_post.log
