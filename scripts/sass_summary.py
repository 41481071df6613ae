#!/usr/bin/env python
"""Static SASS summary of the shipped library (run here, no GPU needed):
   python scripts/sass_summary.py [lib.so] > profiles/r02_sass_mnemonics.txt
Per kernel: total instructions and the counts that prove how data moves — UTMALDG (cp.async.bulk.tensor), SYNCS (mbarrier),
LDS / STS, LDG.E.128 / STG.E.128, CCTL prefetches, SHFL, local-memory spills — and the FP32 arithmetic share."""
import collections
import re
import subprocess
import sys

lib = sys.argv[1] if len(sys.argv) > 1 else "khronos.jl_b200/lib/libkhronos_b200.so"
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
names = {}
cur = None
counts = collections.OrderedDict()
for ln in out.splitlines():
    m = re.search(r"Function : (\S+)", ln)
    if m:
        cur = m.group(1)
        counts[cur] = collections.Counter()
        continue
    m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P[0-9T]+ )?([A-Z0-9_.]+)", ln)
    if m and cur:
        op = m.group(1)
        c = counts[cur]
        c["total"] += 1
        base = op.split(".")[0]
        if base == "UTMALDG": c["UTMALDG"] += 1
        elif base == "SYNCS": c["SYNCS(mbarrier)"] += 1
        elif base == "LDS": c["LDS.128" if ".128" in op else "LDS"] += 1
        elif base == "STS": c["STS"] += 1
        elif base == "LDG": c["LDG.128" if ".128" in op else ("LDG.64" if ".64" in op else "LDG.32")] += 1
        elif base == "STG": c["STG.128" if ".128" in op else "STG.32/64"] += 1
        elif base == "CCTL": c["CCTL(prefetch)"] += 1
        elif base == "SHFL": c["SHFL"] += 1
        elif base in ("LDL", "STL"): c["local(spill)"] += 1
        elif base in ("FADD", "FMUL", "FFMA"): c["fp32 " + base] += 1
        elif base in ("DADD", "DMUL", "DFMA"): c["fp64"] += 1
        elif base in ("ATOMG", "RED"): c["atomics"] += 1
dem = subprocess.run(["cu++filt"] + list(counts), capture_output=True, text=True).stdout.splitlines()
print("# cuobjdump -sass %s (sm_100a), static instruction counts per kernel (whole body, all paths)" % lib)
for (mangled, c), d in zip(counts.items(), dem):
    if not any(k in d for k in ("step_kernel", "pml_tma_kernel", "step_tma_kernel", "sweep_kernel", "dft_kernel")):
        continue
    if "double" in d:
        continue
    print("%-78s total=%5d  %s" % (d[:78], c["total"], "  ".join("%s=%d" % (k, v) for k, v in sorted(c.items()) if k != "total")))
