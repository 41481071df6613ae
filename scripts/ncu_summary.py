#!/usr/bin/env python
"""Summarise an `ncu --set full` report of the step kernels (run here, no GPU needed):

  python scripts/ncu_summary.py gpurun_out/X.ncu-rep profiles/r01_ncu_X.txt [profiles/traffic_<workload>.json]

Writes one block per captured launch (duration, DRAM bytes, throughput, registers, occupancy,
issue utilisation, top stall reasons) and, optionally, the per-launch DRAM traffic
(dram__bytes_read.sum + dram__bytes_write.sum, averaged over the captured launches of the
same kernel) keyed by the kernel names bench.py reports, which bench.py copies into
`roofline.traffic`.
"""
import csv
import io
import json
import re
import subprocess
import sys

WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "launch__registers_per_thread",
        "launch__grid_size", "launch__occupancy_limit_registers", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct", "sm__cycles_active.avg", "sm__cycles_elapsed.max"]
UNIT = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}


def bench_name(kname):
    m = re.search(r"pml_tma_kernel<(float|double), (?:\(int\))?(\d)", kname)
    if m:   # the persistent TMA half-step kernel: bench.py's name carries a material suffix, matched by prefix there
        return "halfstep_tma_kernel<%s,%s,interior+pml" % ("f32" if m.group(1) == "float" else "f64", "HE"[int(m.group(2))])
    m = re.search(r"step_kernel<(float|double), (?:\(int\))?(\d), (?:\(int\))?(\d), (?:\(int\)|\(bool\))?(\d), (?:\(int\))?(\d)"
                  r"(?:, (?:\(bool\))?(\d))?>", kname)
    if not m:
        return None
    t, g, mode, marr, axm = m.group(1), int(m.group(2)), int(m.group(3)), int(m.group(4)), int(m.group(5))
    mn = {0: "interior", 2: "full"}.get(mode) or {1: "pml-x", 2: "pml-y", 4: "pml-z"}.get(axm, "pml")
    return "step_kernel<%s,%s,%s,%s>" % ("f32" if t == "float" else "f64", "HE"[g], mn, ("mscalar", "marr", "muniform")[marr])


def main():
    rep, out_txt = sys.argv[1], sys.argv[2]
    out_json = sys.argv[3] if len(sys.argv) > 3 else None
    # optional 4th argument: a bench.py JSON line of the same workload; its per-kernel work-item counts identify the launch
    # shape of the persistent TMA kernel (whose grid is always one CTA per SM)
    items = {}
    if len(sys.argv) > 4:
        try:
            bj = json.loads(open(sys.argv[4]).read().strip().splitlines()[-1])
            items = {k["name"]: int(k["ctas"]) for k in bj["details"]["kernels"]}
        except Exception:
            items = {}
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    ik = hdr.index("Kernel Name")
    stall = [i for i, h in enumerate(hdr) if "warps_issue_stalled" in h and h.endswith("_per_issue_active.ratio")]
    traffic = {}
    grids = {}
    with open(out_txt, "w") as f:
        f.write("# %s  (ncu --set full --clock-control none; per-launch, cold cache, serialised)\n" % rep)
        for r in rows[2:]:
            f.write("%s\n" % r[ik])
            for w in WANT:
                if w in hdr:
                    i = hdr.index(w)
                    f.write("    %-62s %s %s\n" % (w, r[i], units[i]))
            top = sorted(((float(r[i].replace(",", "")) if r[i] else 0.0, hdr[i]) for i in stall), reverse=True)[:4]
            f.write("    top stalls (per issue): %s\n" % ", ".join(
                "%s=%.2f" % (h.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", ""), v)
                for v, h in top))
            bn = bench_name(r[ik])
            if bn:
                tot = 0.0
                for w in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
                    i = hdr.index(w)
                    tot += float(r[i].replace(",", "")) * UNIT.get(units[i], 1.0)
                traffic.setdefault(bn, []).append(tot)
                ig = hdr.index("launch__grid_size") if "launch__grid_size" in hdr else -1
                if ig >= 0:
                    grids.setdefault(bn, []).append(int(float(r[ig].replace(",", ""))))
    if out_json:
        with open(out_json, "w") as f:
            # per kernel: DRAM bytes per launch and the launch's grid size (bench.py uses the entry only for the same launch shape)
            out = {}
            for k, v in traffic.items():
                out[k] = {"bytes": sum(v) / len(v), "grid": max(grids.get(k, [0])), "captures": len(v)}
                it = [n for name, n in items.items() if name.startswith(k)]
                if it:
                    out[k]["items"] = it[0]
            json.dump(out, f, indent=1)
    print("wrote", out_txt, out_json or "")


if __name__ == "__main__":
    main()
