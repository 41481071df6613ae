import sys, os
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [R, os.path.join(R, "oracle"), os.path.join(R, "tests")]
import numpy as np
import khronos_b200 as kb
from khronos_b200 import workloads as w
from bridge import oracle_from_simulation
d = w.uled(res=12)
for ob in d["geometry"]:
    ob.material.susceptibilities = []
sim = w.build_simulation(d, np.float32)
o, mids = oracle_from_simulation(sim)
sim.prepare_simulation()
sim.step(150); sim.sync(); o.step(150)
print("decimation", [m.decimation for m in sim.dft_monitors][:3], "dt", sim.grid.dt)
for i, (m, mid) in enumerate(zip(sim.dft_monitors, mids)):
    a, b = sim.get_dft(m), o.get_dft(mid)
    per_f = [np.linalg.norm(a[..., k] - b[..., k]) / max(np.linalg.norm(b[..., k]), 1e-300) for k in range(a.shape[-1])]
    print(i, "comp", m.component, "shape", a.shape, "norm %.3e" % np.linalg.norm(b), "err/f", " ".join("%.1e" % e for e in per_f))
    if i == 0:
        k = int(np.argmax(per_f))
        idx = np.unravel_index(np.argmax(np.abs(a[..., k] - b[..., k])), a.shape[:3])
        print("   worst cell", idx, a[idx + (k,)], b[idx + (k,)], "ratio", a[idx + (k,)] / b[idx + (k,)])
