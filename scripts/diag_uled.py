import sys, os
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [R, os.path.join(R, "oracle"), os.path.join(R, "tests")]
import numpy as np
import khronos_b200 as kb
from khronos_b200 import workloads as w
from bridge import oracle_from_simulation

def run(desc, label, checkpoints=(25, 50, 100, 150, 300)):
    sim = w.build_simulation(desc, np.float32)
    o, mids = oracle_from_simulation(sim)
    sim.prepare_simulation()
    done = 0
    for n in checkpoints:
        sim.step(n - done); sim.sync(); o.step(n - done); done = n
        errs = []
        num = den = 0.0
        for c in range(6):
            a, b = sim.get_field(c).astype(np.float64), o.get_field(c)
            errs.append(np.sqrt(((a - b) ** 2).sum() / max((b ** 2).sum(), 1e-300)))
            num += ((a - b) ** 2).sum(); den += (b ** 2).sum()
        dn = dd = 0.0
        for m, mid in zip(sim.dft_monitors, mids):
            if m.component < 3:
                a, b = sim.get_dft(m), o.get_dft(mid)
                dn += np.sum(np.abs(a - b) ** 2); dd += np.sum(np.abs(b) ** 2)
        print(label, n, "field %.2e" % np.sqrt(num / den), " ".join("%.1e" % e for e in errs), "E-dft %.2e" % np.sqrt(dn / max(dd, 1e-300)))
    sim.close()

d = w.uled(res=12); run(d, "uled both poles")
d = w.uled(res=12, lorentz=False); run(d, "uled drude only")
d = w.uled(res=12)
for ob in d["geometry"]:
    ob.material.susceptibilities = []
run(d, "uled no poles")
