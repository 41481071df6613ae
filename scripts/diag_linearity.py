import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import khronos_b200 as kb
from khronos_b200 import workloads as w

def run(amp, steps=40, res=40):
    d = w.waveguide_mode(res=res)
    for s in d["sources"]:
        s.amplitude = s.amplitude * amp
    sim = w.build_simulation(d, np.float32)
    sim.prepare_simulation()
    sim.step(steps); sim.sync()
    f = [sim.get_field(c) for c in range(6)]
    sim.close()
    return f

a = run(1.0); b = run(1.0); c = run(2.0)
for comp in range(6):
    d1 = np.abs(a[comp] - b[comp]).max()
    diff = np.abs(c[comp] - 2 * a[comp])
    i = np.unravel_index(np.argmax(diff), diff.shape)
    print(comp, "repeat diff", d1, "| lin diff", diff.max(), "at", i, "vals", a[comp][i], c[comp][i], "n_bad", int((diff > 0).sum()),
          "n_bad_big", int(((diff > 0) & (np.abs(a[comp]) > 1e-30)).sum()))
