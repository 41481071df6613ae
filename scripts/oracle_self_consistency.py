import sys
import os; R=os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path[:0]=[R, os.path.join(R,'oracle'), os.path.join(R,'tests')]
import numpy as np
import khronos_b200 as kb
from khronos_b200 import workloads as w
import oracle as ko
from bridge import oracle_from_simulation
d = w.uled(res=12)
for ob in d["geometry"]:
    ob.material.susceptibilities = []
def build(mode):
    sim = w.build_simulation(d, np.float32)
    sim.host_prepare()
    g = sim.grid
    o = ko.OracleSim(sim.T, g.cell_size_user, g.cell_center, g.resolution, g.courant, sim.boundaries)
    for key in ("eps_inv","mu_inv","sigma_D","sigma_B"):
        arr = sim.material_arrays[key]
        if arr is not None:
            for dd in range(3): o.set_material_array(key, dd, arr[dd])
    for sd in sim.source_data:
        tp = sd["src"].time_profile
        o.add_source(sd["comp"], sd["start"], sd["amp"], tp.kind, tp.params(sim.T))
    mids=[o.add_dft(m.component, m.start, m.end, [float(sim.T(f)) for f in m.frequencies], m.decimation) for m in sim.dft_monitors]
    o.prepare(mode)
    return o, mids, sim
a, ma, sim = build("single")
b, mb, _ = build("chunked")
for n in (50,100,150):
    a.step(50); b.step(50)
    num=den=0
    for c in range(6):
        x,y=a.get_field(c),b.get_field(c); num+=((x-y)**2).sum(); den+=(y**2).sum()
    errs=[]
    for i,(p,q) in enumerate(zip(ma,mb)):
        x,y=a.get_dft(p),b.get_dft(q)
        errs.append(np.linalg.norm(x-y)/np.linalg.norm(y))
    print(n, "field %.2e"%np.sqrt(num/den), "dft", " ".join("%.1e"%e for e in errs[:8]))
