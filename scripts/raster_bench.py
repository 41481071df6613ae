"""prepare-time of the geometry step: host point sampler + upload vs khr_geometry_rasterize.
   python scripts/raster_bench.py [sphere|metalens]"""
import os, sys, time, json
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path[:0] = [R]
import numpy as np
import khronos_b200 as kb
from khronos_b200 import workloads as w

def run(desc, rasterizer, smoothing=None):
    sim = w.build_simulation(desc, np.float32, rasterizer=rasterizer, subpixel_smoothing=smoothing)
    t0 = time.perf_counter(); sim.host_prepare(); t1 = time.perf_counter()
    sim.prepare_simulation(); sim.sync(); t2 = time.perf_counter()
    out = dict(workload=desc["name"], objects=len(desc["geometry"]), grid=[sim.Nx, sim.Ny, sim.Nz], rasterizer=rasterizer,
               smoothing=smoothing, host_prepare_s=round(t1 - t0, 3), device_prepare_s=round(t2 - t1, 3), total_s=round(t2 - t0, 3),
               smoothed=getattr(sim, "smoothed_voxels", None))
    e = sim.get_material("eps_inv", 0)
    out["eps_inv_x_mean"] = float(e.mean())
    sim.close()
    return out, e

which = sys.argv[1] if len(sys.argv) > 1 else "sphere"
if which == "sphere":
    desc = w.sphere(res=64)
else:
    desc = w.metalens(nx=1024, ny=1024, nz=256, res=32, pillars=72, rotate=True)   # 5184 rotated pillars + substrate
for d in desc["monitors"]:
    pass
desc["monitors"] = []            # geometry timing only
b, eb = run(desc, "device")
print(json.dumps(b), flush=True)
c, _ = run(desc, "device", "anisotropic")
print(json.dumps(c), flush=True)
if which == "sphere":            # the numpy point sampler visits every voxel for every object: only sane for few objects
    a, ea = run(desc, "host")
    print(json.dumps(a))
    print("fraction of voxels where host and device arrays differ (Float32 coordinate rounding on a face):", float((ea != eb).mean()))
