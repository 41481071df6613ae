"""Build the CPU oracle from the same host-side plan the product hands to the GPU library.

TEST INFRASTRUCTURE ONLY (tests/, __graft_entry__.smoke(), bench.py's CPU legs).
The product's `Simulation.host_prepare()` is pure host code (index maps, rasterised
material arrays, source amplitudes); this module feeds exactly those inputs to
oracle/khronos_oracle.cpp so both sides see identical data.
"""
import numpy as np

import oracle as ko


def oracle_from_simulation(sim):
    """sim: khronos_b200.Simulation (need not be device-prepared). Returns (OracleSim, monitor ids)."""
    sim.host_prepare()
    g = sim.grid
    o = ko.OracleSim(sim.T, g.cell_size_user, g.cell_center, g.resolution, g.courant, sim.boundaries)
    assert tuple(o.N) == tuple(g.N)
    for a in range(3):
        if g.dlv[a] is not None:
            o.set_grid_spacing(a, g.dlv[a])
    assert o.dt == float(g.dt), (o.dt, float(g.dt))
    if getattr(sim, "bc_codes", None) is not None:
        o.set_boundary_conditions(sim.bc_codes)
    if getattr(sim, "complex_fields", False):
        for a in range(3):
            o.set_bloch(a, sim.bloch_k[a])
    if getattr(sim, "rasterizer", "host") == "device":
        # the oracle rasterises the same object list itself (its own restatement of Geometry.jl)
        objs, mask = sim.geometry_objects()
        rows = [[o.kind] + list(o.center) + list(o.size) + list(o.axes) + list(o.eps_inv) + list(o.mu_inv)
                + list(o.sigma_d) + list(o.sigma_b) for o in objs]
        mode = {None: 0, "volume": 1, "anisotropic": 2}[sim.subpixel_smoothing]
        o.smoothed_voxels = o.rasterize(rows, mask, mode) if mask else [0, 0, 0]
    for key in ("eps_inv", "mu_inv", "sigma_D", "sigma_B"):
        arr = sim.material_arrays[key]
        if arr is not None:
            for d in range(3):
                o.set_material_array(key, d, arr[d])
    if sim.material_arrays.get("chi3") is not None:
        o.set_material_array("chi3", 0, sim.material_arrays["chi3"])
    for (w0, gam, s) in sim.poles:
        o.add_pole(w0, gam, s)
    for sd in sim.source_data:
        tp = sd["src"].time_profile
        if tp.kind not in (0, 1):
            raise ValueError("the oracle evaluates CW / Gaussian time profiles only")
        o.add_source(sd["comp"], sd["start"], sd["amp"], tp.kind, tp.params(sim.T))
    mids = []
    for m in sim.dft_monitors:
        mids.append(o.add_dft(m.component, m.start, m.end, [float(sim.T(f)) for f in m.frequencies], m.decimation))
    o.prepare("single")
    return o, mids
