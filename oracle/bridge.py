"""Build the CPU oracle from the USER-LEVEL description of a simulation.

TEST INFRASTRUCTURE ONLY (tests/, __graft_entry__.smoke(), bench.py's CPU legs).

The oracle derives every input of the hot path itself, with its own restatement of the
reference (oracle/khronos_oracle.cpp): grid and sigma profiles, GridVolume index boxes,
source interpolation weights and amplitude arrays (Sources.jl:43-135), time-profile constants
(TimeSources.jl:89-112), the geometry raster + subpixel smoothing (Geometry.jl:150-246, 795-972),
absorber ramps (Geometry.jl:708-789), pole sigma rasterisation, its zeroing inside the PML and
the chi1 fold into eps^-1 (Geometry.jl:1059-1353), the Kerr array (Geometry.jl:610-635) and the
auto-decimation (Monitors.jl:33-78).  Nothing the product's `host_prepare()` computed is handed
to the oracle; with `check=True` (default) the two derivations are compared BIT FOR BIT and an
AssertionError names the first input that differs — so a wrong map, weight or ramp in the
product's host code fails every parity test instead of cancelling out.

What the two sides still share are genuine user inputs: dense arrays the caller supplies
(`eps_inv=...`, `poles=[(w0, gamma, sigma_array)]`, `chi3=...`) and Python callables (a source's
spatial `profile`, evaluated here at the oracle's own point coordinates).
"""
import numpy as np

import oracle as ko

_SHAPE_SPHERE, _SHAPE_CUBOID, _SHAPE_CYLINDER = 0, 1, 2


def object_rows(sim):
    """Rows of 28 numbers per geometry object (kind, centre, size, axes, eps_inv, mu_inv, sigma_D,
    sigma_B) from the user-level Object / Material values: get_perm_inv = one(T) / T(perm),
    get_sigma = T(sigma) (Geometry.jl:64-81); mask = which kinds the scene needs (:317-352)."""
    T = sim.T
    rows = []
    for ob in sim.geometry:
        sh, m = ob.shape, ob.material
        name = type(sh).__name__
        if name == "Ball":
            kind, size, axes = _SHAPE_SPHERE, [sh.radius, 0.0, 0.0], [0.0] * 9
        elif name == "Cuboid":
            kind, size = _SHAPE_CUBOID, list(sh.size)
            axes = [0.0] * 9 if sh.axes is None else [float(v) for v in np.asarray(sh.axes).reshape(9)]
        elif name == "Cylinder":
            kind, size = _SHAPE_CYLINDER, [sh.radius, sh.height, 0.0]
            axes = [float(v) for v in sh.axis] + [0.0] * 6
        else:
            raise ValueError("the oracle knows Ball, Cuboid and Cylinder shapes")
        rows.append([kind] + list(sh.center) + size + axes + [float(T(1) / T(m.epsilon))] * 3 + [float(T(1) / T(m.mu))] * 3
                    + [float(T(m.sigma_D))] * 3 + [float(T(m.sigma_B))] * 3)
    mask = 0
    for bit, attr, neutral in ((0, "epsilon", 1.0), (1, "mu", 1.0), (2, "sigma_D", 0.0), (3, "sigma_B", 0.0)):
        if any(getattr(o.material, attr) != neutral for o in sim.geometry):
            mask |= 1 << bit
    return rows, mask


def _time_params(tp, T):
    """(kind, [fcen, width, peak_time, cutoff] as T, f_cen for decimation, fwidth for decimation)."""
    name = type(tp).__name__
    if name == "ContinuousWaveSource":
        return 0, [float(T(tp.fcen)), 0.0, 0.0, 0.0], tp.fcen, 0.0
    if name == "GaussianPulseSource":
        g = ko.gaussian_pulse(*tp.ctor_args)
        return 1, [float(T(g["fcen"])), float(T(g["width"])), float(T(g["peak_time"])), float(T(g["cutoff"]))], g["fcen"], g["fwidth"]
    raise ValueError("the oracle evaluates CW / Gaussian time profiles only")


def _eq(name, a, b):
    a, b = np.asarray(a), np.asarray(b)
    if a.shape != b.shape or not np.array_equal(a, b):
        bad = "shape %s vs %s" % (a.shape, b.shape) if a.shape != b.shape else "%d of %d values differ (max |d| %.3e)" % (
            int(np.count_nonzero(a != b)), a.size, float(np.max(np.abs(a.astype(np.complex128) - b.astype(np.complex128)))))
        raise AssertionError("host plan differs from the oracle's own derivation: %s: %s" % (name, bad))


def oracle_from_simulation(sim, check=True):
    """sim: khronos_b200.Simulation (need not be device-prepared). Returns (OracleSim, monitor ids)."""
    g, T = sim.grid, sim.T
    ct = np.complex64 if T is np.float32 else np.complex128
    o = ko.OracleSim(T, g.cell_size_user, g.cell_center, g.resolution, g.courant, sim.boundaries)
    for a in range(3):
        if g.dlv[a] is not None:
            o.set_grid_spacing(a, g.dlv[a])
    if getattr(sim, "bc_codes", None) is not None:
        o.set_boundary_conditions(sim.bc_codes)
    if getattr(sim, "complex_fields", False):
        for a in range(3):
            o.set_bloch(a, sim.bloch_k[a])
    # ---- materials: raster (+ smoothing), user arrays, Kerr array, absorbers, poles
    rows, mask = object_rows(sim) if sim.geometry else ([], 0)
    mode = {None: 0, "volume": 1, "anisotropic": 2}[sim.subpixel_smoothing]
    o.smoothed_voxels = [0, 0, 0]
    if rows and mask:
        o.smoothed_voxels = o.rasterize(rows, mask, mode)
    for key, arr in sim.user_arrays.items():
        if arr is not None:
            for d in range(3):
                o.set_material_array(key, d, np.array(arr[d], dtype=T))
    if sim.user_chi3 is not None:
        o.set_material_array("chi3", 0, np.array(sim.user_chi3, dtype=T))
    elif rows:
        o.chi3_from_geometry(rows, [ob.material.chi3 for ob in sim.geometry])
    if sim.absorbers is not None:
        for axis, ax in enumerate(sim.absorbers):
            for side, ab in enumerate(ax or []):
                if ab is not None:
                    o.add_absorber(axis, side, ab.num_layers, ab.sigma_order, ab.sigma_max)
    pole_keys = []
    if rows:
        sus = [[(s.omega_0, s.gamma, s.sigma) for s in ob.material.susceptibilities] for ob in sim.geometry]
        o.poles_from_geometry(rows, sus)
        for lst in sus:
            for (w0, gam, _) in lst:
                if (w0, gam) not in pole_keys:
                    pole_keys.append((w0, gam))
    for (w0, gam, s) in sim.user_poles:
        o.add_pole(w0, gam, np.asarray(s, dtype=T))
        pole_keys.append((w0, gam))
    o.finish_poles()
    # ---- sources: box, weights, amplitude = weight * amplitude * profile (Complex{Float64} -> Complex{T})
    o_sources, kinds, fcs, fws = [], [], [], []
    for src in sim.sources:
        kind, tpar, fc, fw = _time_params(src.time_profile, T)
        kinds.append(kind), fcs.append(fc), fws.append(fw)
        for comp in src.components:
            start, dims, w, pts = o.source_weights(comp, src.center, src.size)
            amp = w.astype(np.complex128) * complex(src.amplitude)
            if src.profile is not None:
                X, Y, Z = np.meshgrid(*pts, indexing="ij", sparse=True)
                amp = amp * src.profile([X, Y, Z], comp)
            else:
                amp = amp * 1.0
            amp = amp.astype(ct)
            o.add_source(comp, start, amp, kind, tpar)
            o_sources.append((comp, start, dims, amp, kind, tpar))
    # ---- monitors: FluxMonitor -> its four tangential DFT monitors, index boxes, auto-decimation
    dms = []
    for m in sim.monitors:
        dms.extend(m.monitors if hasattr(m, "monitors") else [m])
    d_max = o.auto_decimation(kinds, fcs, fws) if kinds else 1
    user_dec = [m.user_decimation for m in dms]
    mids, o_mons = [], []
    for m, dec0 in zip(dms, user_dec):
        st, en = o.grid_volume(m.center, m.size, m.component)
        dec = d_max if (dec0 == 1 and d_max > 1) else dec0
        fr = [float(T(f)) for f in m.frequencies]
        mids.append(o.add_dft(m.component, st, en, fr, dec))
        o_mons.append(([int(v) for v in st], [int(v) for v in en], dec))
    o.prepare("single")
    if check:
        check_host_plan(sim, o, o_sources, o_mons, pole_keys)
    return o, mids


def check_host_plan(sim, o, o_sources, o_mons, pole_keys):
    """The product's host_prepare() against the oracle's own derivations, bit for bit."""
    sim.host_prepare()
    g = sim.grid
    assert tuple(o.N) == tuple(g.N), (o.N, g.N)
    assert o.dt == float(g.dt), (o.dt, float(g.dt))
    assert [float(v) for v in g.dl] == list(o.dl)
    if sim.sigma is not None:
        for grp in range(2):
            for a in range(3):
                _eq("sigma profile group %d axis %d" % (grp, a), sim.sigma[grp][a], o.sigma(a, grp))
    for key in ("eps_inv", "mu_inv", "sigma_D", "sigma_B"):
        arr = sim.material_arrays[key]
        if sim.rasterizer == "device" and arr is None:
            continue   # produced on the device; tests/test_geometry.py compares that raster with the oracle's
        for d in range(3):
            mine = None if arr is None else arr[d]
            ref = o.get_material_array(key, d)
            if mine is None and ref is None:
                continue
            if mine is None or ref is None:
                # one side keeps a scalar: compare against the constant array
                const = np.full(tuple(g.N), sim.T(1 if key in ("eps_inv", "mu_inv") else 0), dtype=sim.T)
                mine = const if mine is None else mine
                ref = const if ref is None else ref
            _eq("%s[%d]" % (key, d), mine, ref)
    mine, ref = sim.material_arrays.get("chi3"), o.get_chi3()
    if mine is not None or ref is not None:
        z = np.zeros(tuple(g.N), dtype=sim.T)
        _eq("chi3", z if mine is None else mine, z if ref is None else ref)
    assert len(sim.poles) == o.num_poles() == len(pole_keys), (len(sim.poles), o.num_poles())
    for q, ((w0, gam, s), key) in enumerate(zip(sim.poles, pole_keys)):
        assert (w0, gam) == key, ("pole order", q, (w0, gam), key)
        _eq("pole %d sigma" % q, s, o.get_pole_sigma(q))
    assert len(sim.source_data) == len(o_sources)
    for q, (sd, (comp, start, dims, amp, kind, tpar)) in enumerate(zip(sim.source_data, o_sources)):
        assert sd["comp"] == comp and list(sd["start"]) == list(start) and list(sd["dims"]) == list(dims), ("source box", q)
        _eq("source %d amplitude" % q, sd["amp"], amp)
        tp = sd["src"].time_profile
        assert tp.kind == kind and list(tp.params(sim.T)) == list(tpar), ("time profile", q, tp.params(sim.T), tpar)
    assert len(sim.dft_monitors) == len(o_mons)
    for q, (m, (st, en, dec)) in enumerate(zip(sim.dft_monitors, o_mons)):
        assert list(m.start) == st and list(m.end) == en, ("monitor box", q, m.start, m.end, st, en)
        assert m.decimation == dec, ("monitor decimation", q, m.decimation, dec)
