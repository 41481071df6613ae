"""ctypes front-end of the CPU parity oracle (oracle/khronos_oracle.cpp).

TEST INFRASTRUCTURE ONLY.  Only tests/, __graft_entry__.smoke() and the
cpu_baseline / --impl reference legs of bench.py may import this module; the
product package (khronos.jl_b200/) never does.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

EX, EY, EZ, HX, HY, HZ, CENTER = range(7)
COMP_NAMES = ["Ex", "Ey", "Ez", "Hx", "Hy", "Hz"]


def build(force=False):
    so = os.path.join(_HERE, "libkhronos_oracle.so")
    src = os.path.join(_HERE, "khronos_oracle.cpp")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s"], env={k: v for k, v in os.environ.items() if k != "CXX"})
    return so


def lib():
    global _LIB
    if _LIB is None:
        so = os.path.join(_HERE, "libkhronos_oracle.so")
        if not os.path.exists(so):
            build()
        L = C.CDLL(so)
        dp = C.POINTER(C.c_double)
        ip = C.POINTER(C.c_int)
        L.ko_create.restype = C.c_void_p
        L.ko_create.argtypes = [C.c_int, dp, dp, C.c_double, C.c_double, C.c_int, dp]
        L.ko_destroy.argtypes = [C.c_void_p]
        L.ko_grid.argtypes = [C.c_void_p, ip, dp, dp]
        L.ko_gridvolume.argtypes = [C.c_void_p, dp, dp, C.c_int, ip, ip]
        L.ko_component_origin.argtypes = [C.c_void_p, C.c_int, dp]
        L.ko_sigma.restype = C.c_int
        L.ko_sigma.argtypes = [C.c_void_p, C.c_int, C.c_int, dp]
        L.ko_plan_pml_grid.restype = C.c_int
        L.ko_plan_pml_grid.argtypes = [C.c_void_p, C.c_int, ip, ip, C.c_int]
        L.ko_adjacency.restype = C.c_int
        L.ko_adjacency.argtypes = [C.c_int, ip, ip, C.c_int]
        L.ko_halo_ranges.argtypes = [ip, ip, C.c_int, C.c_int, C.c_int, ip, ip]
        L.ko_interp_weight.restype = C.c_double
        L.ko_interp_weight.argtypes = [dp, dp, dp, dp, C.c_int, dp]
        L.ko_ade_coefficients.argtypes = [C.c_double, C.c_double, C.c_double, dp]
        L.ko_eval_time_source.argtypes = [C.c_int, C.c_int, dp, C.c_double, dp]
        L.ko_set_material_scalar.argtypes = [C.c_void_p, C.c_int, C.c_double]
        L.ko_set_material_array.argtypes = [C.c_void_p, C.c_int, dp]
        L.ko_get_material_array.restype = C.c_int
        L.ko_get_material_array.argtypes = [C.c_void_p, C.c_int, dp]
        L.ko_add_absorber.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_double]
        L.ko_add_pole.argtypes = [C.c_void_p, C.c_double, C.c_double, dp]
        L.ko_add_source.restype = C.c_int
        L.ko_add_source.argtypes = [C.c_void_p, C.c_int, ip, ip, dp, C.c_int, dp]
        L.ko_add_dft.restype = C.c_int
        L.ko_add_dft.argtypes = [C.c_void_p, C.c_int, ip, ip, C.c_int, dp, C.c_int]
        L.ko_prepare.argtypes = [C.c_void_p, C.c_int, C.c_int]
        L.ko_num_chunks.restype = C.c_int
        L.ko_num_chunks.argtypes = [C.c_void_p]
        L.ko_chunk_aux_pattern.argtypes = [C.c_void_p, C.c_int, ip, ip, ip, ip]
        L.ko_chunk_sigma.restype = C.c_int
        L.ko_chunk_sigma.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, dp]
        L.ko_step.argtypes = [C.c_void_p, C.c_int]
        L.ko_set_sources_active.argtypes = [C.c_void_p, C.c_int]
        L.ko_set_chi3_literal_order.argtypes = [C.c_void_p, C.c_int]
        L.ko_set_bloch.argtypes = [C.c_void_p, C.c_int, C.c_double]
        L.ko_set_grid_spacing.argtypes = [C.c_void_p, C.c_int, dp, C.c_int]
        L.ko_rasterize.argtypes = [C.c_void_p, C.c_int, dp, C.c_int, C.c_int, C.POINTER(C.c_long)]
        L.ko_set_boundary_conditions.argtypes = [C.c_void_p, ip]
        L.ko_timestep.restype = C.c_long
        L.ko_num_threads.restype = C.c_int
        L.ko_set_num_threads.argtypes = [C.c_int]
        L.ko_timestep.argtypes = [C.c_void_p]
        L.ko_get_field.argtypes = [C.c_void_p, C.c_int, C.c_int, dp]
        L.ko_set_field.argtypes = [C.c_void_p, C.c_int, dp]
        L.ko_get_dft.restype = C.c_size_t
        L.ko_get_dft.argtypes = [C.c_void_p, C.c_int, dp]
        L.ko_flux.argtypes = [C.c_void_p, C.c_int, ip, dp]
        L.ko_near2far.argtypes = [C.c_void_p, C.c_int, ip, C.c_double, C.c_double, C.c_double, dp, dp, dp, C.c_int, dp]
        L.ko_mode_amplitudes.argtypes = [C.c_void_p, C.c_int, ip, dp, dp, dp, dp]
        L.ko_green3d_many.argtypes = [C.c_int, dp, C.c_int, dp, C.c_double, C.c_double, C.c_double, dp]
        L.ko_diffraction.argtypes = [C.c_void_p, C.c_int, ip, C.c_int, C.c_double, C.c_double, C.c_double, C.c_double, dp, dp, ip]
        L.ko_source_weights.argtypes = [C.c_void_p, C.c_int, dp, dp, ip, ip, dp, dp]
        L.ko_gaussian_pulse.argtypes = [C.c_double, C.c_double, C.c_double, C.c_double, dp]
        L.ko_auto_decimation.restype = C.c_int
        L.ko_auto_decimation.argtypes = [C.c_void_p, C.c_int, ip, dp, dp]
        L.ko_poles_from_geometry.restype = C.c_int
        L.ko_poles_from_geometry.argtypes = [C.c_void_p, C.c_int, dp, ip, dp]
        L.ko_finish_poles.argtypes = [C.c_void_p]
        L.ko_chi3_from_geometry.argtypes = [C.c_void_p, C.c_int, dp, dp, ip]
        L.ko_num_poles.restype = C.c_int
        L.ko_num_poles.argtypes = [C.c_void_p]
        L.ko_get_pole_sigma.argtypes = [C.c_void_p, C.c_int, dp]
        L.ko_get_chi3.restype = C.c_int
        L.ko_get_chi3.argtypes = [C.c_void_p, dp]
        _LIB = L
    return _LIB


def num_threads():
    return int(lib().ko_num_threads())


def set_num_threads(n):
    lib().ko_set_num_threads(int(n))


def _d(a):
    a = np.ascontiguousarray(a, dtype=np.float64)
    return a, a.ctypes.data_as(C.POINTER(C.c_double))


def _i(a):
    a = np.ascontiguousarray(a, dtype=np.int32)
    return a, a.ctypes.data_as(C.POINTER(C.c_int))


def interp_weight(p, lo, hi, size, ndims, delta):
    args = [_d(x) for x in (p, lo, hi, size)]
    dl = _d(delta)
    return lib().ko_interp_weight(args[0][1], args[1][1], args[2][1], args[3][1], int(ndims), dl[1])


def ade_coefficients(omega0, gamma, dt):
    out = np.zeros(6)
    lib().ko_ade_coefficients(float(omega0), float(gamma), float(dt), out.ctypes.data_as(C.POINTER(C.c_double)))
    return dict(gamma1_inv=out[0], gamma1=out[1], omega0_dt_sq=out[2], sigma_omega0_dt_sq=out[3],
                drude_coeff=out[4], is_drude=bool(out[5]))


def eval_time_source(dtype, kind, params4, t):
    p = _d(params4)
    out = np.zeros(2)
    lib().ko_eval_time_source(0 if dtype in ("f32", np.float32) else 1, int(kind), p[1], float(t),
                              out.ctypes.data_as(C.POINTER(C.c_double)))
    return complex(out[0], out[1])


def adjacency(regions):
    r, rp = _i(np.asarray(regions).reshape(-1, 6))
    out = np.zeros((4096, 3), dtype=np.int32)
    m = lib().ko_adjacency(len(r), rp, out.ctypes.data_as(C.POINTER(C.c_int)), 4096)
    return out[:m].copy()


def halo_ranges(src6, dst6, axis, src_upper, dst_lower):
    s, sp = _i(src6)
    d, dp_ = _i(dst6)
    sr = np.zeros(6, dtype=np.int32)
    dr = np.zeros(6, dtype=np.int32)
    lib().ko_halo_ranges(sp, dp_, int(axis), int(src_upper), int(dst_lower),
                         sr.ctypes.data_as(C.POINTER(C.c_int)), dr.ctypes.data_as(C.POINTER(C.c_int)))
    return sr, dr


def gaussian_pulse(fcen, fwidth, start_time=0.0, cutoff_scale=5.0):
    """GaussianPulseSource constructor (TimeSources.jl:89-112): dict of fcen, fwidth, width, peak_time, cutoff (Float64)."""
    out = np.zeros(5)
    lib().ko_gaussian_pulse(float(fcen), float(fwidth), float(start_time), float(cutoff_scale),
                            out.ctypes.data_as(C.POINTER(C.c_double)))
    return dict(fcen=out[0], fwidth=out[1], width=out[2], peak_time=out[3], cutoff=out[4])


def green3d(obs, sources, freq, eps=1.0, mu=1.0):
    """green3d! (Near2Far.jl:40-96) summed over point currents.  obs: (nobs, 3); sources: rows
    (x, y, z, c0, f0) with c0 = 1..3 electric Jx,Jy,Jz, 4..6 magnetic, f0 complex.  Returns (nobs, 6) complex."""
    o, op = _d(np.asarray(obs, dtype=np.float64).reshape(-1, 3))
    rows = np.array([[r[0], r[1], r[2], float(r[3]), complex(r[4]).real, complex(r[4]).imag] for r in sources], dtype=np.float64)
    s, sp = _d(rows)
    out = np.zeros(12 * o.shape[0])
    lib().ko_green3d_many(o.shape[0], op, s.shape[0], sp, float(freq), float(eps), float(mu),
                          out.ctypes.data_as(C.POINTER(C.c_double)))
    z = out[0::2] + 1j * out[1::2]
    return z.reshape(o.shape[0], 6)


class OracleSim:
    """One simulation in the CPU oracle.  Arrays cross as float64 and are cast
    to the working type T inside (values given must already be T-representable
    where bit-exactness matters)."""

    def __init__(self, dtype, cell_size, cell_center, resolution, courant=0.5, boundaries=None):
        self.L = lib()
        self.dtype = np.float32 if dtype in ("f32", np.float32, "Float32") else np.float64
        cs, csp = _d(cell_size)
        cc, ccp = _d(cell_center)
        if boundaries is None:
            self.h = self.L.ko_create(0 if self.dtype == np.float32 else 1, csp, ccp, float(resolution),
                                      float(courant), 0, None)
        else:
            b, bp = _d(np.asarray(boundaries, dtype=np.float64).reshape(6))
            self.h = self.L.ko_create(0 if self.dtype == np.float32 else 1, csp, ccp, float(resolution),
                                      float(courant), 1, bp)
        N = np.zeros(3, dtype=np.int32)
        dl = np.zeros(3)
        dt = C.c_double()
        self.L.ko_grid(self.h, N.ctypes.data_as(C.POINTER(C.c_int)), dl.ctypes.data_as(C.POINTER(C.c_double)),
                       C.byref(dt))
        self.N = tuple(int(x) for x in N)
        self.dl = tuple(float(x) for x in dl)
        self.dt = dt.value
        self._monitors = []

    def __del__(self):
        try:
            if self.h:
                self.L.ko_destroy(self.h)
                self.h = None
        except Exception:
            pass

    # ---- index maps -------------------------------------------------------
    def grid_volume(self, center, size, comp):
        c, cp = _d(center)
        s, sp = _d(size)
        st = np.zeros(3, dtype=np.int32)
        en = np.zeros(3, dtype=np.int32)
        self.L.ko_gridvolume(self.h, cp, sp, int(comp), st.ctypes.data_as(C.POINTER(C.c_int)),
                             en.ctypes.data_as(C.POINTER(C.c_int)))
        return st, en

    def component_origin(self, comp):
        o = np.zeros(3)
        self.L.ko_component_origin(self.h, int(comp), o.ctypes.data_as(C.POINTER(C.c_double)))
        return o

    def sigma(self, axis, group=0):
        n = self.L.ko_sigma(self.h, axis, group, None)
        if n == 0:
            return None
        out = np.zeros(n)
        self.L.ko_sigma(self.h, axis, group, out.ctypes.data_as(C.POINTER(C.c_double)))
        return out.astype(self.dtype)

    def plan_pml_grid(self, nranks=0):
        out = np.zeros((4096, 6), dtype=np.int32)
        fl = np.zeros((4096, 3), dtype=np.int32)
        n = self.L.ko_plan_pml_grid(self.h, int(nranks), out.ctypes.data_as(C.POINTER(C.c_int)),
                                    fl.ctypes.data_as(C.POINTER(C.c_int)), 4096)
        return out[:n].copy(), fl[:n].copy()

    # ---- problem description ----------------------------------------------
    def set_material_scalar(self, kind, v):
        self.L.ko_set_material_scalar(self.h, {"eps_inv": 0, "mu_inv": 1}[kind], float(v))

    _KINDS = {"eps_inv": 0, "mu_inv": 3, "sigma_D": 6, "sigma_B": 9, "chi3": 12}

    def set_material_array(self, kind, comp, arr):
        """arr: (Nx,Ny,Nz) numpy array (any order; converted to x-fastest)."""
        a = np.asarray(arr, dtype=np.float64)
        assert a.shape == self.N, (a.shape, self.N)
        flat = np.ascontiguousarray(a.transpose(2, 1, 0)).ravel()
        self.L.ko_set_material_array(self.h, self._KINDS[kind] + comp, flat.ctypes.data_as(C.POINTER(C.c_double)))

    def get_material_array(self, kind, comp):
        out = np.zeros(self.N[::-1])
        ok = self.L.ko_get_material_array(self.h, self._KINDS[kind] + comp, out.ctypes.data_as(C.POINTER(C.c_double)))
        return out.transpose(2, 1, 0).astype(self.dtype) if ok else None

    def add_absorber(self, axis, side, num_layers=40, sigma_order=3, sigma_max=0.0):
        self.L.ko_add_absorber(self.h, axis, side, num_layers, sigma_order, float(sigma_max))

    def add_pole(self, omega0, gamma, sigma):
        a = np.asarray(sigma, dtype=np.float64)
        assert a.shape == self.N
        flat = np.ascontiguousarray(a.transpose(2, 1, 0)).ravel()
        self.L.ko_add_pole(self.h, float(omega0), float(gamma), flat.ctypes.data_as(C.POINTER(C.c_double)))

    def source_weights(self, comp, center, size):
        """The oracle's own GridVolume + interpolation weights of a source volume (Sources.jl:43-135):
        (start, dims, weights (nx,ny,nz) Float64, [xs, ys, zs] point coordinates)."""
        c, cp = _d(center)
        s, sp = _d(size)
        st = np.zeros(3, dtype=np.int32)
        dm = np.zeros(3, dtype=np.int32)
        ipt = C.POINTER(C.c_int)
        dpt = C.POINTER(C.c_double)
        self.L.ko_source_weights(self.h, int(comp), cp, sp, st.ctypes.data_as(ipt), dm.ctypes.data_as(ipt), None, None)
        w = np.zeros(int(dm[0]) * int(dm[1]) * int(dm[2]))
        pts = np.zeros(int(dm.sum()))
        self.L.ko_source_weights(self.h, int(comp), cp, sp, st.ctypes.data_as(ipt), dm.ctypes.data_as(ipt),
                                 w.ctypes.data_as(dpt), pts.ctypes.data_as(dpt))
        w = w.reshape(int(dm[2]), int(dm[1]), int(dm[0])).transpose(2, 1, 0)
        xs, ys, zs = pts[:dm[0]], pts[dm[0]:dm[0] + dm[1]], pts[dm[0] + dm[1]:]
        return [int(v) for v in st], [int(v) for v in dm], w, [xs, ys, zs]

    def auto_decimation(self, kinds, fcens, fwidths):
        """auto_decimate! (Monitors.jl:33-78): D_max from the sources' time profiles."""
        k, kp = _i(kinds)
        f, fp = _d(fcens)
        w, wp = _d(fwidths)
        return int(self.L.ko_auto_decimation(self.h, len(k), kp, fp, wp))

    def poles_from_geometry(self, objects, sus_lists):
        """_collect_unique_poles + _rasterize_pole_sigma! (Geometry.jl:1059-1125).  objects: rows of 28
        numbers as for rasterize(); sus_lists[q] = [(omega_0, gamma, sigma), ...] of object q."""
        o, op = _d(np.asarray(objects, dtype=np.float64).reshape(-1, 28))
        ns, nsp = _i([len(x) for x in sus_lists])
        flat = [v for x in sus_lists for t in x for v in t] or [0.0]
        f, fp = _d(flat)
        return int(self.L.ko_poles_from_geometry(self.h, o.shape[0], op, nsp, fp))

    def finish_poles(self):
        """Rest of init_polarization! (Geometry.jl:1180-1353): PML zeroing of sigma, chi1 fold into eps_inv."""
        self.L.ko_finish_poles(self.h)

    def chi3_from_geometry(self, objects, chi3_values):
        """chi3 painting (Geometry.jl:610-635); chi3_values[q] = None or the Kerr coefficient of object q."""
        o, op = _d(np.asarray(objects, dtype=np.float64).reshape(-1, 28))
        v, vp = _d([0.0 if x is None else float(x) for x in chi3_values])
        hs, hp = _i([0 if x is None else 1 for x in chi3_values])
        self.L.ko_chi3_from_geometry(self.h, o.shape[0], op, vp, hp)

    def num_poles(self):
        return int(self.L.ko_num_poles(self.h))

    def get_pole_sigma(self, q):
        out = np.zeros(self.N[::-1])
        self.L.ko_get_pole_sigma(self.h, int(q), out.ctypes.data_as(C.POINTER(C.c_double)))
        return out.transpose(2, 1, 0).astype(self.dtype)

    def get_chi3(self):
        out = np.zeros(self.N[::-1])
        ok = self.L.ko_get_chi3(self.h, out.ctypes.data_as(C.POINTER(C.c_double)))
        return out.transpose(2, 1, 0).astype(self.dtype) if ok else None

    def add_source(self, comp, start, amp, time_kind, time_params):
        """amp: complex (nx,ny,nz) array; start: 1-based global start index of
        the source's GridVolume; time_params = (fcen, width, peak_time, cutoff)."""
        amp = np.asarray(amp, dtype=np.complex128)
        dims = amp.shape
        flat = np.ascontiguousarray(amp.transpose(2, 1, 0)).ravel()
        ri = np.empty(2 * flat.size)
        ri[0::2] = flat.real
        ri[1::2] = flat.imag
        s, sp = _i(start)
        d, dp_ = _i(dims)
        tp, tpp = _d(time_params)
        return self.L.ko_add_source(self.h, int(comp), sp, dp_, ri.ctypes.data_as(C.POINTER(C.c_double)),
                                    int(time_kind), tpp)

    def add_dft(self, comp, start, end, freqs, decimation=1):
        s, sp = _i(start)
        e, ep = _i(end)
        f, fp = _d(freqs)
        mid = self.L.ko_add_dft(self.h, int(comp), sp, ep, len(f), fp, int(decimation))
        self._monitors.append((tuple(int(x) for x in (e - s + 1)), len(f)))
        return mid

    def set_chi3_literal_order(self, on):
        """True: apply the Kerr correction after the halo / wrap copies, the literal order of step!
        (Kernels.jl:76-79); default False = before them (neighbours see the corrected E)."""
        self.L.ko_set_chi3_literal_order(self.h, int(bool(on)))

    def set_grid_spacing(self, axis, spacing):
        """Non-uniform grid (DataStructures.jl:737-739): one spacing per cell of `axis`; updates dt."""
        d, dp_ = _d(spacing)
        assert len(d) == self.N[axis]
        self.L.ko_set_grid_spacing(self.h, int(axis), dp_, len(d))
        N = np.zeros(3, dtype=np.int32)
        dl = np.zeros(3)
        dt = C.c_double()
        self.L.ko_grid(self.h, N.ctypes.data_as(C.POINTER(C.c_int)), dl.ctypes.data_as(C.POINTER(C.c_double)),
                       C.byref(dt))
        self.dl = tuple(float(x) for x in dl)
        self.dt = dt.value

    def set_bloch(self, axis, k):
        """Bloch(k) on both sides of `axis` (DataStructures.jl:158-160): complex fields; the axis must
        also be flagged periodic through set_boundary_conditions."""
        self.L.ko_set_bloch(self.h, int(axis), float(k))

    def set_boundary_conditions(self, bc6):
        """bc6: per (axis, side) 0 = PML, 1 = Periodic, 2 = PEC, 3 = PMC (before prepare)."""
        b, bp = _i(np.asarray(bc6, dtype=np.int32).reshape(6))
        self.L.ko_set_boundary_conditions(self.h, bp)

    # ---- run ----------------------------------------------------------------
    def prepare(self, mode="single", nranks=0):
        self.L.ko_prepare(self.h, 0 if mode == "single" else 1, int(nranks))

    def num_chunks(self):
        return self.L.ko_num_chunks(self.h)

    def chunk_info(self, q):
        pat = np.zeros(18, dtype=np.int32)
        st = np.zeros(3, dtype=np.int32)
        n = np.zeros(3, dtype=np.int32)
        pml = np.zeros(3, dtype=np.int32)
        ip = C.POINTER(C.c_int)
        self.L.ko_chunk_aux_pattern(self.h, q, pat.ctypes.data_as(ip), st.ctypes.data_as(ip), n.ctypes.data_as(ip),
                                    pml.ctypes.data_as(ip))
        names = ["CB", "UB", "WB", "CD", "UD", "WD"]
        aux = {names[g] + "xyz"[d]: bool(pat[3 * g + d]) for g in range(6) for d in range(3)}
        return dict(start=st, n=n, pml=pml.astype(bool), aux=aux)

    def chunk_sigma(self, q, group, axis):
        n = self.L.ko_chunk_sigma(self.h, q, group, axis, None)
        if n == 0:
            return None
        out = np.zeros(n)
        self.L.ko_chunk_sigma(self.h, q, group, axis, out.ctypes.data_as(C.POINTER(C.c_double)))
        return out.astype(self.dtype)

    def step(self, n=1):
        self.L.ko_step(self.h, int(n))

    @property
    def timestep(self):
        return self.L.ko_timestep(self.h)

    def get_field(self, comp, which="EH"):
        out = np.zeros(self.N[::-1])
        self.L.ko_get_field(self.h, {"EH": 0, "DB": 1, "imag": 2}[which], int(comp),
                            out.ctypes.data_as(C.POINTER(C.c_double)))
        return out.transpose(2, 1, 0)

    def set_field(self, comp, arr):
        a = np.asarray(arr, dtype=np.float64)
        flat = np.ascontiguousarray(a.transpose(2, 1, 0)).ravel()
        self.L.ko_set_field(self.h, int(comp), flat.ctypes.data_as(C.POINTER(C.c_double)))

    def get_dft(self, mid):
        dims, nf = self._monitors[mid]
        n = self.L.ko_get_dft(self.h, mid, None)
        out = np.zeros(2 * n)
        self.L.ko_get_dft(self.h, mid, out.ctypes.data_as(C.POINTER(C.c_double)))
        z = out[0::2] + 1j * out[1::2]
        return z.reshape((nf,) + dims[::-1]).transpose(3, 2, 1, 0)  # (nx,ny,nz,nf)

    def flux(self, normal_axis, ids4):
        i4, ip = _i(ids4)
        nf = self._monitors[ids4[0]][1]
        out = np.zeros(nf)
        self.L.ko_flux(self.h, int(normal_axis), ip, out.ctypes.data_as(C.POINTER(C.c_double)))
        return out

    def near2far(self, normal_axis, ids4, normal_sign, eps, mu, bases, freqs, obs):
        """_compute_far_field_cpu (Near2Far.jl:254-371): EH complex (nobs, 6, nf)."""
        i4, ip = _i(ids4)
        b, bp = _d(np.asarray(bases, dtype=np.float64).reshape(12))
        f, fp = _d(freqs)
        o, op = _d(np.asarray(obs, dtype=np.float64).reshape(-1, 3))
        nobs, nf = o.shape[0], len(f)
        out = np.zeros(2 * nobs * 6 * nf)
        self.L.ko_near2far(self.h, int(normal_axis), ip, float(normal_sign), float(eps), float(mu), bp, fp, op, nobs,
                           out.ctypes.data_as(C.POINTER(C.c_double)))
        z = out[0::2] + 1j * out[1::2]
        return z.reshape(nf, 6, nobs).transpose(2, 1, 0)

    def mode_amplitudes(self, normal_axis, ids4, mode_fields):
        """compute_mode_amplitudes (ModeMonitor.jl:345-515); mode_fields complex (4, n1, n2, nf) on the DFT grid."""
        i4, ip = _i(ids4)
        m = np.asarray(mode_fields, dtype=np.complex128)
        nf = m.shape[3]
        flat = np.ascontiguousarray(m.transpose(0, 3, 2, 1)).ravel()  # [4][nf][n2][n1]
        ri = np.empty(2 * flat.size)
        ri[0::2] = flat.real
        ri[1::2] = flat.imag
        ap, am, pm = np.zeros(2 * nf), np.zeros(2 * nf), np.zeros(nf)
        dp = C.POINTER(C.c_double)
        self.L.ko_mode_amplitudes(self.h, int(normal_axis), ip, ri.ctypes.data_as(dp), ap.ctypes.data_as(dp),
                                  am.ctypes.data_as(dp), pm.ctypes.data_as(dp))
        return ap[0::2] + 1j * ap[1::2], am[0::2] + 1j * am[1::2], pm

    def rasterize(self, objects, kinds_mask, smoothing=0):
        """init_geometry rasterisation + _apply_subpixel_smoothing! (Geometry.jl:150-246, 450-605, 795-972).
        objects: rows of 28 numbers (kind, centre, size, axes(9), eps_inv(3), mu_inv(3), sigma_D(3), sigma_B(3)).
        Returns the number of smoothed interface voxels per E component."""
        o, op = _d(np.asarray(objects, dtype=np.float64).reshape(-1, 28))
        cnt = (C.c_long * 3)()
        self.L.ko_rasterize(self.h, o.shape[0], op, int(kinds_mask), int(smoothing), cnt)
        return list(cnt)

    def diffraction(self, normal_axis, ids4, max_order, L1, L2, freqs, kinc1=0.0, kinc2=0.0):
        """get_diffraction_efficiencies (DiffractionMonitor.jl:87-165): (power, propagating), each (nf, 2M+1, 2M+1)."""
        i4, ip = _i(ids4)
        f, fp = _d(freqs)
        nord = 2 * int(max_order) + 1
        power = np.zeros(len(f) * nord * nord)
        prop = np.zeros(len(f) * nord * nord, dtype=np.int32)
        self.L.ko_diffraction(self.h, int(normal_axis), ip, int(max_order), float(L1), float(L2), float(kinc1), float(kinc2), fp,
                              power.ctypes.data_as(C.POINTER(C.c_double)), prop.ctypes.data_as(C.POINTER(C.c_int)))
        return power.reshape(len(f), nord, nord), prop.reshape(len(f), nord, nord)
