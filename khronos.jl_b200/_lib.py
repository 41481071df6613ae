"""ctypes binding of libkhronos_b200.so (the C ABI in include/khronos_b200.h).

The product has no CPU path: if the shared library is missing or cannot be
loaded this module raises, it never substitutes another implementation.
"""
import ctypes as C
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("KHRONOS_B200_LIB") or os.path.join(_HERE, "lib", "libkhronos_b200.so")
CSRC = os.path.join(_HERE, "csrc")

KHR_F32, KHR_F64 = 0, 1
GROUP_H, GROUP_E = 0, 1
MAT_EPS_INV, MAT_MU_INV, MAT_SIGMA_D, MAT_SIGMA_B, MAT_CHI3 = 0, 1, 2, 3, 4
TIME_CW, TIME_GAUSSIAN, TIME_HOST = 0, 1, 2


class GridDesc(C.Structure):
    _fields_ = [
        ("dtype", C.c_int32),
        ("n", C.c_int32 * 3),
        ("dl", C.c_double * 3),
        ("dt", C.c_double),
        ("z_start", C.c_int32),
        ("nz_local", C.c_int32),
        ("rank", C.c_int32),
        ("nranks", C.c_int32),
    ]


SHAPE_SPHERE, SHAPE_CUBOID, SHAPE_CYLINDER = 0, 1, 2


class KhrObject(C.Structure):
    _fields_ = [
        ("kind", C.c_int32),
        ("pad_", C.c_int32),
        ("center", C.c_double * 3),
        ("size", C.c_double * 3),
        ("axes", C.c_double * 9),
        ("eps_inv", C.c_double * 3),
        ("mu_inv", C.c_double * 3),
        ("sigma_d", C.c_double * 3),
        ("sigma_b", C.c_double * 3),
    ]


class KernelStat(C.Structure):
    _fields_ = [
        ("name", C.c_char * 96),
        ("launches", C.c_int64),
        ("total_ms", C.c_double),
        ("cells_per_launch", C.c_int64),
        ("alg_bytes_per_launch", C.c_double),
        ("ctas", C.c_int64),
        ("uniform_ctas", C.c_int64),
        ("ref_model_bytes_per_launch", C.c_double),
    ]


# every exported symbol of include/khronos_b200.h with its signature
_P = C.c_void_p
_I = C.c_int32
_SIGNATURES = {
    "khr_last_error": (C.c_char_p, []),
    "khr_version": (_I, []),
    "khr_ctx_create": (_I, [_I, C.POINTER(GridDesc), C.POINTER(_P)]),
    "khr_ctx_destroy": (_I, [_P]),
    "khr_set_pml_sigma": (_I, [_P, _I, _I, _P, _I]),
    "khr_set_grid_spacing": (_I, [_P, _I, _P, _I]),
    "khr_geometry_rasterize": (_I, [_P, C.POINTER(KhrObject), _I, _I, _I, C.POINTER(C.c_double), C.POINTER(C.c_int64)]),
    "khr_material_read": (_I, [_P, _I, _I, _P]),
    "khr_set_material_scalar": (_I, [_P, _I, C.c_double]),
    "khr_set_material_array": (_I, [_P, _I, _I, _P]),
    "khr_pole_register": (_I, [_P, C.c_double, C.c_double, _P, C.POINTER(_I)]),
    "khr_source_register": (_I, [_P, _I, C.POINTER(_I), C.POINTER(_I), _P, _I, C.POINTER(C.c_double), C.POINTER(_I)]),
    "khr_source_set_amplitude": (_I, [_P, _I, C.c_double, C.c_double]),
    "khr_set_sources_active": (_I, [_P, _I]),
    "khr_monitor_register": (_I, [_P, _I, C.POINTER(_I), C.POINTER(_I), _I, C.POINTER(C.c_double), _I, C.POINTER(_I)]),
    "khr_finalize_plan": (_I, [_P]),
    "khr_set_periodic": (_I, [_P, _I, _I]),
    "khr_set_complex_fields": (_I, [_P]),
    "khr_set_bloch": (_I, [_P, _I, C.c_double]),
    "khr_field_read_imag": (_I, [_P, _I, _P]),
    "khr_flux": (_I, [_P, C.POINTER(_I), _I, C.POINTER(C.c_double), _I]),
    "khr_near2far": (_I, [_P, C.POINTER(_I), _I, C.c_double, C.c_double, C.c_double, C.POINTER(C.c_double),
                          C.POINTER(C.c_double), _I, C.POINTER(C.c_double), _I, C.POINTER(C.c_double)]),
    "khr_mode_overlap": (_I, [_P, C.POINTER(_I), _I, C.POINTER(C.c_double), _I, _I, _I, C.POINTER(C.c_double)]),
    "khr_diffraction": (_I, [_P, C.POINTER(_I), _I, _I, C.c_double, C.c_double, C.c_double, C.c_double, C.POINTER(C.c_double), _I,
                             C.POINTER(C.c_double), C.POINTER(_I)]),
    "khr_step": (_I, [_P, _I]),
    "khr_step_h": (_I, [_P]),
    "khr_step_e": (_I, [_P]),
    "khr_dft_update": (_I, [_P, _I, C.c_double]),
    "khr_get_timestep": (_I, [_P, C.POINTER(C.c_int64)]),
    "khr_set_timestep": (_I, [_P, C.c_int64]),
    "khr_reset_fields": (_I, [_P]),
    "khr_comm_unique_id": (_I, [_P]),
    "khr_comm_init": (_I, [_P, _P, _I, _I]),
    "khr_halo_exchange": (_I, [_P, _I]),
    "khr_field_read": (_I, [_P, _I, _P]),
    "khr_field_write": (_I, [_P, _I, _P]),
    "khr_field_view": (_I, [_P, _I, C.POINTER(_P), C.POINTER(C.c_int64), C.POINTER(C.c_int64)]),
    "khr_monitor_read": (_I, [_P, _I, _P]),
    "khr_monitor_view": (_I, [_P, _I, C.POINTER(_P), C.POINTER(C.c_int64)]),
    "khr_monitor_norm": (_I, [_P, _I, C.POINTER(C.c_double)]),
    "khr_monitor_norms": (_I, [_P, C.POINTER(C.c_double), _I]),
    "khr_sync": (_I, [_P]),
    "khr_get_stream": (_I, [_P, C.POINTER(_P)]),
    "khr_last_step_timing": (_I, [_P, C.POINTER(C.c_double), C.POINTER(C.c_int64)]),
    "khr_voxel_census": (_I, [_P, C.POINTER(C.c_int64)]),
    "khr_device_bytes": (_I, [_P, C.POINTER(C.c_int64)]),
    "khr_set_profiling": (_I, [_P, _I]),
    "khr_kernel_stat_get": (_I, [_P, _I, C.POINTER(KernelStat), C.POINTER(_I)]),
    "khr_comm_stat_get": (_I, [_P, C.POINTER(C.c_double), C.POINTER(C.c_int64)]),
    "khr_graph_info": (_I, [_P, C.POINTER(C.c_int64), C.POINTER(C.c_int64)]),
}
EXPORTED_SYMBOLS = tuple(_SIGNATURES)

_LIB = None


def build(force=False):
    """Compile libkhronos_b200.so in-tree with nvcc for sm_100a (no GPU needed)."""
    srcs = [os.path.join(CSRC, f) for f in ("khronos_b200.cu", "step_kernels.cuh", "pml_tma.cuh", "post_kernels.cuh", "geom_kernels.cuh")]
    srcs.append(os.path.join(_HERE, "..", "include", "khronos_b200.h"))
    stale = (not os.path.exists(LIB_PATH)) or any(os.path.getmtime(s) > os.path.getmtime(LIB_PATH) for s in srcs)
    if force or stale:
        subprocess.check_call(["make", "-C", CSRC, "-s"] + (["-B"] if force else []))
    return LIB_PATH


class KhronosError(RuntimeError):
    pass


def lib():
    """Load the shared library; raises if it is missing (no fallback)."""
    global _LIB
    if _LIB is None:
        if not os.path.exists(LIB_PATH):
            raise KhronosError(
                "libkhronos_b200.so is not built (%s). Run `python -c 'import __graft_entry__ as g; g.build()'` "
                "or `make -C khronos.jl_b200/csrc`. There is no CPU fallback." % LIB_PATH
            )
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in _SIGNATURES.items():
            fn = getattr(L, name)  # AttributeError if the symbol is missing
            fn.restype = res
            fn.argtypes = args
        _LIB = L
    return _LIB


def check(status):
    if status != 0:
        raise KhronosError(lib().khr_last_error().decode("utf-8", "replace"))
