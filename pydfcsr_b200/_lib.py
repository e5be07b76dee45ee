"""ctypes binding of libdfcsr_b200.so (include/dfcsr_b200.h).

The product path has no CPU fallback: if the shared library is missing or does not export the ABI
declared in the header, importing this module raises.
"""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("DFCSR_LIB", os.path.join(HERE, "libdfcsr_b200.so"))   # DFCSR_LIB: developer override

VOXEL_DOUBLES = 6
VOXEL_FLOATS = 8
VOXEL_F64, VOXEL_F32 = 0, 1
LATTICE_DOUBLES = 6
STATS_DOUBLES = 16
DF_SCALARS = 8
ABI_VERSION = 4
MAX_PEERS = 8
SKIP_MODES = {"auto": 0, "on": 1, "off": 2}

# indices of dfcsr_stat
(S_MEAN_X, S_MEAN_Z, S_SIGMA_X, S_SIGMA_Z, S_SLOPE, S_INTERCEPT, S_MEAN_XT, S_SIGMA_XT,
 S_SLICE_SIGMA_X, S_SLICE_COUNT, S_MEAN_PZ, S_SIGMA_PZ, S_N, S_ABSMAX_PX) = range(14)
STAT_BLOCKS = 1024


class Axis(C.Structure):
    _fields_ = [("start", C.c_double), ("stop", C.c_double), ("n", C.c_int32), ("_pad", C.c_int32)]

    @classmethod
    def make(cls, start, stop, n):
        return cls(float(start), float(stop), int(n), 0)


class History(C.Structure):
    _fields_ = [("d_ring", C.c_void_p), ("slice_elems", C.c_int64), ("cap", C.c_int32), ("head", C.c_int32),
                ("T", C.c_int32), ("X", C.c_int32), ("Z", C.c_int32), ("format", C.c_int32),
                ("min_t", C.c_double), ("min_x", C.c_double), ("min_z", C.c_double),
                ("delta_t", C.c_double), ("delta_x", C.c_double), ("delta_z", C.c_double),
                ("d_row_support", C.c_void_p)]


class Lattice(C.Structure):
    _fields_ = [("d_table", C.c_void_p), ("ns", C.c_int32), ("n_elements", C.c_int32),
                ("min_s", C.c_double), ("delta_s", C.c_double), ("d_rho", C.c_void_p), ("d_distance", C.c_void_p)]


class Element(C.Structure):
    _fields_ = [("kind", C.c_int32), ("fringe_entrance", C.c_int32), ("fringe_exit", C.c_int32), ("n_step", C.c_int32),
                ("L", C.c_double), ("g", C.c_double), ("e1", C.c_double), ("e2", C.c_double), ("k1", C.c_double),
                ("k2", C.c_double)]


class WakeParams(C.Structure):
    _fields_ = [("t", C.c_double), ("sigma_x", C.c_double), ("sigma_z", C.c_double), ("slope0", C.c_double),
                ("mean_x", C.c_double), ("formation_window", C.c_double), ("csr_scaling", C.c_double),
                ("nx", C.c_int32), ("nz", C.c_int32), ("skip_mode", C.c_int32), ("reserved", C.c_int32)]


class XGroupPlan(C.Structure):
    _fields_ = [("n_groups", C.c_int64), ("unit_nodes", C.c_int32), ("max_units", C.c_int32),
                ("workspace_bytes_per_group", C.c_int64), ("group_points", C.c_int32), ("reserved", C.c_int32)]


_P = C.c_void_p
_D = C.c_double
_I = C.c_int32
_L = C.c_int64

# name -> (restype, argtypes); must list every symbol declared in include/dfcsr_b200.h
SIGNATURES = {
    "dfcsr_abi_version": (C.c_int, []),
    "dfcsr_last_error": (C.c_char_p, []),
    "dfcsr_launch_count": (_L, []),
    "dfcsr_beam_stats_workspace": (_L, []),
    "dfcsr_stat_chunk": (_L, [_L]),
    "dfcsr_beam_stats": (C.c_int, [_P, _P, _P, _P, _L, C.POINTER(C.c_double), _P, _P, _P]),
    "dfcsr_beam_stats_partial": (C.c_int, [_I, _P, _P, _P, _P, _L, _L, _I, _I, C.POINTER(C.c_double), _P, _P,
                                           C.POINTER(C.c_uint64), _I, _P]),
    "dfcsr_beam_stats_final": (C.c_int, [_I, _P, _L, C.POINTER(C.c_double), _I, _I, _P, _P]),
    "dfcsr_mirror_to_host": (C.c_int, [_P, _P, _I, _P]),
    "dfcsr_beam_cov_workspace": (_L, []),
    "dfcsr_beam_cov": (C.c_int, [_P, _P, _P, _P, _P, _P, _L, C.POINTER(C.c_double), _P, _P, _P]),
    "dfcsr_beam_cov_partial": (C.c_int, [_P, _P, _P, _P, _P, _P, _L, _L, _I, _I, C.POINTER(C.c_double), _P,
                                         C.POINTER(C.c_uint64), _I, _P]),
    "dfcsr_beam_cov_final": (C.c_int, [_P, _L, C.POINTER(C.c_double), _P, _P]),
    "dfcsr_deposit_cic": (C.c_int, [_P, _P, _P, _L, _I, _D, _D, _I, _D, _D, _P, _P, _I, _P]),
    "dfcsr_deposit_cic_q": (C.c_int, [_P, _P, _P, _L, _L, _I, _D, _D, _I, _D, _D, _D, _P, _P]),
    "dfcsr_deposit_cic_finish": (C.c_int, [C.POINTER(C.c_uint64), _I, _I, _I, _L, _D, _P, _P, _P, _P]),
    "dfcsr_deposit_ngp": (C.c_int, [_P, _P, _L, _I, _D, _D, _I, _D, _D, _P, _P]),
    "dfcsr_make_df_workspace": (_L, [_I, _I]),
    "dfcsr_make_df": (C.c_int, [_P, _P, Axis, Axis, _I, _P, _P, _P, _D, _P, _P, _P, _P, _P]),
    "dfcsr_get_df": (C.c_int, [_P, _P, _P, _L, Axis, Axis, _D, _P, _P, _P, _P, _I, _P, _P, _P, _D, _P, _P, _P, _P]),
    "dfcsr_get_df_from_stats": (C.c_int, [_P, _P, _P, _L, _P, _D, _D, _I, _I, _P, _P, _P, _P, _P, _I, _P, _P, _P, _D, _P, _P,
                                          _P, _P]),
    "dfcsr_df_limits": (C.c_int, [_P, _D, _D, _P, _P]),
    "dfcsr_deposit_cic_q_dev": (C.c_int, [_P, _P, _P, _L, _L, _I, _I, _P, _P, _P, _P]),
    "dfcsr_deposit_cic_finish_dev": (C.c_int, [C.POINTER(C.c_uint64), _I, _I, _I, _L, _P, _P, _P, _P, _P]),
    "dfcsr_make_df_dev": (C.c_int, [_P, _P, _I, _I, _P, _I, _P, _P, _P, _D, _P, _P, _P, _P, _P]),
    "dfcsr_history_regrid": (C.c_int, [_P, Axis, Axis, Axis, Axis, _D, _P, _I, _P, _P, _P]),
    "dfcsr_history_row_support": (C.c_int, [_P, _I, _I, _I, _P, _P]),
    "dfcsr_history_pack": (C.c_int, [_P, _I, _I, _I, _P, _P]),
    "dfcsr_history_unpack": (C.c_int, [_P, _I, _I, _I, _P, _P]),
    "dfcsr_wake_mesh": (C.c_int, [C.POINTER(History), C.POINTER(Lattice), C.POINTER(WakeParams),
                                  _P, _P, _L, _L, _P, _P, _P, _P]),
    "dfcsr_wake_grid": (C.c_int, [C.POINTER(History), C.POINTER(Lattice), C.POINTER(WakeParams), Axis, Axis, _D, _D,
                                  _L, _L, _P, _P, _P, _P]),
    "dfcsr_wake_grid_peers": (C.c_int, [C.POINTER(History), C.POINTER(Lattice), C.POINTER(WakeParams), Axis, Axis, _D, _D,
                                        _L, _L, _L, C.POINTER(C.c_uint64), _I, _P, _P]),
    "dfcsr_wake_xgroup_plan": (C.c_int, [C.POINTER(History), C.POINTER(WakeParams), Axis, Axis, C.POINTER(XGroupPlan)]),
    "dfcsr_wake_grid_xgroups": (C.c_int, [C.POINTER(History), C.POINTER(Lattice), C.POINTER(WakeParams), Axis, Axis, _D, _D,
                                          _L, _L, _L, _P, _P, C.POINTER(C.c_uint64), _I, _P, _L, _P, _P]),
    "dfcsr_wake_preload": (C.c_int, []),
    "dfcsr_wake_uses_skipping": (C.c_int, [C.POINTER(History), C.POINTER(WakeParams)]),
    "dfcsr_wake_point_debug": (C.c_int, [C.POINTER(History), C.POINTER(Lattice), C.POINTER(WakeParams),
                                         _D, _D, _P, _P, _L, _P, _P, _P]),
    "dfcsr_apply_kick": (C.c_int, [_P, _P, _P, _P, _L, _D, _D, _P, _P, Axis, Axis, _D, _D, _I, _P]),
    "dfcsr_track_element": (C.c_int, [_P, _P, _P, _P, _P, _P, _L, C.POINTER(Element), _D, _D, _P]),
    "dfcsr_track_linear": (C.c_int, [_P, _P, _P, _P, _P, _P, _L, C.POINTER(C.c_double), _P]),
    "dfcsr_sgolay2d": (C.c_int, [_P, _I, _I, _I, _P, _I, _P, _P]),
    "dfcsr_selftest_sqrt": (C.c_int, [_L, C.c_uint64, _D, _D, C.POINTER(C.c_uint64), _P]),
}


class DfcsrError(RuntimeError):
    pass


def _load():
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: build it with `python -m pydfcsr_b200.build` "
            "(pydfcsr_b200 has no CPU fallback)")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        try:
            fn = getattr(lib, name)
        except AttributeError as e:
            raise ImportError(f"libdfcsr_b200.so does not export {name}") from e
        fn.restype = res
        fn.argtypes = args
    if lib.dfcsr_abi_version() != ABI_VERSION:
        raise ImportError("libdfcsr_b200.so ABI version mismatch; rebuild")
    return lib


lib = _load()


def check(rc: int, what: str = "") -> None:
    if rc != 0:
        msg = lib.dfcsr_last_error().decode("utf-8", "replace")
        raise DfcsrError(f"{what or 'libdfcsr_b200'} failed ({rc}): {msg}")
