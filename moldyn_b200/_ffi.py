"""ctypes binding of include/moldyn_b200.h.  There is no fallback: if the CUDA library is missing or no
CUDA device is present, the first call raises."""
from __future__ import annotations

import ctypes as C
import os

from . import build as _build

MD_OK = 0
ERR_NAMES = {
    1: "MD_ERR_INVALID_ARGUMENT", 2: "MD_ERR_CUDA", 3: "MD_ERR_NCCL", 4: "MD_ERR_UNSUPPORTED",
    5: "MD_ERR_NEIGHBOUR_OVERFLOW", 6: "MD_ERR_NO_STATE", 7: "MD_ERR_NONFINITE", 8: "MD_ERR_DECOMPOSITION",
}
UNIQUE_ID_BYTES = 128
FORCE_FAST, FORCE_EXACT = 0, 1
LOOP_AUTO, LOOP_HOST, LOOP_CHUNK = 0, 1, 2
CELL_UNIFORM, CELL_FCC = 0, 1
CROSS_REFERENCE, CROSS_SYMMETRIC = 0, 1  # md_set_cross_type_mode
MAX_TYPES = 8
THERMOSTAT_NONE, THERMOSTAT_BERENDSEN, THERMOSTAT_NOSE_HOOVER = 0, 1, 2
BAROSTAT_NONE, BAROSTAT_BERENDSEN = 0, 1


class MdError(RuntimeError):
    def __init__(self, code, message):
        super().__init__(f"{ERR_NAMES.get(code, code)}: {message}")
        self.code = code


class Config(C.Structure):
    _fields_ = [("device", C.c_int32), ("force_mode", C.c_int32), ("loop_mode", C.c_int32),
                ("max_neighbours", C.c_int32), ("cell_subdiv", C.c_int32), ("reserved0", C.c_int32),
                ("skin", C.c_double), ("cell_atoms", C.c_double)]


class ThermostatC(C.Structure):
    _fields_ = [("kind", C.c_int32), ("reserved0", C.c_int32), ("tau", C.c_double), ("target", C.c_double),
                ("lambda_", C.c_double), ("psi", C.c_double)]


class BarostatC(C.Structure):
    _fields_ = [("kind", C.c_int32), ("reserved0", C.c_int32), ("beta", C.c_double), ("tau", C.c_double),
                ("target", C.c_double), ("myu", C.c_double)]


class MacroOut(C.Structure):
    _fields_ = [("kinetic_energy", C.c_double), ("thermal_energy", C.c_double), ("potential_energy", C.c_double),
                ("temperature", C.c_double), ("pressure", C.c_double), ("vcom", C.c_double * 3),
                ("momentum", C.c_double * 3), ("box", C.c_double * 3), ("lambda_", C.c_double),
                ("myu", C.c_double), ("n", C.c_int64)]


class Stats(C.Structure):
    _fields_ = [("steps", C.c_int64), ("rebuilds", C.c_int64), ("kernel_launches", C.c_int64),
                ("graph_launches", C.c_int64), ("loop_launches", C.c_int64), ("loop_steps", C.c_int64),
                ("cells", C.c_int32 * 3), ("nbr_capacity", C.c_int32), ("nbr_max", C.c_int32),
                ("peer_memory", C.c_int32), ("persistent_loop", C.c_int32), ("tile_lists", C.c_int32),
                ("skin", C.c_double), ("nbr_mean", C.c_double), ("n_owned", C.c_int64), ("n_ghost", C.c_int64),
                ("migrated", C.c_int64), ("wait_halo_ms", C.c_double), ("wait_sums_ms", C.c_double),
                ("force_atoms_ms", C.c_double), ("force_tail_ms", C.c_double), ("rebuild_ms", C.c_double),
                ("loop_phase_ms", C.c_double * 4)]


# every symbol include/moldyn_b200.h declares (tests check the library exports all of them)
SYMBOLS = [
    "md_create", "md_destroy", "md_last_error", "md_version", "md_lj_new", "md_lj_potential_and_force",
    "md_set_potential_lj", "md_upload_state", "md_download_state", "md_update_force", "md_step", "md_macro",
    "md_update_force_host", "md_calculate_host", "md_download_cells", "md_neighbour_counts",
    "md_neighbour_lists", "md_get_stats", "md_stream", "md_synchronize", "md_invalidate_lists", "md_time_kernels",
    "md_comm_unique_id", "md_comm_init", "md_local_count", "md_download_local", "md_plan_decomposition",
    "md_initialize_lattice", "md_measure_fp64_peak", "md_set_potential_pair", "md_set_cross_type_mode",
    "md_upload_state_typed", "md_macro_type",
]

_lib = None


def library_path() -> str:
    """The in-tree library; MOLDYN_B200_LIBRARY points at another build of the same sources (kernel A/B runs on the
    GPU box, see scripts/gpu_round1_final.sh) — still a CUDA build of this repo, never a fallback."""
    return os.environ.get("MOLDYN_B200_LIBRARY") or _build.LIB


def lib():
    """Loads libmoldyn_b200.so (built in-tree by moldyn_b200.build). Raises if it is not there."""
    global _lib
    if _lib is None:
        path = library_path()
        if not os.path.exists(path):
            raise ImportError(f"{path} is missing: run `python -m moldyn_b200.build` (needs nvcc); "
                              "moldyn_b200 has no CPU fallback")
        L = C.CDLL(path)
        vp, i64, f64, pd = C.c_void_p, C.c_int64, C.c_double, C.c_void_p
        sig = {
            "md_create": (C.c_int, [C.POINTER(Config), C.POINTER(vp)]),
            "md_destroy": (None, [vp]),
            "md_last_error": (C.c_char_p, [vp]),
            "md_version": (C.c_char_p, []),
            "md_lj_new": (C.c_int, [f64, f64, C.POINTER(f64), C.POINTER(f64)]),
            "md_lj_potential_and_force": (C.c_int, [f64, f64, f64, f64, f64, C.POINTER(f64), C.POINTER(f64)]),
            "md_set_potential_lj": (C.c_int, [vp, f64, f64, f64, f64]),
            "md_upload_state": (C.c_int, [vp, i64, pd, pd, pd, pd, pd, f64, pd]),
            "md_download_state": (C.c_int, [vp, pd, pd, pd, pd, pd, pd]),
            "md_update_force": (C.c_int, [vp]),
            "md_step": (C.c_int, [vp, i64, f64, C.POINTER(ThermostatC), C.POINTER(BarostatC)]),
            "md_macro": (C.c_int, [vp, C.POINTER(MacroOut)]),
            "md_macro_type": (C.c_int, [vp, C.c_int32, C.POINTER(MacroOut)]),
            "md_set_potential_pair": (C.c_int, [vp, C.c_int32, C.c_int32, f64, f64, f64, f64]),
            "md_set_cross_type_mode": (C.c_int, [vp, C.c_int32]),
            "md_upload_state_typed": (C.c_int, [vp, i64, pd, pd, pd, pd, pd, C.c_int32, pd, pd, pd]),
            "md_update_force_host": (C.c_int, [vp, i64, pd, f64, pd, pd, pd, pd]),
            "md_calculate_host": (C.c_int, [vp, i64, pd, pd, pd, pd, pd, f64, pd, f64, C.POINTER(ThermostatC),
                                            C.POINTER(BarostatC)]),
            "md_download_cells": (C.c_int, [vp, pd, pd]),
            "md_neighbour_counts": (C.c_int, [vp, pd]),
            "md_neighbour_lists": (C.c_int, [vp, pd, pd]),
            "md_get_stats": (C.c_int, [vp, C.POINTER(Stats)]),
            "md_stream": (vp, [vp]),
            "md_synchronize": (C.c_int, [vp]),
            "md_invalidate_lists": (C.c_int, [vp]),
            "md_time_kernels": (C.c_int, [vp, i64, f64, C.POINTER(ThermostatC), C.POINTER(BarostatC), pd, pd]),
            "md_comm_unique_id": (C.c_int, [pd]),
            "md_comm_init": (C.c_int, [vp, C.c_int, C.c_int, pd]),
            "md_local_count": (C.c_int, [vp, C.POINTER(i64), C.POINTER(i64)]),
            "md_download_local": (C.c_int, [vp, pd, pd, pd, pd, pd, pd, pd]),
            "md_plan_decomposition": (C.c_int, [i64, pd, f64, C.c_int, C.c_int, C.POINTER(f64), C.POINTER(f64),
                                                C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(i64)]),
            "md_initialize_lattice": (C.c_int, [vp, C.c_int, pd, pd, f64, f64, f64, C.c_uint64]),
            "md_measure_fp64_peak": (C.c_int, [vp, C.POINTER(f64)]),
        }
        assert set(sig) == set(SYMBOLS)
        for name, (res, args) in sig.items():
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


def check(ctx, rc):
    if rc != MD_OK:
        msg = lib().md_last_error(ctx)
        raise MdError(rc, msg.decode() if msg else "")
