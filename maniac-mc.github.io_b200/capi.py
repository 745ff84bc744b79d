"""ctypes binding of the C ABI declared in ``include/maniac_gpu.h`` (``libmaniac_gpu.so``).

The library is built in-tree by ``__graft_entry__.build()`` (nvcc, sm_100a).  Loading it
needs the CUDA runtime only; calling ``mgpu_init`` needs a GPU -- there is no CPU fallback.
"""
from __future__ import annotations

import ctypes as C
from pathlib import Path

MAX_RES = 8
MAX_SITES = 16

_HERE = Path(__file__).resolve().parent
import os as _os
# MANIAC_GPU_LIB: another build of the SAME library (kernel-variant experiments); still CUDA, still no CPU path
LIB_PATH = Path(_os.environ["MANIAC_GPU_LIB"]) if _os.environ.get("MANIAC_GPU_LIB") else _HERE / "libmaniac_gpu.so"
_LIB = None


class MgpuResidue(C.Structure):
    _fields_ = [("natom", C.c_int32), ("is_active", C.c_int32), ("nmol", C.c_int32), ("capacity", C.c_int32),
                ("charges", C.POINTER(C.c_double)), ("types", C.POINTER(C.c_int32)),
                ("com", C.POINTER(C.c_double)), ("offset", C.POINTER(C.c_double)),
                ("mass", C.c_double), ("fugacity", C.c_double), ("chemical_potential", C.c_double)]


class MgpuSystem(C.Structure):
    _fields_ = [("matrix", C.c_double * 9), ("lo", C.c_double * 3), ("nres", C.c_int32),
                ("residues", C.POINTER(MgpuResidue)), ("ntypes", C.c_int32),
                ("epsilon", C.POINTER(C.c_double)), ("sigma", C.POINTER(C.c_double)),
                ("temperature", C.c_double), ("ewald_tolerance", C.c_double), ("real_space_cutoff", C.c_double),
                ("translation_step", C.c_double), ("rotation_step_angle", C.c_double),
                ("p_translation", C.c_double), ("p_rotation", C.c_double), ("p_swap", C.c_double),
                ("p_insertion_deletion", C.c_double), ("p_widom", C.c_double),
                ("n_walkers", C.c_int32), ("device", C.c_int32)]


class MgpuStepTrace(C.Structure):
    _fields_ = [("move", C.c_int32), ("res", C.c_int32), ("mol", C.c_int32), ("accepted", C.c_int32),
                ("dE", C.c_double), ("prob", C.c_double), ("e_old", C.c_double * 6), ("e_new", C.c_double * 6)]


_pd = C.POINTER(C.c_double)
_pi = C.POINTER(C.c_int32)
_pl = C.POINTER(C.c_int64)
_pu = C.POINTER(C.c_uint64)
I, D, L64, U64 = C.c_int32, C.c_double, C.c_int64, C.c_uint64

# every symbol include/maniac_gpu.h declares: name -> (restype, argtypes)
SIGNATURES = {
    "mgpu_init": (C.c_int, [C.POINTER(MgpuSystem)]),
    "mgpu_finalize": (None, []),
    "mgpu_last_error": (C.c_char_p, []),
    "mgpu_device_info": (C.c_int, [C.c_char_p, C.c_int, C.POINTER(C.c_int), _pd]),
    "mgpu_get_ewald": (C.c_int, [_pd, _pi, _pi, _pd]),
    "mgpu_get_kvectors": (C.c_int, [_pi, _pi, _pi, _pd, _pd]),
    "mgpu_get_box": (C.c_int, [_pd, _pd, _pd, _pi]),
    "mgpu_get_launch_info": (C.c_int, [_pi, _pl, _pi]),
    "mgpu_get_triclinic_candidates": (C.c_int, [_pi]),
    "mgpu_get_sweep_shape": (C.c_int, [C.c_int32, _pi, _pi]),
    "mgpu_plan_sweep_shape": (C.c_int, [C.c_int32, C.c_int32, C.c_int32, C.c_int32, _pi, _pi]),
    "mgpu_plan_triclinic": (C.c_int, [_pd, _pi, _pd, _pi, _pi, _pi, C.POINTER(C.c_uint32)]),
    "mgpu_plan_framework_order": (C.c_int, [_pd, _pd, C.c_int32, _pd, _pi, _pi]),
    "mgpu_get_thermo": (C.c_int, [I, _pd, _pd, _pd]),
    "mgpu_set_molecule": (C.c_int, [I, I, I, _pd, _pd]),
    "mgpu_get_molecule": (C.c_int, [I, I, I, _pd, _pd]),
    "mgpu_set_count": (C.c_int, [I, I, I]),
    "mgpu_get_count": (C.c_int, [I, I, _pi]),
    "mgpu_set_chemical_potential": (C.c_int, [I, I, D]),
    "mgpu_set_fugacity": (C.c_int, [I, I, D]),
    "mgpu_set_chemical_potentials": (C.c_int, [I, I, I, _pd]),
    "mgpu_get_counts": (C.c_int, [I, I, I, _pi]),
    "mgpu_get_Ak": (C.c_int, [I, _pd]),
    "mgpu_get_energy": (C.c_int, [I, _pd]),
    "mgpu_set_option": (C.c_int, [I, I]),
    "mgpu_total_energy": (C.c_int, [I, _pd]),
    "mgpu_pairwise_energy_for_molecule": (C.c_int, [I, I, I, I, _pd, _pd, _pd, _pd]),
    "mgpu_ewald_self_energy_single_mol": (C.c_int, [I, _pd]),
    "mgpu_intra_res_real_coulomb_energy": (C.c_int, [I, I, I, _pd, _pd, _pd]),
    "mgpu_reciprocal_ewald_energy": (C.c_int, [I, _pd]),
    "mgpu_old_energy": (C.c_int, [I, I, I, I, _pd]),
    "mgpu_new_energy": (C.c_int, [I, I, I, I, _pd, _pd, _pd]),
    "mgpu_swap_energy": (C.c_int, [I, I, I, I, _pd, _pd, _pd, _pd]),
    "mgpu_commit": (C.c_int, [I]),
    "mgpu_rollback": (C.c_int, [I]),
    "mgpu_trial_batch": (C.c_int, [I, _pi, _pi, _pi, _pi, _pd, _pd, _pd, _pd]),
    "mgpu_commit_batch": (C.c_int, [I, _pi, _pi]),
    "mgpu_seed": (C.c_int, [U64]),
    "mgpu_get_rng_state": (C.c_int, [I, _pu]),
    "mgpu_sweep": (C.c_int, [I, I, L64, I, C.c_void_p]),
    "mgpu_host_alloc": (C.c_void_p, [C.c_size_t]),
    "mgpu_host_free": (None, [C.c_void_p]),
    "mgpu_record_doubles_max": (L64, []),
    "mgpu_save_walkers": (C.c_int, [I, I, C.c_void_p, L64, C.c_void_p]),
    "mgpu_load_walkers": (C.c_int, [I, I, C.c_void_p, C.c_void_p]),
    "mgpu_block": (C.c_int, [I, I, L64, C.c_void_p, C.c_void_p, C.c_void_p, L64, C.c_void_p]),
    "mgpu_get_traffic": (C.c_int, [_pl, _pl, I]),
    "mgpu_adjust_move_step_sizes": (C.c_int, [I, I]),
    "mgpu_get_step_sizes": (C.c_int, [I, _pd]),
    "mgpu_set_step_sizes": (C.c_int, [I, D, D]),
    "mgpu_get_counters": (C.c_int, [I, _pl]),
    "mgpu_get_widom": (C.c_int, [I, I, _pd, _pl]),
    "mgpu_get_averages": (C.c_int, [I, I, _pd]),
    "mgpu_reset_averages": (C.c_int, []),
    "mgpu_widom_batch": (C.c_int, [I, I, L64, L64, U64, _pd, _pd, _pl]),
    "mgpu_nccl_unique_id": (C.c_int, [C.c_char_p]),
    "mgpu_nccl_init": (C.c_int, [C.c_char_p, I, I]),
    "mgpu_reduce_averages": (C.c_int, [_pd, I]),
    "mgpu_nccl_finalize": (None, []),
    "mgpu_nccl_last_error": (C.c_char_p, []),
    "mgpu_timing_reset": (C.c_int, []),
    "mgpu_timing_get": (C.c_int, [C.c_char_p, _pd, _pl]),
    "mgpu_get_pair_counts": (C.c_int, [_pl]),
    "mgpu_get_screened_pairs": (C.c_int, [_pl]),
    "mgpu_reset_pair_counts": (C.c_int, []),
    "mgpu_measure_fp64_peak": (C.c_int, [_pd, _pd]),
    "mgpu_measure_l2_peak": (C.c_int, [_pd]),
    "mgpu_selftest_math": (C.c_int, [_pd, _pd]),
    "mgpu_coulomb_table_check": (C.c_int, [D, D, D, I, _pd, _pd]),
}


class EngineUnavailable(RuntimeError):
    """libmaniac_gpu.so is missing or cannot be loaded.  There is no CPU fallback."""


def lib():
    global _LIB
    if _LIB is not None:
        return _LIB
    if not LIB_PATH.exists():
        raise EngineUnavailable(
            f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(nvcc, sm_100a).  The energy path has no CPU fallback.")
    try:
        L = C.CDLL(str(LIB_PATH), mode=C.RTLD_GLOBAL)
    except OSError as e:  # pragma: no cover - depends on the box
        raise EngineUnavailable(f"cannot load {LIB_PATH}: {e}") from e
    for name, (res, args) in SIGNATURES.items():
        if _os.environ.get("MANIAC_GPU_LIB") and not hasattr(L, name):
            continue                # kernel-variant experiments may load an older build of the library
        fn = getattr(L, name)       # AttributeError here = header / library mismatch
        fn.restype, fn.argtypes = res, args
    _LIB = L
    return L
