"""Host-driven Monte Carlo over the C ABI (``include/maniac_host.h``): the part MANIAC's
Fortran move drivers play.  The host proposes and decides, the GPU computes energies;
every MC step is one ``mgpu_trial_batch`` + one ``mgpu_commit_batch`` over all walkers,
with host<->device copies of the proposals / energies inside the call."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import capi
from .engine import TRACE_DTYPE, Engine, ManiacAbort

_SIGS = {
    "mhost_create": (C.c_void_p, [C.POINTER(capi.MgpuSystem), C.c_uint64]),
    "mhost_destroy": (None, [C.c_void_p]),
    "mhost_last_error": (C.c_char_p, []),
    "mhost_run": (C.c_int, [C.c_void_p, C.c_int64, C.c_int32, C.c_void_p]),
    "mhost_set_chemical_potential": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.c_double]),
    "mhost_get_count": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32]),
    "mhost_get_energy": (C.c_int, [C.c_void_p, C.c_int32, C.POINTER(C.c_double)]),
    "mhost_get_counters": (C.c_int, [C.c_void_p, C.c_int32, C.POINTER(C.c_int64)]),
    "mhost_get_molecule": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.POINTER(C.c_double), C.POINTER(C.c_double)]),
    "mhost_adjust_move_step_sizes": (C.c_int, [C.c_void_p]),
    "mhost_get_step_sizes": (C.c_int, [C.c_void_p, C.c_int32, C.POINTER(C.c_double)]),
    "mhost_get_traffic": (C.c_int, [C.c_void_p, C.POINTER(C.c_int64), C.POINTER(C.c_int64), C.POINTER(C.c_int64)]),
}


class HostMonteCarlo:
    def __init__(self, engine: Engine, seed: int):
        self.eng = engine
        self.L = capi.lib()
        for name, (res, args) in _SIGS.items():
            fn = getattr(self.L, name)
            fn.restype, fn.argtypes = res, args
        self.h = self.L.mhost_create(C.byref(engine._sys_struct), int(seed) & 0xFFFFFFFFFFFFFFFF)
        if not self.h:
            raise ManiacAbort(self.L.mhost_last_error().decode())

    def close(self):
        if self.h:
            self.L.mhost_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def run(self, n_steps, trace_walker=None):
        tr = ptr = None
        tw = -1
        if trace_walker is not None:
            tr = np.zeros(n_steps, dtype=TRACE_DTYPE)
            ptr = tr.ctypes.data_as(C.c_void_p)
            tw = trace_walker
        if self.L.mhost_run(self.h, n_steps, tw, ptr):
            raise ManiacAbort(self.L.mhost_last_error().decode())
        return tr

    def set_chemical_potential(self, res, mu, walker=0):
        if self.L.mhost_set_chemical_potential(self.h, walker, res, float(mu)):
            raise ManiacAbort(self.L.mhost_last_error().decode())

    def count(self, res, walker=0):
        return self.L.mhost_get_count(self.h, walker, res)

    def energy(self, walker=0):
        out = np.zeros(6)
        self.L.mhost_get_energy(self.h, walker, out.ctypes.data_as(C.POINTER(C.c_double)))
        return out

    def adjust_move_step_sizes(self):
        self.L.mhost_adjust_move_step_sizes(self.h)

    def step_sizes(self, walker=0):
        out = (C.c_double * 2)()
        self.L.mhost_get_step_sizes(self.h, walker, out)
        return float(out[0]), float(out[1])

    def counters(self, walker=0):
        out = (C.c_int64 * 12)()
        self.L.mhost_get_counters(self.h, walker, out)
        return np.array(out[:]).reshape(6, 2)

    def get_molecule(self, res, mol, walker=0):
        com, off = np.zeros(3), np.zeros((self.eng.natom[res], 3))
        self.L.mhost_get_molecule(self.h, walker, res, mol, com.ctypes.data_as(C.POINTER(C.c_double)),
                                  off.ctypes.data_as(C.POINTER(C.c_double)))
        return com, off

    def traffic(self):
        a, b, c = C.c_int64(), C.c_int64(), C.c_int64()
        self.L.mhost_get_traffic(self.h, C.byref(a), C.byref(b), C.byref(c))
        return dict(h2d_bytes=a.value, d2h_bytes=b.value, trials=c.value)
