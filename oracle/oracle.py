"""ctypes binding of the CPU oracle (``oracle/maniac_oracle.c``).

TEST INFRASTRUCTURE.  Import this only from ``tests/``, ``__graft_entry__.smoke()`` and
``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs.  The product package never
imports it.
"""
from __future__ import annotations

import ctypes as C
import subprocess
from pathlib import Path

import numpy as np

_HERE = Path(__file__).resolve().parent
_LIB = None

E_NON_COULOMB, E_COULOMB, E_RECIP, E_SELF, E_INTRA, E_TOTAL = range(6)
KIND_MOVE, KIND_CREATE, KIND_DELETE = 0, 1, 2
MV_NONE, MV_TRANSLATE, MV_ROTATE, MV_CREATE, MV_DELETE, MV_SWAP, MV_WIDOM = range(7)


class StepTrace(C.Structure):
    _fields_ = [("move", C.c_int32), ("res", C.c_int32), ("mol", C.c_int32), ("accepted", C.c_int32),
                ("dE", C.c_double), ("prob", C.c_double), ("e_old", C.c_double * 6), ("e_new", C.c_double * 6)]


TRACE_DTYPE = np.dtype([("move", "i4"), ("res", "i4"), ("mol", "i4"), ("accepted", "i4"),
                        ("dE", "f8"), ("prob", "f8"), ("e_old", "f8", 6), ("e_new", "f8", 6)])
assert TRACE_DTYPE.itemsize == C.sizeof(StepTrace)


def build(force: bool = False) -> Path:
    """Builds both variants; returns the one in use: -O2 (parity tests) or, with MANIAC_ORACLE_VARIANT=o3 in the
    environment, the -O3 build of the same source that bench.py times as the CPU baseline."""
    import os
    src = _HERE / "maniac_oracle.c"
    newest = max(src.stat().st_mtime, (_HERE / "maniac_oracle.h").stat().st_mtime)
    for name in ("libmaniac_oracle.so", "libmaniac_oracle_o3.so"):
        so = _HERE / name
        if force or not so.exists() or so.stat().st_mtime < newest:
            subprocess.run(["make", "-C", str(_HERE), "-B", name], check=True, stdout=subprocess.DEVNULL)
    return _HERE / ("libmaniac_oracle_o3.so" if os.environ.get("MANIAC_ORACLE_VARIANT", "") == "o3" else "libmaniac_oracle.so")


def lib():
    global _LIB
    if _LIB is not None:
        return _LIB
    L = C.CDLL(str(build()))
    P, D, I, L64, U64 = C.c_void_p, C.c_double, C.c_int, C.c_int64, C.c_uint64
    pd, pi = C.POINTER(C.c_double), C.POINTER(C.c_int)
    sig = {
        "orc_create": (P, []), "orc_destroy": (None, [P]), "orc_last_error": (C.c_char_p, [P]),
        "orc_const_PI": (D, []), "orc_const_TWOPI": (D, []), "orc_const_SQRTPI": (D, []),
        "orc_const_EPS0_INV_real": (D, []), "orc_const_KB_kcalmol": (D, []),
        "orc_set_box": (I, [P, pd, pd]),
        "orc_add_residue": (I, [P, I, I, pd, pi, D, I]),
        "orc_lj_begin": (I, [P, I]), "orc_lj_pair_coeff": (I, [P, I, I, D, D]), "orc_lj_finalize": (I, [P]),
        "orc_lj_get": (I, [P, pd, pd]),
        "orc_set_molecule": (I, [P, I, I, pd, pd]), "orc_get_molecule": (I, [P, I, I, pd, pd]),
        "orc_set_count": (I, [P, I, I]), "orc_get_count": (I, [P, I]),
        "orc_setup_ewald": (I, [P, D, D]),
        "orc_get_ewald": (I, [P, pd, pi, pi, pd]),
        "orc_get_kvectors": (I, [P, pi, pi, pi, pd, pd]),
        "orc_get_box": (I, [P, pd, pd, pd, pi]),
        "orc_set_thermo": (I, [P, D]), "orc_set_fugacity": (I, [P, I, D]),
        "orc_set_chemical_potential": (I, [P, I, D]),
        "orc_set_mc_input": (I, [P, D, D, D, D, D, D, D]),
        "orc_get_beta": (D, [P]), "orc_get_lambda": (D, [P, I]), "orc_get_mu": (D, [P, I]),
        "orc_minimum_image_distance": (D, [P, I, I, I, I, I, I]),
        "orc_pairwise_energy_for_molecule": (I, [P, I, I, I, pd, pd]),
        "orc_update_system_energy": (I, [P, pd]), "orc_get_energy": (I, [P, pd]),
        "orc_ewald_self_energy_single_mol": (D, [P, I]),
        "orc_intra_res_real_coulomb_energy": (D, [P, I, I]),
        "orc_reciprocal_ewald_energy": (D, [P]), "orc_get_Ak": (I, [P, pd]),
        "orc_compute_ewald_phase_factors": (I, [P, I, I]),
        "orc_save_single_mol_fourier_terms": (I, [P, I, I]),
        "orc_restore_single_mol_fourier": (I, [P, I, I]),
        "orc_update_reciprocal_amplitude_single_mol": (I, [P, I, I, I]),
        "orc_compute_old_energy": (I, [P, I, I, I, pd]), "orc_compute_new_energy": (I, [P, I, I, I, pd]),
        "orc_apply_PBC": (None, [P, pd]), "orc_wrap_into_box": (None, [P, pd]),
        "orc_seed_rng": (None, [P, U64]), "orc_rand_uniform": (D, [P]),
        "orc_get_rng_state": (None, [P, C.POINTER(U64)]), "orc_set_rng_state": (None, [P, C.POINTER(U64)]),
        "orc_set_uniform_stream": (I, [P, pd, L64]), "orc_uniform_stream_used": (L64, [P]),
        "orc_attempt_translation_move": (I, [P, I, I, C.POINTER(StepTrace)]),
        "orc_attempt_rotation_move": (I, [P, I, I, C.POINTER(StepTrace)]),
        "orc_attempt_creation_move": (I, [P, I, I, C.POINTER(StepTrace)]),
        "orc_attempt_deletion_move": (I, [P, I, I, C.POINTER(StepTrace)]),
        "orc_attempt_swap_move": (I, [P, I, I, C.POINTER(StepTrace)]),
        "orc_widom_trial": (I, [P, I, I, C.POINTER(StepTrace)]),
        "orc_monte_carlo_steps": (I, [P, L64, C.c_void_p]),
        "orc_get_counters": (I, [P, C.POINTER(L64)]),
        "orc_adjust_move_step_sizes": (I, [P]),
        "orc_get_step_sizes": (I, [P, C.POINTER(D)]),
        "orc_get_widom": (I, [P, I, pd, C.POINTER(L64)]), "orc_reset_widom": (I, [P]),
        "orc_widom_batch": (I, [P, I, L64, L64, U64, pd, pd, C.POINTER(L64)]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(L, name)
        fn.restype, fn.argtypes = res, args
    _LIB = L
    return L


def _pd(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def _pi(a):
    return a.ctypes.data_as(C.POINTER(C.c_int))


class OracleError(RuntimeError):
    pass


class Oracle:
    """One walker of the reference algorithm on the CPU."""

    def __init__(self, system=None, capacity=None):
        self.L = lib()
        self.h = self.L.orc_create()
        self._keep = []
        if system is not None:
            self.load(system, capacity)

    def __del__(self):
        try:
            if self.h:
                self.L.orc_destroy(self.h)
                self.h = None
        except Exception:
            pass

    def _ck(self, rc):
        if rc:
            raise OracleError(self.L.orc_last_error(self.h).decode())

    # -- construction ---------------------------------------------------------------
    def load(self, sysm, capacity=None):
        L, h = self.L, self.h
        m = np.ascontiguousarray(sysm.matrix, dtype=np.float64)
        lo = np.ascontiguousarray(sysm.lo, dtype=np.float64)
        self._ck(L.orc_set_box(h, _pd(m), _pd(lo)))
        self.natom = []
        for r in sysm.residues:
            cap = r.nmol if not r.active else max(r.nmol + 1, capacity or (r.nmol + 64))
            cap = max(cap, 1)
            ch = np.ascontiguousarray(r.charges, dtype=np.float64)
            ty = np.ascontiguousarray(r.types, dtype=np.int32)
            rid = L.orc_add_residue(h, r.natom, int(r.active), _pd(ch), _pi(ty), float(r.mass), cap)
            if rid < 0:
                self._ck(1)
            self.natom.append(r.natom)
            for mi in range(r.nmol):
                self.set_molecule(rid, mi, r.com[mi], r.offset[mi])
            self._ck(L.orc_set_count(h, rid, r.nmol))
            if r.active:
                if r.fugacity >= 0.0:
                    L.orc_set_fugacity(h, rid, float(r.fugacity))
                else:
                    L.orc_set_chemical_potential(h, rid, float(r.chemical_potential))
        self.nres = len(sysm.residues)
        self._ck(L.orc_lj_begin(h, sysm.ntypes))
        for (ti, tj, eps, sig) in sysm.pair_coeff:
            self._ck(L.orc_lj_pair_coeff(h, ti, tj, eps, sig))
        self._ck(L.orc_lj_finalize(h))
        self.ntypes = sysm.ntypes
        self._ck(L.orc_setup_ewald(h, sysm.ewald_tolerance, sysm.real_space_cutoff))
        self._ck(L.orc_set_thermo(h, sysm.temperature))
        self._ck(L.orc_set_mc_input(h, sysm.translation_step, sysm.rotation_step_angle, sysm.p_translation,
                                    sysm.p_rotation, sysm.p_swap, sysm.p_insertion_deletion, sysm.p_widom))
        return self

    # -- accessors ------------------------------------------------------------------
    def set_molecule(self, res, mol, com, offset):
        com = np.ascontiguousarray(com, dtype=np.float64)
        off = np.ascontiguousarray(offset, dtype=np.float64)
        self._ck(self.L.orc_set_molecule(self.h, res, mol, _pd(com), _pd(off)))

    def get_molecule(self, res, mol):
        com = np.zeros(3)
        off = np.zeros((self.natom[res], 3))
        self._ck(self.L.orc_get_molecule(self.h, res, mol, _pd(com), _pd(off)))
        return com, off

    def set_count(self, res, n):
        self._ck(self.L.orc_set_count(self.h, res, n))

    def count(self, res):
        return self.L.orc_get_count(self.h, res)

    def ewald(self):
        alpha, rc = C.c_double(), C.c_double()
        kmax = (C.c_int * 3)()
        nk = C.c_int()
        self.L.orc_get_ewald(self.h, C.byref(alpha), kmax, C.byref(nk), C.byref(rc))
        return dict(alpha=alpha.value, kmax=list(kmax), nk=nk.value, rc=rc.value)

    def lj_table(self):
        eps = np.zeros((self.ntypes, self.ntypes))
        sig = np.zeros((self.ntypes, self.ntypes))
        self.L.orc_lj_get(self.h, _pd(eps), _pd(sig))
        return eps, sig

    def box(self):
        m, r = np.zeros((3, 3)), np.zeros((3, 3))
        v, sh = C.c_double(), C.c_int()
        self.L.orc_get_box(self.h, _pd(m), _pd(r), C.byref(v), C.byref(sh))
        return dict(matrix=m, reciprocal=r, volume=v.value, shape=sh.value)

    def Ak(self):
        nk = self.ewald()["nk"]
        a = np.zeros(2 * nk)
        self.L.orc_get_Ak(self.h, _pd(a))
        return a[0::2] + 1j * a[1::2]

    @property
    def beta(self):
        return self.L.orc_get_beta(self.h)

    # -- energies ---------------------------------------------------------------------
    def update_system_energy(self):
        out = np.zeros(6)
        self._ck(self.L.orc_update_system_energy(self.h, _pd(out)))
        return out

    def energy(self):
        out = np.zeros(6)
        self.L.orc_get_energy(self.h, _pd(out))
        return out

    def pairwise_energy_for_molecule(self, res, mol, skip_ordering_check=True):
        a, b = C.c_double(), C.c_double()
        self._ck(self.L.orc_pairwise_energy_for_molecule(self.h, res, mol, int(skip_ordering_check), C.byref(a), C.byref(b)))
        return a.value, b.value

    def minimum_image_distance(self, r1, m1, a1, r2, m2, a2):
        return self.L.orc_minimum_image_distance(self.h, r1, m1, a1, r2, m2, a2)

    def compute_old_energy(self, res, mol, kind=KIND_MOVE):
        out = np.zeros(6)
        self._ck(self.L.orc_compute_old_energy(self.h, res, mol, kind, _pd(out)))
        return out

    def compute_new_energy(self, res, mol, kind=KIND_MOVE):
        out = np.zeros(6)
        self._ck(self.L.orc_compute_new_energy(self.h, res, mol, kind, _pd(out)))
        return out

    def save_fourier(self, res, mol):
        self._ck(self.L.orc_save_single_mol_fourier_terms(self.h, res, mol))

    def restore_fourier(self, res, mol):
        self._ck(self.L.orc_restore_single_mol_fourier(self.h, res, mol))

    def reciprocal_ewald_energy(self):
        return self.L.orc_reciprocal_ewald_energy(self.h)

    def self_energy_single_mol(self, res):
        return self.L.orc_ewald_self_energy_single_mol(self.h, res)

    def intra_energy(self, res, mol):
        return self.L.orc_intra_res_real_coulomb_energy(self.h, res, mol)

    # -- RNG / MC -----------------------------------------------------------------------
    def seed(self, seed):
        self.L.orc_seed_rng(self.h, int(seed) & 0xFFFFFFFFFFFFFFFF)

    def rng_state(self):
        st = (C.c_uint64 * 4)()
        self.L.orc_get_rng_state(self.h, st)
        return [int(x) for x in st]

    def rand_uniform(self):
        return self.L.orc_rand_uniform(self.h)

    def set_uniform_stream(self, u):
        u = np.ascontiguousarray(u, dtype=np.float64)
        self._keep = [u]
        self.L.orc_set_uniform_stream(self.h, _pd(u), u.size)

    def monte_carlo_steps(self, n, trace=True):
        tr = np.zeros(n, dtype=TRACE_DTYPE) if trace else None
        ptr = tr.ctypes.data_as(C.c_void_p) if trace else None
        self._ck(self.L.orc_monte_carlo_steps(self.h, n, ptr))
        return tr

    def _move(self, fn, res, mol):
        t = StepTrace()
        self._ck(fn(self.h, res, mol, C.byref(t)))
        return t

    def attempt_translation_move(self, res, mol):
        return self._move(self.L.orc_attempt_translation_move, res, mol)

    def attempt_rotation_move(self, res, mol):
        return self._move(self.L.orc_attempt_rotation_move, res, mol)

    def attempt_creation_move(self, res, mol):
        return self._move(self.L.orc_attempt_creation_move, res, mol)

    def attempt_deletion_move(self, res, mol):
        return self._move(self.L.orc_attempt_deletion_move, res, mol)

    def attempt_swap_move(self, res, mol):
        return self._move(self.L.orc_attempt_swap_move, res, mol)

    def widom_trial(self, res, mol):
        return self._move(self.L.orc_widom_trial, res, mol)

    def adjust_move_step_sizes(self):
        self._ck(self.L.orc_adjust_move_step_sizes(self.h))

    def step_sizes(self):
        out = (C.c_double * 2)()
        self.L.orc_get_step_sizes(self.h, out)
        return float(out[0]), float(out[1])

    def counters(self):
        out = (C.c_int64 * 12)()
        self.L.orc_get_counters(self.h, out)
        return np.array(out[:]).reshape(6, 2)

    def widom(self, res):
        w, n = C.c_double(), C.c_int64()
        self.L.orc_get_widom(self.h, res, C.byref(w), C.byref(n))
        return w.value, n.value

    def widom_batch(self, res, first_id, n, seed, want_dE=True):
        dE = np.zeros(n) if want_dE else None
        sw, nok = C.c_double(), C.c_int64()
        self._ck(self.L.orc_widom_batch(self.h, res, first_id, n, int(seed) & 0xFFFFFFFFFFFFFFFF,
                                        _pd(dE) if want_dE else None, C.byref(sw), C.byref(nok)))
        return dE, sw.value, nok.value

    def set_chemical_potential(self, res, mu):
        self.L.orc_set_chemical_potential(self.h, res, float(mu))

    def set_fugacity(self, res, f):
        self.L.orc_set_fugacity(self.h, res, float(f))
