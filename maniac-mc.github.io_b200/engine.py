"""Host-side mirror of the reference's energy interface on top of the C ABI.

Method names follow the Fortran routines they stand for (``update_system_energy``,
``pairwise_energy_for_molecule``, ``compute_old_energy`` / ``compute_new_energy``, ...),
argument meaning is the same with 0-based indices, and errors surface the way the
reference reports them: a non-zero status from the library becomes :class:`ManiacAbort`
carrying the library's message (the Fortran shim calls ``abort_run(msg, code)``,
``src/output_utils.f90:581-605``).

Everything here calls the CUDA library; nothing falls back to the CPU.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Sequence

import numpy as np

from . import capi
from .inputs import ERROR, System

E_NON_COULOMB, E_COULOMB, E_RECIP, E_SELF, E_INTRA, E_TOTAL = range(6)
KIND_MOVE, KIND_CREATE, KIND_DELETE, KIND_SWAP = 0, 1, 2, 3
OPT_HOST_CACHE, OPT_PHASE_SYNC, OPT_BLOCK_SLICES, OPT_SWEEP_TEAM = 1, 2, 3, 4
MV_NONE, MV_TRANSLATE, MV_ROTATE, MV_CREATE, MV_DELETE, MV_SWAP, MV_WIDOM = range(7)

TRACE_DTYPE = np.dtype([("move", "i4"), ("res", "i4"), ("mol", "i4"), ("accepted", "i4"),
                        ("dE", "f8"), ("prob", "f8"), ("e_old", "f8", 6), ("e_new", "f8", 6)])
assert TRACE_DTYPE.itemsize == C.sizeof(capi.MgpuStepTrace)


class ManiacAbort(RuntimeError):
    """The engine reported an error (``abort_run`` in the reference)."""


def lj_table(system: System):
    """Per-type-pair epsilon / sigma after the Lorentz-Berthelot fill.

    ``read_parameters`` + ``apply_lorentz_berthelot`` (src/parameters_parser.f90:19-178):
    explicit ``pair_coeff i j eps sigma`` lines are stored symmetrically in file order; a
    pair whose epsilon and sigma are both < 1e-10 afterwards receives
    sigma = (s_ii + s_jj)/2, eps = sqrt(e_ii e_jj) if both results exceed 1e-10.  The
    reference's 4-D arrays depend on the two atom types only, so the table is ntypes^2."""
    n = system.ntypes
    eps = np.zeros((n, n))
    sig = np.zeros((n, n))
    for ti, tj, e, s in system.pair_coeff:
        if not (0 <= ti < n and 0 <= tj < n):
            raise ManiacAbort("Failed to read pair_coeff value")
        eps[ti, tj] = e
        sig[ti, tj] = s
        eps[tj, ti] = e
        sig[tj, ti] = s
    for ri in system.residues:
        for ti in ri.types:
            for rj in system.residues:
                for tj in rj.types:
                    if abs(eps[ti, tj]) < ERROR and abs(sig[ti, tj]) < ERROR:
                        s = (sig[ti, ti] + sig[tj, tj]) / 2
                        e = float(np.sqrt(eps[ti, ti] * eps[tj, tj]))
                        if s > ERROR and e > ERROR:
                            sig[ti, tj] = s
                            eps[ti, tj] = e
    return eps, sig


def _pd(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def _pi(a):
    return a.ctypes.data_as(C.POINTER(C.c_int32))


class Engine:
    """One process, one GPU, ``n_walkers`` independent copies of the system."""

    def __init__(self, system: System, n_walkers: int = 1, capacity: Optional[int] = None, device: int = 0):
        self.L = capi.lib()
        self.system = system
        self.n_walkers = int(n_walkers)
        self.natom = [r.natom for r in system.residues]
        self.active = [bool(r.active) for r in system.residues]
        eps, sig = lj_table(system)
        self._keep = [eps, sig]
        res_arr = (capi.MgpuResidue * len(system.residues))()
        self.capacity = []
        for i, r in enumerate(system.residues):
            cap = r.nmol if not r.active else max(r.nmol + 1, capacity or (r.nmol + 64))
            cap = max(cap, 1)
            self.capacity.append(cap)
            ch = np.ascontiguousarray(r.charges, dtype=np.float64)
            ty = np.ascontiguousarray(r.types, dtype=np.int32)
            com = np.ascontiguousarray(r.com if r.nmol else np.zeros((1, 3)), dtype=np.float64)
            off = np.ascontiguousarray(r.offset if r.nmol else np.zeros((1, r.natom, 3)), dtype=np.float64)
            self._keep += [ch, ty, com, off]
            res_arr[i] = capi.MgpuResidue(r.natom, int(r.active), r.nmol, cap, _pd(ch), _pi(ty), _pd(com), _pd(off),
                                          float(r.mass), float(r.fugacity), float(r.chemical_potential))
        s = capi.MgpuSystem()
        s.matrix = (C.c_double * 9)(*np.asarray(system.matrix, dtype=np.float64).reshape(9))
        s.lo = (C.c_double * 3)(*np.asarray(system.lo, dtype=np.float64))
        s.nres = len(system.residues)
        s.residues = res_arr
        s.ntypes = system.ntypes
        s.epsilon = _pd(eps)
        s.sigma = _pd(sig)
        s.temperature = system.temperature
        s.ewald_tolerance = system.ewald_tolerance
        s.real_space_cutoff = system.real_space_cutoff
        s.translation_step = system.translation_step
        s.rotation_step_angle = system.rotation_step_angle
        s.p_translation, s.p_rotation, s.p_swap = system.p_translation, system.p_rotation, system.p_swap
        s.p_insertion_deletion, s.p_widom = system.p_insertion_deletion, system.p_widom
        s.n_walkers = self.n_walkers
        s.device = device
        self._sys_struct = s
        self._keep.append(res_arr)
        self._pinned = []
        self._ck(self.L.mgpu_init(C.byref(s)))
        self._open = True

    # -- plumbing -------------------------------------------------------------------------
    def _ck(self, rc):
        if rc:
            raise ManiacAbort(self.L.mgpu_last_error().decode())

    def close(self):
        if getattr(self, "_open", False):
            for p in self._pinned:
                self.L.mgpu_host_free(C.c_void_p(p))
            self._pinned = []
            self.L.mgpu_finalize()
            self._open = False

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    # -- derived parameters -----------------------------------------------------------------
    def ewald(self):
        alpha, rc = C.c_double(), C.c_double()
        kmax = (C.c_int32 * 3)()
        nk = C.c_int32()
        self._ck(self.L.mgpu_get_ewald(C.byref(alpha), kmax, C.byref(nk), C.byref(rc)))
        return dict(alpha=alpha.value, kmax=list(kmax), nk=nk.value, rc=rc.value)

    def kvectors(self):
        nk = self.ewald()["nk"]
        kx, ky, kz = (np.zeros(nk, dtype=np.int32) for _ in range(3))
        k2, ffw = np.zeros(nk), np.zeros(nk)
        self._ck(self.L.mgpu_get_kvectors(_pi(kx), _pi(ky), _pi(kz), _pd(k2), _pd(ffw)))
        return kx, ky, kz, k2, ffw

    def box(self):
        m, r = np.zeros((3, 3)), np.zeros((3, 3))
        v, t = C.c_double(), C.c_int32()
        self._ck(self.L.mgpu_get_box(_pd(m), _pd(r), C.byref(v), C.byref(t)))
        return dict(matrix=m, reciprocal=r, volume=v.value, triclinic=bool(t.value))

    def launch_info(self):
        w, s, n = C.c_int32(0), C.c_int64(0), C.c_int32(0)
        self._ck(self.L.mgpu_get_launch_info(C.byref(w), C.byref(s), C.byref(n)))
        return dict(walkers_per_cta=w.value, smem_bytes_per_cta=s.value, sm_count=n.value)

    def triclinic_candidates(self):
        n = C.c_int32(0)
        self._ck(self.L.mgpu_get_triclinic_candidates(C.byref(n)))
        return n.value

    def sweep_shape(self, n_walkers):
        """Launch shape mgpu_sweep / mgpu_block would pick for n_walkers walkers in flight."""
        t, p = C.c_int32(0), C.c_int32(0)
        self._ck(self.L.mgpu_get_sweep_shape(int(n_walkers), C.byref(t), C.byref(p)))
        return dict(threads_per_walker=t.value, walkers_per_cta=p.value,
                    text={32: "warp / walker", 64: "team (2 warps / walker)", 128: "team (4 warps / walker)"}.get(t.value, "?") + ", %d walkers / CTA" % p.value)

    def thermo(self, res):
        b, l, m = C.c_double(), C.c_double(), C.c_double()
        self._ck(self.L.mgpu_get_thermo(res, C.byref(b), C.byref(l), C.byref(m)))
        return dict(beta=b.value, lambda_=l.value, mu=m.value)

    def device_info(self):
        name = C.create_string_buffer(128)
        sm, mem = C.c_int(), C.c_double()
        self._ck(self.L.mgpu_device_info(name, 128, C.byref(sm), C.byref(mem)))
        return dict(name=name.value.decode(), sm_count=sm.value, mem_gb=mem.value)

    # -- state ---------------------------------------------------------------------------------
    def set_molecule(self, res, mol, com, offset, walker=0):
        com = np.ascontiguousarray(com, dtype=np.float64)
        off = np.ascontiguousarray(offset, dtype=np.float64)
        self._ck(self.L.mgpu_set_molecule(walker, res, mol, _pd(com), _pd(off)))

    def get_molecule(self, res, mol, walker=0):
        com, off = np.zeros(3), np.zeros((self.natom[res], 3))
        self._ck(self.L.mgpu_get_molecule(walker, res, mol, _pd(com), _pd(off)))
        return com, off

    def set_count(self, res, n, walker=0):
        self._ck(self.L.mgpu_set_count(walker, res, n))

    def count(self, res, walker=0):
        n = C.c_int32()
        self._ck(self.L.mgpu_get_count(walker, res, C.byref(n)))
        return n.value

    def set_chemical_potential(self, res, mu, walker=0):
        self._ck(self.L.mgpu_set_chemical_potential(walker, res, float(mu)))

    def set_fugacity(self, res, f, walker=0):
        self._ck(self.L.mgpu_set_fugacity(walker, res, float(f)))

    def set_fugacities(self, res, f, first_walker=0):
        """thermo%fugacity(res) of a run of walkers in one copy (mu = ln f / beta, prepare_utils.f90:245-247)."""
        mu = np.ascontiguousarray(np.log(np.asarray(f, dtype=np.float64)) / self.thermo(res)["beta"])
        self._ck(self.L.mgpu_set_chemical_potentials(first_walker, mu.size, res, _pd(mu)))

    def counts(self, res, first_walker=0, n_walkers=None):
        n = self.n_walkers - first_walker if n_walkers is None else n_walkers
        out = np.zeros(n, dtype=np.int32)
        self._ck(self.L.mgpu_get_counts(first_walker, n, res, _pi(out)))
        return out

    def Ak(self, walker=0):
        nk = self.ewald()["nk"]
        a = np.zeros(2 * nk)
        self._ck(self.L.mgpu_get_Ak(walker, _pd(a)))
        return a[0::2] + 1j * a[1::2]

    def energy(self, walker=0):
        out = np.zeros(6)
        self._ck(self.L.mgpu_get_energy(walker, _pd(out)))
        return out

    # -- the reference's energy routines ---------------------------------------------------------
    def set_option(self, option, value):
        """mgpu_set_option; OPT_HOST_CACHE = 1 (per-molecule framework-energy cache on / off)."""
        self._ck(self.L.mgpu_set_option(int(option), int(value)))

    def update_system_energy(self, walker=0):
        out = np.zeros(6)
        self._ck(self.L.mgpu_total_energy(walker, _pd(out)))
        return out

    def pairwise_energy_for_molecule(self, res, mol, skip_ordering_check=True, walker=0, com=None, offset=None):
        a, b = C.c_double(), C.c_double()
        pc = po = None
        if com is not None:
            com = np.ascontiguousarray(com, dtype=np.float64)
            offset = np.ascontiguousarray(offset, dtype=np.float64)
            pc, po = _pd(com), _pd(offset)
        self._ck(self.L.mgpu_pairwise_energy_for_molecule(walker, res, mol, int(skip_ordering_check), pc, po,
                                                          C.byref(a), C.byref(b)))
        return a.value, b.value

    def ewald_self_energy_single_mol(self, res):
        e = C.c_double()
        self._ck(self.L.mgpu_ewald_self_energy_single_mol(res, C.byref(e)))
        return e.value

    def intra_res_real_coulomb_energy(self, res, mol, walker=0, com=None, offset=None):
        e = C.c_double()
        pc = po = None
        if com is not None:
            com = np.ascontiguousarray(com, dtype=np.float64)
            offset = np.ascontiguousarray(offset, dtype=np.float64)
            pc, po = _pd(com), _pd(offset)
        self._ck(self.L.mgpu_intra_res_real_coulomb_energy(walker, res, mol, pc, po, C.byref(e)))
        return e.value

    def reciprocal_ewald_energy(self, walker=0):
        e = C.c_double()
        self._ck(self.L.mgpu_reciprocal_ewald_energy(walker, C.byref(e)))
        return e.value

    def compute_old_energy(self, res, mol, kind=KIND_MOVE, walker=0):
        out = np.zeros(6)
        self._ck(self.L.mgpu_old_energy(walker, res, mol, kind, _pd(out)))
        return out

    def compute_new_energy(self, res, mol, kind=KIND_MOVE, com=None, offset=None, walker=0):
        out = np.zeros(6)
        pc = po = None
        if com is not None:
            com = np.ascontiguousarray(com, dtype=np.float64)
            offset = np.ascontiguousarray(offset, dtype=np.float64)
            pc, po = _pd(com), _pd(offset)
        self._ck(self.L.mgpu_new_energy(walker, res, mol, kind, pc, po, _pd(out)))
        return out

    def swap_energy(self, res_old, mol_old, res_new, com, offset, walker=0):
        """attempt_swap_move's two energy calls (src/swapping.f90:62,88): deletion-style ``old`` of
        (res_old, mol_old), creation-style ``new`` of a res_new molecule with geometry (com, offset)."""
        com = np.ascontiguousarray(com, dtype=np.float64)
        offset = np.ascontiguousarray(offset, dtype=np.float64)
        e_old, e_new = np.zeros(6), np.zeros(6)
        self._ck(self.L.mgpu_swap_energy(walker, res_old, mol_old, res_new, _pd(com), _pd(offset),
                                         _pd(e_old), _pd(e_new)))
        return e_old, e_new

    def commit(self, walker=0):
        self._ck(self.L.mgpu_commit(walker))

    def rollback(self, walker=0):
        self._ck(self.L.mgpu_rollback(walker))

    # -- batched host-driven trials -----------------------------------------------------------------
    def trial_batch(self, walkers, res, mol, kind, com, offset):
        """``offset`` is (n, MAX_SITES, 3) or (n, natom, 3) (padded here)."""
        walkers = np.ascontiguousarray(walkers, dtype=np.int32)
        n = walkers.size
        res = np.ascontiguousarray(np.broadcast_to(res, (n,)), dtype=np.int32)
        mol = np.ascontiguousarray(mol, dtype=np.int32)
        kind = np.ascontiguousarray(np.broadcast_to(kind, (n,)), dtype=np.int32)
        com = np.ascontiguousarray(com, dtype=np.float64).reshape(n, 3)
        offset = np.asarray(offset, dtype=np.float64)
        if offset.shape[1] != capi.MAX_SITES:
            pad = np.zeros((n, capi.MAX_SITES, 3))
            pad[:, :offset.shape[1]] = offset
            offset = pad
        offset = np.ascontiguousarray(offset)
        e_old, e_new = np.zeros((n, 6)), np.zeros((n, 6))
        self._ck(self.L.mgpu_trial_batch(n, _pi(walkers), _pi(res), _pi(mol), _pi(kind), _pd(com), _pd(offset),
                                         _pd(e_old), _pd(e_new)))
        return e_old, e_new

    def commit_batch(self, walkers, accept):
        walkers = np.ascontiguousarray(walkers, dtype=np.int32)
        accept = np.ascontiguousarray(accept, dtype=np.int32)
        self._ck(self.L.mgpu_commit_batch(walkers.size, _pi(walkers), _pi(accept)))

    # -- device-resident Monte Carlo ------------------------------------------------------------------
    def seed(self, seed):
        self._ck(self.L.mgpu_seed(int(seed) & 0xFFFFFFFFFFFFFFFF))

    def rng_state(self, walker=0):
        st = (C.c_uint64 * 4)()
        self._ck(self.L.mgpu_get_rng_state(walker, st))
        return list(st)

    def sweep(self, n_steps, first_walker=0, n_walkers=None, trace_walker=None):
        n = self.n_walkers - first_walker if n_walkers is None else n_walkers
        tr = None
        ptr = None
        tw = -1
        if trace_walker is not None:
            tr = np.zeros(n_steps, dtype=TRACE_DTYPE)
            ptr = tr.ctypes.data_as(C.c_void_p)
            tw = trace_walker
        self._ck(self.L.mgpu_sweep(first_walker, n, n_steps, tw, ptr))
        return tr

    # ---- walker records / block-level entry (host state in, host state out) ------------------
    def host_buffer(self, n_doubles):
        """Page-locked host array of ``n_doubles`` float64 (mgpu_host_alloc); freed by close()."""
        p = self.L.mgpu_host_alloc(C.c_size_t(int(n_doubles) * 8))
        if not p:
            raise ManiacAbort("mgpu_host_alloc failed")
        self._pinned.append(p)
        return np.ctypeslib.as_array((C.c_double * int(n_doubles)).from_address(p))

    def record_doubles_max(self):
        return int(self.L.mgpu_record_doubles_max())

    def save_walkers(self, blob, first_walker=0, n_walkers=None):
        n = self.n_walkers - first_walker if n_walkers is None else n_walkers
        off = np.zeros(n + 1, dtype=np.int64)
        self._ck(self.L.mgpu_save_walkers(first_walker, n, blob.ctypes.data, blob.size, off.ctypes.data))
        return off

    def load_walkers(self, blob, offsets, first_walker=0):
        offsets = np.ascontiguousarray(offsets, dtype=np.int64)
        self._ck(self.L.mgpu_load_walkers(first_walker, len(offsets) - 1, blob.ctypes.data, offsets.ctypes.data))

    def block(self, n_steps, blob_in, offsets_in, blob_out, first_walker=0, n_walkers=None):
        """One block of the MC loop from host records to host records (mgpu_block)."""
        n = self.n_walkers - first_walker if n_walkers is None else n_walkers
        off_out = np.zeros(n + 1, dtype=np.int64)
        pin = pof = None
        if blob_in is not None:
            offsets_in = np.ascontiguousarray(offsets_in, dtype=np.int64)
            pin, pof = blob_in.ctypes.data, offsets_in.ctypes.data
        self._ck(self.L.mgpu_block(first_walker, n, n_steps, pin, pof, blob_out.ctypes.data, blob_out.size, off_out.ctypes.data))
        return off_out

    def traffic(self, reset=False):
        a, b = C.c_int64(0), C.c_int64(0)
        self._ck(self.L.mgpu_get_traffic(C.byref(a), C.byref(b), 1 if reset else 0))
        return dict(h2d_bytes=a.value, d2h_bytes=b.value)

    @staticmethod
    def parse_record(rec, nk, residues):
        """Decode one walker record (see include/maniac_gpu.h) into a dict; ``residues`` = [(active, natom)]."""
        out = dict(length=int(rec[0]), count=rec[1:9].astype(int), energy=rec[9:15].copy(),
                   rng=rec[15:19].view(np.uint64).copy(), counters=rec[19:31].astype(np.int64).reshape(6, 2),
                   averages=rec[32:64].reshape(8, 4).copy(), step_sizes=(float(rec[64]), float(rec[65])),
                   Ak=rec[72:72 + 2 * nk].copy(), molecules={})
        p = 72 + 2 * nk
        for r, (active, na) in enumerate(residues):
            if not active:
                continue
            ms, cnt = 3 + 3 * na + 2, int(rec[1 + r])
            stored = max(cnt, 1)                   # slot 1 is stored even when the type is empty (insertion template)
            m = rec[p:p + stored * ms].reshape(stored, ms)
            out["molecules"][r] = dict(com=m[:cnt, :3].copy(), offset=m[:cnt, 3:3 + 3 * na].reshape(cnt, na, 3).copy(), cache=m[:cnt, -2:].copy(),
                                       template_offset=m[0, 3:3 + 3 * na].reshape(na, 3).copy())
            p += stored * ms
        return out

    # ---- multi-GPU: the one exchange of the path -----------------------------------
    def nccl_init_from_torch(self):
        """Bootstrap the library's own NCCL communicator: rank 0 makes the id, torch.distributed carries it."""
        import torch.distributed as dist
        world, rank = dist.get_world_size(), dist.get_rank()
        ident = C.create_string_buffer(128)
        if rank == 0:
            self._ck_nccl(self.L.mgpu_nccl_unique_id(ident))
        box = [ident.raw]
        dist.broadcast_object_list(box, src=0)
        ident = C.create_string_buffer(box[0], 128)
        self._ck_nccl(self.L.mgpu_nccl_init(ident, world, rank))
        self.nccl_ready = True

    def _ck_nccl(self, rc):
        if rc:
            raise ManiacAbort(self.L.mgpu_nccl_last_error().decode())

    def reduce_averages(self, buf):
        """In-place sum over ranks of a float64 host array (mgpu_reduce_averages, one ncclAllReduce)."""
        assert buf.dtype == np.float64 and buf.flags.c_contiguous
        self._ck_nccl(self.L.mgpu_reduce_averages(_pd(buf), buf.size))

    def nccl_finalize(self):
        if getattr(self, "nccl_ready", False):
            self.L.mgpu_nccl_finalize()
            self.nccl_ready = False

    def all_averages(self, res):
        """[n_walkers, 4] block accumulators (sum N, sum N^2, sum E, samples) of residue ``res``."""
        out = np.zeros((self.n_walkers, 4))
        tmp = np.zeros(4)
        for w in range(self.n_walkers):
            self._ck(self.L.mgpu_get_averages(w, res, _pd(tmp)))
            out[w] = tmp
        return out

    def adjust_move_step_sizes(self, first_walker=0, n_walkers=None):
        n = self.n_walkers - first_walker if n_walkers is None else n_walkers
        self._ck(self.L.mgpu_adjust_move_step_sizes(first_walker, n))

    def step_sizes(self, walker=0):
        out = np.zeros(2)
        self._ck(self.L.mgpu_get_step_sizes(walker, _pd(out)))
        return float(out[0]), float(out[1])

    def set_step_sizes(self, translation_step, rotation_step_angle, walker=0):
        self._ck(self.L.mgpu_set_step_sizes(walker, translation_step, rotation_step_angle))

    def counters(self, walker=0):
        out = (C.c_int64 * 12)()
        self._ck(self.L.mgpu_get_counters(walker, out))
        return np.array(out[:]).reshape(6, 2)

    def widom(self, res, walker=0):
        w, n = C.c_double(), C.c_int64()
        self._ck(self.L.mgpu_get_widom(walker, res, C.byref(w), C.byref(n)))
        return w.value, n.value

    def averages(self, res, walker=0):
        out = np.zeros(4)
        self._ck(self.L.mgpu_get_averages(walker, res, _pd(out)))
        return out

    def reset_averages(self):
        self._ck(self.L.mgpu_reset_averages())

    def widom_batch(self, res, n, seed, first_id=0, walker=0, want_dE=False):
        dE = np.zeros(n) if want_dE else None
        sw, nok = C.c_double(), C.c_int64()
        self._ck(self.L.mgpu_widom_batch(walker, res, first_id, n, int(seed) & 0xFFFFFFFFFFFFFFFF,
                                         _pd(dE) if want_dE else None, C.byref(sw), C.byref(nok)))
        return dE, sw.value, nok.value

    # -- measurement -----------------------------------------------------------------------------------
    def timing_reset(self):
        self.L.mgpu_timing_reset()

    def timing(self, kernel):
        ms, n = C.c_double(), C.c_int64()
        self.L.mgpu_timing_get(kernel.encode(), C.byref(ms), C.byref(n))
        return ms.value, n.value

    def pair_counts(self):
        out = (C.c_int64 * 3)()
        self._ck(self.L.mgpu_get_pair_counts(out))
        scr = C.c_int64(0)
        if hasattr(self.L, "mgpu_get_screened_pairs"):
            self._ck(self.L.mgpu_get_screened_pairs(C.byref(scr)))
        return dict(pairs=out[0], lj=out[1], coulomb=out[2], screened=scr.value)

    def reset_pair_counts(self):
        self._ck(self.L.mgpu_reset_pair_counts())

    def selftest_math(self):
        """(max relative error of the fast 1/r^2, max error of the erfc(alpha r)/r table relative to
        1/r) measured on the device against the exact forms."""
        a, b = C.c_double(), C.c_double()
        self._ck(self.L.mgpu_selftest_math(C.byref(a), C.byref(b)))
        return a.value, b.value

    def measure_l2_peak(self):
        gb = C.c_double()
        self._ck(self.L.mgpu_measure_l2_peak(C.byref(gb)))
        return gb.value

    def measure_fp64_peak(self):
        tf, s = C.c_double(), C.c_double()
        self._ck(self.L.mgpu_measure_fp64_peak(C.byref(tf), C.byref(s)))
        return tf.value, s.value
