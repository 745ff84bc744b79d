"""Synthetic configurations of the named systems: loaded pores built by
random insertion with a minimum-distance rejection, isotherm fugacity grids.  Host logic."""
from __future__ import annotations

import numpy as np

from .inputs import System


def _rotation(axis: int, theta: float) -> np.ndarray:
    c, s = np.cos(theta), np.sin(theta)
    M = np.eye(3)
    if axis == 0:
        M[1, 1], M[1, 2], M[2, 1], M[2, 2] = c, -s, s, c
    elif axis == 1:
        M[0, 0], M[0, 2], M[2, 0], M[2, 2] = c, s, -s, c
    else:
        M[0, 0], M[0, 1], M[1, 0], M[1, 1] = c, -s, s, c
    return M


def load_pore(system: System, res: int, n_target: int, seed: int = 12345, min_dist: float = 2.0,
              max_tries: int = 200000) -> System:
    """Return a copy of ``system`` whose guest residue ``res`` holds ``n_target`` molecules:
    uniform random centre + random axis rotation of the template geometry (molecule 1), rejected
    if any site is closer than ``min_dist`` to any atom already present (orthorhombic cells)."""
    s = system.copy()
    r = s.residues[res]
    rng = np.random.default_rng(seed)
    L = np.array([s.matrix[0, 0], s.matrix[1, 1], s.matrix[2, 2]])
    pts = []
    for rr in s.residues:
        if rr is r or rr.nmol == 0:
            continue
        pts.append((rr.com[:, None, :] + rr.offset).reshape(-1, 3))
    template = r.offset[0].copy()
    coms, offs = [], []
    pts = np.concatenate(pts) if pts else np.zeros((0, 3))
    tries = 0
    while len(coms) < n_target:
        tries += 1
        if tries > max_tries:
            raise RuntimeError(f"load_pore: could not place {n_target} molecules")
        com = s.lo + L * rng.random(3)
        off = template @ _rotation(int(rng.integers(3)), rng.random() * 2 * np.pi).T
        pos = com + off
        d = pts[None, :, :] - pos[:, None, :]
        d -= L * np.rint(d / L)
        if pts.shape[0] and (np.einsum("ijk,ijk->ij", d, d).min() < min_dist ** 2):
            continue
        coms.append(com)
        offs.append(off)
        pts = np.concatenate([pts, pos])
    r.com = np.array(coms).reshape(-1, 3)
    r.offset = np.array(offs).reshape(-1, r.natom, 3)
    return s


def isotherm_fugacities(n_points: int = 64, lo: float = 1e-2, hi: float = 1e4) -> np.ndarray:
    """Log-spaced fugacity grid of the isotherm sweep (BASELINE.json configs[3])."""
    return np.logspace(np.log10(lo), np.log10(hi), n_points)


def shard_walkers(n_global: int, world_size: int, rank: int):
    """Contiguous block partition of global walker ids over ranks; walkers are
    independent, so there is no data-path collective."""
    base, rem = divmod(n_global, world_size)
    start = rank * base + min(rank, rem)
    return range(start, start + base + (1 if rank < rem else 0))


def mixture_supercell(base: System, reps=(2, 2, 2), tilt_xy: float = 3.0, n_co2: int = 24, n_n2: int = 24,
                      seed: int = 777, min_dist: float = 2.2) -> System:
    """BASELINE.json configs[4]: binary CO2 / N2 mixture with identity swaps in a
    large TRICLINIC supercell.  ``base`` is the CO2-in-ZIF-8 Widom system (framework of
    testcase-widom/ZIF8-MET + CO2 of molecule-reservoir/CO2).  The 2208-atom framework block is
    replicated ``reps`` times and sheared by ``tilt_xy`` so that the cell matrix gets one
    off-diagonal element, which sends every distance through the reference's triclinic branch
    (geometry_utils.f90:263-280; the cell vectors are the COLUMNS of the matrix there, so the shear
    follows column 1).  N2 is a synthetic 3-site TraPPE-like molecule (N, massless centre charge,
    N); cross terms come from the Lorentz-Berthelot fill.  Guests are placed at random with a
    minimum-distance rejection, like ``load_pore``."""
    s = base.copy()
    zif = [r for r in s.residues if not r.active][0]
    co2 = [r for r in s.residues if r.active][0]
    a0 = np.array([s.matrix[0, 0], s.matrix[1, 1], s.matrix[2, 2]])
    nx, ny, nz = reps
    L = a0 * np.array(reps)
    # framework block replicated, then sheared: y += (t / Lx) * x  (lattice column 1 = (Lx, t, 0))
    pos0 = (zif.com[0] + zif.offset[0])
    blocks = [pos0 + a0 * np.array([i, j, k]) for i in range(nx) for j in range(ny) for k in range(nz)]
    pos = np.concatenate(blocks)
    pos[:, 1] += (tilt_xy / L[0]) * (pos[:, 0] - s.lo[0])
    ntile = nx * ny * nz
    zif.natom = int(pos.shape[0])
    zif.types = np.tile(zif.types, ntile).astype(np.int32)
    zif.charges = np.tile(zif.charges, ntile)
    zif.mass = zif.mass * ntile
    com = pos.mean(axis=0)
    zif.com = com[None, :]
    zif.offset = (pos - com)[None, :, :]
    s.matrix = np.array([[L[0], 0.0, 0.0], [tilt_xy, L[1], 0.0], [0.0, 0.0, L[2]]])
    C = s.matrix                                     # columns = cell vectors (reference quirk)
    Cinv = np.linalg.inv(C)
    # N2: two new atom types (N_n2, COM_n2)
    nt = s.ntypes
    n2 = type(co2)(name="n2", active=True, natom=3, site_types=[nt + 1, nt + 2], site_names=["Nn", "Cm"],
                   fugacity=-1.0, chemical_potential=0.0)
    n2.types = np.array([nt, nt + 1, nt], dtype=np.int32)
    n2.charges = np.array([-0.482, 0.964, -0.482])
    n2.mass = 3 * 14.0067                            # natom x mass(last atom's type), the reference's COM-frame quirk
    s.pair_coeff = list(s.pair_coeff) + [(nt, nt, 0.0715393, 3.31), (nt + 1, nt + 1, 0.0, 0.0)]
    s.ntypes = nt + 2
    g_n2 = np.array([[0.55, 0.0, 0.0], [0.0, 0.0, 0.0], [-0.55, 0.0, 0.0]])
    g_co2 = co2.offset[0].copy()
    rng = np.random.default_rng(seed)
    pts = pos.copy()

    def place(template, n):
        nonlocal pts
        coms, offs = [], []
        tries = 0
        while len(coms) < n:
            tries += 1
            if tries > 400000:
                raise RuntimeError("mixture_supercell: could not place the guests")
            c = s.lo + C @ rng.random(3)
            off = template @ _rotation(int(rng.integers(3)), rng.random() * 2 * np.pi).T
            p = c + off
            d = pts[None, :, :] - p[:, None, :]
            f = d @ Cinv.T
            d = d - np.rint(f) @ C.T
            if np.einsum("ijk,ijk->ij", d, d).min() < min_dist ** 2:
                continue
            coms.append(c)
            offs.append(off)
            pts = np.concatenate([pts, p])
        return np.array(coms).reshape(-1, 3), np.array(offs).reshape(-1, 3, 3)

    co2.com, co2.offset = place(g_co2, max(1, n_co2))
    n2.com, n2.offset = place(g_n2, max(1, n_n2))
    co2.fugacity, n2.fugacity = 5.0, 20.0
    s.residues = [zif, co2, n2]
    s.p_translation, s.p_rotation, s.p_swap, s.p_insertion_deletion, s.p_widom = 0.3, 0.3, 0.2, 0.2, 0.0
    s.translation_step, s.rotation_step_angle = 1.0, 0.685
    return s
