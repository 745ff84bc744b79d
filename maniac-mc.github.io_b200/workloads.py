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
