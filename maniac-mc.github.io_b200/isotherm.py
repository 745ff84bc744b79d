"""Isotherm sweep over many walkers and ranks (BASELINE.json configs[3]).

Units = walkers (isotherm point x replica).  They share nothing but read-only data, so they are
sharded over ranks with NO data-path collective; the only exchange is one sum, at the end, of the
per-point accumulators {sum N, sum N^2, sum E, samples, sum w, n_w}.  That sum goes through
``torch.distributed`` (gloo on CPU in the tests, nccl on the GPU box) or through the library's own
``mgpu_reduce_averages`` (NCCL, dlopen'ed) -- both are exercised.

Layout rule inside one rank: a CTA of the sweep kernel holds 16 walkers (one warp each) and the
four warps w, w+4, w+8, w+12 share an SM sub-partition and are phase-aligned by a barrier per MC
step; they are made replicas of the SAME point so their steps have statistically equal length.
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np

from .workloads import isotherm_fugacities, shard_walkers

N_ACC = 6        # sum N, sum N^2, sum E, samples, sum widom weight, widom samples


@dataclass
class IsothermPlan:
    n_points: int = 64
    walkers_per_rank: int = 4736
    world_size: int = 1
    rank: int = 0
    f_lo: float = 1e-2
    f_hi: float = 1e4

    @property
    def n_global(self) -> int:
        return self.walkers_per_rank * self.world_size

    def global_ids(self) -> range:
        """Global walker ids of this rank: contiguous block (weak scaling: fixed walkers per rank)."""
        return shard_walkers(self.n_global, self.world_size, self.rank)

    def point_of(self, global_id: int) -> int:
        """Isotherm point of a walker: quartets (same id mod 4 inside a block of 16) share a point."""
        return ((global_id // 16) * 4 + global_id % 4) % self.n_points

    def points(self) -> np.ndarray:
        return np.array([self.point_of(g) for g in self.global_ids()], dtype=np.int64)

    def fugacities(self) -> np.ndarray:
        return isotherm_fugacities(self.n_points, self.f_lo, self.f_hi)[self.points()]

    def accumulate(self, per_walker: np.ndarray) -> np.ndarray:
        """[walkers_of_this_rank, N_ACC] -> per-point sums [n_points, N_ACC] of this rank."""
        per_walker = np.asarray(per_walker, dtype=np.float64)
        out = np.zeros((self.n_points, N_ACC))
        np.add.at(out, self.points(), per_walker)
        return out


def reduce_sums(local: np.ndarray, engine=None) -> np.ndarray:
    """Sum the per-point accumulators over ranks.  ``engine`` given and the library's NCCL communicator
    initialised -> ``mgpu_reduce_averages``; else torch.distributed (any backend); single process -> identity."""
    import torch
    import torch.distributed as dist
    buf = np.ascontiguousarray(local, dtype=np.float64).copy()
    if engine is not None and getattr(engine, "nccl_ready", False):
        engine.reduce_averages(buf.reshape(-1))
        return buf
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dev = "cuda" if dist.get_backend() == "nccl" else "cpu"
        t = torch.from_numpy(buf).to(dev)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        buf = t.cpu().numpy()
    return buf


def summarize(sums: np.ndarray, beta: float) -> dict:
    """Per-point ensemble averages from the reduced accumulators: <N>, var N, <E>, mu_ex = -kT ln(<w>)
    (calculate_excess_mu, src/monte_carlo_utils.f90:598-636)."""
    n = np.maximum(sums[:, 3], 1.0)
    mean_n = sums[:, 0] / n
    out = dict(mean_N=mean_n, var_N=sums[:, 1] / n - mean_n ** 2, mean_E=sums[:, 2] / n, samples=sums[:, 3])
    with np.errstate(divide="ignore", invalid="ignore"):
        out["mu_ex"] = np.where(sums[:, 5] > 0, -np.log(sums[:, 4] / np.maximum(sums[:, 5], 1.0)) / beta, np.nan)
    return out
