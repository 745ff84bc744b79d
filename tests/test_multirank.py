"""The N > 1 path on CPU: walkers sharded over 2 ranks (gloo), no data-path collective, one final
sum of per-point accumulators.  The per-walker engine here is the CPU oracle (tests may use it);
on the GPU box the same plan / reduction code drives the CUDA engine (bench.py under torchrun)."""
import os
import socket
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parents[1]


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _walker_accumulators(system, plan, n_steps):
    """Run this rank's walkers on the oracle; return [walkers, N_ACC] accumulators."""
    from maniac_b200.isotherm import N_ACC
    from oracle.oracle import Oracle
    fug = plan.fugacities()
    acc = np.zeros((len(fug), N_ACC))
    for i, g in enumerate(plan.global_ids()):
        s = system.copy()
        s.residues[0].fugacity = float(fug[i])
        o = Oracle(s, capacity=64)
        o.update_system_energy()
        o.seed(12345 + 104729 * g)                      # disjoint streams by GLOBAL walker id
        sn = sn2 = se = 0.0
        for _ in range(n_steps):
            o.monte_carlo_steps(1, trace=False)
            n = o.count(0)
            sn += n; sn2 += n * n; se += o.energy()[5]
        acc[i] = [sn, sn2, se, n_steps, 0.0, 0.0]
    return acc


def _rank_main(rank, world, port, out_dir):
    sys.path.insert(0, str(ROOT))
    import torch.distributed as dist
    import maniac_b200  # noqa: F401
    from maniac_b200.isotherm import IsothermPlan, reduce_sums
    from maniac_b200.snapshot import load_snapshot
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    s = load_snapshot(ROOT / "tests" / "golden" / "methanol.npz")
    s.p_translation, s.p_rotation, s.p_insertion_deletion = 0.4, 0.3, 0.3
    plan = IsothermPlan(n_points=4, walkers_per_rank=8, world_size=world, rank=rank, f_lo=1e-3, f_hi=1e-1)
    local = plan.accumulate(_walker_accumulators(s, plan, 40))
    total = reduce_sums(local)
    np.save(Path(out_dir) / f"total_{rank}.npy", total)
    np.save(Path(out_dir) / f"ids_{rank}.npy", np.array(list(plan.global_ids())))
    dist.barrier()
    dist.destroy_process_group()


def test_two_ranks_gloo_match_single_process(tmp_path):
    import torch.multiprocessing as mp
    sys.path.insert(0, str(ROOT))
    import maniac_b200  # noqa: F401
    from maniac_b200.isotherm import IsothermPlan, summarize
    from maniac_b200.snapshot import load_snapshot
    port = _free_port()
    mp.spawn(_rank_main, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    t0, t1 = np.load(tmp_path / "total_0.npy"), np.load(tmp_path / "total_1.npy")
    np.testing.assert_array_equal(t0, t1)                       # every rank holds the reduced sums
    ids = np.concatenate([np.load(tmp_path / "ids_0.npy"), np.load(tmp_path / "ids_1.npy")])
    assert sorted(ids.tolist()) == list(range(16))              # disjoint cover of the global walkers
    # the same 16 walkers in one process
    s = load_snapshot(ROOT / "tests" / "golden" / "methanol.npz")
    s.p_translation, s.p_rotation, s.p_insertion_deletion = 0.4, 0.3, 0.3
    ref = np.zeros_like(t0)
    for r in range(2):
        plan = IsothermPlan(n_points=4, walkers_per_rank=8, world_size=2, rank=r, f_lo=1e-3, f_hi=1e-1)
        ref += plan.accumulate(_walker_accumulators(s, plan, 40))
    np.testing.assert_allclose(t0, ref, rtol=1e-13, atol=0)
    summ = summarize(t0, beta=1.6773875557828624)
    assert summ["samples"].sum() == 16 * 40 and np.isfinite(summ["mean_N"]).all()


def test_plan_layout():
    sys.path.insert(0, str(ROOT))
    import maniac_b200  # noqa: F401
    from maniac_b200.isotherm import IsothermPlan
    for world in (1, 2, 4, 8):
        seen = []
        for r in range(world):
            p = IsothermPlan(n_points=64, walkers_per_rank=4736, world_size=world, rank=r)
            ids = p.global_ids()
            assert len(ids) == 4736
            seen += [ids[0], ids[-1]]
            pts = p.points()
            # the four warps of a CTA that share an SM sub-partition are replicas of one point
            blk = pts[:16].reshape(4, 4)
            assert (blk == blk[0]).all()
            assert len(set(pts.tolist())) == 64
        assert seen[0] == 0 and seen[-1] == 4736 * world - 1


def test_summarize_and_isotherm_table(tmp_path):
    sys.path.insert(0, str(ROOT))
    import maniac_b200  # noqa: F401
    from maniac_b200.isotherm import IsothermPlan, summarize
    from maniac_b200.outputs import write_isotherm
    plan = IsothermPlan(n_points=4, walkers_per_rank=32, world_size=1, rank=0)
    per_walker = np.zeros((32, 6))
    pts = plan.points()
    per_walker[:, 0] = 10.0 * (pts + 1) * 5          # sum N over 5 samples, N = 10 (point + 1)
    per_walker[:, 1] = (10.0 * (pts + 1)) ** 2 * 5
    per_walker[:, 2] = -3.0 * 5
    per_walker[:, 3] = 5
    per_walker[:, 4] = 0.25 * 7
    per_walker[:, 5] = 7
    sums = plan.accumulate(per_walker)
    s = summarize(sums, beta=2.0)
    np.testing.assert_allclose(s["mean_N"], [10, 20, 30, 40])
    np.testing.assert_allclose(s["var_N"], 0.0, atol=1e-9)
    np.testing.assert_allclose(s["mean_E"], -3.0)
    np.testing.assert_allclose(s["mu_ex"], -np.log(0.25) / 2.0)
    assert s["samples"].sum() == 32 * 5
    write_isotherm(tmp_path / "isotherm.dat", [1e-2, 1e-1, 1.0, 10.0], s)
    lines = (tmp_path / "isotherm.dat").read_text().splitlines()
    assert len(lines) == 5 and lines[0].startswith("#") and lines[2].split()[1] == "20.000000"
