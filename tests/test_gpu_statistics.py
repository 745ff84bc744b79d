"""North-star statistical gate: ensemble uptake <N> and Widom mu_ex of the GPU path agree with the CPU oracle
(the restatement of the reference's drivers) within 2 sigma -- INDEPENDENT random streams on the two sides, so this
is a test of the sampled distribution, not of arithmetic parity (that is tests/test_gpu_parity.py).

sigma comes from independent replicas / blocks on each side: sigma^2 = s_gpu^2 / n_gpu + s_cpu^2 / n_cpu.  Seeds are
fixed, so the outcome is reproducible; the margin actually observed is printed."""
import copy
import threading

import numpy as np
import pytest

from oracle.oracle import Oracle

pytestmark = pytest.mark.gpu


def _gcmc_system(load):
    """CO2 in ZIF-8 2x2x2 (BASELINE configs[2] framework), GCMC mix, a fugacity where the cell holds a few molecules."""
    s = copy.deepcopy(load("zif8_co2_widom"))
    s.residues[1].fugacity = 2.0e-8                    # <N> ~ 7-8 molecules, equilibrated after ~2500 MC steps from empty
    s.p_translation, s.p_rotation, s.p_insertion_deletion, s.p_widom = 0.3, 0.3, 0.4, 0.0
    return s


def test_uptake_within_two_sigma(load):
    from maniac_b200.engine import Engine
    s = _gcmc_system(load)
    cap, n_equil, n_prod = 64, 3000, 8000
    # ---- CPU oracle: independent replicas, one per thread
    n_rep = 12
    means = np.zeros(n_rep)

    def run(i):
        o = Oracle(s, capacity=cap)
        o.set_count(1, 0)
        o.update_system_energy()
        o.seed(900001 + 7919 * i)                       # streams disjoint from the GPU walkers' (seed + 104729 w)
        o.monte_carlo_steps(n_equil, trace=False)
        acc, blk = 0.0, 50
        for _ in range(n_prod // blk):
            o.monte_carlo_steps(blk, trace=False)
            acc += o.count(1)
        means[i] = acc / (n_prod // blk)
    ths = [threading.Thread(target=run, args=(i,)) for i in range(n_rep)]
    for t in ths:
        t.start()
    # ---- GPU: 256 walkers, per-walker block averages accumulated on the device every MC step
    nW = 256
    with Engine(s, n_walkers=nW, capacity=cap) as eng:
        for w in range(nW):
            eng.set_count(1, 0, walker=w)
        for w in range(nW):
            eng.update_system_energy(w)
        eng.seed(2024)
        eng.sweep(n_equil)
        eng.reset_averages()
        eng.sweep(n_prod)
        av = eng.all_averages(1)[:nW]                   # [sum N, sum N^2, sum E, samples] per walker
        g = av[:, 0] / av[:, 3]
    for t in ths:
        t.join()
    mean_g, mean_c = g.mean(), means.mean()
    sigma = np.sqrt(g.var(ddof=1) / nW + means.var(ddof=1) / n_rep)
    print(f"uptake: GPU {mean_g:.4f} (256 walkers), oracle {mean_c:.4f} ({n_rep} replicas), diff {mean_g - mean_c:+.4f} = {(mean_g - mean_c) / sigma:+.2f} sigma")
    assert mean_c > 0.5, "the chosen fugacity should hold a few molecules"
    assert abs(mean_g - mean_c) <= 2.0 * sigma


def test_widom_mu_ex_within_two_sigma(load):
    """Batched test-particle insertions (k_widom_batch, 2^20 insertions in 16 blocks) against the oracle's widom trials
    on OTHER insertion ids (10 blocks of 1500): mu_ex = -kT ln <exp(-beta dU)> per block, block sigma on both sides."""
    from maniac_b200.engine import Engine
    s = copy.deepcopy(load("zif8_co2_widom"))
    s.p_translation, s.p_rotation, s.p_insertion_deletion, s.p_widom = 0.0, 0.0, 0.0, 1.0
    nb_c, per_c = 10, 1500
    mu_c = np.zeros(nb_c)

    def run(b):
        o = Oracle(s, capacity=8)
        o.set_count(1, 0)
        o.update_system_energy()
        beta = o.beta
        _, sw, n_ok = o.widom_batch(1, 10_000_000 + b * per_c, per_c, 777, want_dE=False)
        mu_c[b] = -np.log(sw / per_c) / beta
    ths = [threading.Thread(target=run, args=(b,)) for b in range(nb_c)]
    for t in ths:
        t.start()
    nb_g, per_g = 16, 65536
    mu_g = np.zeros(nb_g)
    with Engine(s, n_walkers=1, capacity=8) as eng:
        eng.set_count(1, 0)
        eng.update_system_energy()
        beta = eng.thermo(1)["beta"]
        tot_w = 0.0
        for b in range(nb_g):
            _, sw, n_ok = eng.widom_batch(1, per_g, seed=777, first_id=b * per_g)
            mu_g[b] = -np.log(sw / per_g) / beta
            tot_w += sw
        mu_all = -np.log(tot_w / (nb_g * per_g)) / beta
    for t in ths:
        t.join()
    sigma = np.sqrt(mu_g.var(ddof=1) / nb_g + mu_c.var(ddof=1) / nb_c)
    d = mu_g.mean() - mu_c.mean()
    print(f"mu_ex: GPU {mu_g.mean():.4f} (all {mu_all:.4f}), oracle {mu_c.mean():.4f}, diff {d:+.4f} = {d / sigma:+.2f} sigma (sigma {sigma:.4f} kcal/mol)")
    assert abs(d) <= 2.0 * sigma
