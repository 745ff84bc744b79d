"""Parity of the CUDA path (through the C ABI) against the CPU oracle.  GPU only.

Tolerances are the north-star's: total / component energies within 1e-10 relative
(|d| <= 1e-10 * max(1, |ref|)), per-move dE within 1e-9 (|d| <= 1e-9 * max(1, |dE_ref|)),
identical accept / reject sequence over the first 10^4 moves on the shared RNG contract.
"""
import numpy as np
import pytest

from oracle.oracle import KIND_CREATE, KIND_DELETE, KIND_MOVE, Oracle

pytestmark = pytest.mark.gpu

REL_E = 1e-10
REL_DE = 1e-9


def close_e(a, b, rel=REL_E):
    a, b = np.asarray(a, dtype=float), np.asarray(b, dtype=float)
    return np.all(np.abs(a - b) <= rel * np.maximum(1.0, np.abs(b)))


def assert_e(a, b, rel=REL_E, what=""):
    assert close_e(a, b, rel), f"{what}: gpu={np.asarray(a)!r} ref={np.asarray(b)!r} diff={np.asarray(a) - np.asarray(b)!r}"


def _engine_variant(param):
    from maniac_b200.engine import OPT_HOST_CACHE, OPT_SWEEP_TEAM, Engine

    class EngineVariant(Engine):
        def __init__(self, *a, **k):
            super().__init__(*a, **k)
            # the device-resident drivers in their throughput shape (one warp per walker) unless the variant asks for
            # the team shape; left to itself the engine would pick teams for the handful of walkers a test uses
            self.set_option(OPT_SWEEP_TEAM, {"team": 1, "team2": 2}.get(param, 0))
            if param == "nocache":
                self.set_option(OPT_HOST_CACHE, 0)
    EngineVariant.__name__ = f"Engine_{param}"
    return EngineVariant


@pytest.fixture(params=["hcache", "nocache"])
def engine_cls(request):
    """Every parity test runs twice: with the per-molecule framework-energy cache (default: the old
    geometry of a move / deletion reads the molecule's cached framework sum) and without it (the
    framework is swept for old and new geometry, like the reference does)."""
    return _engine_variant(request.param)


@pytest.fixture(params=["hcache", "nocache", "team", "team2"])
def sweep_engine_cls(request):
    """Tests of the device-resident drivers run again in the two team shapes of the sweep kernel (four / two warps per
    walker, MGPU_OPT_SWEEP_TEAM = 1 / 2): the shapes launches with few walkers per GPU use."""
    return _engine_variant(request.param)


ALL = ["lj_gas", "zif8_h2o", "h2o_gas", "methanol", "two_atoms", "dipole", "two_dipole", "dipole_triclinic",
       "zif8_co2_widom"]


@pytest.mark.parametrize("name", ALL)
def test_total_energy_and_Sk(name, load, engine_cls):
    s = load(name)
    o = Oracle(s)
    e_ref = o.update_system_energy()
    with engine_cls(s, n_walkers=2) as eng:
        ew_o, ew_g = o.ewald(), eng.ewald()
        assert ew_o["nk"] == ew_g["nk"] and ew_o["kmax"] == ew_g["kmax"]
        assert ew_o["alpha"] == ew_g["alpha"] and ew_o["rc"] == ew_g["rc"]      # same setup arithmetic, bit for bit
        for w in (0, 1):
            assert_e(eng.update_system_energy(w), e_ref, what=f"{name} total energy walker {w}")
        ak_o, ak_g = o.Ak(), eng.Ak(0)
        scale = max(1.0, np.abs(ak_o).max())
        assert np.abs(ak_o - ak_g).max() <= 1e-11 * scale
        bo, bg = o.box(), eng.box()
        np.testing.assert_array_equal(bo["reciprocal"], bg["reciprocal"])
        assert bo["volume"] == bg["volume"]


def test_known_answers_on_gpu(load, kat, engine_cls):
    """The reference's KATs, checked directly on the GPU path with the reference's tolerances."""
    for name in ["lj_gas", "zif8_h2o", "h2o_gas", "methanol", "two_atoms", "dipole", "two_dipole"]:
        ref = kat[name]["reference"]
        with engine_cls(load(name)) as eng:
            e = eng.update_system_energy()
        assert abs(e[5] - ref["total"]) < ref["tol"], (name, e, ref)


@pytest.mark.parametrize("name", ["zif8_h2o_gcmc", "h2o_gas", "dipole_triclinic"])
def test_pairwise_energy_for_molecule(name, load, engine_cls):
    s = load(name)
    o = Oracle(s)
    o.update_system_energy()
    with engine_cls(s) as eng:
        res = [i for i, r in enumerate(s.residues) if r.active][0]
        n = s.residues[res].nmol
        for mol in sorted(set([0, n // 2, n - 1])):
            for skip in (True, False):
                ref = o.pairwise_energy_for_molecule(res, mol, skip)
                got = eng.pairwise_energy_for_molecule(res, mol, skip)
                assert_e(got, ref, what=f"{name} pair energy mol {mol} skip {skip}")
        assert eng.ewald_self_energy_single_mol(res) == pytest.approx(o.self_energy_single_mol(res), rel=1e-14)
        assert_e(eng.intra_res_real_coulomb_energy(res, 0), o.intra_energy(res, 0), what="intra")


def _rot(axis, theta, off):
    c, s_ = np.cos(theta), np.sin(theta)
    M = np.eye(3)
    if axis == 0:
        M[1, 1], M[1, 2], M[2, 1], M[2, 2] = c, -s_, s_, c
    elif axis == 1:
        M[0, 0], M[0, 2], M[2, 0], M[2, 2] = c, s_, -s_, c
    else:
        M[0, 0], M[0, 1], M[1, 0], M[1, 1] = c, -s_, s_, c
    return off @ M.T


@pytest.mark.parametrize("name", ["zif8_h2o_gcmc", "methanol", "dipole_triclinic", "lj_gas"])
def test_old_new_energy_all_kinds(name, load, engine_cls):
    """compute_old_energy / compute_new_energy for move, creation and deletion, then
    commit / rollback, against the oracle doing the same through the reference's sequence."""
    s = load(name)
    rng = np.random.default_rng(5)
    res = [i for i, r in enumerate(s.residues) if r.active][0]
    na = s.residues[res].natom
    o = Oracle(s, capacity=32)
    o.update_system_energy()
    with engine_cls(s, capacity=32) as eng:
        eng.update_system_energy()
        for it in range(6):
            n = o.count(res)
            mol = int(rng.integers(n))
            com, off = o.get_molecule(res, mol)
            # ---- translation / rotation
            new_com = com + rng.uniform(-0.5, 0.5, 3)
            new_off = _rot(int(rng.integers(3)), rng.uniform(-0.3, 0.3), off) if na > 1 else off
            o.save_fourier(res, mol)
            old_r = o.compute_old_energy(res, mol, KIND_MOVE)
            o.set_molecule(res, mol, new_com, new_off)
            new_r = o.compute_new_energy(res, mol, KIND_MOVE)
            old_g = eng.compute_old_energy(res, mol, KIND_MOVE)
            new_g = eng.compute_new_energy(res, mol, KIND_MOVE, new_com, new_off)
            assert_e(old_g, old_r, what="old(move)")
            assert_e(new_g, new_r, what="new(move)")
            d_r, d_g = new_r[5] - old_r[5], new_g[5] - old_g[5]
            assert abs(d_r - d_g) <= REL_DE * max(1.0, abs(d_r))
            if it % 2 == 0:      # accept
                eng.commit()
                # oracle: accept_molecule_move bookkeeping == recompute
                o.update_system_energy()
            else:                # reject
                eng.rollback()
                o.set_molecule(res, mol, com, off)
                o.restore_fourier(res, mol)
            assert_e(eng.energy(), o.energy(), rel=1e-9, what="running energy after move")
            ak_o, ak_g = o.Ak(), eng.Ak()
            assert np.abs(ak_o - ak_g).max() <= 1e-10 * max(1.0, np.abs(ak_o).max())
        # ---- creation (through the oracle's own driver, forced acceptance)
        bx = o.box()
        n = o.count(res)
        o.set_chemical_potential(res, 50.0)
        o.seed(99)
        t = o.attempt_creation_move(res, n)
        com_new, off_new = o.get_molecule(res, n)
        old_g = eng.compute_old_energy(res, n, KIND_CREATE)
        new_g = eng.compute_new_energy(res, n, KIND_CREATE, com_new, off_new)
        assert_e(old_g, np.array(t.e_old[:]), what="old(create)")
        # real-space, self and intramolecular components at the component tolerance (1e-10); the reciprocal term (and the
        # total that contains it) is a full sum over k of |S + dS|^2 whose value, not its change, sets the rounding: 1e-9
        assert_e(new_g[[0, 1, 3, 4]], np.array(t.e_new[:])[[0, 1, 3, 4]], what="new(create) components")
        assert_e(new_g, np.array(t.e_new[:]), rel=1e-9, what="new(create)")
        assert abs((new_g[5] - old_g[5]) - t.dE) <= REL_DE * max(1.0, abs(t.dE))
        if t.accepted:
            eng.commit()
            assert eng.count(res) == n + 1 == o.count(res)
        else:
            eng.rollback()
        assert_e(eng.energy(), o.energy(), rel=1e-9, what="running energy after creation")
        # ---- deletion
        n = o.count(res)
        mol = 0
        o.set_chemical_potential(res, -50.0)
        t = o.attempt_deletion_move(res, mol)
        old_g = eng.compute_old_energy(res, mol, KIND_DELETE)
        new_g = eng.compute_new_energy(res, mol, KIND_DELETE)
        assert_e(old_g[[0, 1, 3, 4]], np.array(t.e_old[:])[[0, 1, 3, 4]], what="old(delete) components")
        assert_e(new_g[[0, 1, 3, 4]], np.array(t.e_new[:])[[0, 1, 3, 4]], what="new(delete) components")
        assert_e(old_g, np.array(t.e_old[:]), rel=1e-9, what="old(delete)")
        assert_e(new_g, np.array(t.e_new[:]), rel=1e-9, what="new(delete)")
        assert abs((new_g[5] - old_g[5]) - t.dE) <= REL_DE * max(1.0, abs(t.dE))
        assert t.accepted == 1
        eng.commit()
        assert eng.count(res) == n - 1 == o.count(res)
        for m in range(o.count(res)):           # swap-with-last compaction, remove_molecule
            co, oo = o.get_molecule(res, m)
            cg, og = eng.get_molecule(res, m)
            np.testing.assert_array_equal(co, cg)
            np.testing.assert_array_equal(oo, og)
        full = eng.update_system_energy()
        assert_e(full, o.update_system_energy(), what="full energy after create/delete")


def test_trial_batch_many_walkers(load, engine_cls):
    s = load("zif8_h2o_gcmc")
    W = 12
    rng = np.random.default_rng(1)
    oracles = []
    for w in range(W):
        o = Oracle(s, capacity=16)
        o.update_system_energy()
        oracles.append(o)
    with engine_cls(s, n_walkers=W, capacity=16) as eng:
        mols = rng.integers(0, 3, W)
        coms, offs, refs = [], [], []
        for w in range(W):
            com, off = oracles[w].get_molecule(0, int(mols[w]))
            nc = com + rng.uniform(-0.5, 0.5, 3)
            coms.append(nc)
            offs.append(off)
            oracles[w].save_fourier(0, int(mols[w]))
            old = oracles[w].compute_old_energy(0, int(mols[w]), KIND_MOVE)
            oracles[w].set_molecule(0, int(mols[w]), nc, off)
            new = oracles[w].compute_new_energy(0, int(mols[w]), KIND_MOVE)
            refs.append((old, new))
        e_old, e_new = eng.trial_batch(np.arange(W), 0, mols, KIND_MOVE, np.array(coms), np.array(offs))
        for w in range(W):
            assert_e(e_old[w], refs[w][0], what=f"walker {w} old")
            assert_e(e_new[w], refs[w][1], what=f"walker {w} new")
        accept = (np.arange(W) % 2).astype(np.int32)
        eng.commit_batch(np.arange(W), accept)
        for w in range(W):
            if accept[w]:
                oracles[w].update_system_energy()
                assert_e(eng.energy(w), oracles[w].energy(), rel=1e-9, what=f"walker {w} committed")
                cg, _ = eng.get_molecule(0, int(mols[w]), walker=w)
                np.testing.assert_array_equal(cg, coms[w])
            else:
                cg, _ = eng.get_molecule(0, int(mols[w]), walker=w)
                assert not np.array_equal(cg, coms[w])
        with pytest.raises(Exception):
            eng.commit_batch([0], [1])          # nothing pending any more


def _compare_traces(tr, ref, n_check=None):
    n = len(ref) if n_check is None else n_check
    assert (tr["move"][:n] == ref["move"][:n]).all()
    assert (tr["res"][:n] == ref["res"][:n]).all()
    assert (tr["mol"][:n] == ref["mol"][:n]).all()
    assert (tr["accepted"][:n] == ref["accepted"][:n]).all(), np.nonzero(tr["accepted"][:n] != ref["accepted"][:n])
    d = np.abs(tr["dE"][:n] - ref["dE"][:n])
    lim = REL_DE * np.maximum(1.0, np.abs(ref["dE"][:n]))
    assert (d <= lim).all(), (d.max(), np.argmax(d - lim))


def test_sweep_10k_moves_accept_reject_sequence(load, sweep_engine_cls):
    engine_cls = sweep_engine_cls
    """North-star gate: identical accept/reject sequence over the first 10^4 moves and
    per-move dE within 1e-9, GCMC mix of BASELINE configs[1] (0.4/0.4/0.2)."""
    s = load("zif8_h2o_gcmc")
    n = 10_000
    o = Oracle(s, capacity=256)
    o.update_system_energy()
    o.seed(12345)
    ref = o.monte_carlo_steps(n)
    with engine_cls(s, n_walkers=3, capacity=256) as eng:
        eng.seed(12345)
        tr = eng.sweep(n, trace_walker=0)
        _compare_traces(tr, ref)
        assert eng.count(0, walker=0) == o.count(0)
        assert_e(eng.energy(0), o.energy(), rel=1e-9, what="running energy after 10^4 moves")
        np.testing.assert_array_equal(eng.counters(0), o.counters())
        assert eng.rng_state(0) == o.rng_state()            # same number of draws consumed
        # drift audit (output_utils.f90:258-275 analogue): incremental == full recompute
        inc = eng.energy(0)
        full = eng.update_system_energy(0)
        assert_e(inc, full, rel=1e-9, what="incremental vs full after sweep")
        # the other walkers ran their own streams
        assert eng.counters(1).sum() > 0 and not np.array_equal(eng.counters(1), eng.counters(0))


@pytest.mark.parametrize("name,steps", [("methanol", 3000), ("dipole_triclinic", 400), ("lj_gas", 2000)])
def test_sweep_other_systems(name, steps, load, sweep_engine_cls):
    engine_cls = sweep_engine_cls
    s = load(name)
    if name == "lj_gas":
        s.p_translation, s.p_rotation, s.p_insertion_deletion = 0.5, 0.0, 0.5
        for r in s.residues:
            if r.active:
                r.fugacity, r.chemical_potential = 50.0, 0.0
    cap = 512 if name == "lj_gas" else 128
    o = Oracle(s, capacity=cap)
    o.update_system_energy()
    o.seed(777)
    ref = o.monte_carlo_steps(steps)
    with engine_cls(s, capacity=cap) as eng:
        eng.seed(777)
        tr = eng.sweep(steps, trace_walker=0)
        _compare_traces(tr, ref)
        res = [i for i, r in enumerate(s.residues) if r.active][0]
        assert eng.count(res) == o.count(res)
        assert_e(eng.energy(), o.energy(), rel=1e-9)


def test_sweep_from_empty_and_widom_moves(load, sweep_engine_cls):
    engine_cls = sweep_engine_cls
    """Edge cases: N = 0 (moves return early, creations use the slot-1 template) and Widom trials
    inside the loop (state never changes; statistic%weight / sample)."""
    s = load("zif8_co2_widom")
    s.residues[1].fugacity = 5.0e4
    s.p_translation, s.p_rotation, s.p_insertion_deletion, s.p_widom = 0.3, 0.3, 0.4, 0.0
    o = Oracle(s, capacity=64)
    o.set_count(1, 0)
    o.update_system_energy()
    o.seed(31)
    ref = o.monte_carlo_steps(1500)
    with engine_cls(s, capacity=64) as eng:
        eng.set_count(1, 0)
        eng.update_system_energy()
        eng.seed(31)
        tr = eng.sweep(1500, trace_walker=0)
        _compare_traces(tr, ref)
        assert eng.count(1) == o.count(1)
    s.p_translation, s.p_rotation, s.p_insertion_deletion, s.p_widom = 0.25, 0.25, 0.0, 0.5
    o = Oracle(s, capacity=64)
    o.update_system_energy()
    o.seed(32)
    ref = o.monte_carlo_steps(800)
    with engine_cls(s, capacity=64) as eng:
        eng.seed(32)
        tr = eng.sweep(800, trace_walker=0)
        _compare_traces(tr, ref)
        w_g, n_g = eng.widom(1)
        w_o, n_o = o.widom(1)
        assert n_g == n_o and w_g == pytest.approx(w_o, rel=1e-8)


@pytest.mark.parametrize("loaded", [0, 6])
def test_widom_batch(loaded, load, engine_cls):
    """K3: per-insertion dE vs the oracle on the same counter-based draws, empty and loaded pore."""
    s = load("zif8_co2_widom")
    s.residues[1].fugacity = 1.0e5
    s.p_translation, s.p_rotation, s.p_insertion_deletion, s.p_widom = 0.0, 0.0, 1.0, 0.0
    o = Oracle(s, capacity=64)
    o.set_count(1, 0)
    o.update_system_energy()
    o.seed(5)
    with engine_cls(s, capacity=64) as eng:
        eng.set_count(1, 0)
        eng.update_system_energy()
        if loaded:
            # grow a loaded pore with creations only (device and oracle follow the same stream)
            eng.seed(5)
            steps = 0
            while o.count(1) < loaded and steps < 4000:
                o.monte_carlo_steps(50, trace=False)
                eng.sweep(50)
                steps += 50
            assert eng.count(1) == o.count(1) >= 1
        n = 300
        dE_o, sw_o, ok_o = o.widom_batch(1, 1000, n, seed=424242)
        dE_g, sw_g, ok_g = eng.widom_batch(1, n, seed=424242, first_id=1000, want_dE=True)
        lim = REL_DE * np.maximum(1.0, np.abs(dE_o))
        assert (np.abs(dE_g - dE_o) <= lim).all(), np.abs(dE_g - dE_o).max()
        assert ok_g == ok_o and sw_g == pytest.approx(sw_o, rel=1e-9)
        # state untouched
        assert eng.count(1) == o.count(1)
        assert_e(eng.update_system_energy(), o.update_system_energy())
        # size-independent property: a batch equals the sum of its halves
        _, a, na_ = eng.widom_batch(1, 5000, seed=9, first_id=0)
        _, b, nb_ = eng.widom_batch(1, 5000, seed=9, first_id=5000)
        _, c, nc_ = eng.widom_batch(1, 10000, seed=9, first_id=0)
        assert na_ + nb_ == nc_ and (a + b) == pytest.approx(c, rel=1e-12)


def test_error_behaviour(load, engine_cls):
    from maniac_b200.engine import ManiacAbort
    s = load("methanol")
    with engine_cls(s, capacity=2) as eng:
        with pytest.raises(ManiacAbort):
            eng.update_system_energy(walker=5)
        with pytest.raises(ManiacAbort):
            eng.pairwise_energy_for_molecule(3, 0)
        with pytest.raises(ManiacAbort):
            eng.compute_new_energy(0, 0, KIND_MOVE)            # geometry missing
        with pytest.raises(ManiacAbort):
            eng.commit()                                       # nothing pending
        com, off = eng.get_molecule(0, 0)
        with pytest.raises(ManiacAbort):
            eng.compute_new_energy(0, 7, KIND_CREATE, com, off)  # beyond capacity (check_molecule_index)
        # the engine is still usable afterwards
        assert np.isfinite(eng.update_system_energy()).all()
    # capacity overflow inside the device-resident loop is reported, not ignored
    s2 = load("methanol")
    s2.residues[0].chemical_potential = 0.0
    s2.residues[0].fugacity = 1.0e12
    s2.p_translation, s2.p_rotation, s2.p_insertion_deletion = 0.0, 0.0, 1.0
    with engine_cls(s2, capacity=3) as eng:
        with pytest.raises(ManiacAbort):
            eng.sweep(400)


def test_host_driven_drivers_match_oracle_and_sweep(load, engine_cls):
    """Third implementation of the drivers: C++ host drivers over the C ABI (the Fortran
    drivers' role, include/maniac_host.h).  Same RNG contract => same trajectory as the
    oracle and as the device-resident sweep."""
    from maniac_b200.hostmc import HostMonteCarlo
    s = load("zif8_h2o_gcmc")
    n = 1500
    o = Oracle(s, capacity=128)
    o.update_system_energy()
    o.seed(4242)
    ref = o.monte_carlo_steps(n)
    with engine_cls(s, n_walkers=4, capacity=128) as eng:
        hm = HostMonteCarlo(eng, seed=4242)
        tr = hm.run(n, trace_walker=0)
        _compare_traces(tr, ref)
        assert hm.count(0) == o.count(0) == eng.count(0)
        assert_e(hm.energy(), o.energy(), rel=1e-9, what="host-side running energy")
        assert_e(eng.energy(0), o.energy(), rel=1e-9, what="device running energy")
        np.testing.assert_array_equal(hm.counters(), o.counters())
        for m in range(o.count(0)):
            co, oo = o.get_molecule(0, m)
            ch, oh = hm.get_molecule(0, m)
            cg, og = eng.get_molecule(0, m)
            np.testing.assert_allclose(ch, co, rtol=0, atol=1e-12)
            np.testing.assert_array_equal(ch, cg)
            np.testing.assert_array_equal(oh, og)
        t = hm.traffic()
        assert t["trials"] > 0 and t["h2d_bytes"] > 0 and t["d2h_bytes"] == 96 * t["trials"]
        hm.close()
    with engine_cls(s, n_walkers=2, capacity=128) as eng:
        eng.seed(4242)
        tr2 = eng.sweep(n, trace_walker=0)
        assert (tr2["accepted"] == tr["accepted"]).all() and (tr2["move"] == tr["move"]).all()


def test_fast_pair_math_on_device(load, engine_cls):
    """1/r^2 from MUFU.RCP64H + Newton and the erfc(alpha r)/r table, measured on the GPU against
    the exact double-precision forms over the whole tabulated range."""
    with engine_cls(load("zif8_h2o_gcmc")) as eng:
        e_rcp, e_tab = eng.selftest_math()
        assert e_rcp < 4e-16, e_rcp
        assert e_tab < 2e-14, e_tab


# ---------------------------------------------------------------------------------------------
# BASELINE configs[4]: binary CO2 / N2 mixture with identity swaps in a triclinic supercell
# ---------------------------------------------------------------------------------------------
def _mixture(load, reps=(1, 1, 1), tilt=2.5, n_co2=6, n_n2=6):
    from maniac_b200.workloads import mixture_supercell
    return mixture_supercell(load("zif8_co2_widom"), reps=reps, tilt_xy=tilt, n_co2=n_co2, n_n2=n_n2)


def test_mixture_triclinic_energy_and_fast_min_image(load, engine_cls):
    """Mildly tilted cell: min_image takes the rounded image + a few lattice candidates instead of the
    27-image loop and must still give the reference's minimum (total energy, S(k), every molecule's pair energy)."""
    s = _mixture(load)
    o = Oracle(s, capacity=64)
    e_ref = o.update_system_energy()
    with engine_cls(s, capacity=64) as eng:
        assert 0 < eng.triclinic_candidates() <= 8
        assert_e(eng.update_system_energy(), e_ref, what="mixture total energy")
        ak_o = o.Ak()
        assert np.abs(eng.Ak() - ak_o).max() <= 1e-11 * max(1.0, np.abs(ak_o).max())
        for res in (1, 2):
            for m in range(o.count(res)):
                assert_e(eng.pairwise_energy_for_molecule(res, m), o.pairwise_energy_for_molecule(res, m), what=f"pair {res},{m}")


def test_very_skewed_cell_keeps_27_image_search(load, engine_cls):
    s = load("dipole_triclinic")
    with engine_cls(s) as eng:
        n = eng.triclinic_candidates()
        assert n == -1 or n > 0
        o = Oracle(s)
        assert_e(eng.update_system_energy(), o.update_system_energy())


def test_swap_energy_commit_rollback(load, engine_cls):
    """attempt_swap_move through the single-trial C ABI: mgpu_swap_energy + mgpu_commit / mgpu_rollback
    against the oracle's driver (forced accept and forced reject)."""
    s = _mixture(load)
    s.p_translation, s.p_rotation, s.p_swap, s.p_insertion_deletion = 0.0, 0.0, 1.0, 0.0
    o = Oracle(s, capacity=64)
    o.update_system_energy()
    o.seed(99)
    with engine_cls(s, capacity=64) as eng:
        eng.update_system_energy()
        done = 0
        for step in range(60):
            n1, n2 = o.count(1), o.count(2)
            tr = o.monte_carlo_steps(1)
            if tr["move"][0] != 5:
                continue
            ra, ma = int(tr["res"][0]), int(tr["mol"][0])
            rb = 3 - ra
            nb_before = n2 if rb == 2 else n1
            if tr["accepted"][0]:
                com, off = o.get_molecule(rb, nb_before)       # the new molecule sits in slot N+1 of its type
            else:
                continue                                        # rejected: the oracle restored everything; geometry not observable
            e_old, e_new = eng.swap_energy(ra, ma, rb, com, off)
            assert_e(e_old, tr["e_old"][0], rel=REL_DE, what="swap old")
            assert_e(e_new, tr["e_new"][0], rel=REL_DE, what="swap new")
            d_ref = tr["e_new"][0][5] - tr["e_old"][0][5]
            assert abs((e_new[5] - e_old[5]) - d_ref) <= REL_DE * max(1.0, abs(d_ref))
            # a rollback first (state must be untouched), then the same trial again and commit
            eng.rollback()
            assert eng.count(1) == n1 and eng.count(2) == n2
            eng.swap_energy(ra, ma, rb, com, off)
            eng.commit()
            assert eng.count(1) == o.count(1) and eng.count(2) == o.count(2)
            assert_e(eng.energy(), o.energy(), rel=1e-9, what="running energy after swap")
            np.testing.assert_allclose(eng.Ak(), o.Ak(), rtol=0, atol=1e-9)     # incl. the reference's un-removed dS of the old molecule
            done += 1
        assert done >= 3


@pytest.mark.parametrize("steps", [1500])
def test_sweep_mixture_with_swaps(steps, load, sweep_engine_cls):
    engine_cls = sweep_engine_cls
    """Device-resident drivers incl. swapping.f90 on the triclinic mixture: same move / accept sequence,
    per-move energies within 1e-9, same final state as the oracle."""
    s = _mixture(load)
    o = Oracle(s, capacity=64)
    o.update_system_energy()
    o.seed(31337)
    ref = o.monte_carlo_steps(steps)
    assert (ref["move"] == 5).sum() > 50 and ((ref["move"] == 5) & (ref["accepted"] == 1)).sum() > 5
    with engine_cls(s, n_walkers=2, capacity=64) as eng:
        eng.seed(31337)
        tr = eng.sweep(steps, trace_walker=0)
        _compare_traces(tr, ref)
        for res in (1, 2):
            assert eng.count(res) == o.count(res)
        assert_e(eng.energy(), o.energy(), rel=1e-9)
        np.testing.assert_array_equal(eng.counters(0), o.counters())
        assert eng.rng_state(0) == o.rng_state()
        ak_o = o.Ak()                     # after `steps` incremental updates on both sides
        assert np.abs(eng.Ak() - ak_o).max() <= 1e-10 * max(1.0, np.abs(ak_o).max())


def test_host_driven_mixture_with_swaps(load, engine_cls):
    from maniac_b200.hostmc import HostMonteCarlo
    s = _mixture(load)
    n = 600
    o = Oracle(s, capacity=64)
    o.update_system_energy()
    o.seed(2718)
    ref = o.monte_carlo_steps(n)
    with engine_cls(s, n_walkers=3, capacity=64) as eng:
        hm = HostMonteCarlo(eng, seed=2718)
        tr = hm.run(n, trace_walker=0)
        _compare_traces(tr, ref)
        for res in (1, 2):
            assert hm.count(res) == o.count(res) == eng.count(res)
        assert_e(eng.energy(0), o.energy(), rel=1e-9)
        np.testing.assert_array_equal(hm.counters(), o.counters())
        hm.close()


def test_large_triclinic_supercell(load, sweep_engine_cls):
    engine_cls = sweep_engine_cls
    """configs[4] at full size: 2x2x2 blocks = 17 664 framework atoms, dense k set; total energy and a short
    trajectory against the oracle, then a longer device run audited by a full recompute (drift gate)."""
    s = _mixture(load, reps=(2, 2, 2), tilt=3.0, n_co2=24, n_n2=24)
    o = Oracle(s, capacity=128)
    e_ref = o.update_system_energy()
    o.seed(5)
    ref = o.monte_carlo_steps(40)
    with engine_cls(s, n_walkers=8, capacity=128) as eng:
        assert eng.ewald()["nk"] == o.ewald()["nk"] > 2000
        assert eng.triclinic_candidates() == 6        # (in cells > ~82 A the list is provably irrelevant and dropped, see mgpu_init)
        assert_e(eng.update_system_energy(), e_ref, what="17.7k-atom triclinic total energy")
        eng.seed(5)
        tr = eng.sweep(40, trace_walker=0)
        _compare_traces(tr, ref)
        eng.sweep(400)
        for w in (0, 7):
            inc = eng.energy(w)
            full = eng.update_system_energy(w)
            # swaps leave the old molecule's dS in S(k) (reference behaviour), so only the real-space parts can be audited
            assert_e(inc[[0, 1, 3, 4]], full[[0, 1, 3, 4]], rel=1e-9, what=f"drift walker {w}")


# ---------------------------------------------------------------------------------------------
# walker records and the block-level entry (host state in, host state out)
# ---------------------------------------------------------------------------------------------
def test_walker_records_roundtrip_and_block(load, sweep_engine_cls):
    engine_cls = sweep_engine_cls
    """mgpu_save_walkers / mgpu_load_walkers / mgpu_block: the record is the whole state, so
    (a) save -> load into another walker reproduces that walker exactly, (b) a trajectory cut into
    blocks that travel through host memory equals the uninterrupted device-resident one and the
    oracle's, (c) malformed records are refused."""
    from maniac_b200.engine import Engine, ManiacAbort
    s = load("zif8_h2o_gcmc")
    nW, cap = 6, 96
    o = Oracle(s, capacity=cap)
    o.update_system_energy()
    o.seed(2025)
    ref = o.monte_carlo_steps(900)
    with engine_cls(s, n_walkers=nW, capacity=cap) as eng:
        eng.seed(2025)
        blob_a = eng.host_buffer(nW * eng.record_doubles_max())
        blob_b = eng.host_buffer(nW * eng.record_doubles_max())
        tr1 = eng.sweep(300, trace_walker=0)
        off = eng.save_walkers(blob_a)
        nk = eng.ewald()["nk"]
        resinfo = [(r.active, r.natom) for r in s.residues]
        rec0 = Engine.parse_record(blob_a[off[0]:off[1]], nk, resinfo)
        assert rec0["length"] == off[1] - off[0]
        assert rec0["count"][0] == eng.count(0, walker=0) == len(rec0["molecules"][0]["com"])
        np.testing.assert_array_equal(rec0["energy"], eng.energy(0))
        np.testing.assert_array_equal(rec0["counters"], eng.counters(0))
        assert tuple(int(x) for x in rec0["rng"]) == tuple(eng.rng_state(0))
        com, offs = eng.get_molecule(0, 2, walker=0)
        np.testing.assert_array_equal(rec0["molecules"][0]["com"][2], com)
        np.testing.assert_array_equal(rec0["molecules"][0]["offset"][2], offs)
        ak = eng.Ak(0)
        np.testing.assert_array_equal(rec0["Ak"][:nk], np.asarray(ak).real)
        np.testing.assert_array_equal(rec0["Ak"][nk:], np.asarray(ak).imag)
        # (a) walker 0's record loaded into walker 3: identical state, identical continuation
        eng.load_walkers(blob_a[off[0]:off[1]], np.array([0, off[1] - off[0]]), first_walker=3)
        assert eng.count(0, walker=3) == eng.count(0, walker=0)
        np.testing.assert_array_equal(eng.energy(3), eng.energy(0))
        off = eng.save_walkers(blob_a)                       # records again, walker 3 now being the clone
        # (b) two more blocks through host memory: records in, records out
        off2 = eng.block(300, blob_a, off, blob_b)
        off3 = eng.block(300, blob_b, off2, blob_a)
        rec_end = Engine.parse_record(blob_a[off3[0]:off3[1]], nk, resinfo)
        rec_end3 = Engine.parse_record(blob_a[off3[3]:off3[4]], nk, resinfo)
        assert rec_end["count"][0] == o.count(0)
        assert_e(rec_end["energy"], o.energy(), rel=1e-9, what="energy after 3 blocks through host records")
        np.testing.assert_array_equal(rec_end["counters"], o.counters())
        assert tuple(int(x) for x in rec_end["rng"]) == tuple(o.rng_state())
        # walker 3 was a clone of walker 0 after block 1 (same RNG state): it must have followed the same path
        np.testing.assert_array_equal(rec_end3["energy"], rec_end["energy"])
        np.testing.assert_array_equal(rec_end3["molecules"][0]["com"], rec_end["molecules"][0]["com"])
        # the uninterrupted run of a fresh engine state gives the same trajectory as the oracle (sanity of ref)
        assert (tr1["accepted"] == ref["accepted"][:300]).all()
        # drift audit on the restored state
        inc = eng.energy(0)
        full = eng.update_system_energy(0)
        assert_e(inc, full, rel=1e-9, what="restored state vs full recompute")
        t = eng.traffic()
        assert t["h2d_bytes"] > 8 * off[-1] and t["d2h_bytes"] > 8 * off[-1]
        # (c) malformed: wrong length word, count beyond capacity
        bad = blob_b[off2[0]:off2[1]].copy()
        bad[0] += 1
        with pytest.raises(ManiacAbort):
            eng.load_walkers(bad, np.array([0, len(bad)]), first_walker=1)
        bad = blob_b[off2[0]:off2[1]].copy()
        bad[1] = cap + 5
        with pytest.raises(ManiacAbort):
            eng.load_walkers(bad, np.array([0, len(bad)]), first_walker=1)
        small = eng.host_buffer(100)
        with pytest.raises(ManiacAbort):
            eng.save_walkers(small)


def test_step_size_adaptation_per_block(load, sweep_engine_cls):
    engine_cls = sweep_engine_cls
    """adjust_move_step_sizes (src/monte_carlo_utils.f90:98-134) at the end of every block, on the device, per walker:
    same step sizes and same trajectory as the oracle doing the same."""
    s = load("zif8_h2o_gcmc")
    o = Oracle(s, capacity=96)
    o.update_system_energy()
    o.seed(808)
    with engine_cls(s, n_walkers=2, capacity=96) as eng:
        eng.seed(808)
        assert eng.step_sizes(0) == o.step_sizes()
        for blk in range(4):
            ref = o.monte_carlo_steps(700)
            tr = eng.sweep(700, trace_walker=0)
            _compare_traces(tr, ref)
            o.adjust_move_step_sizes()
            eng.adjust_move_step_sizes()
            t_ref, r_ref = o.step_sizes()
            t_gpu, r_gpu = eng.step_sizes(0)
            assert abs(t_gpu - t_ref) <= 1e-14 * t_ref and abs(r_gpu - r_ref) <= 1e-14 * r_ref
        assert eng.step_sizes(0) != (s.translation_step, s.rotation_step_angle)       # it did adapt
        assert eng.step_sizes(1) != eng.step_sizes(0)                                  # every walker its own
        eng.set_step_sizes(0.5, 0.3, walker=1)
        assert eng.step_sizes(1) == (0.5, 0.3)
    # the host-driven drivers adapt the same way
    from maniac_b200.hostmc import HostMonteCarlo
    o = Oracle(s, capacity=96)
    o.update_system_energy()
    o.seed(909)
    with engine_cls(s, n_walkers=2, capacity=96) as eng:
        hm = HostMonteCarlo(eng, seed=909)
        for blk in range(3):
            ref = o.monte_carlo_steps(700)
            tr = hm.run(700, trace_walker=0)
            _compare_traces(tr, ref)
            o.adjust_move_step_sizes()
            hm.adjust_move_step_sizes()
            assert np.allclose(hm.step_sizes(0), o.step_sizes(), rtol=1e-14, atol=0)
        hm.close()


def test_overlap_sentinel_of_pairs_with_no_interaction(load, engine_cls):
    """pairwise_lj_energy returns 1e20 for r < 1e-10 whatever epsilon is (src/pairwise_energy_utils.f90:116-119), so even an
    atom pair with neither LJ nor Coulomb (TIP4P O on top of another water's M or H) must show up.  The engine screens
    those pair lists per molecule (centre distance vs molecular radii); the screen must let exactly these cases through."""
    s = load("zif8_h2o_gcmc")
    o = Oracle(s, capacity=16)
    o.update_system_energy()
    com0, off0 = o.get_molecule(0, 0)
    com1, off1 = o.get_molecule(0, 1)
    with engine_cls(s, capacity=16) as eng:
        eng.update_system_energy()
        for tgt_atom in (1, 2):                      # O of molecule 1 exactly on M, then on H, of molecule 0
            new_com = (com0 + off0[tgt_atom]) - off1[0]
            o.save_fourier(0, 1)
            old_ref = o.compute_old_energy(0, 1, KIND_MOVE)
            o.set_molecule(0, 1, new_com, off1)
            new_ref = o.compute_new_energy(0, 1, KIND_MOVE)
            o.set_molecule(0, 1, com1, off1)
            o.update_system_energy()
            assert new_ref[0] > 1e19                  # the sentinel is in the non-Coulomb component
            new_gpu = eng.compute_new_energy(0, 1, KIND_MOVE, new_com, off1)
            eng.rollback()
            assert new_gpu[0] > 1e19 and abs(new_gpu[0] - new_ref[0]) <= 1e-9 * abs(new_ref[0])
        # a molecule larger than any seen before, handed in through the API, must still be screened correctly
        big = off1 * 3.0
        eng.set_molecule(0, 2, com1 + np.array([9.0, 0.0, 0.0]), big)
        eng.update_system_energy()
        o.set_molecule(0, 2, com1 + np.array([9.0, 0.0, 0.0]), big)
        o.update_system_energy()
        cb, ob = o.get_molecule(0, 2)
        new_com = (cb + ob[2]) - off1[0]              # O of molecule 1 on the far-out H of the enlarged molecule
        o.save_fourier(0, 1)
        o.compute_old_energy(0, 1, KIND_MOVE)
        o.set_molecule(0, 1, new_com, off1)
        new_ref = o.compute_new_energy(0, 1, KIND_MOVE)
        new_gpu = eng.compute_new_energy(0, 1, KIND_MOVE, new_com, off1)
        assert new_ref[0] > 1e19 and new_gpu[0] > 1e19
