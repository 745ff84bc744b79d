"""Round-2 GPU parity tests (through the C ABI, against the CPU oracle): commits of guests with more than ten
sites (k_commit's row copies span several warps), walker records of EMPTY walkers (slot 1 = insertion
template), pipelined mgpu_block == unpipelined, argument checks of the batched trial entry."""
import copy

import numpy as np
import pytest

from oracle.oracle import KIND_CREATE, KIND_DELETE, KIND_MOVE, Oracle

pytestmark = pytest.mark.gpu
FUG12 = 1.0e-6          # creations and deletions of the fused (strongly self-attracting) guest balance here


def close_e(a, b, rel):
    a, b = np.asarray(a, dtype=float), np.asarray(b, dtype=float)
    return np.all(np.abs(a - b) <= rel * np.maximum(1.0, np.abs(b)))


def _compare_traces(tr, ref):
    assert (tr["move"] == ref["move"]).all()
    assert (tr["mol"] == ref["mol"]).all()
    assert (tr["accepted"] == ref["accepted"]).all()
    assert np.all(np.abs(tr["dE"] - ref["dE"]) <= 1e-9 * np.maximum(1.0, np.abs(ref["dE"])))


def _twelve_site_system(load):
    """Methanol box (tests/integration/gcmc) with every two methanols fused into one rigid 12-site guest:
    natom * 3 + 2 = 38 rows per molecule, more than one warp of k_commit's 128 threads can copy in one go."""
    s = copy.deepcopy(load("methanol"))
    r = [x for x in s.residues if x.active][0]
    na = r.natom
    rng = np.random.default_rng(11)
    shift = np.array([2.9, 0.4, -0.3])
    off1 = np.asarray(r.offset[0])
    geom = np.concatenate([off1, off1 + shift])
    geom = geom - geom.mean(axis=0)                       # COM frame = plain centroid (data_parser.f90:1425 quirk)
    nmol = 6
    L = np.diag(np.asarray(s.matrix))
    grid = np.array([[i, j, k] for i in range(2) for j in range(2) for k in range(2)], dtype=float)[:nmol]
    com = np.asarray(s.lo) + L * (0.25 + 0.5 * grid) + 0.3 * rng.random((nmol, 3))
    r.natom = 2 * na
    r.types = np.concatenate([r.types, r.types]).astype(np.int32)
    r.charges = np.concatenate([r.charges, r.charges])
    r.site_types = list(r.site_types) * 2
    r.site_names = list(r.site_names) * 2
    r.mass = 2.0 * r.mass
    r.com = com
    r.offset = np.repeat(geom[None], nmol, axis=0)
    r.fugacity = FUG12
    s.p_translation, s.p_rotation, s.p_insertion_deletion, s.p_swap, s.p_widom = 0.3, 0.3, 0.4, 0.0, 0.0
    return s


def test_commit_of_guest_with_twelve_sites(load):
    """Host-driven drivers (mgpu_trial_batch + mgpu_commit_batch -> k_commit) on a 12-site guest: accepted
    deletions move the last molecule into the hole (38 rows), creations append; the trajectory, the final
    coordinates and the running energies must follow the oracle."""
    from maniac_b200.engine import Engine
    from maniac_b200.hostmc import HostMonteCarlo
    s = _twelve_site_system(load)
    cap = 32
    o = Oracle(s, capacity=cap)
    e_ref = o.update_system_energy()
    o.seed(4242)
    with Engine(s, n_walkers=3, capacity=cap) as eng:
        assert close_e(eng.update_system_energy(0), e_ref, 1e-10)
        hm = HostMonteCarlo(eng, seed=4242)
        ref = o.monte_carlo_steps(600)
        tr = hm.run(600, trace_walker=0)
        _compare_traces(tr, ref)
        res = [i for i, r in enumerate(s.residues) if r.active][0]
        assert ref["accepted"][ref["move"] == 3].sum() > 0 or ref["accepted"][ref["move"] == 4].sum() > 0   # some creation / deletion went through
        n = o.count(res)
        assert eng.count(res, walker=0) == n
        for m in range(n):
            c_o, off_o = o.get_molecule(res, m)
            c_g, off_g = eng.get_molecule(res, m, walker=0)
            np.testing.assert_array_equal(c_g, c_o)
            np.testing.assert_array_equal(off_g, off_o)
        assert close_e(eng.energy(0), o.energy(), 1e-9)
        inc, full = eng.energy(0), eng.update_system_energy(0)
        assert close_e(inc, full, 1e-9)
        hm.close()
    # the same system through the device-resident drivers (one warp commits: same rows, other code path)
    o = Oracle(s, capacity=cap)
    o.update_system_energy()
    o.seed(99)
    ref = o.monte_carlo_steps(600)
    with Engine(s, n_walkers=2, capacity=cap) as eng:
        eng.seed(99)
        _compare_traces(eng.sweep(600, trace_walker=0), ref)


def test_empty_walker_record_travels_with_its_insertion_template(load):
    """A walker with N = 0 still owns slot 1: the geometry the next insertion starts from, overwritten by every
    rejected creation.  Its record must carry it: saved from walker 0, loaded into walker 2 (whose slot 1 holds
    something else), the continuation must be walker 0's own = the oracle's."""
    from maniac_b200.engine import Engine
    s = copy.deepcopy(load("zif8_co2_widom"))
    s.residues[1].fugacity = 1.0e-8                  # the walker hovers around N = 0 .. 2
    s.p_translation, s.p_rotation, s.p_insertion_deletion, s.p_widom = 0.2, 0.2, 0.6, 0.0
    cap = 48
    o = Oracle(s, capacity=cap)
    o.set_count(1, 0)
    o.update_system_energy()
    o.seed(78)
    ref_a = o.monte_carlo_steps(40)                  # from empty: rejected creations rewrite slot 1
    with Engine(s, n_walkers=3, capacity=cap) as eng:
        for w in range(3):
            eng.set_count(1, 0, walker=w)
            eng.update_system_energy(w)
        eng.seed(78)
        tr_a = eng.sweep(40, trace_walker=0)
        _compare_traces(tr_a, ref_a)
        # find a point where walker 0 is empty again (or still): extend in small blocks until count == 0
        extra = 0
        while eng.count(1, walker=0) != 0 and extra < 400:
            ref_x = o.monte_carlo_steps(1)
            tr_x = eng.sweep(1, trace_walker=0)
            _compare_traces(tr_x, ref_x)
            extra += 1
        assert eng.count(1, walker=0) == 0 and o.count(1) == 0, "no empty state reached; change the seed"
        blob = eng.host_buffer(3 * eng.record_doubles_max())
        off = eng.save_walkers(blob)
        nk = eng.ewald()["nk"]
        rec0 = Engine.parse_record(blob[off[0]:off[1]], nk, [(r.active, r.natom) for r in s.residues])
        assert rec0["count"][1] == 0 and len(rec0["molecules"][1]["com"]) == 0
        _, tmpl = eng.get_molecule(1, 0, walker=0)
        np.testing.assert_array_equal(rec0["molecules"][1]["template_offset"], tmpl)
        # walker 2 has run its own stream: another slot-1 geometry
        _, other = eng.get_molecule(1, 0, walker=2)
        assert not np.array_equal(other, tmpl)
        eng.load_walkers(blob[off[0]:off[1]], np.array([0, off[1] - off[0]]), first_walker=2)
        _, got = eng.get_molecule(1, 0, walker=2)
        np.testing.assert_array_equal(got, tmpl)
        ref_b = o.monte_carlo_steps(300)
        tr_b = eng.sweep(300, trace_walker=2)         # the clone continues exactly like the original would have
        _compare_traces(tr_b, ref_b)
        assert eng.count(1, walker=2) == o.count(1)


def test_pipelined_block_equals_unpipelined(load):
    """mgpu_block in 8 slices on 8 streams (copies under compute) gives bit-identical records to the plain
    load -> sweep -> save (walkers are independent), incl. ragged last slice and empty walkers."""
    from maniac_b200.engine import OPT_BLOCK_SLICES, Engine
    s = copy.deepcopy(load("zif8_h2o_gcmc"))
    nW, cap, steps = 150, 64, 60                      # 150 = 9 CTAs of 16 + a ragged one
    with Engine(s, n_walkers=nW, capacity=cap) as eng:
        for w in range(0, nW, 7):
            eng.set_count(0, 0, walker=w)             # some empty walkers
            eng.update_system_energy(w)
        for w in range(nW):
            eng.set_fugacity(0, 10.0 ** (-2 + 6 * (w % 13) / 12.0), walker=w)
        eng.seed(5150)
        size = nW * eng.record_doubles_max()
        b0, b1, b2 = eng.host_buffer(size), eng.host_buffer(size), eng.host_buffer(size)
        off0 = eng.save_walkers(b0)
        eng.set_option(OPT_BLOCK_SLICES, 1)
        off1 = eng.block(steps, b0, off0, b1)
        eng.set_option(OPT_BLOCK_SLICES, 8)
        off2 = eng.block(steps, b0, off0, b2)
        np.testing.assert_array_equal(off1, off2)
        np.testing.assert_array_equal(b1[:off1[-1]], b2[:off2[-1]])
        assert off1[-1] != off0[-1] or not np.array_equal(b1[:off1[-1]], b0[:off0[-1]])     # something happened
        # a second pipelined block from the pipelined output, against a device-resident continuation
        off3 = eng.block(steps, b2, off2, b1)
        eng.load_walkers(b2[:off2[-1]], off2)
        eng.sweep(steps)
        off4 = eng.save_walkers(b0)
        np.testing.assert_array_equal(off3, off4)
        np.testing.assert_array_equal(b1[:off3[-1]], b0[:off4[-1]])
        # output buffer too small: refused, error names the need
        from maniac_b200.engine import ManiacAbort
        small = eng.host_buffer(int(off2[-1]) // 2)
        with pytest.raises(ManiacAbort):
            eng.block(steps, b2, off2, small)


def test_trial_batch_argument_checks(load):
    """mgpu_trial_batch refuses what would race or read stale slots: a walker listed twice, a walker with a pending
    trial, a molecule index beyond the current count, a creation that is not slot N + 1; mgpu_sweep refuses to run over
    a pending trial; a malformed record batch leaves every walker untouched."""
    from maniac_b200.engine import Engine, ManiacAbort
    s = load("zif8_h2o_gcmc")
    with Engine(s, n_walkers=4, capacity=16) as eng:
        n = eng.count(0)
        com, off = eng.get_molecule(0, 0)
        off16 = np.zeros((16, 3))
        off16[:len(off)] = off
        W = np.array([1, 1], dtype=np.int32)
        z = np.zeros(2, dtype=np.int32)
        with pytest.raises(ManiacAbort, match="twice"):
            eng.trial_batch(W, z, z, np.full(2, KIND_MOVE, dtype=np.int32), np.tile(com, (2, 1)), np.tile(off16, (2, 1, 1)))
        with pytest.raises(ManiacAbort, match="count"):
            eng.compute_new_energy(0, n, KIND_MOVE, com, off)
        with pytest.raises(ManiacAbort, match="count"):
            eng.compute_new_energy(0, n + 1, KIND_CREATE, com, off)
        with pytest.raises(ManiacAbort, match="count"):
            eng.compute_new_energy(0, n, KIND_DELETE, None, None)
        eng.compute_new_energy(0, 0, KIND_MOVE, com + 0.1, off)           # leaves a pending trial on walker 0
        with pytest.raises(ManiacAbort, match="pending"):
            eng.compute_new_energy(0, 1, KIND_MOVE, com + 0.2, off)
        with pytest.raises(ManiacAbort, match="pending"):
            eng.sweep(5)
        eng.rollback()
        eng.compute_new_energy(0, n, KIND_CREATE, com + 3.0, off)          # creation in slot N + 1, accepted: count mirror follows
        eng.commit()
        assert eng.count(0) == n + 1
        eng.compute_new_energy(0, n, KIND_MOVE, com + 3.1, off)            # the new molecule is addressable now
        eng.rollback()
        eng.sweep(20)                                                       # counts change on the device ...
        n2 = eng.count(0)
        with pytest.raises(ManiacAbort, match="count"):                     # ... and the mirror is refreshed before the next check
            eng.compute_new_energy(0, n2, KIND_MOVE, com, off)
        # malformed batch: second record bad -> the first walker must not have been overwritten
        blob = eng.host_buffer(4 * eng.record_doubles_max())
        offs = eng.save_walkers(blob)
        e_before = [eng.energy(w).copy() for w in range(4)]
        eng.sweep(30)
        e_mid = [eng.energy(w).copy() for w in range(4)]
        bad = blob[:offs[-1]].copy()
        bad[offs[1]] += 2.0                                                  # length word of record 1
        with pytest.raises(ManiacAbort, match="malformed"):
            eng.load_walkers(bad, offs)
        for w in range(4):
            np.testing.assert_array_equal(eng.energy(w), e_mid[w])           # nothing was written
        short = offs.copy()
        short[1] = 40                                                        # record 0 shorter than its header
        with pytest.raises(ManiacAbort, match="malformed"):
            eng.load_walkers(blob[:offs[-1]], short)
        eng.load_walkers(blob[:offs[-1]], offs)
        for w in range(4):
            np.testing.assert_array_equal(eng.energy(w), e_before[w])


@pytest.mark.parametrize("n_walkers,expect_threads", [(5, 128), (450, 128), (1000, 64), (2000, 32)])
def test_automatic_sweep_shapes(n_walkers, expect_threads, load):
    """MGPU_OPT_SWEEP_TEAM = -1 (the default): the engine picks the threads per walker and the walkers per CTA from the
    number of walkers in flight (mgpu_get_sweep_shape) -- CTAs that are not full quartets / not the maximal team count,
    e.g. 7 two-warp teams or 14 warps.  The traced walker and the last walker of the launch (a partly filled CTA) must
    follow the oracle whatever the shape."""
    s = load("zif8_h2o_gcmc")
    steps = 600
    from maniac_b200.engine import Engine
    with Engine(s, n_walkers=n_walkers, capacity=64) as eng:
        shape = eng.sweep_shape(n_walkers)
        sm = eng.launch_info()["sm_count"]
        assert shape["threads_per_walker"] == expect_threads or sm != 148
        assert shape["walkers_per_cta"] * shape["threads_per_walker"] <= 512
        eng.seed(4321)
        last = n_walkers - 1
        tr = eng.sweep(steps, trace_walker=last)
        o = Oracle(s, capacity=64)
        o.update_system_energy()
        o.seed(4321 + 104729 * last)                      # the walker's own stream: state = splitmix64(seed + 104729 w)
        ref = o.monte_carlo_steps(steps)
        _compare_traces(tr, ref)
        assert eng.count(0, walker=last) == o.count(0)
        assert close_e(eng.energy(last), o.energy(), 1e-9)
