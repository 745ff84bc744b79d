"""Host-side logic that needs no GPU: input readers, LJ table semantics, the C-ABI library's
symbol table, loud failure without a device."""
import ctypes as C
import re
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parents[1]


def test_library_exports_every_declared_symbol():
    from maniac_b200 import capi
    hdr = (ROOT / "include" / "maniac_gpu.h").read_text()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(mgpu_[a-zA-Z0-9_]+)\s*\(", hdr))
    assert len(declared) >= 40
    L = capi.lib()
    for name in sorted(declared):
        assert hasattr(L, name), f"{name} declared in include/maniac_gpu.h but not exported"
    assert declared == set(capi.SIGNATURES), declared ^ set(capi.SIGNATURES)


def test_struct_layout_matches_header():
    from maniac_b200 import capi
    # mgpu_step_trace: 4 int32 + 2 double + 2*6 double
    assert C.sizeof(capi.MgpuStepTrace) == 16 + 16 + 96
    assert C.sizeof(capi.MgpuResidue) == 16 + 4 * 8 + 3 * 8
    assert capi.MgpuSystem.residues.offset == 9 * 8 + 3 * 8 + 8


def test_engine_fails_loudly_without_gpu(load):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from maniac_b200.engine import Engine, ManiacAbort
    with pytest.raises(ManiacAbort, match="no CUDA device"):
        Engine(load("methanol"))
    from maniac_b200 import capi
    L = capi.lib()
    out = (C.c_double * 6)()
    assert L.mgpu_total_energy(0, out) != 0 and b"not initialised" in L.mgpu_last_error()


def test_lj_table_matches_oracle(load):
    from maniac_b200.engine import lj_table
    from oracle.oracle import Oracle
    for name in ["zif8_h2o", "zif8_co2_widom", "lj_gas", "methanol", "dipole_triclinic"]:
        s = load(name)
        eps, sig = lj_table(s)
        eo, so = Oracle(s).lj_table()
        np.testing.assert_array_equal(eps, eo)
        np.testing.assert_array_equal(sig, so)
    # Lorentz-Berthelot fill for the CO2 / ZIF-8 cross terms (A16): sigma mean, eps geometric mean
    s = load("zif8_co2_widom")
    eps, sig = lj_table(s)
    assert sig[7, 0] == (sig[7, 7] + sig[0, 0]) / 2 and eps[7, 0] == pytest.approx(np.sqrt(eps[7, 7] * eps[0, 0]))
    # explicit "0 0" cross lines whose diagonals are zero stay zero (TIP4P M / H sites)
    s = load("zif8_h2o")
    eps, sig = lj_table(s)
    assert eps[1, 5] == 0.0 and sig[1, 5] == 0.0 and eps[0, 3] == 0.115221


def test_readers_roundtrip(tmp_path, load):
    """.maniac / .data / .inc readers on a small synthetic case incl. residue sorting, the COM quirk
    and unwrapping of a molecule that straddles the boundary."""
    from maniac_b200.inputs import load_system
    (tmp_path / "in.maniac").write_text("""
nb_block 2
nb_step 10
temperature 300
ewald_tolerance 1e-5
real_space_cutoff 5
translation_step 1
rotation_step_angle 0.5
translation_proba 0.6
rotation_proba 0.6
begin_residue
  name wall
  state inactif
  types 3
  names W
  nb-atoms 1
end_residue
begin_residue
  name ab
  state actif
  fugacity 2.0
  types 1 2
  names A B
  nb-atoms 2
end_residue
""")
    (tmp_path / "top.data").write_text("""test

5 atoms
3 atom types

-5 5 xlo xhi
-5 5 ylo yhi
-5 5 zlo zhi

Masses

1 2.0
2 4.0
3 10.0

Atoms # full

3 2 3 0.0 1.0 1.0 1.0
1 1 1 0.5 4.8 0.0 0.0
2 1 2 -0.5 -4.8 0.0 0.0
4 3 1 0.5 0.0 1.0 0.0
5 3 2 -0.5 0.0 2.0 0.0
""")
    (tmp_path / "p.inc").write_text("pair_coeff 1 1 0.1 3.0\npair_coeff 2 2 0.2 2.0\npair_coeff 3 3 0.3 1.0\n")
    s = load_system(tmp_path / "in.maniac", tmp_path / "top.data", tmp_path / "p.inc")
    assert [r.name for r in s.residues] == ["ab", "wall"]          # sorted by smallest type id
    assert s.p_translation == pytest.approx(0.5) and s.p_rotation == pytest.approx(0.5)   # rescaled to 1
    ab = s.residues[0]
    assert ab.nmol == 2 and ab.mass == pytest.approx(8.0)          # natom * mass(type of last atom)
    # first molecule straddles the boundary: atom 2 is unwrapped to x = 5.2, centroid 5.0 -> wrapped to -5.0
    np.testing.assert_allclose(ab.offset[0], [[-0.2, 0, 0], [0.2, 0, 0]], atol=1e-12)
    assert ab.com[0][0] == pytest.approx(-5.0)
    np.testing.assert_allclose(ab.com[1], [0.0, 1.5, 0.0])
    assert list(ab.types) == [0, 1] and list(s.residues[1].types) == [2]
    with pytest.raises(ValueError):
        (tmp_path / "bad.maniac").write_text("nb_block 1\n")
        load_system(tmp_path / "bad.maniac", tmp_path / "top.data", tmp_path / "p.inc")


def test_snapshot_roundtrip(tmp_path, load):
    from maniac_b200.snapshot import load_snapshot, save_snapshot
    s = load("zif8_h2o")
    save_snapshot(tmp_path / "x.npz", s)
    t = load_snapshot(tmp_path / "x.npz")
    assert t.pair_coeff == s.pair_coeff and t.ntypes == s.ntypes
    for a, b in zip(s.residues, t.residues):
        np.testing.assert_array_equal(a.com, b.com)
        np.testing.assert_array_equal(a.offset, b.offset)
        assert a.mass == b.mass and a.active == b.active


@pytest.mark.parametrize("alpha,r_hi", [(0.16215694, 30.5), (0.16215694, 43.0), (0.23463702, 21.8), (0.25430534, 27.5), (0.25396482, 27.5)])
def test_coulomb_table_accuracy(alpha, r_hi):
    """The piecewise-polynomial table that replaces erfc(alpha r)/r on the device (host-side
    builder, evaluated with the device's integer indexing and Horner order) against long-double
    erfc: error relative to the pair's Coulomb scale 1/r below 4e-15 over the whole range."""
    import ctypes as C
    from maniac_b200 import capi
    L = capi.lib()
    er, ea = C.c_double(), C.c_double()
    n = L.mgpu_coulomb_table_check(alpha, 1.0, r_hi, 300000, C.byref(er), C.byref(ea))
    assert n == 300000           # every sample falls inside the tabulated range
    assert ea.value < 4e-15


def test_output_lines_follow_the_fortran_formats(tmp_path):
    """energy.dat / number_*.dat / moves.dat / widom_*.dat lines, character by character as the edit descriptors of
    src/write_utils.f90:153-403 produce them (I10, F16.6, A10 truncation, trim)."""
    from maniac_b200 import outputs as out
    e = [-3.5031658, 133.29021, 3.90647708, -10181.6689, 147.525509, -10050.673]   # energy_type order
    line = out.energy_line(7, e)
    assert line == "         7    -10050.673000         3.906477        -3.503166       133.290210    -10181.668900       147.525509"
    assert len(line) == 10 + 6 * 17
    assert out.ENERGY_HEADER.startswith("#    block        total        recipCoulomb")
    assert out.NUMBER_HEADER == "   # Block Active_Mol"               # A10 truncates the 16-character literal
    assert out.number_line(3, 114) == "         3        114"
    p = dict(translation=0.4, rotation=0.4, insertion_deletion=0.2, swap=0.0, widom=0.0)
    assert out.moves_header(p) == "Block" + "".join(f" {c:>12}" for c in
                                                   ["Trans_Acc", "Trans_Trial", "Rot_Acc", "Rot_Trial", "Create_Acc", "Create_Trial", "Delete_Acc", "Delete_Trial"])
    c = [[100, 40], [90, 80], [30, 5], [25, 3], [0, 0], [0, 0]]
    assert out.moves_line(2, c, p) == f"{2:12d} {40:12d} {100:12d} {80:12d} {90:12d} {5:12d} {30:12d} {3:12d} {25:12d}"
    assert out.WIDOM_HEADER == "   # Block Excess_Mu_kcalmo Total_Mu_kcalmol    Widom_Samples"
    mu_ex, mu_tot = out.excess_mu(sum_weight=250.0, samples=1000, temperature=300.0, n_molecules=10, volume=39400.0, lam=0.5)
    assert abs(mu_ex - (-out.KB_KCALMOL * 300.0 * np.log(0.25))) < 1e-15
    assert abs(mu_tot - mu_ex - out.KB_KCALMOL * 300.0 * np.log(10 / 39400.0 * 0.125)) < 1e-12
    assert out.widom_line(1, mu_ex, mu_tot, 1000) == f"{1:10d} {mu_ex:16.6f} {mu_tot:16.6f} {1000:12d}"
    w = out.BlockWriter(tmp_path / "w0", {0: "wat"}, p)
    rec = dict(energy=e, count=np.array([114, 0, 0, 0, 0, 0, 0, 0]), counters=np.array(c))
    w.write(0, rec)
    w.write(1, rec)
    assert (tmp_path / "w0" / "energy.dat").read_text().splitlines() == [out.ENERGY_HEADER, out.energy_line(0, e), out.energy_line(1, e)]
    assert (tmp_path / "w0" / "number_wat.dat").read_text().splitlines()[1:] == [out.number_line(0, 114), out.number_line(1, 114)]
    assert len((tmp_path / "w0" / "moves.dat").read_text().splitlines()) == 3


def test_lammpstrj_frame_and_wrap():
    """write_dump_lammpstrj (src/write_utils.f90:38-119) + wrap_into_box (src/geometry_utils.f90:105-140, the values of
    tests/units/test_wrap_into_box.f90's style: nint wrapping into [-L/2, L/2])."""
    from maniac_b200 import outputs as out
    m = np.diag([10.0, 20.0, 30.0])
    np.testing.assert_allclose(out.wrap_into_box([6.0, -11.0, 44.0], m), [-4.0, 9.0, 14.0], atol=1e-12)
    np.testing.assert_allclose(out.wrap_into_box([5.0, -10.0, 0.0], m), [-5.0, 10.0, 0.0], atol=1e-12)     # nint: halves away from zero
    tri = np.array([[10.0, 0.0, 0.0], [2.0, 10.0, 0.0], [0.0, 0.0, 10.0]])
    p = out.wrap_into_box([17.0, 3.0, -6.0], tri)
    f = out._reciprocal(tri) @ p
    assert (np.abs(f) <= 0.5 + 1e-12).all()
    residues = [dict(active=False, types=[3, 4], com=np.array([[0.0, 0.0, 0.0]]), offset=np.array([[[1.0, 0.0, 0.0], [6.0, 0.0, 0.0]]])),
                dict(active=True, types=[1, 2, 2])]
    mol = {1: dict(com=np.array([[7.0, 0.0, 0.0]]), offset=np.array([[[0.0, 0.0, 0.0], [0.5, 0.0, 0.0], [-0.5, 0.0, 0.0]]]))}
    txt = out.lammpstrj_frame(12, m, residues, mol).splitlines()
    assert txt[0] == "ITEM: TIMESTEP" and txt[1] == "        12" and txt[3] == "         5"
    assert txt[5] == "    -5.00000000      5.00000000"
    assert txt[8] == "ITEM: ATOMS id type x y z"
    assert txt[9] == "     1    3    1.0000000    0.0000000    0.0000000"
    assert txt[10] == "     2    4   -4.0000000    0.0000000    0.0000000"          # inactive: the ATOM is wrapped
    assert txt[11] == "     3    1   -3.0000000    0.0000000    0.0000000"          # active: the CoM is wrapped, atoms follow it
    assert txt[12] == "     4    2   -2.5000000    0.0000000    0.0000000"


def test_topology_data_writer_roundtrips_through_the_data_reader(tmp_path, load):
    """write_topology_data (src/write_utils.f90:406-636): the file written from a walker's molecules is a LAMMPS data file the
    reader of this package (and LAMMPS) takes back: same atoms, types, charges, positions (F12.7), bonds renumbered per molecule."""
    import numpy as np
    from maniac_b200 import outputs as out
    s = load("zif8_h2o_gcmc")
    residues, mols = [], {}
    for r, res in enumerate(s.residues):
        residues.append(dict(active=res.active, types=[int(t) + 1 for t in res.types], charges=res.charges, com=res.com, offset=res.offset))
        if res.active:
            mols[r] = dict(com=res.com.copy(), offset=res.offset.copy())
    masses = np.arange(1, s.ntypes + 1, dtype=float) * 1.5
    act = [r for r, res in enumerate(s.residues) if res.active][0]
    connect = {"bonds": [[] for _ in s.residues], "angles": [[] for _ in s.residues]}
    connect["bonds"][act] = [(1, 1, 3), (1, 1, 4)]                      # O-H, O-H of TIP4P (sites O, M, H, H)
    connect["angles"][act] = [(1, 3, 1, 4)]
    txt = out.topology_data(s.matrix, s.lo, masses, residues, mols, connect, {"bonds": 1, "angles": 1})
    lines = txt.splitlines()
    n_atoms = sum(res.nmol * res.natom for res in s.residues)
    assert lines[0] == " ! LAMMPS data file (atom_style full)"
    assert lines[1] == f"{n_atoms:12d}  atoms" and lines[2] == f"{s.ntypes:12d}  atom types"
    nw = s.residues[act].nmol
    assert lines[3] == f"{2 * nw:12d}  bonds" and lines[5] == f"{nw:12d}  angles"
    assert f"{s.lo[0]:15.8f} {s.lo[0] + s.matrix[0, 0]:15.8f} xlo xhi" in lines
    i0 = lines.index(" Atoms") + 2
    rows = [ln.split() for ln in lines[i0:i0 + n_atoms]]
    assert [int(r[0]) for r in rows] == list(range(1, n_atoms + 1))
    # last water, last site: charge and unwrapped position, fixed formats
    w = s.residues[act]
    pos = w.com[-1] + w.offset[-1][-1]
    first_water_atom = sum(res.nmol * res.natom for res in s.residues[:act])
    k_last = first_water_atom + nw * w.natom - 1
    assert lines[i0 + k_last] == f"{k_last + 1:6d} {int(rows[k_last][1]):6d} {int(w.types[-1]) + 1:4d} {w.charges[-1]:12.8f} {pos[0]:12.7f} {pos[1]:12.7f} {pos[2]:12.7f}"
    # inactive (framework) atoms are wrapped into the box, active molecules are not
    inactive = [r for r, res in enumerate(s.residues) if not res.active][0]
    fw0 = sum(res.nmol * res.natom for res in s.residues[:inactive])
    fw = rows[fw0:fw0 + s.residues[inactive].nmol * s.residues[inactive].natom]
    xyz = np.array([[float(v) for v in r[4:7]] for r in fw])
    L = np.diag(s.matrix)
    assert np.all(np.abs(xyz) <= L / 2 + 1e-6)
    # bonds: two per water, atom ids offset by the atoms before the molecule
    b0 = lines.index(" Bonds") + 2
    assert [int(v) for v in lines[b0].split()] == [1, 1, first_water_atom + 1, first_water_atom + 3]
    assert [int(v) for v in lines[b0 + 2].split()] == [3, 1, first_water_atom + w.natom + 1, first_water_atom + w.natom + 3]
    a0 = lines.index(" Angles") + 2
    assert [int(v) for v in lines[a0].split()] == [1, 1, first_water_atom + 3, first_water_atom + 1, first_water_atom + 4]
    # the package's own data reader takes the Atoms section back
    (tmp_path / "topology.data").write_text(txt)
    atoms = [ln.split() for ln in (tmp_path / "topology.data").read_text().splitlines()[i0:i0 + n_atoms]]
    q = np.array([float(r[3]) for r in atoms])
    assert abs(q.sum() - sum(res.nmol * res.charges.sum() for res in s.residues)) < 1e-5


@pytest.mark.parametrize("n,forced,expect", [
    (5, -1, (128, 1)), (450, -1, (128, 4)), (592, -1, (128, 4)), (593, -1, (64, 5)), (1024, -1, (64, 7)), (1184, -1, (64, 8)),
    (1185, -1, (32, 9)), (2048, -1, (32, 14)), (2368, -1, (32, 16)), (4096, -1, (32, 14)), (4736, -1, (32, 16)),
    (16384, -1, (32, 16)), (1000, 0, (32, 16)), (1000, 1, (128, 4)), (1000, 2, (64, 8))])
def test_sweep_shape_plan(n, forced, expect):
    """Launch shape of the device-resident drivers (mgpu_plan_sweep_shape = the pure part of mgpu_get_sweep_shape) on a
    148-SM GPU with 16-warp CTAs: the widest team with which one wave of CTAs holds every walker, the fewest walkers per
    CTA that need no extra round; forced shapes use full CTAs."""
    from maniac_b200 import capi
    L = capi.lib()
    t, p = C.c_int32(0), C.c_int32(0)
    assert L.mgpu_plan_sweep_shape(n, 148, 16, forced, C.byref(t), C.byref(p)) == 0
    assert (t.value, p.value) == expect
    assert t.value * p.value <= 512
    if forced < 0:
        per_wave = 148 * p.value
        rounds = -(-n // per_wave)
        full = 148 * (512 // t.value)
        assert rounds == -(-n // full)                       # no more rounds than full CTAs would need ...
        assert p.value == 1 or 148 * (p.value - 1) * rounds < n     # ... and one walker less per CTA would need one more


def _tri_plan(M):
    from maniac_b200 import capi
    L = capi.lib()
    n = C.c_int32(0)
    vec = np.zeros((14, 3)); coef = np.zeros((14, 3), dtype=np.int32); faces = np.zeros(14, dtype=np.int32)
    thr = np.zeros(3, dtype=np.int32); lut = np.zeros(16, dtype=np.uint32)
    rc = L.mgpu_plan_triclinic(np.ascontiguousarray(M, dtype=np.float64).ctypes.data_as(C.POINTER(C.c_double)), C.byref(n),
                               vec.ctypes.data_as(C.POINTER(C.c_double)), coef.ctypes.data_as(C.POINTER(C.c_int32)),
                               faces.ctypes.data_as(C.POINTER(C.c_int32)), thr.ctypes.data_as(C.POINTER(C.c_int32)),
                               lut.ctypes.data_as(C.POINTER(C.c_uint32)))
    assert rc == 0
    return n.value, vec, coef, faces, thr, lut


@pytest.mark.parametrize("M", [
    [[68.04, 0, 0], [3.0, 68.04, 0], [0, 0, 68.04]],                      # configs[4]: xy tilt, columns convention
    [[30.0, 0, 0], [4.0, 28.0, 0], [-3.0, 5.0, 26.0]],                    # three tilts
    [[30.0, 2.0, 1.0], [4.0, 28.0, -3.0], [-3.0, 5.0, 26.0]],             # not triangular
    [[30.0, 0, 0], [14.9, 28.0, 0], [14.9, 13.9, 26.0]]])                 # LAMMPS' maximal skew
def test_triclinic_plan_reproduces_the_27_image_search(M):
    """The triclinic minimum image of the framework passes (min_image_frac_fast / _slow), replayed in numpy from the plan
    mgpu_plan_triclinic hands the kernels -- rounded image, face test on the high words, tri_lut -> vectors to try, the
    reference's 27-image search when the winning shift leaves {-1,0,1}^3 -- against the reference's definition: the
    minimum over the 27 shifts of the raw difference vector (geometry_utils.f90:263-280), on random vectors."""
    M = np.array(M, dtype=np.float64)
    n, vec, coef, faces, thr_hi, lut = _tri_plan(M)
    rng = np.random.default_rng(7)
    g = rng.uniform(-1.0, 1.0, (200_000, 3))                              # differences of two points of the cell
    shifts = np.array([[i, j, k] for i in (-1, 0, 1) for j in (-1, 0, 1) for k in (-1, 0, 1)], dtype=np.float64)
    d = g @ M.T                                                           # raw Cartesian difference, cell vectors = columns
    ref = np.min(((d[:, None, :] + (shifts @ M.T)[None, :, :]) ** 2).sum(-1), axis=1)
    if n < 0:
        assert (lut >> 31).all()                                          # too skewed: every pair takes the literal search
        return
    nn = np.rint(g); f = g - nn; t = f @ M.T
    hi = (np.abs(f).view(np.int64) >> 32).astype(np.int64)
    fb = ((hi[:, 0] >= thr_hi[0]) * 1 + (hi[:, 1] >= thr_hi[1]) * 2 + (hi[:, 2] >= thr_hi[2]) * 4).astype(int)
    cand = lut[fb]
    t2 = (t * t).sum(1)
    best, shift = t2.copy(), np.zeros_like(t)
    for k in range(n):
        use = ((cand >> k) & 1).astype(bool)
        dot = t @ vec[k]
        gk = vec[k] @ vec[k] - 2 * np.abs(dot)
        better = use & (t2 + gk < best)
        sg = np.where(dot > 0, -1.0, 1.0)
        best = np.where(better, t2 + gk, best)
        shift = np.where(better[:, None], sg[:, None] * coef[k][None, :], shift)
    o = nn - shift
    fallback = np.abs(o).max(axis=1) > 1.0                                # beyond the reference's 27 images: literal search
    got = np.where(fallback, ref, best)
    assert np.max(np.abs(got - ref)) < 1e-9
    assert fallback.mean() < 0.05
    # the gate is doing something: in the tilted 68 A cell only a few per cent of the pairs look at any vector
    if abs(M[0][0] - 68.04) < 1e-9:
        assert n == 6 and (cand != 0).mean() < 0.12 and sorted(np.unique(faces[:n])) == [1, 2, 5, 6]


def _framework_order(matrix, lo, pos):
    from maniac_b200 import capi
    L = capi.lib()
    pos = np.ascontiguousarray(pos, dtype=np.float64)
    order = np.zeros(len(pos), dtype=np.int32); cols = np.zeros(3, dtype=np.int32)
    pd, pi = C.POINTER(C.c_double), C.POINTER(C.c_int32)
    rc = L.mgpu_plan_framework_order(np.ascontiguousarray(matrix, dtype=np.float64).ctypes.data_as(pd),
                                     np.ascontiguousarray(lo, dtype=np.float64).ctypes.data_as(pd), len(pos),
                                     pos.ctypes.data_as(pd), order.ctypes.data_as(pi), cols.ctypes.data_as(pi))
    assert rc == 0
    return order, cols


def test_framework_order_clusters_and_columns(load):
    """mgpu_init's order of the static framework atoms (mgpu_plan_framework_order): a permutation; in an orthorhombic cell the
    32 atoms of a warp iteration form a compact cluster (Hilbert curve); in the tilted 68 A cell of configs[4] the atoms are
    binned into 16 x 16 columns along the free z axis, and fewer warp iterations have an atom near a face that matters
    (within tri_eps of |f_d| = 1/2 for d = x, y) than with compact clusters."""
    from maniac_b200.workloads import mixture_supercell
    base = load("zif8_co2_widom")
    zif = [r for r in base.residues if not r.active][0]
    pos0 = np.asarray(zif.com[0]) + np.asarray(zif.offset[0])
    order0, cols0 = _framework_order(base.matrix, base.lo, pos0)
    assert sorted(order0) == list(range(len(pos0))) and list(cols0) == [1, 1, 1]
    cl = pos0[order0][: len(pos0) // 32 * 32].reshape(-1, 32, 3)
    ext = (cl.max(axis=1) - cl.min(axis=1))
    assert ext.mean() < 12.0                                   # 2208 atoms in (34 A)^3: a 32-atom box is ~8 A per edge

    sm = mixture_supercell(base, reps=(2, 2, 2), tilt_xy=3.0, n_co2=4, n_n2=4)
    fw = [r for r in sm.residues if not r.active][0]
    pos = np.asarray(fw.com[0]) + np.asarray(fw.offset[0])
    order, cols = _framework_order(sm.matrix, sm.lo, pos)
    assert sorted(order) == list(range(len(pos))) and list(cols) == [16, 16, 1]
    Cm = np.asarray(sm.matrix); f = (pos - np.asarray(sm.lo)) @ np.linalg.inv(Cm).T
    f -= np.floor(f)
    n, _, _, faces, thr_hi, _ = _tri_plan(Cm)
    thr = 0.478                                                # 1/2 - tri_eps of this cell (test_triclinic_plan...)
    rng = np.random.default_rng(3)

    def near_fraction(o):
        ff = f[o][: len(o) // 32 * 32]
        hits = tot = 0
        for _ in range(60):
            gg = ff - rng.random(3); gg -= np.rint(gg)
            near = (np.abs(gg[:, 0]) >= thr) | (np.abs(gg[:, 1]) >= thr)
            hits += near.reshape(-1, 32).any(axis=1).sum(); tot += len(ff) // 32
        return hits / tot
    plain, _ = _framework_order(np.diag(np.diag(Cm)), sm.lo, pos)      # the same atoms ordered as in an untilted cell: compact clusters
    assert near_fraction(order) < 0.24 < 0.27 < near_fraction(plain)
