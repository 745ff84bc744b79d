"""Generate the committed fixtures under tests/golden/ from the reference's own input files.

Run in the build container, where /root/reference exists:

    python tests/golden/make_fixtures.py

It (1) reads the reference fixtures with the package's readers and stores them as .npz
snapshots (the GPU box has no /root/reference), (2) records the known-answer values the
reference's tests pin for them (LAMMPS logs / analytic values, with the reference's own
tolerances), and (3) records the oracle's component energies for the same systems as a
regression pin of the oracle itself.  Nothing here copies reference source code; the
snapshots are the reference's input DATA (coordinates, charges, pair_coeff lines).
"""
import json
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
import maniac_b200  # noqa: E402
from maniac_b200.inputs import Residue, load_system  # noqa: E402
from maniac_b200.snapshot import save_snapshot  # noqa: E402
from oracle.oracle import Oracle  # noqa: E402

REF = Path("/root/reference")
OUT = Path(__file__).resolve().parent
MT = REF / "mc-topology"

# name -> (dir, input, data, params, KAT dict) ; KATs cite the reference file that pins them
CASES = {
    "lj_gas": (MT / "testcase-energy/LJ-gas", "input.maniac", "topology.data", "parameters.inc",
               dict(total=-38.589253, tol=0.001, e_vdwl=-0.40707303, e_coul=49.708669, e_long=-87.890849,
                    source="mc-topology/testcase-energy/LJ-gas/log.lammps:79-80; tests/integration/energy/LJ-gas/run-test.sh:21")),
    "zif8_h2o": (MT / "testcase-energy/ZIF8-H2O", "input.maniac", "topology.data", "parameters.inc",
                 dict(total=-10050.673, tol=3.0, e_vdwl=-3.5031658, e_coul=133.29021, e_long=-10180.46,
                      source="mc-topology/testcase-energy/ZIF8-H2O/log.lammps:150-151; tests/integration/energy/ZIF8-H2O/run-test.sh:21")),
    "h2o_gas": (MT / "testcase-energy/H2O-gas", "input.maniac", "topology.data", "parameters.inc",
                dict(total=-32.822927, tol=0.015, source="tests/integration/energy/H2O-gas/run-test.sh:15-16")),
    "methanol": (REF / "tests/integration/gcmc", "input.maniac", "topology.data", "parameters.inc",
                 dict(total=-0.012882149, tol=0.1, e_coul=28.911538, e_long=-28.924421,
                      source="tests/integration/gcmc/test_delete_and_create.f90:27-30")),
    "two_atoms": (REF / "tests/integration/energy-analytical/two-atoms", "input.maniac", "topology.data", "parameters.inc",
                  dict(total=-8.3192992, tol=0.001, source="tests/integration/energy-analytical/two-atoms/run-test.sh:16")),
    "dipole": (REF / "tests/integration/energy-analytical/dipole", "input.maniac", "topology.data", "parameters.inc",
               dict(total=-0.0015669252, tol=0.001, source="tests/integration/energy-analytical/dipole/run-test.sh:16")),
    "two_dipole": (REF / "tests/integration/energy-analytical/two-dipole", "input.maniac", "topology.data", "parameters.inc",
                   dict(total=16.167421, tol=0.001, source="tests/integration/energy-analytical/two-dipole/run-test.sh:16")),
    # the GCMC config of BASELINE.json configs[1] (same cell as zif8_h2o, adsorption input)
    "zif8_h2o_gcmc": (MT / "testcase-adsorption/ZIF8-H2O", "zif8-water.maniac", "zif8-water.data", "zif8-water.inc", None),
    # triclinic box (71 dipoles); no reference test pins it (transposed-H quirk)
    "dipole_triclinic": (MT / "testcase-adsorption/DIPOLE-orthorhombic", "dipole.maniac", "dipole.data", "dipole.inc", None),
    "drift_methanol": (REF / "tests/integration/energy-drift", "input.maniac", "topology.data", "parameters.inc", None),
}


def widom_co2_in_zif8():
    """BASELINE.json configs[2]: framework of testcase-widom/ZIF8-MET/zif8.data, CO2 from
    molecule-reservoir/CO2 (co2.mol geometry / charges, parameters.inc), cross terms by the
    Lorentz-Berthelot fill (no explicit cross pair_coeff lines)."""
    d = MT / "testcase-widom/ZIF8-MET"
    s = load_system(d / "input.maniac", d / "zif8.data", d / "parameters.inc")
    zif = [r for r in s.residues if not r.active][0]
    # CO2: types 8 (O) and 9 (C) in the merged numbering (0-based 7, 8)
    co2 = Residue(name="co2", active=True, natom=3, site_types=[8, 9], site_names=["O", "C"], fugacity=-1.0,
                  chemical_potential=0.0)
    co2.types = np.array([7, 8, 7], dtype=np.int32)
    co2.charges = np.array([-0.3256, 0.6512, -0.3256])
    # COM quirk: mass = natom * mass(last atom's type) = 3 * 15.9994
    co2.mass = 3 * 15.9994
    geom = np.array([[1.149, 0.0, 0.0], [0.0, 0.0, 0.0], [-1.149, 0.0, 0.0]])
    co2.com = np.zeros((0, 3))
    co2.offset = np.zeros((0, 3, 3))
    co2.template = geom - geom.mean(axis=0)
    pc = [p for p in s.pair_coeff if p[0] < 7 and p[1] < 7]
    pc += [(7, 7, 0.159961, 3.033), (7, 8, 0.0945081, 2.895), (8, 8, 0.0558373, 2.757)]
    s.residues = [zif, co2]
    s.ntypes = 9
    s.pair_coeff = pc
    s.p_translation, s.p_rotation, s.p_insertion_deletion, s.p_widom = 0.0, 0.0, 0.0, 1.0
    s.translation_step = 1.0
    return s


def main():
    kat = {}
    for name, (d, inp, dat, par, ref) in CASES.items():
        s = load_system(d / inp, d / dat, d / par)
        o = Oracle(s)
        e = o.update_system_energy()
        ew = o.ewald()
        save_snapshot(OUT / f"{name}.npz", s)
        kat[name] = dict(reference=ref, oracle_energy=[float(x) for x in e], alpha=ew["alpha"], kmax=ew["kmax"],
                         nk=ew["nk"], rc=ew["rc"], natoms=int(sum(r.natom * r.nmol for r in s.residues)))
        print(name, "E =", e, "nk", ew["nk"], "ref", ref and ref["total"])
    s = widom_co2_in_zif8()
    # one CO2 is placed as slot-1 template (the reference copies the geometry of molecule 1)
    co2 = s.residues[1]
    box_c = s.lo + 0.5 * np.array([s.matrix[0, 0], s.matrix[1, 1], s.matrix[2, 2]])
    co2.com = np.array([box_c])
    co2.offset = np.array([co2.template])
    del co2.template
    save_snapshot(OUT / "zif8_co2_widom.npz", s)
    o = Oracle(s)
    o.set_count(1, 0)          # empty pore; slot 1 keeps the template geometry
    e = o.update_system_energy()
    ew = o.ewald()
    kat["zif8_co2_widom"] = dict(reference=None, oracle_energy=[float(x) for x in e], alpha=ew["alpha"], kmax=ew["kmax"],
                                 nk=ew["nk"], rc=ew["rc"], natoms=2208)
    print("zif8_co2_widom", e, ew)
    (OUT / "kat.json").write_text(json.dumps(kat, indent=1))


if __name__ == "__main__":
    main()
