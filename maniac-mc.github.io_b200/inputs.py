"""Readers for MANIAC's on-disk inputs (.maniac keyword file, LAMMPS .data, pair_coeff .inc).

They produce a :class:`System` -- the host-side image of ``simulation_state`` that the
energy engine is initialised from (``mgpu_init``).  Behaviour follows the reference
readers, including their quirks, so that energies computed from the result match what
MANIAC itself would compute from the same files:

* keyword file: ``src/input_parser.f90:213-550`` (keywords), ``:555-627`` (residues are
  sorted by their smallest atom-type id), ``:59-92`` (probabilities rescaled to 1).
* data file: ``src/readers_utils.f90:170-260`` (box; the matrix rows are a, b, c),
  ``src/data_parser.f90:551-669`` (Atoms, ``full`` style), ``:1068-1140`` (sort by id),
  ``:1147-1294`` (molecule detection by type pattern), ``:1301-1374`` + ``readers_utils.f90:266-331``
  (active molecules are made whole), ``:1380-1486`` (COM frame).
* COM quirk (``data_parser.f90:1425``): the mass vector of a residue is a scalar
  broadcast of the mass of the *last* atom's type, so the "centre of mass" is the plain
  centroid and ``res%mass = natom * mass(type of last atom)``.
* LJ table: ``src/parameters_parser.f90:19-178`` (``pair_coeff`` lines in file order are
  handed to the engine / oracle, which applies the Lorentz-Berthelot fill).

This is host logic: pure Python / NumPy, no GPU, no oracle.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from pathlib import Path
from typing import List, Optional, Sequence, Tuple

import numpy as np

ERROR = 1.0e-10  # src/parameters.f90:57


@dataclass
class Residue:
    """One residue type (``res%...`` / ``thermo%...`` entries of one index)."""

    name: str
    active: bool
    natom: int
    site_types: List[int]                 # atom-type ids listed in the input (1-based, as in the files)
    site_names: List[str] = field(default_factory=list)
    fugacity: float = -1.0                # input_parser.f90:262 (unset)
    chemical_potential: float = 0.0       # input_parser.f90:264 (unset)
    # filled by the data reader
    types: Optional[np.ndarray] = None    # (natom,) atom type per site, 0-based
    charges: Optional[np.ndarray] = None  # (natom,)
    mass: float = 0.0                     # res%mass (quirk above)
    com: Optional[np.ndarray] = None      # (nmol, 3)
    offset: Optional[np.ndarray] = None   # (nmol, natom, 3)

    @property
    def nmol(self) -> int:
        return 0 if self.com is None else int(self.com.shape[0])


@dataclass
class System:
    """Everything ``mgpu_init`` needs; mirrors the reference's module-level state."""

    matrix: np.ndarray                    # (3,3) box%cell%matrix (rows a,b,c)
    lo: np.ndarray                        # (3,)  box%cell%bounds(:,1)
    residues: List[Residue]
    ntypes: int
    pair_coeff: List[Tuple[int, int, float, float]]   # (type_i, type_j, eps, sigma), 0-based, file order
    temperature: float = 300.0
    ewald_tolerance: float = 1e-5
    real_space_cutoff: float = 12.0
    translation_step: float = 1.0
    rotation_step_angle: float = 0.685
    recalibrate_moves: bool = False
    p_translation: float = 0.0
    p_rotation: float = 0.0
    p_swap: float = 0.0
    p_insertion_deletion: float = 0.0
    p_widom: float = 0.0
    nb_block: int = 0
    nb_step: int = 0
    seed: int = 0
    masses: Optional[np.ndarray] = None   # per atom type

    def copy(self) -> "System":
        import copy as _copy
        return _copy.deepcopy(self)


# ----------------------------------------------------------------------------------
# .maniac keyword file
# ----------------------------------------------------------------------------------
def _strip_comment(line: str) -> str:
    return line.split("#", 1)[0].strip()


def read_maniac_input(path) -> dict:
    """Parse the keyword file (input_parser.f90:213-550).  Returns a dict of scalars plus
    ``residues`` (already sorted like ``sort_residues``)."""
    out = dict(
        nb_block=None, nb_step=None, temperature=None, seed=0, ewald_tolerance=None,
        real_space_cutoff=None, translation_step=None, rotation_step_angle=None,
        recalibrate_moves=False, translation_proba=0.0, rotation_proba=0.0,
        insertion_deletion_proba=0.0, swap_proba=0.0, widom_proba=0.0,
    )
    residues: List[Residue] = []
    cur: Optional[dict] = None
    real_keys = {"temperature", "ewald_tolerance", "real_space_cutoff", "translation_step",
                 "rotation_step_angle", "translation_proba", "rotation_proba",
                 "insertion_deletion_proba", "swap_proba", "widom_proba"}
    positive = {"temperature", "ewald_tolerance", "real_space_cutoff", "translation_step",
                "rotation_step_angle"}
    for raw in Path(path).read_text().splitlines():
        if raw[:1] == "#" or not raw.strip():
            continue
        line = _strip_comment(raw)
        if not line:
            continue
        tok = line.split()
        key, rest = tok[0], tok[1:]
        if key == "begin_residue":
            cur = dict(name="", active=None, natom=0, types=[], names=[], fugacity=-1.0, mu=0.0)
            continue
        if key == "end_residue":
            if cur is None:
                raise ValueError("end_residue without begin_residue")
            if cur["active"] is None:
                raise ValueError("Unknown residue state")
            residues.append(Residue(name=cur["name"], active=cur["active"], natom=cur["natom"],
                                    site_types=cur["types"], site_names=cur["names"],
                                    fugacity=cur["fugacity"], chemical_potential=cur["mu"]))
            cur = None
            continue
        if cur is not None:
            if key == "name":
                cur["name"] = rest[0]
            elif key == "state":
                if rest[0] == "actif":
                    cur["active"] = True
                elif rest[0] == "inactif":
                    cur["active"] = False
                else:
                    raise ValueError("Unknown residue state")      # input_parser.f90:414-416
            elif key == "fugacity":
                cur["fugacity"] = float(rest[0])
            elif key == "chemical_potential":
                cur["mu"] = float(rest[0])
            elif key == "nb-atoms":
                cur["natom"] = int(rest[0])
            elif key == "types":
                cur["types"] = [int(x) for x in rest]
            elif key == "names":
                cur["names"] = list(rest)
            continue
        if key in ("nb_block", "nb_step", "seed"):
            out[key] = int(rest[0])
        elif key in real_keys:
            v = float(rest[0].replace("d", "e").replace("D", "e"))
            if key in positive and v <= 0.0:
                raise ValueError(f"Invalid {key}: must be > 0")
            if key.endswith("_proba") and not (0.0 <= v <= 1.0):
                raise ValueError(f"Invalid {key}: must be in [0,1]")
            out[key] = v
        elif key == "recalibrate_moves":
            out[key] = rest[0].strip(".").lower() in ("true", "t")
    for req in ("nb_block", "nb_step", "temperature", "real_space_cutoff", "ewald_tolerance",
                "translation_step", "rotation_step_angle"):
        if out[req] is None:
            raise ValueError(f"Missing required parameter: {req}")
    # validation of fugacity / chemical potential, input_parser.f90:489-509
    for r in residues:
        if not r.active:
            continue
        has_f, has_mu = r.fugacity >= 0.0, r.chemical_potential < 0.0
        if not (has_f or has_mu) and out["insertion_deletion_proba"] > 0:
            raise ValueError(f"Chemical potential not provided for GCMC insert of active residue: {r.name}")
        if has_f and has_mu:
            raise ValueError(f"Both fugacity and chemical potential were specified for active residue: {r.name}")
    tot = sum(out[k] for k in ("translation_proba", "rotation_proba", "insertion_deletion_proba",
                               "swap_proba", "widom_proba"))
    if tot < ERROR:
        raise ValueError("Invalid move probabilities: all enabled moves have zero probability")
    if out["widom_proba"] > 0 and out["insertion_deletion_proba"] > 0:
        raise ValueError("Cannot enable both Widom insertions and physical insertion/deletion moves")
    # rescale_move_probabilities, input_parser.f90:59-92
    if abs(tot - 1.0) > ERROR:
        sf = 1.0 / tot
        for k in ("translation_proba", "rotation_proba", "insertion_deletion_proba", "swap_proba", "widom_proba"):
            out[k] = out[k] * sf
    # sort_residues, input_parser.f90:555-627: stable insertion sort on min atom type
    keys = [min(r.site_types) for r in residues]
    order = list(range(len(residues)))
    for i in range(1, len(order)):
        k = order[i]
        j = i - 1
        while j >= 0 and keys[order[j]] > keys[k]:
            order[j + 1] = order[j]
            j -= 1
        order[j + 1] = k
    out["residues"] = [residues[i] for i in order]
    return out


# ----------------------------------------------------------------------------------
# LAMMPS .data
# ----------------------------------------------------------------------------------
def f_modulo(a: float, p: float) -> float:
    """gfortran MODULO for reals: fmod plus sign fix."""
    r = math.fmod(a, p)
    if r != 0.0 and ((r < 0.0) != (p < 0.0)):
        r += p
    return r


def box_shape_is_triclinic(matrix: np.ndarray) -> bool:
    """determine_box_symmetry, geometry_utils.f90:341-361."""
    off = [matrix[0, 1], matrix[0, 2], matrix[1, 0], matrix[1, 2], matrix[2, 0], matrix[2, 1]]
    return max(abs(x) for x in off) > ERROR


def box_reciprocal(matrix: np.ndarray) -> np.ndarray:
    """compute_box_determinant_and_inverse, geometry_utils.f90:148-200 (adjugate columns)."""
    a, b, c = matrix[:, 0], matrix[:, 1], matrix[:, 2]
    adj = np.empty((3, 3))
    adj[:, 0] = np.cross(b, c)
    adj[:, 1] = np.cross(c, a)
    adj[:, 2] = np.cross(a, b)
    det = float(np.dot(a, adj[:, 0]))
    return adj * (1.0 / det)


def apply_pbc(pos: np.ndarray, matrix: np.ndarray, lo: np.ndarray) -> np.ndarray:
    """apply_PBC, geometry_utils.f90:45-97."""
    pos = np.array(pos, dtype=np.float64)
    if not box_shape_is_triclinic(matrix):
        for d in range(3):
            pos[d] = lo[d] + f_modulo(pos[d] - lo[d], matrix[d, d])
        return pos
    rec = box_reciprocal(matrix)
    frac = rec @ (pos - lo)
    frac = np.array([f_modulo(x, 1.0) for x in frac])
    return lo + matrix @ frac


def _wrap_nearest(x: float, L: float) -> float:
    return f_modulo(x + 0.5 * L, L) - 0.5 * L          # readers_utils.f90:324-331


def _repair_molecule(xyz: np.ndarray, matrix: np.ndarray) -> np.ndarray:
    """repair_molecule, readers_utils.f90:266-318."""
    xyz = xyz.copy()
    tri = box_shape_is_triclinic(matrix)
    rec = box_reciprocal(matrix) if tri else None
    for i in range(1, xyz.shape[0]):
        d = xyz[i] - xyz[i - 1]
        if not tri:
            d = np.array([_wrap_nearest(d[k], matrix[k, k]) for k in range(3)])
        else:
            f = rec @ d
            f = np.array([_wrap_nearest(x, 1.0) for x in f])
            d = matrix @ f
        xyz[i] = xyz[i - 1] + d
    return xyz


def read_lammps_data(path, residues: Sequence[Residue]):
    """Read box, masses and atoms of a LAMMPS data file (atom style ``full``) and split the
    atoms into molecules of the given residue types.  Fills the residues in place and
    returns ``(matrix, lo, ntypes, masses)``."""
    lines = Path(path).read_text().splitlines()
    natoms = ntypes = None
    lo = np.zeros(3)
    hi = np.zeros(3)
    tilt = np.zeros(3)
    for ln in lines:
        t = ln.split("#", 1)[0].split()
        if len(t) >= 2 and t[1] == "atoms" and natoms is None:
            natoms = int(t[0])
        elif len(t) >= 3 and t[1] == "atom" and t[2] == "types":
            ntypes = int(t[0])
        elif len(t) >= 4 and t[2:4] == ["xlo", "xhi"]:
            lo[0], hi[0] = float(t[0]), float(t[1])
        elif len(t) >= 4 and t[2:4] == ["ylo", "yhi"]:
            lo[1], hi[1] = float(t[0]), float(t[1])
        elif len(t) >= 4 and t[2:4] == ["zlo", "zhi"]:
            lo[2], hi[2] = float(t[0]), float(t[1])
        elif len(t) >= 6 and t[3:6] == ["xy", "xz", "yz"]:
            tilt[:] = [float(t[0]), float(t[1]), float(t[2])]
    if natoms is None or ntypes is None:
        raise ValueError(f"{path}: header without atoms / atom types")
    for d, nm in enumerate("xyz"):
        if abs(lo[d]) < 1e-11 and abs(hi[d]) < 1e-11:
            raise ValueError(f"parse_lammps_box: {nm}lo {nm}hi not found in input file!")
    lx, ly, lz = hi - lo
    # readers_utils.f90:256-258 -- rows are the cell vectors
    matrix = np.array([[lx, 0.0, 0.0], [tilt[0], ly, 0.0], [tilt[1], tilt[2], lz]], dtype=np.float64)

    def section(name):
        for i, ln in enumerate(lines):
            if ln.strip().startswith(name):
                return i
        return None

    masses = np.zeros(ntypes)
    im = section("Masses")
    if im is None:
        raise ValueError("No masses found in data file")
    found = 0
    j = im + 1
    while found < ntypes and j < len(lines):
        t = lines[j].split("#", 1)[0].split()
        j += 1
        if not t:
            if found:
                break
            continue
        masses[int(t[0]) - 1] = float(t[1])
        found += 1
    if found != ntypes:
        raise ValueError("Number of masses found in data file differs from declared atom types")

    ia = section("Atoms")
    if ia is None:
        raise ValueError(f"No atoms found in data file: {path}")
    ids = np.empty(natoms, dtype=np.int64)
    types = np.empty(natoms, dtype=np.int64)
    q = np.empty(natoms)
    xyz = np.empty((natoms, 3))
    k = 0
    j = ia + 1
    while k < natoms:
        if j >= len(lines):
            raise ValueError(f"Unexpected end of file at atom line {k + 1} in: {path}")
        t = lines[j].split()
        j += 1
        if not t:
            continue
        ids[k], types[k], q[k] = int(t[0]), int(t[2]), float(t[3])
        if types[k] < 1 or types[k] > ntypes:
            raise ValueError(f"Invalid atom type {types[k]} (max allowed: {ntypes}) in: {path}")
        xyz[k] = [float(t[4]), float(t[5]), float(t[6])]
        k += 1
    order = np.argsort(ids, kind="stable")               # sort_atoms_by_original_ID
    types, q, xyz = types[order], q[order], xyz[order]

    # detect_residue_pattern, data_parser.f90:1147-1185
    pattern = [[0] * r.natom for r in residues]
    cpt = [0] * len(residues)
    for t in types:
        for i, r in enumerate(residues):
            if int(t) in r.site_types:
                pattern[i][cpt[i]] = int(t)
                cpt[i] += 1
                if cpt[i] >= r.natom:
                    cpt[i] = 0
                break

    type_mass_index = []
    i_m = 0
    for r in residues:                                    # read_lammps_masses mapping, data_parser.f90:266-277
        type_mass_index.append(list(range(i_m, i_m + len(r.site_types))))
        i_m += len(r.site_types)

    for i, r in enumerate(residues):
        coms, offs = [], []
        kk = 0
        r.types, r.charges = None, None
        while kk < natoms:
            if int(types[kk]) == pattern[i][0]:
                if kk + r.natom > natoms:
                    raise ValueError("Not enough atoms left in box to complete residue type")
                mt = types[kk:kk + r.natom]
                if r.active and list(map(int, mt)) != pattern[i]:
                    raise ValueError("Issue with atom order in data file")
                mol_xyz = xyz[kk:kk + r.natom].copy()
                if r.active:
                    mol_xyz = _repair_molecule(mol_xyz, matrix)
                r.types = np.array(mt, dtype=np.int32) - 1           # last molecule wins, like the reference
                r.charges = q[kk:kk + r.natom].copy()
                original_com = mol_xyz.sum(axis=0) * 0.0
                # compute_COM with equal masses m: sum(m*x)/sum(m), accumulated in atom order
                last_type = int(mt[-1])
                m = 0.0
                for l, st in enumerate(r.site_types):
                    if st == last_type:
                        m = float(masses[type_mass_index[i][l]]) if type_mass_index[i][l] < ntypes else 0.0
                acc = np.zeros(3)
                tot = 0.0
                for a in range(r.natom):
                    acc = acc + m * mol_xyz[a]
                    tot = tot + m
                if tot <= 0.0:
                    raise ValueError("Total mass is zero or negative")
                original_com = acc / tot
                r.mass = tot
                com = apply_pbc(original_com, matrix, lo)
                coms.append(com)
                offs.append(mol_xyz - original_com)
                kk += r.natom
            else:
                kk += 1
        if r.types is None:
            # no molecule of this type in the file: take the site list as the pattern
            r.types = np.array([r.site_types[min(a, len(r.site_types) - 1)] for a in range(r.natom)], dtype=np.int32) - 1
            r.charges = np.zeros(r.natom)
        r.com = np.array(coms, dtype=np.float64).reshape(-1, 3)
        r.offset = np.array(offs, dtype=np.float64).reshape(-1, r.natom, 3)
    return matrix, lo, ntypes, masses


def read_pair_coeff(path) -> List[Tuple[int, int, float, float]]:
    """``pair_coeff i j eps sigma`` lines in file order (parameters_parser.f90:59-79), 0-based types."""
    out = []
    for raw in Path(path).read_text().splitlines():
        if raw[:1] == "#" or not raw.strip():
            continue
        t = raw.split("#", 1)[0].split()
        if t and t[0] == "pair_coeff":
            if len(t) < 5:
                raise ValueError("Failed to read pair_coeff value")
            out.append((int(t[1]) - 1, int(t[2]) - 1, float(t[3]), float(t[4])))
    return out


def load_system(input_file, data_file, parameter_file) -> System:
    """The reference's start-up sequence ``read_input_file`` -> ``read_system_data`` ->
    ``read_parameters`` (src/main.f90:20-23) for a primary box without reservoir."""
    inp = read_maniac_input(input_file)
    residues = inp["residues"]
    matrix, lo, ntypes, masses = read_lammps_data(data_file, residues)
    pc = read_pair_coeff(parameter_file)
    return System(
        matrix=matrix, lo=lo, residues=residues, ntypes=ntypes, pair_coeff=pc,
        temperature=inp["temperature"], ewald_tolerance=inp["ewald_tolerance"],
        real_space_cutoff=inp["real_space_cutoff"], translation_step=inp["translation_step"],
        rotation_step_angle=inp["rotation_step_angle"], recalibrate_moves=inp["recalibrate_moves"],
        p_translation=inp["translation_proba"], p_rotation=inp["rotation_proba"], p_swap=inp["swap_proba"],
        p_insertion_deletion=inp["insertion_deletion_proba"], p_widom=inp["widom_proba"],
        nb_block=inp["nb_block"], nb_step=inp["nb_step"], seed=inp["seed"], masses=masses,
    )
