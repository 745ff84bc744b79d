"""Per-block output files in the reference's formats, written from walker records.

The reference appends one line per block to ``energy.dat``, ``number_<res>.dat``, ``moves.dat`` and
``widom_<res>.dat`` (``src/write_utils.f90:153-403``).  With thousands of walkers per GPU the same
lines are produced per walker (or per isotherm point) from the record ``mgpu_save_walkers`` /
``mgpu_block`` return -- the host stays in charge of I/O, the device never formats text.  The
strings below reproduce the Fortran edit descriptors (``I10``, ``F16.6``, ``A10`` truncation of
longer literals, ``trim``) character by character.  Host logic only: no GPU, no oracle.
"""
from __future__ import annotations

import math
from pathlib import Path
from typing import Dict, Iterable, Sequence

import numpy as np

KB_KCALMOL = 1.9872171591124124e-3      # KB * NA * J_to_kcal with the reference's literals


def _a(text: str, width: int) -> str:
    """Fortran ``Aw`` output editing: right-justified, truncated on the right if longer."""
    return text[:width].rjust(width)


# ---- energy.dat (write_dat_energy, src/write_utils.f90:153-186) --------------------------------------
ENERGY_HEADER = ("#    block        total        recipCoulomb"
                 "     non-coulomb      coulomb     ewald_self    intramolecular-coulomb")


def energy_line(block: int, energy: Sequence[float]) -> str:
    """``energy`` in energy_type order (non_coulomb, coulomb, recip, self, intra, total)."""
    nc, c, rec, self_, intra, tot = (float(x) for x in energy)
    return f"{block:10d} {tot:16.6f} {rec:16.6f} {nc:16.6f} {c:16.6f} {self_:16.6f} {intra:16.6f}".rstrip()


# ---- number_<res>.dat (write_dat_number :247-285) --------------------------------------------------
NUMBER_HEADER = (_a("# Block", 10) + " " + _a("Active_Molecules", 10)).rstrip()


def number_line(block: int, count: int) -> str:
    return f"{block:10d} {count:10d}"


# ---- moves.dat (write_dat_mcmove :291-403) ---------------------------------------------------------
def _enabled(p: Dict[str, float]):
    return [k for k in ("translation", "rotation", "insertion_deletion", "swap", "widom") if p.get(k, 0.0) > 0]


def moves_header(proba: Dict[str, float]) -> str:
    cols = []
    names = {"translation": ("Trans_Acc", "Trans_Trial"), "rotation": ("Rot_Acc", "Rot_Trial"),
             "insertion_deletion": ("Create_Acc", "Create_Trial", "Delete_Acc", "Delete_Trial"),
             "swap": ("Swap_Acc", "Swap_Trial"), "widom": ("Widom_Trial",)}
    for k in _enabled(proba):
        cols += names[k]
    return (_a("Block", 12).strip() + "".join(" " + _a(c, 12) for c in cols)).rstrip()


def moves_line(block: int, counters, proba: Dict[str, float]) -> str:
    """``counters`` = [6][2] (translations, rotations, creations, deletions, swaps, widom) x (trials, successes)."""
    c = np.asarray(counters).reshape(6, 2)
    out = f"{block:12d}"
    idx = {"translation": [0], "rotation": [1], "insertion_deletion": [2, 3], "swap": [4], "widom": [5]}
    for k in _enabled(proba):
        for i in idx[k]:
            out += f" {int(c[i, 1]):12d} {int(c[i, 0]):12d}"
    return out.rstrip()


# ---- widom_<res>.dat (write_dat_widom :191-241, calculate_excess_mu monte_carlo_utils.f90:598-636) ----
WIDOM_HEADER = (_a("# Block", 10) + " " + _a("Excess_Mu_kcalmol", 16) + " " + _a("Total_Mu_kcalmol", 16) + " "
                + _a("Widom_Samples", 16)).rstrip()


def excess_mu(sum_weight: float, samples: int, temperature: float, n_molecules: int, volume: float, lam: float):
    """(mu_ex, mu_tot) in kcal/mol; mu_tot = mu_ideal + mu_ex with mu_ideal = kT ln(rho Lambda^3)."""
    avg = sum_weight / float(samples)
    mu_ex = -KB_KCALMOL * temperature * math.log(avg)
    rho = float(n_molecules) / volume
    mu_ideal = KB_KCALMOL * temperature * math.log(rho * lam ** 3) if rho > 0 else -math.inf
    return mu_ex, mu_ideal + mu_ex


def widom_line(block: int, mu_ex: float, mu_tot: float, samples: int) -> str:
    return f"{block:10d} {mu_ex:16.6f} {mu_tot:16.6f} {samples:12d}".rstrip()


# ---- one directory per walker / isotherm point -------------------------------------------------------
class BlockWriter:
    """Appends the reference's per-block lines for one walker.  ``names`` = residue names of the active types."""

    def __init__(self, out_dir, names: Dict[int, str], proba: Dict[str, float]):
        self.dir = Path(out_dir)
        self.dir.mkdir(parents=True, exist_ok=True)
        self.names, self.proba = names, proba

    def _append(self, name: str, header: str, line: str, first: bool):
        with open(self.dir / name, "w" if first else "a") as f:
            if first:
                f.write(header + "\n")
            f.write(line + "\n")

    def write(self, block: int, record: dict, widom: Dict[int, tuple] | None = None):
        """``record`` = Engine.parse_record(...) of this walker; ``widom`` = {res: (mu_ex, mu_tot, samples)}."""
        first = block == 0
        self._append("energy.dat", ENERGY_HEADER, energy_line(block, record["energy"]), first)
        for r, nm in self.names.items():
            self._append(f"number_{nm}.dat", NUMBER_HEADER, number_line(block, int(record["count"][r])), first)
        self._append("moves.dat", moves_header(self.proba), moves_line(block, record["counters"], self.proba), first)
        if self.proba.get("widom", 0.0) > 0 and widom:
            for r, (mu_ex, mu_tot, ns) in widom.items():
                self._append(f"widom_{self.names[r]}.dat", WIDOM_HEADER, widom_line(block, mu_ex, mu_tot, ns), first)


def write_isotherm(path, fugacity: Iterable[float], summary: dict):
    """Isotherm table (new: the reference runs one point per process): fugacity, <N>, sqrt(var N), <E>, samples."""
    with open(path, "w") as f:
        f.write("#       fugacity           mean_N            std_N           mean_E      samples\n")
        for i, fu in enumerate(fugacity):
            f.write(f"{fu:16.6e} {summary['mean_N'][i]:16.6f} {math.sqrt(max(summary['var_N'][i], 0.0)):16.6f} "
                    f"{summary['mean_E'][i]:16.6f} {int(summary['samples'][i]):12d}\n")


# ---- trajectory.lammpstrj (write_dump_lammpstrj, src/write_utils.f90:38-119) ---------------------------
def _reciprocal(matrix: np.ndarray) -> np.ndarray:
    """box%cell%reciprocal as compute_box_determinant_and_inverse builds it (geometry_utils.f90:148-200): column j of the
    adjugate = cross product of the other two COLUMNS of the matrix, divided by the determinant."""
    m = np.asarray(matrix, dtype=float)
    adj = np.column_stack([np.cross(m[:, 1], m[:, 2]), np.cross(m[:, 2], m[:, 0]), np.cross(m[:, 0], m[:, 1])])
    det = m[:, 0] @ adj[:, 0]
    return adj / det


def wrap_into_box(pos, matrix) -> np.ndarray:
    """wrap_into_box (geometry_utils.f90:105-140): into [-L/2, L/2] per axis (orthorhombic) or fractional [-1/2, 1/2)."""
    m = np.asarray(matrix, dtype=float)
    p = np.array(pos, dtype=float)
    off = m - np.diag(np.diag(m))
    if np.abs(off).max() <= 1e-10:
        for d in range(3):
            p[d] = p[d] - m[d, d] * np.floor(p[d] / m[d, d] + 0.5) if p[d] / m[d, d] >= 0 else p[d] - m[d, d] * np.ceil(p[d] / m[d, d] - 0.5)
        return p
    f = _reciprocal(m) @ p
    f = f - np.where(f >= 0, np.floor(f + 0.5), np.ceil(f - 0.5))       # Fortran nint: halves away from zero
    return m @ f


def lammpstrj_frame(timestep: int, matrix, residues, record_molecules: dict) -> str:
    """One frame.  ``residues`` = list of dict(active, types (1-based ids per site), com [nmol,3], offset [nmol,natom,3]) in
    residue order, giving the static (inactive) residues; active residues take their molecules from ``record_molecules``
    (Engine.parse_record(...)["molecules"]).  CoMs of active molecules are wrapped, atoms of inactive ones are wrapped,
    exactly like the reference does."""
    m = np.asarray(matrix, dtype=float)
    rows = []
    atom_id = 0
    for r, res in enumerate(residues):
        if res["active"]:
            mol = record_molecules.get(r)
            coms, offs = (mol["com"], mol["offset"]) if mol is not None else (np.zeros((0, 3)), np.zeros((0, len(res["types"]), 3)))
        else:
            coms, offs = np.asarray(res["com"]), np.asarray(res["offset"])
        for k in range(len(coms)):
            com = wrap_into_box(coms[k], m) if res["active"] else np.asarray(coms[k], dtype=float)
            for a, t in enumerate(res["types"]):
                atom_id += 1
                pos = com + offs[k][a]
                if not res["active"]:
                    pos = wrap_into_box(pos, m)
                rows.append(f"{atom_id:6d} {int(t):4d} {pos[0]:12.7f} {pos[1]:12.7f} {pos[2]:12.7f}")
    head = ["ITEM: TIMESTEP", f"{timestep:10d}", "ITEM: NUMBER OF ATOMS", f"{atom_id:10d}", "ITEM: BOX BOUNDS pp pp pp"]
    head += [f"{-m[d, d] / 2:15.8f} {m[d, d] / 2:15.8f}" for d in range(3)]
    head.append("ITEM: ATOMS id type x y z")
    return "\n".join(head + rows) + "\n"


# ---- topology.data (write_topology_data, src/write_utils.f90:406-636) -------------------------------
def _ld_int(n: int) -> str:
    """gfortran list-directed output of a default integer: right-justified in 12 columns (the record's leading blank included)."""
    return f"{int(n):12d}"


def topology_data(matrix, lo, masses, residues, record_molecules: dict, connect: dict | None = None, n_types: dict | None = None) -> str:
    """The LAMMPS data file (atom_style full) the reference rewrites after every block.

    ``residues`` = list of dict(active, types (1-based atom-type id per site), charges, com [nmol,3], offset [nmol,natom,3]) in
    residue order; active residues take their molecules from ``record_molecules`` (Engine.parse_record(...)["molecules"]).
    ``connect`` = optional {"bonds": [per residue: list of (type, i, j)], "angles": [... (type, i, j, k)], "dihedrals": [...],
    "impropers": [...]} with 1-based atom indices inside the molecule (connect%bonds(res, k, 1:3) etc.); ``n_types`` = declared
    numbers of bond / angle / dihedral / improper types.  Atoms of INACTIVE residues are wrapped into the box, active molecules
    are left whole across the boundary (:538-541).  Fixed-format fields (F15.8, I5/F12.6, I6/I6/I4/F12.8/F12.7) are reproduced
    character by character; the list-directed records (`write(unit,*)`) follow gfortran's conventions (leading blank, integers
    in 12 columns), which no test of the reference pins and which LAMMPS reads whitespace-insensitively."""
    m = np.asarray(matrix, dtype=float)
    lo = np.asarray(lo, dtype=float)
    connect = connect or {}
    n_types = n_types or {}
    kinds = ("bonds", "angles", "dihedrals", "impropers")
    nmol = []
    for r, res in enumerate(residues):
        if res["active"]:
            mol = record_molecules.get(r)
            nmol.append(0 if mol is None else len(mol["com"]))
        else:
            nmol.append(len(res["com"]))
    per_res = {k: [len(c) for c in connect.get(k, [[] for _ in residues])] for k in kinds}
    totals = {k: sum(n * per_res[k][r] for r, n in enumerate(nmol)) for k in kinds}
    n_atoms = sum(n * len(res["types"]) for n, res in zip(nmol, residues))
    out = [" ! LAMMPS data file (atom_style full)",
           _ld_int(n_atoms) + "  atoms", _ld_int(len(masses)) + "  atom types"]
    for k, label in zip(kinds, ("bond", "angle", "dihedral", "improper")):
        out += [_ld_int(totals[k]) + f"  {label}s", _ld_int(n_types.get(k, 0)) + f"  {label} types"]
    out.append("")
    hi = lo + np.array([m[0, 0], m[1, 1], m[2, 2]])
    for d, name in enumerate("xyz"):
        out.append(f"{lo[d]:15.8f} {hi[d]:15.8f} {name}lo {name}hi")
    tilt = np.array([m[1, 0], m[2, 0], m[2, 1]])                       # readers_utils.f90:256-258: rows of the matrix are a, b, c
    if np.abs(np.array([m[0, 1], m[0, 2], m[1, 0], m[1, 2], m[2, 0], m[2, 1]])).max() > 1e-10:
        out += [f"{tilt[0]:15.8f} {tilt[1]:15.8f} {tilt[2]:15.8f} ", "xy xz yz"]
    out += ["", " Masses", ""]
    out += [f"{t + 1:5d} {float(mass):12.6f}" for t, mass in enumerate(masses)]
    out += ["", " Atoms", ""]
    atom_id = mol_id = 0
    for r, res in enumerate(residues):
        if res["active"]:
            mol = record_molecules.get(r)
            coms, offs = (mol["com"], mol["offset"]) if mol is not None else (np.zeros((0, 3)), np.zeros((0, len(res["types"]), 3)))
        else:
            coms, offs = np.asarray(res["com"]), np.asarray(res["offset"])
        for k in range(len(coms)):
            mol_id += 1
            for a, t in enumerate(res["types"]):
                atom_id += 1
                pos = np.asarray(coms[k], dtype=float) + offs[k][a]
                if not res["active"]:
                    pos = wrap_into_box(pos, m)
                out.append(f"{atom_id:6d} {mol_id:6d} {int(t):4d} {float(res['charges'][a]):12.8f} {pos[0]:12.7f} {pos[1]:12.7f} {pos[2]:12.7f}")
    for k, title in zip(kinds, ("Bonds", "Angles", "Dihedrals", "Impropers")):
        if totals[k] == 0:
            continue
        out += ["", " " + title, ""]
        cpt, cpt_atom = 1, 0
        for r, res in enumerate(residues):
            for _ in range(nmol[r]):
                for entry in connect[k][r]:
                    out.append("".join(_ld_int(v) for v in (cpt, entry[0], *[cpt_atom + i for i in entry[1:]])))
                    cpt += 1
                cpt_atom += len(res["types"])
    return "\n".join(out) + "\n"
