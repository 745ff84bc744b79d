"""Single-file (.npz) snapshots of a :class:`System` (host logic, NumPy only)."""
from __future__ import annotations

import json
from pathlib import Path

import numpy as np

from .inputs import Residue, System

_SCALARS = ["ntypes", "temperature", "ewald_tolerance", "real_space_cutoff", "translation_step",
            "rotation_step_angle", "recalibrate_moves", "p_translation", "p_rotation", "p_swap",
            "p_insertion_deletion", "p_widom", "nb_block", "nb_step", "seed"]


def save_snapshot(path, system: System, extra: dict | None = None) -> None:
    meta = {k: getattr(system, k) for k in _SCALARS}
    meta["pair_coeff"] = [list(p) for p in system.pair_coeff]
    meta["residues"] = [dict(name=r.name, active=bool(r.active), natom=int(r.natom), site_types=list(map(int, r.site_types)),
                             site_names=list(r.site_names), fugacity=float(r.fugacity),
                             chemical_potential=float(r.chemical_potential), mass=float(r.mass)) for r in system.residues]
    meta["extra"] = extra or {}
    arrays = dict(matrix=np.asarray(system.matrix, dtype=np.float64), lo=np.asarray(system.lo, dtype=np.float64))
    for i, r in enumerate(system.residues):
        arrays[f"r{i}_types"] = np.asarray(r.types, dtype=np.int32)
        arrays[f"r{i}_charges"] = np.asarray(r.charges, dtype=np.float64)
        arrays[f"r{i}_com"] = np.asarray(r.com, dtype=np.float64).reshape(-1, 3)
        arrays[f"r{i}_offset"] = np.asarray(r.offset, dtype=np.float64).reshape(-1, r.natom, 3)
    np.savez_compressed(path, meta=np.frombuffer(json.dumps(meta).encode(), dtype=np.uint8), **arrays)


def load_snapshot(path, with_extra: bool = False):
    z = np.load(Path(path))
    meta = json.loads(bytes(z["meta"]).decode())
    residues = []
    for i, rm in enumerate(meta["residues"]):
        r = Residue(name=rm["name"], active=rm["active"], natom=rm["natom"], site_types=rm["site_types"],
                    site_names=rm["site_names"], fugacity=rm["fugacity"], chemical_potential=rm["chemical_potential"])
        r.mass = rm["mass"]
        r.types = z[f"r{i}_types"].astype(np.int32)
        r.charges = z[f"r{i}_charges"].astype(np.float64)
        r.com = z[f"r{i}_com"].astype(np.float64)
        r.offset = z[f"r{i}_offset"].astype(np.float64)
        residues.append(r)
    s = System(matrix=z["matrix"].astype(np.float64), lo=z["lo"].astype(np.float64), residues=residues,
               ntypes=meta["ntypes"], pair_coeff=[tuple(p) for p in meta["pair_coeff"]])
    for k in _SCALARS:
        setattr(s, k, meta[k])
    s.pair_coeff = [(int(a), int(b), float(c), float(d)) for a, b, c, d in meta["pair_coeff"]]
    return (s, meta.get("extra", {})) if with_extra else s
