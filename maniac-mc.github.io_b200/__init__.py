"""B200-native energy engine for MANIAC-MC's per-trial-move energy change.

The directory name carries the reference's repository name and is not a valid Python
identifier; ``maniac_b200.py`` at the repository root registers this package under the
importable name ``maniac_b200``.

Modules
  inputs    readers for .maniac / LAMMPS .data / pair_coeff files -> :class:`System`
  snapshot  save / load a :class:`System` as a single .npz
  capi      ctypes binding of include/maniac_gpu.h (libmaniac_gpu.so, CUDA sm_100a)
  engine    host-side mirror of the reference's energy interface on top of the C ABI
  hostmc    host-driven Monte Carlo drivers (the Fortran drivers' role) over the C ABI
  csrc/     CUDA kernels + C ABI;  fortran/  ISO_C_BINDING shim for the MANIAC sources
"""
from .inputs import Residue, System, load_system  # noqa: F401
from .snapshot import load_snapshot, save_snapshot  # noqa: F401

__all__ = ["Residue", "System", "load_system", "load_snapshot", "save_snapshot"]
