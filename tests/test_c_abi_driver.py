"""The drop-in boundary exercised from plain C (tests/c_abi_driver.c): the calls, argument conventions and
sequence of the Fortran shim (integration/fortran/*.f90) -- `mgpu_system` by pointer, scalars by value, 1-based ids
converted at the call, geometry as com(3) + off(3, MGPU_MAX_SITES) -- checked number by number against the CPU oracle
playing the reference's drivers (src/monte_carlo_utils.f90:300-423)."""
import struct
import subprocess
from pathlib import Path

import numpy as np
import pytest

from oracle.oracle import KIND_CREATE, KIND_DELETE, KIND_MOVE, Oracle

ROOT = Path(__file__).resolve().parents[1]
PKG = ROOT / "maniac-mc.github.io_b200"
MAX_SITES = 16


def build_driver(tmp_path):
    exe = tmp_path / "c_abi_driver"
    cmd = ["gcc", "-O1", "-Wall", "-Wextra", "-Werror", "-std=c99", "-I", str(ROOT / "include"), str(ROOT / "tests" / "c_abi_driver.c"),
           "-L", str(PKG), "-lmaniac_gpu", f"-Wl,-rpath,{PKG}", "-o", str(exe)]
    subprocess.run(cmd, check=True, capture_output=True, text=True)
    return exe


def write_input(path, s, capacity, trials):
    from maniac_b200.engine import lj_table
    eps, sig = lj_table(s)
    with open(path, "wb") as f:
        f.write(struct.pack("i", 0x4D475055))
        f.write(np.asarray(s.matrix, dtype=np.float64).tobytes())
        f.write(np.asarray(s.lo, dtype=np.float64).tobytes())
        f.write(struct.pack("ii", len(s.residues), s.ntypes))
        for r in s.residues:
            cap = capacity if r.active else r.nmol
            f.write(struct.pack("iiii", r.natom, int(r.active), r.nmol, cap))
            f.write(struct.pack("ddd", float(r.mass), float(r.fugacity), float(r.chemical_potential)))
            f.write(np.asarray(r.charges, dtype=np.float64).tobytes())
            f.write(np.asarray(r.types, dtype=np.int32).tobytes())
            f.write(np.asarray(r.com, dtype=np.float64).reshape(-1).tobytes())
            f.write(np.asarray(r.offset, dtype=np.float64).reshape(-1).tobytes())
        f.write(np.asarray(eps, dtype=np.float64).tobytes())
        f.write(np.asarray(sig, dtype=np.float64).tobytes())
        f.write(struct.pack("10d", s.temperature, s.ewald_tolerance, s.real_space_cutoff, s.translation_step, s.rotation_step_angle,
                            s.p_translation, s.p_rotation, s.p_swap, s.p_insertion_deletion, s.p_widom))
        f.write(struct.pack("i", len(trials)))
        for t in trials:
            f.write(struct.pack("iiii", t["kind"], t["res"] + 1, t["mol"] + 1, int(t["accept"])))      # 1-based on the driver's side
            f.write(np.asarray(t["com"], dtype=np.float64).tobytes())
            off = np.zeros((MAX_SITES, 3))
            off[:len(t["off"])] = t["off"]
            f.write(off.tobytes())


def test_driver_builds_links_and_fails_loudly_without_a_gpu(tmp_path, load):
    """gcc -Wall -Wextra -Werror against include/maniac_gpu.h, linked with the shared library: the header is valid
    C99 and every call the driver makes resolves.  On a box without a CUDA device mgpu_init must fail with a message
    (there is no CPU fallback) and the driver aborts like abort_run."""
    import torch
    exe = build_driver(tmp_path)
    if torch.cuda.is_available():
        pytest.skip("GPU present: the loud-failure leg only applies to CPU boxes")
    s = load("two_atoms")
    write_input(tmp_path / "in.bin", s, 4, [])
    p = subprocess.run([str(exe), str(tmp_path / "in.bin")], capture_output=True, text=True)
    assert p.returncode == 1
    assert p.stdout.startswith("ABORT mgpu_init") and "no CUDA device" in p.stdout


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["zif8_h2o_gcmc", "methanol"])
def test_c_driver_follows_the_oracle(name, tmp_path, load):
    s = load(name)
    cap = 32
    rng = np.random.default_rng(17)
    res = [i for i, r in enumerate(s.residues) if r.active][0]
    o = Oracle(s, capacity=cap)
    e0 = o.update_system_energy()
    trials, expect = [], []
    for it in range(8):                                      # translations with a little rotation, alternately accepted
        n = o.count(res)
        mol = int(rng.integers(n))
        com, off = o.get_molecule(res, mol)
        new_com = com + rng.uniform(-0.4, 0.4, 3)
        th = rng.uniform(-0.2, 0.2)
        c, sn = np.cos(th), np.sin(th)
        new_off = off @ np.array([[c, -sn, 0.0], [sn, c, 0.0], [0.0, 0.0, 1.0]]).T
        o.save_fourier(res, mol)
        old_r = o.compute_old_energy(res, mol, KIND_MOVE)
        o.set_molecule(res, mol, new_com, new_off)
        new_r = o.compute_new_energy(res, mol, KIND_MOVE)
        acc = it % 2 == 0
        if acc:
            o.update_system_energy()
        else:
            o.set_molecule(res, mol, com, off)
            o.restore_fourier(res, mol)
        trials.append(dict(kind=KIND_MOVE, res=res, mol=mol, accept=acc, com=new_com, off=new_off))
        expect.append((old_r.copy(), new_r.copy(), o.energy().copy(), o.count(res)))
    # a creation through the oracle's own driver (forced acceptance), then a deletion of molecule 1
    n = o.count(res)
    o.set_chemical_potential(res, 50.0)
    o.seed(99)
    t = o.attempt_creation_move(res, n)
    com_new, off_new = o.get_molecule(res, n)
    trials.append(dict(kind=KIND_CREATE, res=res, mol=n, accept=bool(t.accepted), com=com_new, off=off_new))
    expect.append((np.array(t.e_old[:]), np.array(t.e_new[:]), o.energy().copy(), o.count(res)))
    o.set_chemical_potential(res, -50.0)
    t = o.attempt_deletion_move(res, 0)
    assert t.accepted == 1
    trials.append(dict(kind=KIND_DELETE, res=res, mol=0, accept=True, com=np.zeros(3), off=np.zeros((1, 3))))
    expect.append((np.array(t.e_old[:]), np.array(t.e_new[:]), o.energy().copy(), o.count(res)))
    e_end = o.update_system_energy()

    write_input(tmp_path / "in.bin", s, cap, trials)
    exe = build_driver(tmp_path)
    p = subprocess.run([str(exe), str(tmp_path / "in.bin")], capture_output=True, text=True, timeout=600)
    assert p.returncode == 0, p.stdout + p.stderr
    lines = p.stdout.strip().splitlines()
    assert lines[-1] == "OK"
    got = {}
    for ln in lines:
        w = ln.split()
        if w[0] in ("TOTAL", "OLD", "NEW", "RUN"):
            got[(w[0], int(w[1]))] = np.array([float(x) for x in w[2:8]])
        elif w[0] == "COUNT":
            got[("COUNT", int(w[1]))] = int(w[2])
        elif w[0] == "EWALD":
            ew = o.ewald()
            assert float(w[1]) == ew["alpha"] and [int(x) for x in w[2:5]] == list(ew["kmax"]) and int(w[5]) == ew["nk"]
        elif w[0] == "BADRES":
            assert w[1] == "1" and "residue" in ln

    def close(a, b, rel):
        return np.all(np.abs(a - b) <= rel * np.maximum(1.0, np.abs(b)))
    assert close(got[("TOTAL", 0)], e0, 1e-10)
    assert close(got[("TOTAL", 1)], e_end, 1e-10)
    for i, (old_r, new_r, run_r, cnt) in enumerate(expect):
        assert close(got[("OLD", i)][:2], old_r[:2], 1e-10) and close(got[("OLD", i)], old_r, 1e-9), (i, got[("OLD", i)], old_r)
        assert close(got[("NEW", i)][:2], new_r[:2], 1e-10) and close(got[("NEW", i)], new_r, 1e-9), (i, got[("NEW", i)], new_r)
        d_r, d_g = new_r[5] - old_r[5], got[("NEW", i)][5] - got[("OLD", i)][5]
        assert abs(d_r - d_g) <= 1e-9 * max(1.0, abs(d_r))
        assert close(got[("RUN", i)], run_r, 1e-9)
        assert got[("COUNT", i)] == cnt
