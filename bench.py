#!/usr/bin/env python
"""bench.py -- MC trial moves/s (+ Widom insertions/s) of the MANIAC energy path on B200.

Contract (driver):  python bench.py --gpus N --steps K --warmup W     (N > 1 via torchrun)
prints ONE JSON line on rank 0.

  step      one launch of the device-resident sweep kernel: INNER Monte Carlo steps for every
            walker of this rank (zero host round trips inside the launch)
  workload  BASELINE.json configs[1]: ZIF-8 2x2x2 + TIP4P water GCMC, move mix 0.4/0.4/0.2
            (translate / rotate / insert-delete), EVERY WALKER PINNED AT 64 WATERS: before anything is
            timed a pool of walkers is run under a feedback on ln(fugacity) until <N> sits at the target
            and the configurations have relaxed, then the fugacity is frozen and the pool is cloned
            (records through host memory, fresh RNG streams) into the timed walkers.  The loading at the
            start and at the end of the timed region is in the line (`stationarity`): moves/s does not
            depend on --steps
  value     trial moves / s over all ranks (weak scaling: --walkers per GPU), state resident in HBM,
            device-timed (CUDA events on the launching stream), max over ranks; the L2 is flushed
            between timed launches
  e2e       the same metric through the block-level C-ABI call with HOST buffers (mgpu_block): every
            step the walkers' whole state travels pinned host -> device, the MC steps run, and the
            updated state travels back, in slices on several streams (copies under compute); wall clock
  roofline  the sweep kernel against the FP64-pipe peak measured on this GPU in this run (DFMA loop);
            achieved = convention-C1 FLOPs of the pairs the launch EVALUATED / event time;
            fp64_pipe_pct / issue_active_pct = executed-instruction utilisation from the committed ncu
            capture of the same kernel; roofline_k2 = the k-space traffic against a measured L2 peak
  grid      SURVEY 8d M1: loadings N in {0,16,64,128} x walkers W in {64,1024,4096,16384}, each pinned
            and relaxed the same way (N = 1 only)
  strong_scaling  SURVEY 8d M3 / BASELINE configs[3]: a FIXED isotherm of 64 fugacity points x 64
            replicas = 4096 walkers split over the N GPUs (few walkers per GPU -> team shape of the
            sweep kernel: four warps per walker); per-point sums reduced once with NCCL
  widom     BASELINE configs[2]: 10^6 CO2 test insertions in empty ZIF-8 in 16 blocks, mu_ex +- sigma;
            sum w / n reduced over ranks through mgpu_reduce_averages
  mixture   BASELINE configs[4]: CO2/N2 with swaps in the 17 664-atom triclinic supercell
  host_driven / single_walker_dropin   the Fortran drivers' role over the batched / single-trial ABI
  cpu_baseline   the CPU oracle (C restatement of the reference's serial Fortran, gcc -O3) on all host
            cores, bounded sample of the same workload state (same loading, same fugacity)
  --impl reference   times that CPU path as the main line (no Fortran compiler in the image, so the
            reference itself cannot be built; kind = "port")
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

INNER_DEFAULT = 256
CAPACITY = 320           # molecules per walker (NB_MAX_MOLECULE analogue); loadings up to 128 pinned
POOL = 2368              # walkers of the equilibration pool: one full wave of the warp-per-walker shape
FLOP_GEOM, FLOP_LJ, FLOP_COUL, FLOP_SINCOS = 29.0, 8.0, 69.0, 64.0      # SURVEY.md 8d, convention C1
FUG_FILE = ROOT / "tests" / "golden" / "bench_fugacity.json"              # ln f that pins each loading (written by a GPU run)
STATES_FILE = ROOT / "tests" / "golden" / "bench_states.npz"              # relaxed configurations of 32 GPU walkers at the headline loading (same run)


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=8)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--walkers", type=int, default=4736, help="walkers per GPU (weak scaling); 4736 = 2 x 148 SMs x 16 warps")
    ap.add_argument("--inner", type=int, default=INNER_DEFAULT, help="MC steps per walker per launch")
    ap.add_argument("--loading", type=int, default=64, help="waters per walker the headline is pinned at")
    ap.add_argument("--equil", type=int, default=80, help="launches of the pinning / relaxation stage")
    ap.add_argument("--grid", type=int, default=-1, help="SURVEY 8d M1 grid: 1 on, 0 off, -1 = on for N = 1")
    ap.add_argument("--strong-walkers", type=int, default=4096, help="walkers of the fixed isotherm (64 points x 64 replicas), split over the GPUs (0 = skip)")
    ap.add_argument("--e2e-walkers", type=int, default=2368, help="walkers of the host-driven leg (one warp-per-trial wave)")
    ap.add_argument("--e2e-steps", type=int, default=20)
    ap.add_argument("--widom", type=int, default=1_000_000, help="insertions in the Widom batch (0 = skip)")
    ap.add_argument("--mixture-walkers", type=int, default=2368, help="walkers of the configs[4] leg (0 = skip)")
    ap.add_argument("--mixture-steps", type=int, default=32)
    ap.add_argument("--cpu-seconds", type=float, default=12.0)
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--phase-sync", type=int, default=-1, help="experiment: MGPU_OPT_PHASE_SYNC value")
    ap.add_argument("--quick", action="store_true", help="kernel-variant experiments: pinning + the timed sweep only, prints a short line")
    ap.add_argument("--write-fugacity", action="store_true", help="store the pinning fugacities found by this run in tests/golden/bench_fugacity.json")
    return ap.parse_args()


def workload(loading):
    import maniac_b200  # noqa: F401
    from maniac_b200.snapshot import load_snapshot
    from maniac_b200.workloads import load_pore
    s = load_snapshot(ROOT / "tests" / "golden" / "zif8_h2o_gcmc.npz")
    s = load_pore(s, 0, max(1, loading), seed=12345)
    s.p_translation, s.p_rotation, s.p_insertion_deletion, s.p_swap, s.p_widom = 0.4, 0.4, 0.2, 0.0, 0.0
    return s


def stored_states(loading):
    """[(com[n,3], offset[n,natom,3])] relaxed at `loading` (written by a GPU run with --write-fugacity), or None"""
    try:
        z = np.load(STATES_FILE)
        if int(z["loading"]) != loading:
            return None
        return [(z[f"com{i}"], z[f"off{i}"]) for i in range(int(z["n"]))]
    except Exception:
        return None


def stored_lnf():
    try:
        return {int(k): float(v) for k, v in json.loads(FUG_FILE.read_text())["ln_fugacity"].items()}
    except Exception:
        return {}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.rows, self.proc, self.idx = [], None, gpu_index
        self.t_begin = self.t_end = None      # perf_counter bounds of the timed region (mark_begin / mark_end)

    def mark_begin(self):
        self.t_begin = time.perf_counter()

    def mark_end(self):
        self.t_end = time.perf_counter()

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "50", "-i", str(self.idx)], stdout=subprocess.PIPE, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), [x.strip() for x in line.split(",")]))

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        rows = self.rows
        if self.t_begin is not None and self.t_end is not None:
            inside = [r for r in rows if self.t_begin <= r[0] <= self.t_end + 0.05]
            rows = inside if inside else rows
        for _, r in rows:
            try:
                sm.append(float(r[1]))
                mx.append(float(r[2]))
                for name, v in zip(["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"], r[4:8]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                pass
        busy = [x for x in sm if x > 0.5 * max(sm)] if sm else []
        return {"sm_mhz": float(np.median(busy)) if busy else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def kspace_flops(counters_delta, na, kmax, nk):
    """Algorithmic FLOPs of K2: moves 2 phase-table sets + nk(30 na + 7);
    creations / deletions 1 set + nk(14 na + 7)."""
    tr = counters_delta[0, 0] + counters_delta[1, 0]
    cr = counters_delta[2, 0] - counters_delta[2, 1]
    de = counters_delta[3, 0] - counters_delta[3, 1]
    tab = sum(k + 1 for k in kmax) * na * FLOP_SINCOS
    return tr * (2 * tab + nk * (30 * na + 7)) + (cr + de) * (tab + nk * (14 * na + 7))


def kspace_bytes(counters_delta, nk):
    """Algorithmic bytes of K2 unique to a walker: S(k) read by every trial that reaches k-space
    (16 nk), S_trial written back when the trial is accepted (16 nk)."""
    trials = counters_delta[0, 0] + counters_delta[1, 0] + (counters_delta[2, 0] - counters_delta[2, 1]) + (counters_delta[3, 0] - counters_delta[3, 1])
    commits = counters_delta[0, 1] + counters_delta[1, 1] + counters_delta[2, 1] + counters_delta[3, 1]
    return 16.0 * nk * trials + 16.0 * nk * commits


def cpu_sample(system, seconds, threads, capacity, target, lnf, states=None, chunk=256):
    """Aggregate moves/s of the CPU oracle on `threads` host threads, one walker each (-O3 build of the oracle source),
    under the same loading controller as the GPU legs (ln f moved between chunks of `chunk` MC steps from the mean loading
    of the walkers).  states: optional list of (com[n,3], offset[n,natom,3]) -- relaxed configurations handed over from
    the GPU walkers, so that both sides sample the same state."""
    os.environ["MANIAC_ORACLE_VARIANT"] = "o3"
    from oracle.oracle import Oracle
    oracles = []
    for t in range(threads):
        o = Oracle(system, capacity=capacity)
        if states:
            com, off = states[t % len(states)]
            o.set_count(0, len(com))
            for m in range(len(com)):
                o.set_molecule(0, m, com[m], off[m])
        o.update_system_energy()
        o.seed(12345 + 104729 * t)
        oracles.append(o)
    ctl = LoadingController(target, -14.0 if lnf is None else lnf)
    ctl.prev = float(np.mean([o.count(0) for o in oracles]))
    for o in oracles:
        o.set_fugacity(0, float(np.exp(ctl.lnf)))

    def run(i):
        oracles[i].monte_carlo_steps(chunk, trace=False)
    done, t0 = 0, time.perf_counter()
    while True:
        ths = [threading.Thread(target=run, args=(i,)) for i in range(threads)]
        for th in ths:
            th.start()
        for th in ths:
            th.join()
        done += chunk * threads
        ctl.update(float(np.mean([o.count(0) for o in oracles])))
        for o in oracles:
            o.set_fugacity(0, float(np.exp(ctl.lnf)))
        if time.perf_counter() - t0 >= seconds:
            break
    dt = time.perf_counter() - t0
    return done / dt, done // threads, dt, ctl.hist[-1]


# ---------------------------------------------------------------------------------------------------------
# pinning a loading: feedback on ln(fugacity), then frozen
# ---------------------------------------------------------------------------------------------------------
class LoadingController:
    """PD feedback on ln(fugacity), one value for all walkers of a leg, applied BETWEEN launches: holds <N> at the
    target.  It has to stay on: TIP4P water in the hydrophobic ZIF-8 cage adsorbs cooperatively (the pinning fugacity
    FALLS with loading: ln f = -13.5 / -14.7 / -15.0 at N = 16 / 64 / 128), so at a frozen fugacity an intermediate
    loading runs away to the empty or to the filled branch.  target = 0: a vanishing fugacity keeps the walkers empty."""

    def __init__(self, target, lnf0):
        self.target, self.scale = target, float(max(target, 4))
        self.lnf = float(lnf0) if target > 0 else -60.0
        self.prev = None
        self.hist, self.lnfs = [], []

    def apply(self, eng, n_walkers):
        eng.set_fugacities(0, np.full(n_walkers, np.exp(self.lnf)), 0)

    def update(self, nbar):
        self.hist.append(float(nbar))
        self.lnfs.append(self.lnf)
        if self.target > 0:
            d = 0.0 if self.prev is None else nbar - self.prev
            e = (nbar - self.target) / self.scale
            # P + D only: an integral term (tried: 0.08 sum e) made the loading overshoot by 25 % within ten launches on the
            # unstable branch; the small steady offset P + D leaves (<N> a few per cent off the target) is reported instead
            self.lnf -= 0.5 * e + 4.0 * d / self.scale
        self.prev = float(nbar)

    def step(self, eng, n_walkers):
        """after a launch: read <N>, move ln f, hand it to the walkers"""
        self.update(float(eng.counts(0, 0, n_walkers).mean()))
        self.apply(eng, n_walkers)


def pin_loading(eng, n_walkers, target, launches, inner, lnf0):
    """Run walkers [0, n_walkers) of `eng` for `launches` launches of `inner` MC steps under the controller; returns it."""
    ctl = LoadingController(target, lnf0)
    ctl.prev = float(eng.counts(0, 0, n_walkers).mean())
    ctl.apply(eng, n_walkers)
    for it in range(launches):
        eng.sweep(inner, 0, n_walkers)
        ctl.step(eng, n_walkers)
    return ctl


def clone_records(eng, blob_pool, off_pool, n_pool, n_target):
    """Tile the pool's records into a pinned buffer for n_target walkers (record i <- pool record i mod n_pool)."""
    reps = (n_target + n_pool - 1) // n_pool
    lens = np.diff(off_pool)
    lens_t = np.tile(lens, reps)[:n_target]
    off = np.zeros(n_target + 1, dtype=np.int64)
    np.cumsum(lens_t, out=off[1:])
    buf = eng.host_buffer(int(off[-1]))
    full = int(off_pool[-1])
    for r in range(reps):
        a = r * full
        b = min(a + full, int(off[-1]))
        buf[a:b] = blob_pool[:b - a]
    return buf, off


def main():
    a = parse()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    cores = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    lnf_known = stored_lnf()
    cfg = {"workload": "ZIF-8 2x2x2 (2208 atoms, 34.02 A cubic) + TIP4P H2O GCMC, BASELINE configs[1] / SURVEY 8d M1; "
                       f"every walker pinned at {a.loading} waters (fugacity found by feedback during an untimed relaxation "
                       "stage, then frozen), mix 0.4/0.4/0.2 translate/rotate/insert-delete",
           "walkers_per_gpu": a.walkers, "mc_steps_per_launch": a.inner, "r_cut": 17.0, "nkvec": 297,
           "l2": "flushed between timed launches (256 MB write); per-walker state (~25 KB x walkers) is of the order of the 126 MB L2"}
    s = workload(a.loading)

    # ------------------------------------------------------------------ reference arm (CPU)
    if a.impl == "reference":
        if rank != 0:
            return
        lnf = lnf_known.get(a.loading)
        states = stored_states(a.loading)
        per_s = []
        for i in range(a.warmup + a.steps):
            v, n, dt, nbar = cpu_sample(s, max(2.0, a.cpu_seconds / max(1, a.steps)), cores, 512, a.loading, lnf, states)
            if i >= a.warmup:
                per_s.append((v, n, dt, nbar))
        v = float(np.mean([x[0] for x in per_s]))
        line = {"impl": "reference", "metric": "mc_trial_moves_per_s", "value": v, "unit": "moves/s", "n_gpus": a.gpus,
                "steps": a.steps, "warmup": a.warmup, "ms_per_step": 1e3 * float(np.mean([x[2] for x in per_s])),
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": cfg,
                "cpu_baseline": {"value": v, "unit": "moves/s", "cores": cores, "kind": "port",
                                 "sample": f"{cores} independent walkers (one per host thread) x {per_s[0][1]} MC steps per step from "
                                           + ("relaxed configurations of GPU walkers (tests/golden/bench_states.npz) " if states else f"the unrelaxed {a.loading}-water configuration ")
                                           + f"under the same loading controller as the GPU legs (mean loading at the end {per_s[-1][3]:.1f}); C restatement of the "
                                           "reference's serial Fortran algorithm with a compact per-type-pair LJ table, gcc -O3 -ffp-contract=off, "
                                           "no -march=native (no Fortran compiler in the image: the reference itself cannot be built)"},
                "e2e": {"value": v, "unit": "moves/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        print(json.dumps(line))
        return

    # ------------------------------------------------------------------ B200 arm
    import torch
    import torch.distributed as dist
    from maniac_b200.engine import OPT_HOST_CACHE, OPT_PHASE_SYNC, Engine
    from maniac_b200.hostmc import HostMonteCarlo
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the energy path has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(*vals):
        if world == 1:
            return vals if len(vals) > 1 else vals[0]
        t = torch.tensor(list(vals), device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        out = [float(x) for x in t]
        return out if len(out) > 1 else out[0]

    flush_buf = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")

    def flush_l2():
        flush_buf.zero_()
        torch.cuda.synchronize()

    do_grid = (a.grid == 1) or (a.grid < 0 and world == 1 and not a.quick)
    W = a.walkers
    Wmax = max(W, 16384 if do_grid else 0, POOL)
    n_pool = min(POOL, Wmax)
    eng = Engine(s, n_walkers=Wmax, capacity=CAPACITY, device=local)
    if a.phase_sync >= 0:
        eng.set_option(OPT_PHASE_SYNC, a.phase_sync)
    peak_tf, _ = eng.measure_fp64_peak()
    l2_peak = eng.measure_l2_peak()
    ew = eng.ewald()
    beta = eng.thermo(0)["beta"]
    na = 4

    def prepare(loading, n_target, seed):
        """Pin `loading` on the pool (walkers [0, n_pool)), then clone the relaxed pool into walkers [0, n_target)."""
        ctl = pin_loading(eng, n_pool, loading, a.equil, a.inner, lnf_known.get(loading, -14.0))
        blob_pool = eng.host_buffer(n_pool * (72 + 2 * ew["nk"] + (int(1.3 * max(loading, 8)) + 48) * (3 + 3 * na + 2)))
        off_pool = eng.save_walkers(blob_pool, 0, n_pool)
        buf, off = clone_records(eng, blob_pool, off_pool, n_pool, n_target)
        eng.load_walkers(buf, off, 0)
        ctl.apply(eng, n_target)
        eng.seed(seed)
        eng.reset_averages()
        return ctl

    sampler = ClockSampler(local)
    sampler.start()
    t_prep = time.perf_counter()
    ctl_head = prepare(a.loading, Wmax if do_grid else W, 12345 + 7919 * rank)
    t_prep = time.perf_counter() - t_prep
    hist_head = list(ctl_head.hist)
    found_lnf = {a.loading: float(np.mean(ctl_head.lnfs[-max(1, a.equil // 3):]))}
    lnf_head = found_lnf[a.loading]

    def timed_sweeps(ctl, n_walkers, n_launch, n_warm):
        for _ in range(n_warm):
            eng.sweep(a.inner, 0, n_walkers)
            ctl.step(eng, n_walkers)
        eng.reset_pair_counts()
        eng.timing_reset()
        for _ in range(n_launch):
            flush_l2()
            eng.sweep(a.inner, 0, n_walkers)
            ctl.step(eng, n_walkers)
        ms, launches = eng.timing("sweep")
        return ms * 1e-3, int(launches)

    # ---- headline: W walkers pinned at the loading --------------------------------------------------
    for _ in range(a.warmup):
        eng.sweep(a.inner, 0, W)
        ctl_head.step(eng, W)
    stride = max(1, W // 256)
    c0 = np.array([eng.counters(w) for w in range(0, W, stride)]).sum(axis=0)
    n_sampled = len(range(0, W, stride))
    n_begin = float(eng.counts(0, 0, W).mean())
    eng.reset_pair_counts()
    eng.timing_reset()
    barrier()
    sampler.mark_begin()
    t0 = time.perf_counter()
    for _ in range(a.steps):
        flush_l2()
        eng.sweep(a.inner, 0, W)
        ctl_head.step(eng, W)                   # between launches: 19 KB of counts D2H, 38 KB of mu H2D (outside the event-timed kernel)
    barrier()
    wall = time.perf_counter() - t0
    sampler.mark_end()
    clocks = sampler.stop()
    ms_total, launches = eng.timing("sweep")
    pc = eng.pair_counts()
    c1 = np.array([eng.counters(w) for w in range(0, W, stride)]).sum(axis=0)
    n_end = float(eng.counts(0, 0, W).mean())
    dcount = (c1 - c0) * (W / n_sampled)
    t_dev, wall = max_over_ranks(ms_total * 1e-3, wall)
    moves = float(W) * a.inner * a.steps * world
    value = moves / t_dev

    if a.quick:
        if rank == 0:
            q_flops = FLOP_GEOM * pc["pairs"] + FLOP_LJ * pc["lj"] + FLOP_COUL * pc["coulomb"] + kspace_flops(dcount, na, ew["kmax"], ew["nk"])
            print(json.dumps({"quick": True, "lib": os.environ.get("MANIAC_GPU_LIB", "default"), "walkers": W, "phase_sync": a.phase_sync, "moves_per_s": value,
                              "frac_c1": q_flops / (ms_total * 1e-3) / 1e12 / peak_tf,
                              "ms_per_step": 1e3 * t_dev / a.steps, "loading": [n_begin, n_end], "ln_fugacity": ctl_head.lnf,
                              "clocks": clocks}))
        eng.close()
        return
    flops = FLOP_GEOM * pc["pairs"] + FLOP_LJ * pc["lj"] + FLOP_COUL * pc["coulomb"] + kspace_flops(dcount, na, ew["kmax"], ew["nk"])
    ach_tf = flops / (ms_total * 1e-3) / 1e12
    peaks = {}
    try:
        peaks = json.loads((ROOT / "MEASURED_PEAKS.json").read_text())
    except Exception:
        pass
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    cap_info, traffic, traffic_note = {}, None, "no ncu capture committed"
    try:
        cap_info = json.loads((ROOT / "profiles" / "ncu_traffic.json").read_text())["k_sweep"]
        traffic = cap_info["dram_bytes_per_move"] * float(W) * a.inner
        traffic_note = (f"dram__bytes_read.sum + dram__bytes_write.sum of {cap_info['capture']} ({cap_info['walkers']} walkers x {cap_info['steps']} MC steps, "
                        f"{cap_info['dram_bytes']:.4g} B) scaled per move to this launch")
    except Exception:
        pass
    roofline = {"bound": "fp64", "achieved": ach_tf, "peak": peak_tf, "unit": "TFLOP/s", "frac": ach_tf / peak_tf if peak_tf else None,
                "fp64_pipe_pct": cap_info.get("fp64_pipe_pct"), "issue_active_pct": cap_info.get("issue_active_pct"),
                "traffic": traffic, "traffic_note": traffic_note, "kernel": "k_sweep<false, 32>",
                "note": "K1 is FP64-pipe bound, not HBM/tensor bound (SURVEY 8d): peak = DFMA loop measured on this GPU in this run "
                        "(mgpu_measure_fp64_peak, 2 FLOP per FMA); achieved = convention-C1 FLOPs (29/pair geometry, 8/LJ term, "
                        "69/erfc-Coulomb term, k-space per SURVEY 8d) of the pairs this launch EVALUATED / event-timed kernel time. C1 counts the "
                        "reference's arithmetic (erfc = 64), the kernel evaluates erfc(ar)/r with 8 FP64 instructions: fp64_pipe_pct / "
                        "issue_active_pct (ncu, committed capture) are the executed-instruction utilisations",
                "pairs_per_launch": pc["pairs"] / max(1, launches), "screened_pairs_per_launch": pc.get("screened", 0) / max(1, launches),
                "lj_terms_per_launch": pc["lj"] / max(1, launches), "coulomb_terms_per_launch": pc["coulomb"] / max(1, launches)}
    k2_bytes = kspace_bytes(dcount, ew["nk"])
    roofline_k2 = {"bound": "l2", "achieved": k2_bytes / (ms_total * 1e-3) / 1e9, "peak": l2_peak, "unit": "GB/s",
                   "frac": k2_bytes / (ms_total * 1e-3) / 1e9 / l2_peak if l2_peak else None, "nkvec": ew["nk"],
                   "note": "K2 (k-space) is fused into the sweep kernel, so its algorithmic bytes (16 nk read per trial + 16 nk written per commit, "
                           "SURVEY 8d) are divided by the WHOLE kernel's time: an upper bound of how far K2 is from its roof; peak = "
                           "mgpu_measure_l2_peak (48 MB L2-resident buffer read by all SMs, this run)"}
    stationarity = {"target_loading": a.loading, "mean_loading_begin": n_begin, "mean_loading_end": n_end,
                    "ln_fugacity_begin": ctl_head.lnfs[-a.steps] if len(ctl_head.lnfs) >= a.steps else None, "ln_fugacity_end": ctl_head.lnf,
                    "controller": "PD feedback on ln f between launches (all legs at a pinned loading): d ln f = -(0.5 e + 4 de), e = (<N> - N0) / N0",
                    "pinning_launches": a.equil, "pinning_seconds": t_prep,
                    "loading_during_pinning": [hist_head[i] for i in range(0, len(hist_head), max(1, len(hist_head) // 8))]}

    # relaxed configurations for the CPU baseline (rank 0): the first `cores` walkers' records
    cpu_states = None
    if rank == 0 and world == 1 and not a.no_cpu:
        nst = min(max(cores, 32), W)
        tmp = eng.host_buffer(nst * eng.record_doubles_max())
        off_st = eng.save_walkers(tmp, 0, nst)
        cpu_states = []
        for i in range(nst):
            rec = Engine.parse_record(tmp[off_st[i]:off_st[i + 1]], ew["nk"], [(r.active, r.natom) for r in s.residues])
            cpu_states.append((rec["molecules"][0]["com"].copy(), rec["molecules"][0]["offset"].copy()))

    # ---- the same sweep with the per-molecule framework-energy cache off (the reference's operation count) -----
    eng.set_option(OPT_HOST_CACHE, 0)
    t_nc, l_nc = timed_sweeps(ctl_head, W, 2, 1)
    pc_nc = eng.pair_counts()
    ms_nc = t_nc * 1e3
    t_nc = max_over_ranks(t_nc)
    fl_nc = FLOP_GEOM * (pc_nc["pairs"] + pc_nc.get("screened", 0)) + FLOP_LJ * pc_nc["lj"] + FLOP_COUL * pc_nc["coulomb"] + kspace_flops(dcount, na, ew["kmax"], ew["nk"]) * 2 / a.steps
    no_cache = {"moves_per_s": float(W) * a.inner * 2 * world / t_nc, "launches": int(l_nc),
                "roofline_frac_c1": fl_nc / (ms_nc * 1e-3) / 1e12 / peak_tf if peak_tf else None,
                "note": "mgpu_set_option(MGPU_OPT_HOST_CACHE, 0): old-geometry framework sums recomputed every trial, like the reference"}
    eng.set_option(OPT_HOST_CACHE, 1)
    eng.set_option(OPT_PHASE_SYNC, 0)
    t_ns, _ = timed_sweeps(ctl_head, W, 2, 1)
    t_ns = max_over_ranks(t_ns)
    no_sync = {"moves_per_s": float(W) * a.inner * 2 * world / t_ns,
               "note": "mgpu_set_option(MGPU_OPT_PHASE_SYNC, 0): walkers of a CTA free-running (no phase-alignment barriers: the default takes one at the top of every MC step and one before the guest pass, CTA-wide)"}
    eng.set_option(OPT_PHASE_SYNC, 1 if a.phase_sync < 0 else a.phase_sync)

    # ---- e2e: the block-level C-ABI entry with HOST buffers (mgpu_block, pipelined) ---------------------
    rec_max = eng.record_doubles_max()
    per_walker = 72 + 2 * ew["nk"] + (int(1.3 * max(a.loading, 8)) + 48) * (3 + 3 * na + 2)
    blob = [eng.host_buffer(min(rec_max, per_walker) * W) for _ in range(2)]
    off = eng.save_walkers(blob[0], 0, W)
    cur = 0
    for i in range(max(1, min(2, a.warmup))):
        off = eng.block(a.inner, blob[cur], off, blob[cur ^ 1], 0, W)
        cur ^= 1
    eng.traffic(reset=True)
    barrier()
    t0 = time.perf_counter()
    for i in range(a.steps):
        off = eng.block(a.inner, blob[cur], off, blob[cur ^ 1], 0, W)
        cur ^= 1
        ctl_head.step(eng, W)                   # the host's part between blocks (inside the wall clock)
    barrier()
    t_e2e = max_over_ranks(time.perf_counter() - t0)
    tr1 = eng.traffic()
    e2e = {"value": float(W) * a.inner * a.steps * world / t_e2e, "unit": "moves/s",
           "h2d_bytes_per_step": tr1["h2d_bytes"] / a.steps, "d2h_bytes_per_step": tr1["d2h_bytes"] / a.steps,
           "walkers_per_gpu": W, "mc_steps_per_call": a.inner, "ms_per_step": 1e3 * t_e2e / a.steps,
           "path": "mgpu_block: walker records (guest%com / guest%offset, ewald%Ak, energy, RNG, counters) H2D from pinned host "
                   "memory -> k_unpack -> k_sweep -> k_pack -> records D2H, every step, in 8 slices on 8 streams (copies under compute); "
                   "wall clock, max over ranks"}
    del blob

    # ---- SURVEY 8d M1 grid: loadings x walker counts, each pinned and relaxed (N = 1 only) --------------
    grid = None
    if do_grid:
        grid = {"loadings": [0, 16, 64, 128], "walkers": [64, 1024, 4096, 16384], "mc_steps_per_launch": a.inner, "cells": []}
        order = [a.loading] + [x for x in grid["loadings"] if x != a.loading]      # the headline state is already on the device
        for loading in order:
            ctl_g = ctl_head
            if loading != a.loading:
                eng.close()                               # the pool restarts from a configuration holding about that many waters
                eng = Engine(workload(loading), n_walkers=Wmax, capacity=CAPACITY, device=local)
                ctl_g = prepare(loading, Wmax, 999 + loading)
                found_lnf[loading] = float(np.mean(ctl_g.lnfs[-max(1, a.equil // 3):]))
            for Wg in grid["walkers"]:
                nb = float(eng.counts(0, 0, Wg).mean())
                ctl_w = LoadingController(loading, ctl_g.lnf)      # its own feedback: the leg's walkers only
                ctl_w.prev = nb
                t_g, l_g = timed_sweeps(ctl_w, Wg, 3, 1)
                pcg = eng.pair_counts()
                fl_g = FLOP_GEOM * pcg["pairs"] + FLOP_LJ * pcg["lj"] + FLOP_COUL * pcg["coulomb"]      # real-space part only
                grid["cells"].append({"loading": loading, "walkers": Wg, "moves_per_s": float(Wg) * a.inner * l_g / t_g,
                                      "frac_c1_realspace": fl_g / t_g / 1e12 / peak_tf if peak_tf else None,
                                      "mean_loading": [nb, float(eng.counts(0, 0, Wg).mean())], "ln_fugacity": found_lnf[loading],
                                      "shape": eng.sweep_shape(Wg)["text"]})
        grid["cells"].sort(key=lambda c: (c["loading"], c["walkers"]))
        eng.close()                                       # back to the headline state for the legs below
        eng = Engine(s, n_walkers=max(a.strong_walkers // world, POOL), capacity=CAPACITY, device=local)
        ctl_head = prepare(a.loading, max(a.strong_walkers // world, POOL), 4242 + rank)
        lnf_head = ctl_head.lnf

    # ---- a FIXED isotherm (64 points x replicas) split over the GPUs ----------------------
    strong = None
    if a.strong_walkers > 0:
        from maniac_b200.isotherm import IsothermPlan, reduce_sums, summarize
        Ws = a.strong_walkers // world
        plan = IsothermPlan(n_points=64, walkers_per_rank=Ws, world_size=world, rank=rank)
        # one decade of fugacity centred on the pinning value: the points drift towards their own loadings while timed
        fug = np.exp(lnf_head + (np.arange(64) / 63.0 - 0.5) * np.log(10.0))
        eng.set_fugacities(0, fug[plan.points()], 0)
        eng.seed(31337 + 7919 * rank)
        eng.reset_averages()
        for _ in range(2):
            eng.sweep(a.inner, 0, Ws)
        eng.timing_reset()
        barrier()
        for _ in range(a.steps):
            flush_l2()
            eng.sweep(a.inner, 0, Ws)
        barrier()
        ms_s, l_s = eng.timing("sweep")
        t_s = max_over_ranks(ms_s * 1e-3)
        per_walker_acc = np.zeros((Ws, 6))
        per_walker_acc[:, :4] = eng.all_averages(0)[:Ws]
        local_sums = plan.accumulate(per_walker_acc)
        if world > 1:
            eng.nccl_init_from_torch()
        total = reduce_sums(local_sums, engine=eng)
        if world > 1:
            eng.nccl_finalize()
        summ = summarize(total, beta)
        strong = {"metric": "mc_trial_moves_per_s", "scaling": "strong", "walkers_total": Ws * world, "walkers_per_gpu": Ws,
                  "value": float(Ws) * world * a.inner * a.steps / t_s, "ms_per_step": 1e3 * t_s / a.steps,
                  "shape": eng.sweep_shape(Ws)["text"],
                  "points": 64, "replicas_per_point": Ws * world / 64.0,
                  "reduction": "mgpu_reduce_averages (one ncclAllReduce of 384 doubles)" if world > 1 else "single rank: none",
                  "fugacity": [float(fug[i]) for i in range(0, 64, 9)],
                  "mean_waters": [float(summ["mean_N"][i]) for i in range(0, 64, 9)]}
    eng.close()

    # ---- Widom batches (configs[2]) in 16 blocks: mu_ex +- sigma, sums reduced over ranks ----------------
    widom = None
    if a.widom > 0:
        from maniac_b200.snapshot import load_snapshot
        sw = load_snapshot(ROOT / "tests" / "golden" / "zif8_co2_widom.npz")
        sw.p_translation, sw.p_rotation, sw.p_insertion_deletion, sw.p_widom = 0.0, 0.0, 0.0, 1.0
        engw = Engine(sw, n_walkers=1, capacity=8, device=local)
        engw.set_count(1, 0)
        engw.update_system_energy()
        engw.widom_batch(1, min(a.widom, 100_000), seed=1)
        engw.timing_reset()
        engw.reset_pair_counts()
        nblk = 16
        per_blk = a.widom // nblk
        blocks = np.zeros((nblk, 2))
        barrier()
        for b in range(nblk):
            _, sw_sum, n_ok = engw.widom_batch(1, per_blk, seed=2024, first_id=(rank * nblk + b) * per_blk)
            blocks[b] = (sw_sum, per_blk)
        barrier()
        ms_w, _ = engw.timing("widom")
        pcw = engw.pair_counts()
        done = nblk * per_blk
        fl = FLOP_GEOM * pcw["pairs"] + FLOP_LJ * pcw["lj"] + FLOP_COUL * pcw["coulomb"] + done * (
            sum(k + 1 for k in engw.ewald()["kmax"]) * 3 * FLOP_SINCOS + engw.ewald()["nk"] * (14 * 3 + 7))
        t_w = max_over_ranks(ms_w * 1e-3)
        beta_w = engw.thermo(1)["beta"]
        if world > 1:                                   # every rank's blocks, through the library's NCCL sum
            allb = np.zeros((world, nblk, 2))
            allb[rank] = blocks
            engw.nccl_init_from_torch()
            flat = allb.reshape(-1).copy()
            engw.reduce_averages(flat)
            engw.nccl_finalize()
            allb = flat.reshape(world * nblk, 2)
        else:
            allb = blocks
        mu_b = -np.log(np.maximum(allb[:, 0], 1e-300) / allb[:, 1]) / beta_w
        mu_all = float(-np.log(allb[:, 0].sum() / allb[:, 1].sum()) / beta_w)
        widom = {"workload": "CO2 test-particle insertion in empty ZIF-8 2x2x2 (BASELINE configs[2])", "batch_per_gpu": done,
                 "insertions_per_s": done * world / t_w, "ms": ms_w,
                 "mu_ex_kcal_mol": mu_all, "mu_ex_sigma": float(mu_b.std(ddof=1) / np.sqrt(len(mu_b))), "blocks": int(len(mu_b)),
                 "insertions_total": int(allb[:, 1].sum()),
                 "roofline": {"bound": "fp64", "achieved": fl / (ms_w * 1e-3) / 1e12, "peak": peak_tf, "unit": "TFLOP/s",
                              "frac": fl / (ms_w * 1e-3) / 1e12 / peak_tf if peak_tf else None, "kernel": "k_widom_batch<false>",
                              "fp64_pipe_pct": json.loads((ROOT / "profiles" / "ncu_traffic.json").read_text()).get("k_widom_batch", {}).get("fp64_pipe_pct")
                              if (ROOT / "profiles" / "ncu_traffic.json").exists() else None}}
        engw.close()

    # ---- configs[4]: CO2 / N2 mixture with swaps in the 17 664-atom triclinic supercell ---------------
    mixture = None
    if a.mixture_walkers > 0:
        from maniac_b200.snapshot import load_snapshot
        from maniac_b200.workloads import mixture_supercell
        sm = mixture_supercell(load_snapshot(ROOT / "tests" / "golden" / "zif8_co2_widom.npz"), reps=(2, 2, 2), tilt_xy=3.0,
                               n_co2=48, n_n2=48)
        Wm = a.mixture_walkers
        engm = Engine(sm, n_walkers=Wm, capacity=256, device=local)
        engm.seed(777 + rank)
        engm.sweep(8)
        engm.timing_reset()
        engm.reset_pair_counts()
        cm0 = np.array([engm.counters(w) for w in range(0, Wm, max(1, Wm // 64))]).sum(axis=0)
        barrier()
        engm.sweep(a.mixture_steps)
        barrier()
        ms_m, _ = engm.timing("sweep")
        cm1 = np.array([engm.counters(w) for w in range(0, Wm, max(1, Wm // 64))]).sum(axis=0)
        pcm = engm.pair_counts()
        t_m = max_over_ranks(ms_m * 1e-3)
        ewm = engm.ewald()
        fl_m = FLOP_GEOM * pcm["pairs"] + FLOP_LJ * pcm["lj"] + FLOP_COUL * pcm["coulomb"]      # geometry counted at the orthorhombic 29 / pair
        dcm = (cm1 - cm0).reshape(6, 2) * (Wm / len(range(0, Wm, max(1, Wm // 64))))
        k2m = kspace_bytes(dcm, ewm["nk"]) + 16.0 * ewm["nk"] * (dcm[4, 0] - dcm[4, 1] + dcm[4, 1])
        mixture = {"workload": "CO2/N2 mixture GCMC with identity swaps (0.3/0.3/0.2/0.2 translate/rotate/swap/insert-delete), ZIF-8 4x4x4 "
                               "unit cells = 17 664 framework atoms, triclinic (xy tilt 3 A), BASELINE configs[4]",
                   "walkers_per_gpu": Wm, "mc_steps_per_launch": a.mixture_steps, "nkvec": ewm["nk"], "kmax": ewm["kmax"],
                   "moves_per_s": float(Wm) * a.mixture_steps * world / t_m, "ms": ms_m,
                   "triclinic_candidates": engm.triclinic_candidates(),
                   "swap_trials_sampled": int(dcm[4, 0] - dcm[4, 1]), "swap_accepted_sampled": int(dcm[4, 1]),
                   "roofline_frac_c1_realspace": fl_m / (ms_m * 1e-3) / 1e12 / peak_tf if peak_tf else None,
                   "roofline_k2": {"bound": "l2", "achieved": k2m / (ms_m * 1e-3) / 1e9, "peak": l2_peak, "unit": "GB/s",
                                   "frac": k2m / (ms_m * 1e-3) / 1e9 / l2_peak if l2_peak else None, "nkvec": ewm["nk"]}}
        engm.close()

    # ---- the Fortran drivers' role: host-driven trials (host RNG / proposal / Metropolis per MC step) ----
    We = min(a.e2e_walkers, W)
    enge = Engine(s, n_walkers=We, capacity=CAPACITY, device=local)
    enge.set_fugacities(0, np.full(We, np.exp(lnf_head)), 0)
    hm = HostMonteCarlo(enge, seed=999 + rank)
    for w in range(0, We):
        hm.set_chemical_potential(0, lnf_head / beta, walker=w)
    hm.run(3)
    tr0 = hm.traffic()
    barrier()
    t0 = time.perf_counter()
    hm.run(a.e2e_steps)
    barrier()
    t_hd = max_over_ranks(time.perf_counter() - t0)
    tr1 = hm.traffic()
    host_driven = {"value": (tr1["trials"] - tr0["trials"]) * world / t_hd, "unit": "moves/s",
                   "h2d_bytes_per_mc_step": (tr1["h2d_bytes"] - tr0["h2d_bytes"]) / a.e2e_steps,
                   "d2h_bytes_per_mc_step": (tr1["d2h_bytes"] - tr0["d2h_bytes"]) / a.e2e_steps,
                   "walkers_per_gpu": We, "mc_steps": a.e2e_steps,
                   "path": "mhost_run -> mgpu_trial_batch + mgpu_commit_batch (host RNG/proposal/Metropolis, pinned staging, "
                           "H2D proposals + D2H energies every MC step), wall clock"}
    hm.close()
    enge.close()

    # ---- the literal drop-in: ONE walker, one trial per call (what a single unchanged Fortran process does) ----
    eng1 = Engine(s, n_walkers=1, capacity=CAPACITY, device=local)
    hm1 = HostMonteCarlo(eng1, seed=5 + rank)
    hm1.set_chemical_potential(0, lnf_head / beta, walker=0)
    hm1.run(200)
    eng1.timing_reset()
    t0 = time.perf_counter()
    n1 = 1500
    hm1.run(n1)
    t_1 = time.perf_counter() - t0
    ms_k1, l_k1 = eng1.timing("trial")
    single = {"moves_per_s": n1 / t_1, "us_per_move_wall": 1e6 * t_1 / n1, "us_per_trial_kernel": 1e3 * ms_k1 / max(1, l_k1),
              "path": "mhost_run with 1 walker -> mgpu_trial_batch(n=1) (one 256-thread CTA per trial) + mgpu_commit_batch, synchronous; "
                      "the latency a single unchanged Fortran process sees per move"}
    hm1.close()
    eng1.close()

    # ---- CPU baseline (rank 0, N = 1 only) ------------------------------------------------------
    cpu = None
    if rank == 0 and world == 1 and not a.no_cpu:
        v, n, dt, nbar = cpu_sample(s, a.cpu_seconds, cores, 512, a.loading, lnf_head, cpu_states)
        cpu = {"value": v, "unit": "moves/s", "cores": cores, "kind": "port",
               "sample": f"{cores} independent walkers (one per host thread) x {n} MC steps, started from relaxed configurations of the GPU walkers "
                         f"(records through host memory) under the same loading controller ({dt:.1f} s, mean loading at the end {nbar:.1f}); CPU oracle = C restatement of the reference's serial "
                         "algorithm with a compact per-type-pair LJ table (friendlier than the reference's 4-D arrays), gcc -O3 "
                         "-ffp-contract=off, no -march=native (no Fortran compiler in the image)"}
    if a.write_fugacity and rank == 0 and cpu_states:
        np.savez_compressed(STATES_FILE, loading=a.loading, n=min(32, len(cpu_states)),
                            **{f"com{i}": c for i, (c, o) in enumerate(cpu_states[:32])}, **{f"off{i}": o for i, (c, o) in enumerate(cpu_states[:32])})
    if a.write_fugacity and rank == 0:
        old = stored_lnf()
        old.update(found_lnf)
        FUG_FILE.write_text(json.dumps({"ln_fugacity": {str(k): v for k, v in sorted(old.items())},
                                        "note": "ln(fugacity) that holds <N> at the keyed loading for ZIF-8 2x2x2 + TIP4P at 300 K with the 0.4/0.4/0.2 "
                                                "mix (bench.py feedback stage, written with --write-fugacity on a B200); the CPU arm and the "
                                                "feedback's starting point read it"}, indent=1) + "\n")
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if rank != 0:
        return
    line = {"metric": "mc_trial_moves_per_s", "value": value, "unit": "moves/s", "n_gpus": world, "steps": a.steps,
            "warmup": a.warmup, "ms_per_step": 1e3 * t_dev / a.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": cfg, "clocks": clocks,
            "e2e": e2e, "gpu_launches": int(launches), "roofline": roofline, "roofline_k2": roofline_k2, "cpu_baseline": cpu,
            "stationarity": stationarity, "strong_scaling": strong, "grid": grid, "widom": widom, "mixture": mixture,
            "host_driven": host_driven, "single_walker_dropin": single, "no_host_cache": no_cache, "no_phase_sync": no_sync,
            "wall_s_timed_region": wall, "fp64_peak_tflops_measured": peak_tf, "l2_peak_gbs_measured": l2_peak}
    print(json.dumps(line))


if __name__ == "__main__":
    main()
