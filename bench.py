#!/usr/bin/env python
"""bench.py -- MC trial moves/s (+ Widom insertions/s) of the MANIAC energy path on B200.

Contract (driver):  python bench.py --gpus N --steps K --warmup W     (N > 1 via torchrun)
prints ONE JSON line on rank 0.

  step      one launch of the device-resident sweep kernel: INNER Monte Carlo steps for every
            walker of this rank (one warp per walker, zero host round trips inside the launch)
  workload  BASELINE.json configs[1]: ZIF-8 2x2x2 + TIP4P water GCMC, move mix 0.4/0.4/0.2
            (translate / rotate / insert-delete), 64 waters per walker initially, walkers
            spread over a 64-point fugacity grid (the isotherm sweep of configs[3])
  value     trial moves / s over all ranks, state resident in HBM, device-timed (CUDA events on
            the launching stream), max over ranks
  e2e       the same metric through the block-level C-ABI call with HOST buffers (mgpu_block):
            every step the walkers' whole state travels pinned host -> device, the MC steps run,
            and the updated state travels back; wall clock
  host_driven  the Fortran drivers' role: host RNG / proposal / Metropolis, one mgpu_trial_batch +
            mgpu_commit_batch per MC step (proposals H2D, energies D2H)
  single_walker_dropin  one walker, one trial per call: the latency one unchanged Fortran process sees
  roofline  the sweep kernel against the FP64-pipe peak measured on this GPU by a DFMA loop
            (the binding roof of K1; SURVEY.md 8d convention C1): frac = C1 FLOPs of the pairs
            EVALUATED; frac_reference_ops = the reference's full operation count per move at this
            rate (the framework-energy cache and the per-molecule screen skip pairs); traffic =
            DRAM bytes per launch from the committed ncu capture
  no_host_cache / no_phase_sync   the same sweep with MGPU_OPT_HOST_CACHE / MGPU_OPT_PHASE_SYNC off
  widom     BASELINE configs[2]: 10^6 CO2 test insertions in empty ZIF-8, one launch
  mixture   BASELINE configs[4]: CO2/N2 with identity swaps in the 17 664-atom triclinic supercell
  isotherm  per-point averages summed over ranks by the library's NCCL reduction (the one exchange)
  cpu_baseline  the CPU oracle (a C restatement of the reference's serial algorithm) on the
            host cores, bounded sample of the same workload
  --impl reference   times that CPU path as the main line (no Fortran compiler in the image,
            so the reference itself cannot be built; kind = "port")
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
CAPACITY = 1024          # molecules per walker (NB_MAX_MOLECULE analogue): room for the pore-filling points of the fugacity grid
FLOP_GEOM, FLOP_LJ, FLOP_COUL, FLOP_SINCOS = 29.0, 8.0, 69.0, 64.0      # SURVEY.md 8d, convention C1


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=8)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--walkers", type=int, default=4736, help="walkers per GPU (weak scaling); 4736 = 2 x 148 SMs x 16 warps")
    ap.add_argument("--inner", type=int, default=INNER_DEFAULT, help="MC steps per walker per launch")
    ap.add_argument("--loading", type=int, default=64, help="initial waters per walker")
    ap.add_argument("--e2e-walkers", type=int, default=2368, help="walkers of the host-driven leg (2368 = one warp-per-trial wave: 148 SMs x 16 warps)")
    ap.add_argument("--e2e-steps", type=int, default=20)
    ap.add_argument("--widom", type=int, default=1_000_000, help="insertions in the Widom batch (0 = skip)")
    ap.add_argument("--mixture-walkers", type=int, default=2368, help="walkers of the configs[4] leg (0 = skip)")
    ap.add_argument("--mixture-steps", type=int, default=32)
    ap.add_argument("--cpu-seconds", type=float, default=12.0)
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--phase-sync", type=int, default=-1, help="experiment: MGPU_OPT_PHASE_SYNC value (2 = top-of-step barrier only, 3 = both)")
    ap.add_argument("--quick", action="store_true", help="kernel-variant experiments: the timed sweep only, prints a short line")
    return ap.parse_args()


def workload(loading):
    import maniac_b200  # noqa: F401
    from maniac_b200.snapshot import load_snapshot
    from maniac_b200.workloads import load_pore
    s = load_snapshot(ROOT / "tests" / "golden" / "zif8_h2o_gcmc.npz")
    s = load_pore(s, 0, loading, seed=12345)
    s.p_translation, s.p_rotation, s.p_insertion_deletion, s.p_swap, s.p_widom = 0.4, 0.4, 0.2, 0.0, 0.0
    return s


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


def cpu_sample(system, seconds, threads, capacity):
    """Aggregate moves/s of the CPU oracle on `threads` host threads, one walker each (-O3 build of the oracle source)."""
    os.environ["MANIAC_ORACLE_VARIANT"] = "o3"
    from oracle.oracle import Oracle
    oracles = []
    for t in range(threads):
        o = Oracle(system, capacity=capacity)
        o.update_system_energy()
        o.seed(12345 + 104729 * t)
        oracles.append(o)
    t0 = time.perf_counter()
    oracles[0].monte_carlo_steps(100, trace=False)
    per = (time.perf_counter() - t0) / 100
    n = max(100, int(seconds / per))
    done = [0] * threads

    def run(i):
        oracles[i].monte_carlo_steps(n, trace=False)
        done[i] = n
    ths = [threading.Thread(target=run, args=(i,)) for i in range(threads)]
    t0 = time.perf_counter()
    for th in ths:
        th.start()
    for th in ths:
        th.join()
    dt = time.perf_counter() - t0
    return sum(done) / dt, n, dt


def main():
    a = parse()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    cores = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    cfg = {"workload": "ZIF-8 2x2x2 (2208 atoms, 34.02 A cubic) + TIP4P H2O GCMC, BASELINE configs[1]; "
                       f"{a.loading} waters/walker initially, mix 0.4/0.4/0.2 translate/rotate/insert-delete, "
                       "walkers over a 64-point log fugacity grid 1e-2..1e4 (configs[3])",
           "walkers_per_gpu": a.walkers, "mc_steps_per_launch": a.inner, "r_cut": 17.0, "nkvec": 297,
           "l2": "per-walker state (~45 KB x walkers) exceeds the 126 MB L2 for >= 3000 walkers; the 88 KB "
                 "framework block is meant to stay cache-resident"}
    s = workload(a.loading)

    # ------------------------------------------------------------------ reference arm (CPU)
    if a.impl == "reference":
        if rank != 0:
            return
        per_s = []
        for i in range(a.warmup + a.steps):
            v, n, dt = cpu_sample(s, max(2.0, a.cpu_seconds / max(1, a.steps)), cores, 512)
            if i >= a.warmup:
                per_s.append((v, n, dt))
        v = float(np.mean([x[0] for x in per_s]))
        line = {"impl": "reference", "metric": "mc_trial_moves_per_s", "value": v, "unit": "moves/s", "n_gpus": a.gpus,
                "steps": a.steps, "warmup": a.warmup, "ms_per_step": 1e3 * float(np.mean([x[2] for x in per_s])),
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": cfg,
                "cpu_baseline": {"value": v, "unit": "moves/s", "cores": cores, "kind": "port",
                                 "sample": f"{cores} independent walkers (one per host thread) x {per_s[0][1]} MC steps per step; "
                                           "C restatement of the reference's serial Fortran algorithm, gcc -O3 -ffp-contract=off "
                                           "(no Fortran compiler in the image: the reference itself cannot be built)"},
                "e2e": {"value": v, "unit": "moves/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        print(json.dumps(line))
        return

    # ------------------------------------------------------------------ B200 arm
    import torch
    import torch.distributed as dist
    from maniac_b200.engine import Engine
    from maniac_b200.hostmc import HostMonteCarlo
    from maniac_b200.workloads import isotherm_fugacities
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the energy path has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    W = a.walkers
    eng = Engine(s, n_walkers=W, capacity=CAPACITY, device=local)
    fug = isotherm_fugacities(64)

    def point(w):
        # replica -> isotherm point: the four walkers (warps) of a CTA that share an SM sub-partition
        # (w, w+4, w+8, w+12) are replicas of the SAME fugacity point, so their loadings -- hence the
        # lengths of their MC steps -- are statistically alike and the per-quartet phase barrier waits little
        return ((rank * W + w) // 16 * 4 + w % 4) % 64
    for w in range(W):
        eng.set_fugacity(0, float(fug[point(w)]), walker=w)
    eng.seed(12345 + 7919 * rank)
    if a.phase_sync >= 0:
        from maniac_b200.engine import OPT_PHASE_SYNC as _OPS
        eng.set_option(_OPS, a.phase_sync)
    peak_tf, _ = eng.measure_fp64_peak()
    ew = eng.ewald()
    na = 4

    sampler = ClockSampler(local)
    sampler.start()                                    # nvidia-smi needs ~0.5 s before its first row: start before the warm-up
    for _ in range(a.warmup):
        eng.sweep(a.inner)
    c0 = np.array([eng.counters(w) for w in range(0, W, max(1, W // 256))]).sum(axis=0)   # sampled walkers
    n_sampled = len(range(0, W, max(1, W // 256)))
    eng.reset_pair_counts()
    eng.timing_reset()
    barrier()
    sampler.mark_begin()
    t0 = time.perf_counter()
    for _ in range(a.steps):
        eng.sweep(a.inner)
    barrier()
    wall = time.perf_counter() - t0
    sampler.mark_end()
    clocks = sampler.stop()
    ms_total, launches = eng.timing("sweep")
    pc = eng.pair_counts()
    c1 = np.array([eng.counters(w) for w in range(0, W, max(1, W // 256))]).sum(axis=0)
    dcount = (c1 - c0) * (W / n_sampled)
    t_dev = ms_total * 1e-3
    if world > 1:
        t = torch.tensor([t_dev, wall], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        t_dev, wall = float(t[0]), float(t[1])
    moves = float(W) * a.inner * a.steps * world
    value = moves / t_dev

    if a.quick:
        if rank == 0:
            print(json.dumps({"quick": True, "lib": os.environ.get("MANIAC_GPU_LIB", "default"), "moves_per_s": value,
                              "ms_per_step": 1e3 * t_dev / a.steps, "clocks": clocks}))
        eng.close()
        return
    flops = FLOP_GEOM * pc["pairs"] + FLOP_LJ * pc["lj"] + FLOP_COUL * pc["coulomb"] + kspace_flops(dcount, na, ew["kmax"], ew["nk"])
    ach_tf = flops / (ms_total * 1e-3) / 1e12
    nmean = float(np.mean([eng.count(0, walker=w) for w in range(0, W, max(1, W // 64))]))
    bytes_alg = float(W) * a.inner * a.steps * (32.0 * ew["nk"] + 8.0 * 15 * nmean + 36.0 * 2208)
    peaks = {}
    try:
        peaks = json.loads((ROOT / "MEASURED_PEAKS.json").read_text())
    except Exception:
        pass
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    traffic, traffic_note = None, "no ncu capture committed"
    try:
        tj = json.loads((ROOT / "profiles" / "ncu_traffic.json").read_text())["k_sweep"]
        traffic = tj["dram_bytes_per_move"] * float(W) * a.inner
        traffic_note = (f"dram__bytes_read.sum + dram__bytes_write.sum of {tj['capture']} ({tj['walkers']} walkers x {tj['steps']} MC steps, "
                        f"{tj['dram_bytes']:.4g} B) scaled per move to this launch")
    except Exception:
        pass
    roofline = {"bound": "fp64", "achieved": ach_tf, "peak": peak_tf, "unit": "TFLOP/s", "frac": ach_tf / peak_tf if peak_tf else None,
                "traffic": traffic, "traffic_note": traffic_note, "kernel": "k_sweep<false>",
                "note": "K1/K2 are FP64-pipe bound, not HBM/tensor bound (SURVEY 8d): peak = DFMA loop measured on this GPU in this "
                        "run (mgpu_measure_fp64_peak, 2 FLOP per FMA); achieved = convention-C1 FLOPs (29/pair geometry, 8/LJ term, "
                        "69/erfc-Coulomb term, k-space per SURVEY 8d) of the pairs this launch EVALUATED / event-timed kernel time; "
                        "the framework-energy cache and the per-molecule screen skip pairs the reference evaluates, "
                        "frac_reference_ops credits the reference's full operation count per move (measured by the no_host_cache leg)",
                "pairs_per_launch": pc["pairs"] / max(1, launches), "screened_pairs_per_launch": pc.get("screened", 0) / max(1, launches), "lj_terms_per_launch": pc["lj"] / max(1, launches),
                "coulomb_terms_per_launch": pc["coulomb"] / max(1, launches),
                "hbm": {"achieved": bytes_alg / t_dev / 1e9 / world, "peak": hbm_peak, "unit": "GB/s",
                        "frac": bytes_alg / t_dev / 1e9 / world / hbm_peak,
                        "peak_source": "MEASURED_PEAKS.json" if peaks else "fallback 6.65 TB/s"}}

    # ---- the one exchange of the path: per-point sums over ranks, once, after the sampling -----
    from maniac_b200.isotherm import IsothermPlan, reduce_sums, summarize
    plan = IsothermPlan(n_points=64, walkers_per_rank=W, world_size=world, rank=rank)
    assert all(plan.point_of(rank * W + w) == point(w) for w in range(0, W, 97))
    per_walker = np.zeros((W, 6))
    per_walker[:, :4] = eng.all_averages(0)
    local_sums = plan.accumulate(per_walker)
    if world > 1:
        eng.nccl_init_from_torch()
    total = reduce_sums(local_sums, engine=eng)
    if world > 1:
        eng.nccl_finalize()
    summ = summarize(total, eng.thermo(0)["beta"])
    isotherm = {"points": 64, "replicas_per_point": float(W * world) / 64.0,
                "reduction": "mgpu_reduce_averages (one ncclAllReduce of 384 doubles)" if world > 1 else "single rank: none",
                "fugacity": [float(fug[i]) for i in range(0, 64, 9)],
                "mean_waters": [float(summ["mean_N"][i]) for i in range(0, 64, 9)],
                "note": "block averages since the start of the run (not equilibrated: a throughput benchmark)"}

    # ---- the same sweep with the per-molecule framework-energy cache off (framework swept for the old
    #      AND the new geometry of every move, the reference's operation count) --------------------
    from maniac_b200.engine import OPT_HOST_CACHE
    eng.set_option(OPT_HOST_CACHE, 0)
    eng.sweep(a.inner)
    eng.timing_reset()
    eng.reset_pair_counts()
    barrier()
    for _ in range(2):
        eng.sweep(a.inner)
    barrier()
    ms_nc, l_nc = eng.timing("sweep")
    pc_nc = eng.pair_counts()
    t_nc = ms_nc * 1e-3
    if world > 1:
        t = torch.tensor([t_nc], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        t_nc = float(t[0])
    fl_nc = FLOP_GEOM * (pc_nc["pairs"] + pc_nc.get("screened", 0)) + FLOP_LJ * pc_nc["lj"] + FLOP_COUL * pc_nc["coulomb"] + kspace_flops(dcount, na, ew["kmax"], ew["nk"]) * 2 / a.steps
    no_cache = {"moves_per_s": float(W) * a.inner * 2 * world / t_nc, "launches": int(l_nc),
                "roofline_frac_c1": fl_nc / (ms_nc * 1e-3) / 1e12 / peak_tf if peak_tf else None,
                "note": "mgpu_set_option(MGPU_OPT_HOST_CACHE, 0): old-geometry framework sums recomputed every trial"}
    eng.set_option(OPT_HOST_CACHE, 1)
    # the reference's own operation count per move on this workload state (no cache: every pair the Fortran loops visit),
    # credited to the cached run: "how fast would the FP64 pipe have to be to do the reference's arithmetic at this rate"
    ref_flops_per_move = fl_nc / (float(W) * a.inner * 2)
    roofline["reference_ops_per_move_c1"] = ref_flops_per_move
    roofline["achieved_reference_ops"] = ref_flops_per_move * value / world / 1e12
    roofline["frac_reference_ops"] = roofline["achieved_reference_ops"] / peak_tf if peak_tf else None

    # ---- the same sweep without the per-quartet phase alignment (MGPU_OPT_PHASE_SYNC) -------------
    from maniac_b200.engine import OPT_PHASE_SYNC
    eng.set_option(OPT_PHASE_SYNC, 0)
    eng.sweep(a.inner)
    eng.timing_reset()
    barrier()
    for _ in range(2):
        eng.sweep(a.inner)
    barrier()
    ms_ns, _ = eng.timing("sweep")
    t_ns = ms_ns * 1e-3
    if world > 1:
        t = torch.tensor([t_ns], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        t_ns = float(t[0])
    no_sync = {"moves_per_s": float(W) * a.inner * 2 * world / t_ns,
               "note": "mgpu_set_option(MGPU_OPT_PHASE_SYNC, 0): walkers of a CTA free-running (no per-quartet barrier)"}
    eng.set_option(OPT_PHASE_SYNC, 1)

    # ---- Widom batch (configs[2]) ------------------------------------------------------------
    widom = None
    eng.close()
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
        barrier()
        _, sw_sum, n_ok = engw.widom_batch(1, a.widom, seed=2024 + rank)
        barrier()
        ms_w, _ = engw.timing("widom")
        pcw = engw.pair_counts()
        fl = FLOP_GEOM * pcw["pairs"] + FLOP_LJ * pcw["lj"] + FLOP_COUL * pcw["coulomb"] + a.widom * (
            sum(k + 1 for k in engw.ewald()["kmax"]) * 3 * FLOP_SINCOS + engw.ewald()["nk"] * (14 * 3 + 7))
        t_w = ms_w * 1e-3
        if world > 1:
            t = torch.tensor([t_w], device="cuda", dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            t_w = float(t[0])
        beta = engw.thermo(1)["beta"]
        widom = {"workload": "CO2 test-particle insertion in empty ZIF-8 2x2x2 (BASELINE configs[2])", "batch": a.widom,
                 "insertions_per_s": a.widom * world / t_w, "ms": ms_w,
                 "mu_ex_kcal_mol": float(-np.log(max(sw_sum, 1e-300) / a.widom) / beta), "accepted_weights": int(n_ok),
                 "roofline": {"bound": "fp64", "achieved": fl / (ms_w * 1e-3) / 1e12, "peak": peak_tf, "unit": "TFLOP/s",
                              "frac": fl / (ms_w * 1e-3) / 1e12 / peak_tf if peak_tf else None, "kernel": "k_widom_batch<false>"}}
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
        t_m = ms_m * 1e-3
        if world > 1:
            t = torch.tensor([t_m], device="cuda", dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            t_m = float(t[0])
        ewm = engm.ewald()
        fl_m = FLOP_GEOM * pcm["pairs"] + FLOP_LJ * pcm["lj"] + FLOP_COUL * pcm["coulomb"]      # geometry counted at the orthorhombic 29 / pair
        dcm = (cm1 - cm0).reshape(6, 2)
        mixture = {"workload": "CO2/N2 mixture GCMC with identity swaps (0.3/0.3/0.2/0.2 translate/rotate/swap/insert-delete), ZIF-8 4x4x4 "
                               "unit cells = 17 664 framework atoms, triclinic (xy tilt 3 A), BASELINE configs[4]",
                   "walkers_per_gpu": Wm, "mc_steps_per_launch": a.mixture_steps, "nkvec": ewm["nk"], "kmax": ewm["kmax"],
                   "moves_per_s": float(Wm) * a.mixture_steps * world / t_m, "ms": ms_m,
                   "triclinic_candidates": engm.triclinic_candidates(),
                   "swap_trials_sampled": int(dcm[4, 0] - dcm[4, 1]), "swap_accepted_sampled": int(dcm[4, 1]),
                   "roofline_frac_c1_realspace": fl_m / (ms_m * 1e-3) / 1e12 / peak_tf if peak_tf else None}
        engm.close()

    # ---- e2e: the block-level C-ABI entry with HOST buffers (mgpu_block) ------------------------
    # every step = one block of the MC loop: the walkers' records (coordinates, S(k), energies, RNG,
    # counters) are copied from pinned host memory to the device, `inner` MC steps run, the updated records
    # are copied back -- the same moves as `value`, plus the host<->device traffic of the whole state.
    enge = Engine(s, n_walkers=W, capacity=CAPACITY, device=local)
    for w in range(W):
        enge.set_fugacity(0, float(fug[point(w)]), walker=w)
    enge.seed(4321 + 7919 * rank)
    for _ in range(2):
        enge.sweep(a.inner)                          # move towards the steady-state loading before records are sized
    rec_max = enge.record_doubles_max()
    nmax = max(enge.count(0, walker=w) for w in range(0, W, max(1, W // 128)))
    per_walker = 72 + 2 * ew["nk"] + int(1.5 * nmax + 64) * (3 + 3 * na + 2)
    blob = [enge.host_buffer(min(rec_max, per_walker) * W) for _ in range(2)]
    off = enge.save_walkers(blob[0])
    for i in range(max(1, a.warmup)):
        off = enge.block(a.inner, blob[i % 2], off, blob[(i + 1) % 2])
        cur = (i + 1) % 2
    enge.traffic(reset=True)
    barrier()
    t0 = time.perf_counter()
    for i in range(a.steps):
        off = enge.block(a.inner, blob[cur], off, blob[cur ^ 1])
        cur ^= 1
    barrier()
    t_e2e = time.perf_counter() - t0
    tr1 = enge.traffic()
    if world > 1:
        t = torch.tensor([t_e2e], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        t_e2e = float(t[0])
    e2e = {"value": float(W) * a.inner * a.steps * world / t_e2e, "unit": "moves/s",
           "h2d_bytes_per_step": tr1["h2d_bytes"] / a.steps, "d2h_bytes_per_step": tr1["d2h_bytes"] / a.steps,
           "walkers_per_gpu": W, "mc_steps_per_call": a.inner, "ms_per_step": 1e3 * t_e2e / a.steps,
           "path": "mgpu_block: walker records (guest%com / guest%offset, ewald%Ak, energy, RNG, counters) H2D from pinned host "
                   "memory -> k_unpack -> k_sweep -> k_pack -> records D2H, every step; wall clock, max over ranks"}
    enge.close()

    # ---- the Fortran drivers' role: host-driven trials (host RNG / proposal / Metropolis per MC step) ----
    We = min(a.e2e_walkers, W)
    enge = Engine(s, n_walkers=We, capacity=CAPACITY, device=local)
    for w in range(We):
        enge.set_fugacity(0, float(fug[(rank * We + w) % 64]), walker=w)
    hm = HostMonteCarlo(enge, seed=999 + rank)
    beta_e = enge.thermo(0)["beta"]
    for w in range(We):
        hm.set_chemical_potential(0, float(np.log(fug[(rank * We + w) % 64]) / beta_e), walker=w)
    hm.run(3)
    tr0 = hm.traffic()
    barrier()
    t0 = time.perf_counter()
    hm.run(a.e2e_steps)
    barrier()
    t_hd = time.perf_counter() - t0
    tr1 = hm.traffic()
    if world > 1:
        t = torch.tensor([t_hd], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        t_hd = float(t[0])
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
        v, n, dt = cpu_sample(s, a.cpu_seconds, cores, 512)
        cpu = {"value": v, "unit": "moves/s", "cores": cores, "kind": "port",
               "sample": f"{cores} independent walkers (one per host thread) x {n} MC steps of the same workload ({dt:.1f} s); "
                         "CPU oracle = C restatement of the reference's serial algorithm, gcc -O3 (no Fortran compiler in the image)"}
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if rank != 0:
        return
    line = {"metric": "mc_trial_moves_per_s", "value": value, "unit": "moves/s", "n_gpus": world, "steps": a.steps,
            "warmup": a.warmup, "ms_per_step": 1e3 * t_dev / a.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": cfg, "clocks": clocks,
            "e2e": e2e, "host_driven": host_driven, "single_walker_dropin": single, "gpu_launches": int(launches), "roofline": roofline, "cpu_baseline": cpu, "widom": widom, "no_host_cache": no_cache, "no_phase_sync": no_sync, "mixture": mixture, "isotherm": isotherm,
            "wall_s_timed_region": wall, "mean_waters_per_walker": nmean, "fp64_peak_tflops_measured": peak_tf}
    print(json.dumps(line))


if __name__ == "__main__":
    main()
