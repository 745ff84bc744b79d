"""The configs[4] leg alone (CO2 / N2 with swaps in the 17 664-atom triclinic supercell): tools/mixture_probe.py [walkers] [steps] [phase_sync]
Used under ncu to capture k_sweep<true>; prints moves/s."""
import sys
sys.path.insert(0, '.')
import maniac_b200  # noqa: F401
from maniac_b200.engine import Engine, OPT_PHASE_SYNC
from maniac_b200.snapshot import load_snapshot
from maniac_b200.workloads import mixture_supercell
W = int(sys.argv[1]) if len(sys.argv) > 1 else 2368
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 32
sm = mixture_supercell(load_snapshot('tests/golden/zif8_co2_widom.npz'), reps=(2, 2, 2), tilt_xy=3.0, n_co2=48, n_n2=48)
eng = Engine(sm, n_walkers=W, capacity=256)
eng.seed(777)
if len(sys.argv) > 3:
    eng.set_option(OPT_PHASE_SYNC, int(sys.argv[3]))
eng.sweep(8)
eng.timing_reset()
eng.reset_pair_counts()
eng.sweep(steps)
ms, _ = eng.timing("sweep")
pc = eng.pair_counts()
import os
print(os.environ.get("MANIAC_GPU_LIB", "default")[-24:], "sync", sys.argv[3] if len(sys.argv) > 3 else "default", end=" ")
print("mixture: %.3f M moves/s, %.2f ms, pairs %d lj %d coul %d, candidates %d" % (W * steps / ms / 1e3, ms, pc["pairs"], pc["lj"], pc["coulomb"], eng.triclinic_candidates()))
eng.close()
