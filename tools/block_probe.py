"""Times the three parts of mgpu_block separately (load / sweep / save), wall clock + device timers."""
import sys, time
sys.path.insert(0, '.')
import numpy as np
import maniac_b200
from maniac_b200.engine import Engine
from maniac_b200.snapshot import load_snapshot
from maniac_b200.workloads import load_pore
W = int(sys.argv[1]) if len(sys.argv) > 1 else 4736
s = load_pore(load_snapshot('tests/golden/zif8_h2o_gcmc.npz'), 0, 64, seed=12345)
s.p_translation, s.p_rotation, s.p_insertion_deletion, s.p_swap, s.p_widom = 0.4, 0.4, 0.2, 0.0, 0.0
eng = Engine(s, n_walkers=W, capacity=1024)
eng.seed(1)
eng.sweep(64)
blob = [eng.host_buffer(W * 40000) for _ in range(2)]
off = eng.save_walkers(blob[0])
for it in range(3):
    eng.timing_reset()
    t0 = time.perf_counter(); eng.load_walkers(blob[0], off); t1 = time.perf_counter()
    eng.sweep(64); t2 = time.perf_counter()
    off = eng.save_walkers(blob[0]); t3 = time.perf_counter()
    print("load %.1f ms  sweep %.1f ms  save %.1f ms | device: unpack %.2f sweep %.2f pack %.2f ms | MB %.1f" % (
        1e3 * (t1 - t0), 1e3 * (t2 - t1), 1e3 * (t3 - t2), eng.timing("unpack")[0], eng.timing("sweep")[0], eng.timing("pack")[0], 8e-6 * off[-1]))
