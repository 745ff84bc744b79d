import sys, time
sys.path.insert(0, '.')
import maniac_b200
from maniac_b200.engine import Engine
from maniac_b200.hostmc import HostMonteCarlo
from maniac_b200.snapshot import load_snapshot
from maniac_b200.workloads import load_pore
s = load_pore(load_snapshot('tests/golden/zif8_h2o_gcmc.npz'), 0, 64, seed=12345)
s.p_translation, s.p_rotation, s.p_insertion_deletion, s.p_swap, s.p_widom = 0.4, 0.4, 0.2, 0.0, 0.0
eng = Engine(s, n_walkers=1, capacity=1024)
hm = HostMonteCarlo(eng, seed=5)
hm.run(300); eng.timing_reset()
t0 = time.perf_counter(); hm.run(2000); t = time.perf_counter() - t0
ms, l = eng.timing("trial")
print("us/move", 1e6 * t / 2000, "kernel us", 1e3 * ms / l, "E", eng.energy()[5])
