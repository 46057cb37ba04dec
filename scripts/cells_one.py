"""One cell-list run for timing/profiling: python scripts/cells_one.py N steps [skin] [pre]
`pre` steps are run first (one launch) so that the timed / captured launches start from the melted
liquid instead of the initial lattice (ncu: -k regex:cells_persistent -s 3 -c 1 with pre > 0)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from jax_tpus_benchmark_physics_simulation_b200 import lattice_jitter
from jax_tpus_benchmark_physics_simulation_b200.md import LJSimulation
N = int(sys.argv[1]); steps = int(sys.argv[2])
skin = float(sys.argv[3]) if len(sys.argv) > 3 else 0.3
pre = int(sys.argv[4]) if len(sys.argv) > 4 else 0
R, V, box = lattice_jitter(N, seed=0)
sim = LJSimulation(N, rc=2.5, dt=0.005, path="cells", skin=skin)
state = (torch.from_numpy(R).cuda(), torch.from_numpy(V).cuda())
if pre > 0:
    state, _ = sim.run(state, pre)
    state = (state[0].tensor, state[1].tensor)
sim.run(state, steps)
for _ in range(2):
    sim.run(state, steps)
    ms = sim.last_run_ms()
    rb = sim.last_rebuilds()
    print(f"N={N} steps={steps} skin={skin} pre={pre} rebuilds={rb} {1e3*ms/steps:.2f} us/step {N*steps/ms*1e3:.3e} particle-steps/s "
          f"alg GB/s {32.0*N*steps/ms/1e6:.1f}")
sim.check()
