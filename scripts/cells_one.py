"""One cell-list run for timing/profiling: python scripts/cells_one.py N steps"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from jax_tpus_benchmark_physics_simulation_b200 import lattice_jitter
from jax_tpus_benchmark_physics_simulation_b200.md import LJSimulation
N = int(sys.argv[1]); steps = int(sys.argv[2])
skin = float(sys.argv[3]) if len(sys.argv) > 3 else 0.3
R, V, box = lattice_jitter(N, seed=0)
sim = LJSimulation(N, rc=2.5, dt=0.005, path="cells", skin=skin)
Rd, Vd = torch.from_numpy(R).cuda(), torch.from_numpy(V).cuda()
sim.run((Rd, Vd), steps)
for _ in range(2):
    sim.run((Rd, Vd), steps)
    ms = sim.last_run_ms()
    rb = sim.last_rebuilds()
    print(f"N={N} steps={steps} skin={skin} rebuilds={rb} {1e3*ms/steps:.2f} us/step {N*steps/ms*1e3:.3e} particle-steps/s "
          f"alg GB/s {32.0*N*steps/ms/1e6:.1f}")
