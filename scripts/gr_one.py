"""g(r) timing: python scripts/gr_one.py N S"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from jax_tpus_benchmark_physics_simulation_b200 import lattice_jitter
from jax_tpus_benchmark_physics_simulation_b200.md import LJSimulation
N = int(sys.argv[1]); S = int(sys.argv[2])
R, V, box = lattice_jitter(N, seed=0)
sim = LJSimulation(N, rc=2.5, dt=0.005, path="allpairs")
(_, _), traj = sim.run((R, V), 10 * S, sample_every=10)
torch.cuda.synchronize()
for _ in range(3):
    t0 = time.perf_counter()
    r, g = sim.calculate_g_r(traj)
    g.block_until_ready()
    dt = time.perf_counter() - t0
print(f"N={N} S={S} g(r) {1e3*dt:.2f} ms  {S*N*(N-1)/2/dt:.3e} pair distances/s  g max {float(np.asarray(g).max()):.3f}")
