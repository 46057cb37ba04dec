"""One all-pairs run for profiling: python scripts/ap_one.py N steps [rc]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from jax_tpus_benchmark_physics_simulation_b200 import lattice_jitter
from jax_tpus_benchmark_physics_simulation_b200.md import LJSimulation
N = int(sys.argv[1]); steps = int(sys.argv[2])
rc = float(sys.argv[3]) if len(sys.argv) > 3 else 2.5
R, V, box = lattice_jitter(N, seed=0)
sim = LJSimulation(N, rc=rc if rc > 0 else None, dt=0.005, path="allpairs")
sim.run((R, V), steps)
sim.run((R, V), steps)
ms = sim.last_run_ms()
print(f"N={N} steps={steps} {1e3*ms/(steps+1):.2f} us/step {N*(N-1)*(steps+1)/ms/1e3:.3e} pairs/s")
