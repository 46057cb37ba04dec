"""Phase profile of the all-pairs persistent kernel (developer tool)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ["LJMD_AP_PROF"] = "1"
import torch
from jax_tpus_benchmark_physics_simulation_b200 import lattice_jitter
from jax_tpus_benchmark_physics_simulation_b200.md import LJSimulation
N = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 1000
R, V, box = lattice_jitter(N, seed=0)
sim = LJSimulation(N, rc=2.5, dt=0.005, path="allpairs")
sim.run((R, V), steps)
os.environ["LJMD_AP_PROF_CTAS"] = "1"
sim.run((R, V), steps)
print("us/step", 1e3 * sim.last_run_ms() / (steps + 1))
