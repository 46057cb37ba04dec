"""One small run of each kernel family for compute-sanitizer (scripts/sanitize.sh):
    python scripts/sanitize_case.py {cluster|ordered|tiles|cells|gr}"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from jax_tpus_benchmark_physics_simulation_b200 import lattice_jitter
from jax_tpus_benchmark_physics_simulation_b200.md import LJSimulation
case = sys.argv[1]
N, path, steps, kw = {"cluster": (400, "allpairs", 12, {}), "ordered": (1024, "allpairs", 8, {}),
                      "tiles": (4096, "allpairs", 6, {}), "cells": (16384, "cells", 30, {"skin": 0.3}),
                      "gr": (1024, "allpairs", 4, {})}[case]
R, V, box = lattice_jitter(N, seed=0)
sim = LJSimulation(N, rc=2.5, dt=0.005, path=path, **kw)
F, pe = sim.force_and_energy(R)
(R1, V1), traj = sim.run((R, V), steps, sample_every=max(1, steps // 2), energy_every=max(1, steps // 2))
R1.block_until_ready()
extra = ""
if case == "cells":
    extra = f" rebuilds {sim.last_rebuilds()}"
if case == "gr":
    c, g = sim.calculate_g_r(traj)
    extra = f" g(r) max {float(np.asarray(g).max()):.3f}"
print(f"{case}: N={N} mode {sim.allpairs_mode()} PE {float(pe):.4f} E {sim.last_energies.numpy()[-1].sum():.4f}{extra}")
