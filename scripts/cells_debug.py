import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from jax_tpus_benchmark_physics_simulation_b200 import lattice_jitter
from jax_tpus_benchmark_physics_simulation_b200.md import LJSimulation
N = int(sys.argv[1]); steps = int(sys.argv[2])
R, V, box = lattice_jitter(N, seed=0)
a = LJSimulation(N, rc=2.5, dt=0.005, path="cells")
b = LJSimulation(N, rc=2.5, dt=0.005, path="allpairs")
for st in (1, 5, steps):
    (Ra, Va), _ = a.run((R, V), st, energy_every=1)
    ea = a.last_energies.numpy().astype(np.float64).sum(1)
    (Rb, Vb), _ = b.run((R, V), st, energy_every=1)
    eb = b.last_energies.numpy().astype(np.float64).sum(1)
    d = np.abs(Ra.numpy() - Rb.numpy()); d = np.minimum(d, float(box) - d)
    print(f"steps={st} rebuilds={a.last_rebuilds()} max|dR|={d.max():.3e} n_bad={(d.max(1)>1e-3).sum()} "
          f"E_cells[0,-1]={ea[0]:.3f},{ea[-1]:.3f} E_ap={eb[0]:.3f},{eb[-1]:.3f} ms={a.last_run_ms():.2f}")
    if (d.max(1) > 1e-3).any():
        bad = np.where(d.max(1) > 1e-3)[0][:10]
        print(" bad idx", bad, Ra.numpy()[bad], Rb.numpy()[bad])
