"""How long would a Verlet list live if every particle had its own half-skin h_i = clamp(beta |v_i|, hmin, hmax)
chosen from its speed at build time?  (rigorous validity: d_i <= h_i for all i; pair radius rc + h_i + h_j)
    python scripts/skin_study.py [N]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from jax_tpus_benchmark_physics_simulation_b200 import lattice_jitter
from jax_tpus_benchmark_physics_simulation_b200.md import LJSimulation
N = int(sys.argv[1]) if len(sys.argv) > 1 else 4194304
R, V, box = lattice_jitter(N, seed=0)
box = float(box)
sim = LJSimulation(N, rc=2.5, dt=0.005, path="cells", skin=0.5)
state, _ = sim.run((R, V), 600)
R0, V0 = state[0].tensor.clone(), state[1].tensor.clone()
speed = V0.norm(dim=1)
print("kT =", float((V0 ** 2).sum() / (2 * N)), "max speed", float(speed.max()))
T = 60
disp = []
st = (R0, V0)
for t in range(T):
    st, _ = sim.run(st, 1)
    st = (st[0].tensor, st[1].tensor)
    d = st[0] - R0
    d = d - box * torch.round(d / box)
    disp.append(d.norm(dim=1))
disp = torch.stack(disp)           # (T, N)
print("uniform half-skin 0.25: first violation at step", int((disp.max(dim=1).values > 0.25).float().argmax()) + 1)
rho = 0.8
for hmin in (0.10, 0.15, 0.20, 0.25):
    for beta in (0.06, 0.08, 0.10, 0.12, 0.15):
        for hmax in (0.5, 0.75):
            h = torch.clamp(beta * speed, hmin, hmax)
            viol = (disp > h[None, :]).any(dim=1)
            first = int(viol.float().argmax()) + 1 if bool(viol.any()) else T + 1
            # mean list length ~ rho pi E[(rc + h_i + h_j)^2] with independent h
            m1, m2 = float(h.mean()), float((h * h).mean())
            mean_list = rho * np.pi * (2.5 ** 2 + 2 * 2.5 * 2 * m1 + 2 * m2 + 2 * m1 * m1)
            print(f"hmin {hmin:.2f} beta {beta:.2f} hmax {hmax:.2f}: lives {first:3d} steps, mean list {mean_list:5.1f}, "
                  f"mean h {m1:.3f}")
