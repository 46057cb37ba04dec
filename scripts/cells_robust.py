"""Robustness probe of the cell-list path on inputs that are not lattice-like:
uniform random positions (the reference's own IC style, MD:133), other densities / cutoffs.
Compares with the all-pairs GPU path and the C oracle (cell grid)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from jax_tpus_benchmark_physics_simulation_b200.md import LJSimulation
from jax_tpus_benchmark_physics_simulation_b200.ic import box_size
from oracle import lj_oracle as O
rng = np.random.default_rng(7)
ok_all = True
for N, rho, rc, kind in [(65536, 0.8, 2.5, "uniform"), (65536, 0.4, 2.5, "uniform"), (16384, 1.1, 2.5, "soft"),
                         (65536, 0.8, 3.5, "soft"), (65536, 0.05, 2.5, "uniform")]:
    box = box_size(N, rho)
    if kind == "uniform":
        R = (rng.random((N, 2)) * float(box)).astype(np.float32)
    else:   # jittered lattice with large jitter: disordered but no hard overlaps
        n = int(round(np.sqrt(N))); a = float(box) / n
        g = (np.stack(np.meshgrid(np.arange(n), np.arange(n), indexing="ij"), -1).reshape(-1, 2) + 0.5) * a
        R = np.mod(g + rng.uniform(-0.3, 0.3, g.shape) * a, float(box)).astype(np.float32)
    try:
        c = LJSimulation(N, rho=rho, rc=rc, dt=0.002, path="cells")
        a_ = LJSimulation(N, rho=rho, rc=rc, dt=0.002, path="allpairs")
        Fc, pec = c.force_and_energy(R)
        Fa, pea = a_.force_and_energy(R)
        Fo, peo = O.c_forces_cells(R, box, rc)
        Fc, Fa = Fc.numpy(), Fa.numpy()
        fin = np.isfinite(Fo).all(axis=1) & np.isfinite(Fa).all(axis=1)
        sc = np.abs(Fo[fin]).max()
        e1 = np.abs(Fc[fin] - Fo[fin]).max() / sc
        e2 = np.abs(Fc[fin] - Fa[fin]).max() / sc
        cnt = c.neighbor_count(R, rc).cpu().numpy()
        cnt_o = O.c_neighbor_count(R, box, rc)
        good = e1 < 1e-5 and e2 < 1e-5 and np.array_equal(cnt, cnt_o)
        print(f"N={N} rho={rho} rc={rc} {kind}: |F|max {sc:.3e} cells-oracle {e1:.2e} cells-allpairs {e2:.2e} "
              f"PE rel {abs(float(pec) - peo) / abs(peo):.2e} nbr max {cnt.max()} counts_equal {np.array_equal(cnt, cnt_o)} -> {'OK' if good else 'BAD'}")
        ok_all &= good
    except Exception as e:
        print(f"N={N} rho={rho} rc={rc} {kind}: EXCEPTION {e}")
        ok_all = False
sys.exit(0 if ok_all else 1)
