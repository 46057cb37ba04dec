"""Perf exploration of the all-pairs persistent kernel (developer tool, not the bench).
usage: python scripts/ap_sweep.py [N ...]   env knobs are set per configuration."""
import os
import sys
import itertools

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from jax_tpus_benchmark_physics_simulation_b200 import lattice_jitter
from jax_tpus_benchmark_physics_simulation_b200.md import LJSimulation

Ns = [int(a) for a in sys.argv[1:]] or [4096, 65536]
prof = os.environ.get("SWEEP_PROF")
for N in Ns:
    R, V, box = lattice_jitter(N, seed=0)
    Rd, Vd = torch.from_numpy(R).cuda(), torch.from_numpy(V).cuda()
    steps = max(2, int(2e10 / (N * N)))
    for ipt, cps in itertools.product((1, 2), (2, 4, 6, 8)):
        os.environ["LJMD_AP_IPT"] = str(ipt)
        os.environ["LJMD_AP_CTAS_PER_SM"] = str(cps)
        if prof:
            os.environ["LJMD_AP_PROF"] = "1"
        sim = LJSimulation(N, rc=2.5, dt=0.005, path="allpairs")
        sim.run((Rd, Vd), steps)
        torch.cuda.synchronize()
        best = 1e30
        for _ in range(3):
            sim.run((Rd, Vd), steps)
            best = min(best, sim.last_run_ms())
        us = 1e3 * best / (steps + 1)
        print(f"N={N} ipt={ipt} ctas/sm={cps} steps={steps}: {us:8.2f} us/step  "
              f"{N * (N - 1) / us * 1e6:.3e} pairs/s", flush=True)
        sim.close()
