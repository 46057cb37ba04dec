"""Writes tests/golden/lj_golden_n64.npz from the torch-autodiff restatement of the reference
(oracle/lj_oracle.py).  These vectors pin the two independent restatements and the CUDA path to each
other (incl. the cutoff variants the reference does not have); they are NOT outputs of the reference.
Vectors made by the reference's own source file are tests/golden/ref_md_*.npz
(make_reference_golden.py).  Run:  python tests/golden/make_golden.py"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import lj_oracle as O  # noqa: E402
from jax_tpus_benchmark_physics_simulation_b200 import lattice_jitter  # noqa: E402

N = 64
R, V, box = lattice_jitter(N, seed=7)
Rt, Vt = torch.from_numpy(R), torch.from_numpy(V)
out = {"R": R, "V": V, "box": np.float32(box)}
for rc, key in ((None, "nocut"), (2.5, "rc25")):
    out[f"F_{key}"] = O.force_autodiff(Rt, float(box), rc=rc).numpy()
    out[f"pe_{key}"] = np.float32(O.total_energy(Rt, float(box), rc=rc))
ff = lambda r: O.force_autodiff(r, float(box), rc=2.5)
state = (Rt, Vt)
for _ in range(10):
    state = O.verlet_step(state, float(box), 0.005, ff)
out["R10_rc25"], out["V10_rc25"] = state[0].numpy(), state[1].numpy()
np.savez(os.path.join(os.path.dirname(os.path.abspath(__file__)), "lj_golden_n64.npz"), **out)
print("written", {k: np.shape(v) for k, v in out.items()})
