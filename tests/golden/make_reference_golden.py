"""Golden vectors from the REFERENCE'S OWN SOURCE FILE, executed unmodified on the torch facade of
the jax API (tests/golden/jax_facade.py: JAX / XLA are not installable in this image).

    python tests/golden/make_reference_golden.py [/root/reference]

For every case the script's `main(args)` (MD:13-192) is run once with the initial conditions
injected through `random.uniform` / `random.normal` (MD:133-135): that executes equilibrate_fn,
production_fn and calculate_g_r exactly as the reference's driver does (MD:138-165).  The closures it
defined (MD:46-131) are then called again on the same inputs for the per-function vectors:
periodic_displacement, total_energy_fn (through the lambda handed to grad, MD:64), force_fn,
verlet_step.  Everything is written to tests/golden/ref_md_<case>.npz; the tests pin the CPU oracle
(oracle/lj_oracle.py, oracle/lj_oracle.c) and the CUDA path to these files.

What the vectors are: the reference's expression sequences in IEEE fp32 (torch CPU), its energy
function differentiated by reverse-mode autodiff, its sampling rule, its histogram call.
What they are not: XLA's reduction order / pow lowering and jax.random's threefry stream.
"""
import argparse
import importlib.util
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)
import jax_facade  # noqa: E402
from jax_tpus_benchmark_physics_simulation_b200 import lattice_jitter  # noqa: E402

REF_DIR = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
REF_FILE = os.path.join(REF_DIR, "molecular_dynamics_jax_single-host_workload.py")

CASES = {
    # the reference's default system, shortened phases (N=400, rho=0.8, dt=1e-3: MD:196-213 defaults)
    "n400": dict(N=400, rho=0.8, kT=1.0, dt=1e-3, eq_steps=100, prod_steps=205, sample_every=50, seed=0),
    # dt = 0.005 as in BASELINE configs 2-5, and a production length that is not a multiple of the
    # sampling period (MD:88-100)
    "n64": dict(N=64, rho=0.8, kT=1.0, dt=5e-3, eq_steps=200, prod_steps=45, sample_every=10, seed=1),
    # BASELINE config 2's size (the Newton's-third-law tile kernel on the GPU side), a handful of steps:
    # a dense autodiff force evaluation of 4096 particles takes seconds on one CPU thread
    "n4096": dict(N=4096, rho=0.8, kT=1.0, dt=5e-3, eq_steps=3, prod_steps=4, sample_every=2, seed=2),
}


def run_case(name, c):
    torch.set_num_threads(1)                      # one thread: the reductions' order is reproducible
    saved = jax_facade.install()
    try:
        spec = importlib.util.spec_from_file_location("ref_md_" + name, REF_FILE)
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)              # defines main(); the argparse block is under __main__
        R0, V0, box = lattice_jitter(c["N"], rho=c["rho"], kT=c["kT"], seed=c["seed"])
        U = torch.from_numpy((R0 / np.float32(box)).astype(np.float32))
        jax_facade.REG.uniform = U
        jax_facade.REG.normal = torch.from_numpy(V0.copy())
        args = argparse.Namespace(output=os.devnull, **c)
        mod.main(args)
        reg = jax_facade.REG
        box_t = torch.tensor(float(np.sqrt(np.float32(c["N"] / c["rho"]))), dtype=torch.float32)       # MD:30
        Ri = U * box_t                                                                     # MD:134
        Vi = reg.normal * torch.tensor(float(np.sqrt(np.float32(c["kT"]))), dtype=torch.float32)      # MD:135
        # what main() blocked on, in order: state_eq[0] (MD:145), R_history (MD:152), g_r (MD:163)
        R_eq, R_hist, g_r = reg.blocked[0], reg.blocked[1], reg.blocked[2]
        # the closures
        pdisp = reg.fn("periodic_displacement")
        force_fn = next(f for n, f in reg.jitted if n.startswith("grad_"))
        neg_energy = reg.grad_sources[0]                                                  # MD:64
        verlet_step = reg.fn("verlet_step")
        equilibrate_fn = reg.fn("equilibrate_fn")
        production_fn = reg.fn("production_fn")
        calc_gr = reg.fn("_calculate_g_r_internal")
        out = {"N": np.int64(c["N"]), "rho": np.float64(c["rho"]), "dt": np.float64(c["dt"]),
               "eq_steps": np.int64(c["eq_steps"]), "prod_steps": np.int64(c["prod_steps"]),
               "sample_every": np.int64(c["sample_every"]), "box": box_t.numpy(),
               "R0": Ri.numpy(), "V0": Vi.numpy()}
        out["E0"] = (-neg_energy(Ri)).numpy()                                              # total_energy_fn(R0)
        out["F0"] = force_fn(Ri).numpy()
        R1, V1 = verlet_step((Ri, Vi))
        out["R1"], out["V1"] = R1.numpy(), V1.numpy()
        out["E1"] = (-neg_energy(R1)).numpy()
        st = equilibrate_fn((Ri, Vi))
        out["R_eq"], out["V_eq"] = st[0].numpy(), st[1].numpy()
        assert torch.equal(st[0], R_eq)            # the same call main() made
        (Rf, Vf), hist = production_fn(st)
        out["R_final"], out["V_final"], out["R_history"] = Rf.numpy(), Vf.numpy(), hist.numpy()
        assert torch.equal(hist, R_hist)
        r_max = box_t / 2.0                                                                # MD:158
        nbins = int(r_max / 0.05)                                                          # MD:159
        centers, g = calc_gr(hist, c["N"], box_t, nbins, r_max)
        assert torch.equal(g, g_r)
        out["gr_centers"], out["g_r"], out["gr_nbins"] = centers.numpy(), g.numpy(), np.int64(nbins)
        # periodic_displacement on a probe set incl. the half-box ties (MD:46-48)
        b = float(box_t)
        probe = np.array([0.0, 0.25 * b, -0.25 * b, 0.5 * b, -0.5 * b, np.nextafter(np.float32(0.5 * b), np.float32(b)),
                          -np.nextafter(np.float32(0.5 * b), np.float32(b)), 0.75 * b, -0.75 * b, b, -b,
                          0.999 * b, 1e-3, -1e-3], dtype=np.float32)
        out["pd_in"] = probe
        out["pd_out"] = pdisp(torch.from_numpy(probe), box_t).numpy()
        return out
    finally:
        jax_facade.uninstall(saved)


if __name__ == "__main__":
    only = sys.argv[2:] or list(CASES)
    for name, c in CASES.items():
        if name not in only:
            continue
        out = run_case(name, c)
        path = os.path.join(HERE, f"ref_md_{name}.npz")
        np.savez_compressed(path, **out)
        print("written", path, {k: np.shape(v) for k, v in out.items()}, "E0", float(out["E0"]),
              "max|F0|", float(np.abs(out["F0"]).max()))
