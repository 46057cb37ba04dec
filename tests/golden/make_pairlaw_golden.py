"""Golden vectors for the gravity pair laws from the reference's own function
source: the `pairwise_forces` definition of nbody_bh_merger_sim_single-host_workload.py (NBODY:53-67)
and the `acceleration` definition of three_particles_em_nonuni_single-host_workload.py (EM3:23-52) are
taken from the files as AST nodes - the scripts themselves prompt / simulate / animate at import, so they
are not executed as a whole - and run unmodified on the torch facade of the jax API
(tests/golden/jax_facade.py).  `acceleration` is called with zero charges, which leaves its gravity term.

    python tests/golden/make_pairlaw_golden.py [/root/reference]   ->  tests/golden/ref_pairlaw.npz
"""
import argparse
import ast
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import jax_facade  # noqa: E402

REF = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"


def function_from(path, name, env):
    tree = ast.parse(open(path).read())
    node = next(n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name == name)
    mod = ast.Module(body=[node], type_ignores=[])
    exec(compile(mod, path, "exec"), env)
    return env[name]


def main():
    torch.set_num_threads(1)
    saved = jax_facade.install()
    try:
        import jax
        import jax.numpy as jnp
        from jax import jit, vmap
        G = 1.5
        env_n = {"jax": jax, "jnp": jnp, "jit": jit, "G": G}
        pairwise_forces = function_from(os.path.join(REF, "nbody_bh_merger_sim_single-host_workload.py"),
                                        "pairwise_forces", env_n)
        env_e = {"jax": jax, "jnp": jnp, "jit": jit, "vmap": vmap, "G": G,
                 "args": argparse.Namespace(Bz=1.0, Bk=0.0, Ex=0.0, Ey=0.0)}
        acceleration = function_from(os.path.join(REF, "three_particles_em_nonuni_single-host_workload.py"),
                                     "acceleration", env_e)
        out = {"G": np.float32(G)}
        rng = np.random.default_rng(11)
        cases = {
            "em3_ic": (np.array([[0.0, 0.0], [1.0, 0.0], [0.5, 0.866]], dtype=np.float32),   # EM3:87-89
                       np.ones(3, dtype=np.float32)),
            "n2": (np.array([[0.0, 0.0], [3.0, 4.0]], dtype=np.float32), np.array([2.0, 7.0], dtype=np.float32)),
            "n5": (rng.uniform(-50, 50, (5, 2)).astype(np.float32), rng.uniform(5, 40, 5).astype(np.float32)),
            "n5_coincident": None,
        }
        p5, m5 = cases["n5"]
        pc = p5.copy(); pc[3] = pc[1]
        cases["n5_coincident"] = (pc, m5)
        for name, (pos, mass) in cases.items():
            pt, mt = torch.from_numpy(pos), torch.from_numpy(mass)
            out[f"{name}_pos"], out[f"{name}_mass"] = pos, mass
            out[f"{name}_nbody"] = pairwise_forces(pt, mt).numpy()
            n = len(pos)
            out[f"{name}_em3"] = acceleration(pt, torch.zeros((n, 2)), mt, torch.zeros(n)).numpy()
        np.savez_compressed(os.path.join(HERE, "ref_pairlaw.npz"), **out)
        print("written", {k: np.shape(v) for k, v in out.items()})
    finally:
        jax_facade.uninstall(saved)


if __name__ == "__main__":
    main()
