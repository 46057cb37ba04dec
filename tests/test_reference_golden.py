"""Parity against vectors produced by the REFERENCE'S OWN SOURCE FILE
(molecular_dynamics_jax_single-host_workload.py, executed unmodified on the torch facade of the jax
API: tests/golden/jax_facade.py + make_reference_golden.py -> tests/golden/ref_md_*.npz).

CPU (not gpu): the oracle - torch restatement and strict-fp32 C restatement - is pinned to them.
GPU: the CUDA path through the C ABI / LJSimulation is compared with them directly.

Tolerances (north_star / SURVEY.md §8c): forces 1e-5 of max|F|, total energy 1e-6 relative,
one step ~ulp, trajectories 1e-4 sigma at 200 steps (dt = 0.005), histogram counts exact.
"""
import os

import numpy as np
import pytest
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
CASES = ["n64", "n400", "n4096"]


def _load(case):
    g = np.load(os.path.join(HERE, "golden", f"ref_md_{case}.npz"))
    return {k: g[k] for k in g.files}


def _pd(a, b, box):
    d = np.abs(np.asarray(a, dtype=np.float64) - np.asarray(b, dtype=np.float64))
    return np.minimum(d, float(box) - d)


# ------------------------------------------------------------------ CPU: the oracle is pinned
@pytest.mark.parametrize("case", CASES)
def test_oracle_box_displacement_energy_forces(oracle, case):
    g = _load(case)
    N, box = int(g["N"]), g["box"]
    assert oracle.box_size(N, float(g["rho"])) == np.float32(box)                     # MD:30, bit for bit
    # periodic_displacement incl. the half-box ties (MD:46-48): bit for bit, torch and C
    pd_t = oracle.periodic_displacement(torch.from_numpy(g["pd_in"]), float(box)).numpy()
    assert np.array_equal(pd_t, g["pd_out"])
    pd_c = np.array([oracle.c_lib().orc_periodic_displacement(float(x), float(box)) for x in g["pd_in"]],
                    dtype=np.float32)
    assert np.array_equal(pd_c, g["pd_out"])
    R0 = torch.from_numpy(g["R0"])
    scale = np.abs(g["F0"]).max()
    if N <= 4096:
        E = float(oracle.total_energy(R0, float(box)))
        assert abs(E - float(g["E0"])) <= 1e-6 * abs(float(g["E0"]))
        F = oracle.force_autodiff(R0, float(box)).numpy()
        assert np.abs(F - g["F0"]).max() <= 2e-6 * scale
    Fc, pe_c = oracle.c_forces(g["R0"], box)
    assert np.abs(Fc - g["F0"]).max() <= 1e-5 * scale
    assert abs(pe_c - float(g["E0"])) <= 1e-6 * abs(float(g["E0"]))


@pytest.mark.parametrize("case", CASES)
def test_oracle_step_equilibrate_production(oracle, case):
    g = _load(case)
    box, dt = g["box"], float(g["dt"])
    eq, prod, se = int(g["eq_steps"]), int(g["prod_steps"]), int(g["sample_every"])
    # one verlet_step (MD:66-75)
    R1, V1, _, e1 = oracle.c_run(g["R0"], g["V0"], box, dt, 1, energy_every=1)
    assert _pd(R1, g["R1"], box).max() <= 2e-6 and np.abs(V1 - g["V1"]).max() <= 2e-5
    assert abs(e1[0, 1] - float(g["E1"])) <= 1e-6 * abs(float(g["E1"]))
    # equilibrate_fn (MD:77-83), then production_fn with its sampling rule (MD:85-106)
    Re, Ve, _, _ = oracle.c_run(g["R0"], g["V0"], box, dt, eq)
    tol = 1e-4 if eq <= 200 else 1e-3
    assert _pd(Re, g["R_eq"], box).max() <= tol
    Rf, Vf, traj, _ = oracle.c_run(g["R_eq"], g["V_eq"], box, dt, prod, sample_every=se)
    assert traj.shape == g["R_history"].shape == (prod // se, int(g["N"]), 2)
    assert _pd(traj, g["R_history"], box).max() <= tol
    assert _pd(Rf, g["R_final"], box).max() <= tol and np.abs(Vf - g["V_final"]).max() <= 20 * tol
    if int(g["N"]) <= 400:   # the torch restatement runs the same loop (dense autodiff)
        (Rt, Vt), trj, _ = oracle.run((torch.from_numpy(g["R_eq"]), torch.from_numpy(g["V_eq"])), float(box), dt,
                                      prod, sample_every=se)
        assert _pd(trj.numpy(), g["R_history"], box).max() <= tol


@pytest.mark.parametrize("case", CASES)
def test_oracle_g_r(oracle, case):
    g = _load(case)
    N, box = int(g["N"]), g["box"]
    nb = int(g["gr_nbins"])
    r_max = np.float32(box) / np.float32(2.0)
    assert nb == int(r_max / 0.05)                                                     # MD:158-159
    if N <= 400:
        centers, gr, hists, _ = oracle.g_r(torch.from_numpy(g["R_history"]), N, box, nb, r_max)
        assert np.array_equal(centers, g["gr_centers"])
        assert np.allclose(gr, g["g_r"], rtol=2e-6, atol=0)
    # the C histogram stage on the reference's own snapshots reproduces its g(r) after the reference's
    # normalisation (MD:111-115,126-128)
    counts = np.stack([oracle.c_gr_hist(R, box, nb, r_max) for R in g["R_history"]])
    edges = np.linspace(0, r_max, nb + 1, dtype=np.float32)
    shell = np.float32(np.pi) * (edges[1:] ** 2 - edges[:-1] ** 2)
    ideal = np.float32(N * (N - 1) / 2.0) / (np.float32(box) ** 2) * shell
    gr_c = counts.astype(np.float32).mean(axis=0) / ideal
    assert np.allclose(gr_c, g["g_r"], rtol=2e-6, atol=0)


# ------------------------------------------------------------------ GPU: the CUDA path, directly
@pytest.mark.gpu
@pytest.mark.parametrize("case", CASES)
def test_gpu_closures_vs_reference_vectors(case):
    from jax_tpus_benchmark_physics_simulation_b200.md import LJSimulation
    g = _load(case)
    N, box, dt = int(g["N"]), g["box"], float(g["dt"])
    eq, prod, se = int(g["eq_steps"]), int(g["prod_steps"]), int(g["sample_every"])
    sim = LJSimulation(N, rho=float(g["rho"]), dt=dt, eq_steps=eq, prod_steps=prod, sample_every=se)   # rc=None
    assert np.float32(sim.box_size) == np.float32(box)
    assert np.array_equal(sim.periodic_displacement(g["pd_in"]).numpy(), g["pd_out"])
    scale = np.abs(g["F0"]).max()
    assert abs(float(sim.total_energy_fn(g["R0"])) - float(g["E0"])) <= 1e-6 * abs(float(g["E0"]))
    assert np.abs(sim.force_fn(g["R0"]).numpy() - g["F0"]).max() <= 1e-5 * scale
    R1, V1 = sim.verlet_step((g["R0"], g["V0"]))
    assert _pd(R1.numpy(), g["R1"], box).max() <= 2e-6 and np.abs(V1.numpy() - g["V1"]).max() <= 2e-5
    assert abs(float(sim.total_energy_fn(R1)) - float(g["E1"])) <= 1e-6 * abs(float(g["E1"]))
    st = sim.equilibrate_fn((g["R0"], g["V0"]))
    tol = 1e-4 if eq <= 200 else 1e-3
    assert _pd(st[0].numpy(), g["R_eq"], box).max() <= tol
    (Rf, Vf), hist = sim.production_fn((g["R_eq"], g["V_eq"]))
    assert hist.shape == g["R_history"].shape
    assert _pd(hist.numpy(), g["R_history"], box).max() <= tol
    assert _pd(Rf.numpy(), g["R_final"], box).max() <= tol
    # calculate_g_r on the reference's own snapshots (MD:108-131)
    nb = int(g["gr_nbins"])
    centers, gr = sim.calculate_g_r(g["R_history"], N, box, nb, np.float32(box) / np.float32(2.0))
    assert np.array_equal(centers.numpy(), g["gr_centers"])
    assert np.allclose(gr.numpy(), g["g_r"], rtol=2e-6, atol=0)
