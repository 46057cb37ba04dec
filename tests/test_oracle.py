"""CPU tests that pin the oracle (oracle/): the C restatement and the torch-autodiff
restatement against each other and against analytic known answers.  The
reference ships no golden vectors and JAX is not installable here, so parity with the
reference running on JAX cannot be; these tests and tests/test_reference_golden.py (vectors made by the
reference's own source on a torch facade of jax) are the pin."""
import math
import os

import numpy as np
import pytest
import torch

from jax_tpus_benchmark_physics_simulation_b200 import box_size, lattice_jitter

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def test_box_size_constants():
    # SURVEY Appendix C exact fp32 values (MD:30)
    assert float(box_size(400, 0.8)) == 22.360679626464844
    assert float(box_size(4096, 0.8)) == 71.5541763305664
    assert float(box_size(65536, 0.8)) == 286.2167053222656
    assert float(box_size(4194304, 0.8)) == 2289.733642578125
    assert float(box_size(16777216, 0.8)) == 4579.46728515625


def test_round_and_mod_semantics(oracle):
    c = oracle.c_lib()
    # round-half-even of d/box (MD:48) and divisor-sign mod that can return exactly box (MD:72)
    assert c.orc_periodic_displacement(4.0, 8.0) == 4.0            # q = 0.5 -> 0
    assert c.orc_periodic_displacement(-4.0, 8.0) == -4.0
    assert c.orc_periodic_displacement(4.0000005, 8.0) < 0.0
    assert c.orc_periodic_displacement(12.0, 8.0) == -4.0          # q = 1.5 -> 2
    box = float(np.float32(22.360679626464844))
    assert c.orc_mod(-1e-9, box) == box
    assert c.orc_mod(box, box) == 0.0
    assert abs(c.orc_mod(-0.1, 5.0) - 4.9) < 1e-6
    t = torch.tensor([-2.5, -1.5, -0.5, 0.5, 1.5, 2.5])
    assert torch.round(t).tolist() == [-2.0, -2.0, -0.0, 0.0, 2.0, 2.0]
    for x in (-1e-9, -0.1, 0.0, 3.3, 22.360679626464844, 30.0, -25.0, 50.0):
        a = c.orc_mod(x, box)
        b = float(torch.remainder(torch.tensor(x, dtype=torch.float32),
                                  torch.tensor(box, dtype=torch.float32)))
        assert a == b, x


@pytest.mark.parametrize("N", [400, 1024])
@pytest.mark.parametrize("rc", [None, 2.5])
def test_c_vs_autodiff(oracle, N, rc):
    R, V, box = lattice_jitter(N, seed=1)
    Rt = torch.from_numpy(R)
    Fa = oracle.force_autodiff(Rt, float(box), rc=rc).numpy()
    Fc, pe_c = oracle.c_forces(R, box, rc=rc)
    Fn, _ = oracle.c_forces(R, box, rc=rc, acc_double=False)
    Fan, pe_an = oracle.force_analytic(Rt, float(box), rc=rc)
    m = np.abs(Fa).max()
    assert np.abs(Fc - Fa).max() / m < 1e-6
    assert np.abs(Fn - Fa).max() / m < 5e-6
    assert np.abs(Fan.numpy() - Fa).max() / m < 1e-6
    pe_a = float(oracle.total_energy(Rt, float(box), rc=rc))
    assert abs(pe_a - pe_c) < 2e-6 * abs(pe_c)
    assert abs(pe_an - pe_c) < 1e-9 * abs(pe_c)
    # fp64 referee: fp32 restatement within the survey's measured fp32-vs-fp64 gap
    F64 = oracle.force_autodiff(Rt.double(), float(box), rc=rc).numpy()
    assert np.abs(Fa - F64).max() / m < 2e-4


def test_two_particle_known_answers(oracle):
    box = np.float32(10.0)
    for r in (1.0, 2.0 ** (1.0 / 6.0), 1.5, 2.4):
        for x0 in (3.0, 9.5):
            R = np.array([[x0, 5.0], [(x0 + r) % 10.0, 5.0]], dtype=np.float32)
            rr = float(R[1, 0]) - float(R[0, 0])
            rr -= 10.0 * round(rr / 10.0)
            F, pe = oracle.c_forces(R, box, rc=None)
            fmag = 24.0 * (2.0 * rr ** -12 - rr ** -6) / abs(rr)
            assert abs(F[1, 0] - math.copysign(1.0, rr) * fmag) <= 5e-5 * max(1.0, abs(fmag))
            assert abs(F[0, 0] + math.copysign(1.0, rr) * fmag) <= 5e-5 * max(1.0, abs(fmag))
            assert abs(pe - 4.0 * (rr ** -12 - rr ** -6)) <= 1e-5
    R = np.array([[1.0, 1.0], [1.0 + 2.0 ** (1.0 / 6.0), 1.0]], dtype=np.float32)
    F, _ = oracle.c_forces(R, box, rc=None)
    assert np.abs(F).max() < 5e-5                                  # zero force at the LJ minimum


def test_lattice_zero_force_and_momentum(oracle):
    N = 400
    Rl, _, box = lattice_jitter(N, seed=0, jitter=0.0)
    F, _ = oracle.c_forces(Rl, box, rc=None)
    assert np.abs(F).max() < 2e-4
    R, V, box = lattice_jitter(N, seed=0)
    F, _ = oracle.c_forces(R, box, rc=None)
    assert np.abs(F.astype(np.float64).sum(axis=0)).max() < 1e-3
    R1, V1, _, _ = oracle.c_run(R, V, box, 1e-3, 200)
    assert np.abs(V1.astype(np.float64).sum(0) - V.astype(np.float64).sum(0)).max() < 1e-3


def test_c_step_matches_torch_step(oracle):
    """verlet_step restated twice (C, carried F; torch, two evaluations as MD:66-75)."""
    N, dt = 400, 0.005
    R, V, box = lattice_jitter(N, seed=2)
    ff = lambda r: oracle.force_autodiff(r, float(box), rc=2.5)
    state = (torch.from_numpy(R), torch.from_numpy(V))
    for _ in range(3):
        state = oracle.verlet_step(state, float(box), dt, ff)
    Rc, Vc, _, _ = oracle.c_run(R, V, box, dt, 3, rc=2.5)
    assert np.abs(state[0].numpy() - Rc).max() < 2e-6
    assert np.abs(state[1].numpy() - Vc).max() < 2e-5
    (Rt, Vt), traj, en = oracle.run((torch.from_numpy(R), torch.from_numpy(V)), float(box), dt, 3,
                                    rc=2.5)
    assert np.abs(Rt.numpy() - Rc).max() < 2e-6


def test_sampling_rule(oracle):
    N = 100
    R, V, box = lattice_jitter(N, seed=0)
    _, _, traj, _ = oracle.c_run(R, V, box, 1e-3, 25, sample_every=10, rc=2.5)
    assert traj.shape == (2, N, 2)                                 # i = 20 -> row 2: dropped
    R1, _, _, _ = oracle.c_run(R, V, box, 1e-3, 1, rc=2.5)
    R11, _, _, _ = oracle.c_run(R, V, box, 1e-3, 11, rc=2.5)
    assert np.array_equal(traj[0], R1) and np.array_equal(traj[1], R11)


def test_energy_drift_of_reference_algorithm(oracle):
    N = 400
    R, V, box = lattice_jitter(N, seed=0)
    _, _, _, ke_pe = oracle.c_run(R, V, box, 1e-3, 2000, energy_every=100)
    e = ke_pe.sum(axis=1)
    assert abs(e[-1] - e[0]) / abs(e[0]) < 5e-5                    # survey: ~5e-6


def test_cell_recount_matches_bruteforce(oracle):
    N = 1024
    R, V, box = lattice_jitter(N, seed=0)
    R[0] = (0.0, box)                                              # closed-interval edge (MD:72)
    cnt = oracle.c_neighbor_count(R, box, 2.8)
    d = R[:, None, :] - R[None, :, :]
    d = d - box * np.round(d / box)
    r2 = (d[..., 0] * d[..., 0] + d[..., 1] * d[..., 1]).astype(np.float32)
    brute = (r2 < np.float32(2.8) * np.float32(2.8)).sum(axis=1) - 1
    assert np.array_equal(cnt, brute.astype(np.int32))
    Fcell, pe_cell = oracle.c_forces_cells(R, box, 2.5)
    Fall, pe_all = oracle.c_forces(R, box, rc=2.5)
    assert np.array_equal(Fcell, Fall) or np.abs(Fcell - Fall).max() < 1e-5 * np.abs(Fall).max()
    assert abs(pe_cell - pe_all) < 1e-9 * abs(pe_all)


def test_gr_restatements_agree(oracle):
    N = 256
    R, V, box = lattice_jitter(N, seed=0)
    r_max = np.float32(box / np.float32(2.0))
    nbins = int(r_max / 0.05)
    c, g, hists, edges = oracle.g_r(torch.from_numpy(R[None]), N, box, nbins, r_max)
    counts = oracle.c_gr_hist(R, box, nbins, r_max)
    assert np.array_equal(hists[0], counts)
    assert g.shape == (nbins,) and c.shape == (nbins,)


def test_golden_fixture(oracle):
    """tests/golden/lj_golden_n64.npz is written by tests/golden/make_golden.py from the torch
    autodiff restatement; the C restatement must reproduce it."""
    z = np.load(os.path.join(GOLDEN, "lj_golden_n64.npz"))
    R, V, box = z["R"], z["V"], np.float32(z["box"])
    for rc, key in ((None, "nocut"), (2.5, "rc25")):
        F, pe = oracle.c_forces(R, box, rc=rc)
        m = np.abs(z[f"F_{key}"]).max()
        assert np.abs(F - z[f"F_{key}"]).max() / m < 1e-6
        assert abs(pe - float(z[f"pe_{key}"])) < 2e-6 * abs(pe)
    R10, V10, _, _ = oracle.c_run(R, V, box, 0.005, 10, rc=2.5)
    assert np.abs(R10 - z["R10_rc25"]).max() < 5e-6
    assert np.abs(V10 - z["V10_rc25"]).max() < 5e-5


def test_initial_conditions_reference_style_uniform():
    """MD:133-135: R ~ U[0,1) * box, V ~ N(0,1) * sqrt(kT) (statistically; threefry bit parity with
    jax.random is version dependent)."""
    from jax_tpus_benchmark_physics_simulation_b200 import box_size, reference_style_uniform
    N, rho, kT = 40000, 0.8, 1.7
    R, V, box = reference_style_uniform(N, rho, kT, seed=42)
    assert R.dtype == np.float32 and V.dtype == np.float32 and R.shape == (N, 2) == V.shape
    assert box == box_size(N, rho) and box.dtype == np.float32
    assert R.min() >= 0.0 and R.max() < float(box)
    assert abs(R.mean() / float(box) - 0.5) < 0.01                       # uniform: mean 1/2, var 1/12
    assert abs(R.var() / float(box) ** 2 - 1.0 / 12.0) < 0.002
    assert abs(V.mean()) < 0.02 and abs(V.var() / kT - 1.0) < 0.02       # Maxwell: <v^2> = kT per component
    R2, V2, _ = reference_style_uniform(N, rho, kT, seed=42)
    assert np.array_equal(R, R2) and np.array_equal(V, V2)               # seeded
    R3, _, _ = reference_style_uniform(N, rho, kT, seed=43)
    assert not np.array_equal(R, R3)
    # nearest pair of the uniform placement is far inside the LJ core (why the run overflows)
    sub = R[:2000].astype(np.float64)
    d = sub[:, None, :] - sub[None, :, :]
    d -= float(box) * np.round(d / float(box))
    r2 = (d ** 2).sum(-1) + np.eye(len(sub)) * 1e9
    assert r2.min() < 0.25 ** 2 * (N / 2000)


def test_lattice_jitter_shape_and_seed():
    from jax_tpus_benchmark_physics_simulation_b200 import lattice_jitter
    R, V, box = lattice_jitter(1024, seed=5)
    assert R.shape == (1024, 2) and R.min() >= 0.0 and R.max() <= float(box)
    a = float(box) / 32
    cell = np.floor(R / a).astype(int)
    assert len({(int(i), int(j)) for i, j in cell}) == 1024                 # one particle per lattice cell
    with pytest.raises(ValueError):
        lattice_jitter(1000)
