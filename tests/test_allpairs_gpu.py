"""GPU parity tests of the all-pairs path (through the C ABI) against the CPU oracle.

Tolerances (BASELINE.json north_star / SURVEY.md §8c):
  forces       max_i |F - F_ref|_inf / max_i |F_ref|_inf <= 1e-5   vs the fp32 oracle
  total energy relative <= 1e-6 after one step
  trajectory   max |dR| <= 1e-4 after 200 steps at dt = 0.005, N = 400 (before chaotic divergence)
"""
import numpy as np
import pytest
import torch

from jax_tpus_benchmark_physics_simulation_b200 import lattice_jitter

pytestmark = pytest.mark.gpu

FORCE_TOL = 1e-5
ENERGY_TOL = 1e-6


def _sim(N, **kw):
    from jax_tpus_benchmark_physics_simulation_b200.md import LJSimulation
    kw.setdefault("path", "allpairs")
    return LJSimulation(N, **kw)


def _rel_force_err(F, Fref):
    return float(np.abs(F - Fref).max() / np.abs(Fref).max())


@pytest.mark.parametrize("N", [400, 4096])
@pytest.mark.parametrize("rc", [None, 2.5])
@pytest.mark.parametrize("seed", [0, 1])
def test_forces_and_energy_vs_oracle(oracle, N, rc, seed):
    R, V, box = lattice_jitter(N, seed=seed)
    sim = _sim(N, rc=rc)
    F, pe = sim.force_and_energy(R)
    F = F.numpy()
    # fp32 autodiff restatement (closest analogue of jit(grad(total_energy_fn)), MD:64)
    Fa = oracle.force_autodiff(torch.from_numpy(R), float(box), rc=rc).numpy()
    Fc, pe_c = oracle.c_forces(R, box, rc=rc)
    assert _rel_force_err(F, Fa) <= FORCE_TOL
    assert _rel_force_err(F, Fc) <= FORCE_TOL
    assert abs(float(pe) - pe_c) <= ENERGY_TOL * abs(pe_c)
    # total_energy_fn alone (MD:50-62)
    pe2 = float(sim.total_energy_fn(R))
    assert pe2 == float(pe)
    # force_fn alone gives the same bits as force_and_energy
    assert np.array_equal(sim.force_fn(R).numpy(), F)


def test_net_force_zero_and_lattice(oracle):
    """SURVEY §8c KAT (2): sum of forces vanishes to rounding; perfect lattice has F_i ~ 0."""
    N = 1024
    R, V, box = lattice_jitter(N, seed=3)
    sim = _sim(N, rc=2.5)
    F = sim.force_fn(R).numpy()
    assert np.abs(F.sum(axis=0)).max() <= 1e-3 * np.abs(F).max()
    # perfect lattice: a*sqrt(5) == rc == 2.5 exactly at rho = 0.8, so use the reference's
    # no-cutoff form here (with rc the (1,2) neighbours sit on the truncation edge)
    Rl, _, _ = lattice_jitter(N, seed=0, jitter=0.0)
    sim0 = _sim(N, rc=None)
    Fl = sim0.force_fn(Rl).numpy()
    assert np.abs(Fl).max() < 5e-4
    # ... and with the cutoff the pair set must still be the oracle's, bit for bit in r2
    Fl_rc = sim.force_fn(Rl).numpy()
    Fc, _ = oracle.c_forces(Rl, box, rc=2.5)
    assert np.abs(Fl_rc - Fc).max() <= 1e-5 * max(np.abs(Fc).max(), 1.0)


def test_two_particles_across_seam():
    """SURVEY §8c KAT (1): F = 24(2 r^-12 - r^-6)/r along x, also across the periodic seam."""
    N = 2
    box = np.float32(10.0)
    sim = _sim(N, box_size=box, rc=None)
    for r in (1.0, 2.0 ** (1.0 / 6.0), 1.5):
        for x0 in (3.0, 9.6):          # second case wraps through x = box
            R = np.array([[x0, 5.0], [np.float32((x0 + r) % 10.0), 5.0]], dtype=np.float32)
            rr = float(np.float32(R[1, 0]) - np.float32(R[0, 0]))
            rr = rr - 10.0 * round(rr / 10.0)
            F, pe = sim.force_and_energy(R)
            F = F.numpy()
            fmag = 24.0 * (2.0 * rr ** -12 - rr ** -6) / abs(rr)
            assert abs(F[0, 0] + np.sign(rr) * fmag) <= 5e-5 * max(1.0, abs(fmag))
            assert abs(F[1, 0] - np.sign(rr) * fmag) <= 5e-5 * max(1.0, abs(fmag))
            assert abs(F[0, 1]) < 1e-6 and abs(F[1, 1]) < 1e-6
            assert abs(float(pe) - 4.0 * (rr ** -12 - rr ** -6)) <= 2e-5


def test_box_edge_particle_equivalence():
    """SURVEY §8c KAT (3): a particle at x == box and one at x == 0 act identically."""
    N = 400
    R, V, box = lattice_jitter(N, seed=0)
    sim = _sim(N, rc=2.5)
    Ra, Rb = R.copy(), R.copy()
    Ra[0, 0] = np.float32(0.0)
    Rb[0, 0] = box
    Fa, Fb = sim.force_fn(Ra).numpy(), sim.force_fn(Rb).numpy()
    assert np.abs(Fa[1:] - Fb[1:]).max() <= 1e-5 * np.abs(Fa).max()


def test_half_box_tie_follows_round_half_even(oracle):
    """SURVEY §8c KAT (4): dr = +-box/2 -> round(0.5) = 0 (half-even): no wrap."""
    box = np.float32(8.0)
    R = np.array([[1.0, 1.0], [5.0, 1.0], [1.0, 5.0], [5.0, 5.0]], dtype=np.float32)
    sim = _sim(4, box_size=box, rc=None)
    F = sim.force_fn(R).numpy()
    Fc, _ = oracle.c_forces(R, box, rc=None)
    assert np.abs(F - Fc).max() <= 1e-6 * max(1e-30, np.abs(Fc).max()) + 1e-12


@pytest.mark.parametrize("N,rc", [(400, None), (400, 2.5), (1024, 2.5), (4096, 2.5)])
def test_one_step_vs_oracle(oracle, N, rc):
    """SURVEY §8c KAT (5): one verlet_step vs the restatement; total energy within 1e-6."""
    dt = 0.005
    for seed in (0, 1, 2):
        R, V, box = lattice_jitter(N, seed=seed)
        sim = _sim(N, rc=rc, dt=dt)
        (R1, V1), _ = sim.run((R, V), 1, energy_every=1)
        ke_pe = sim.last_energies.numpy()
        Rc, Vc, _, ke_pe_c = oracle.c_run(R, V, box, dt, 1, rc=rc, energy_every=1)
        assert np.abs(R1.numpy() - Rc).max() <= 2e-6
        assert np.abs(V1.numpy() - Vc).max() <= 1e-5 * np.abs(Vc).max()
        e_gpu = float(ke_pe[0, 0]) + float(ke_pe[0, 1])
        e_ref = ke_pe_c[0, 0] + ke_pe_c[0, 1]
        assert abs(e_gpu - e_ref) <= ENERGY_TOL * abs(e_ref)
        # verlet_step closure == run(1)
        R1b, V1b = sim.verlet_step((R, V))
        assert np.array_equal(R1b.numpy(), R1.numpy()) and np.array_equal(V1b.numpy(), V1.numpy())


def test_short_trajectory_vs_oracle(oracle):
    """Short-horizon trajectory agreement before chaotic divergence."""
    N, dt, rc = 400, 0.005, 2.5
    R, V, box = lattice_jitter(N, seed=0)
    sim = _sim(N, rc=rc, dt=dt)
    (R200, V200), traj = sim.run((R, V), 200, sample_every=50)
    Rc, Vc, traj_c, _ = oracle.c_run(R, V, box, dt, 200, sample_every=50, rc=rc)
    d = np.abs(R200.numpy() - Rc)
    d = np.minimum(d, float(box) - d)            # periodic distance
    assert d.max() <= 1e-4
    t = traj.numpy()
    assert t.shape == traj_c.shape == (4, N, 2)
    d0 = np.abs(t[0] - traj_c[0])
    assert np.minimum(d0, float(box) - d0).max() <= 1e-6     # snapshot after step 1 (i = 0)


def test_sampling_rule_and_dropped_sample():
    """MD:88-100: S = nsteps // sample_every rows; snapshots after steps 1, 1+k, ...;
    the out-of-range sample when nsteps % sample_every != 0 is silently dropped."""
    N, dt = 400, 0.002
    R, V, box = lattice_jitter(N, seed=1)
    sim = _sim(N, rc=2.5, dt=dt)
    (_, _), traj = sim.run((R, V), 25, sample_every=10)      # S = 2; i = 20 would be row 2: dropped
    t = traj.numpy()
    assert t.shape == (2, N, 2)
    (R1, _), _ = sim.run((R, V), 1)
    (R11, _), _ = sim.run((R, V), 11)
    assert np.array_equal(t[0], R1.numpy())
    assert np.array_equal(t[1], R11.numpy())


def test_run_is_deterministic_and_composable():
    """Same inputs -> same bits; run(a+b) == run(b) o run(a) (F(R_new) carried == recomputed)."""
    N, dt = 1024, 0.005
    R, V, box = lattice_jitter(N, seed=2)
    sim = _sim(N, rc=2.5, dt=dt)
    (Ra, Va), _ = sim.run((R, V), 40)
    (Rb, Vb), _ = sim.run((R, V), 40)
    assert np.array_equal(Ra.numpy(), Rb.numpy()) and np.array_equal(Va.numpy(), Vb.numpy())
    (R1, V1), _ = sim.run((R, V), 15)
    (R2, V2), _ = sim.run((R1, V1), 25)
    assert np.array_equal(R2.numpy(), Ra.numpy()) and np.array_equal(V2.numpy(), Va.numpy())
    # inputs are not modified (the reference closures are pure)
    Rd = torch.from_numpy(R).cuda()
    Rd0 = Rd.clone()
    sim.run((Rd, V), 5)
    assert torch.equal(Rd, Rd0)


@pytest.mark.parametrize("N", [400, 1024, 4096])
def test_kernel_variants_and_launch_chunking(N, monkeypatch):
    """N = 400 runs in one thread-block cluster (mode 4), N = 1024 in the ordered grid kernel (mode 1),
    N = 4096 in the Newton's-third-law tiles (mode 3).  A call split into several launches (state
    handed over through global memory) gives the same bits as one launch, with samples and energies."""
    R, V, box = lattice_jitter(N, seed=3)
    sim = _sim(N, rc=2.5, dt=0.005)
    assert sim.allpairs_mode() == {400: 4, 1024: 1, 4096: 3}[N]
    (Ra, Va), ta = sim.run((R, V), 45, sample_every=10, energy_every=5)
    ea = sim.last_energies.numpy().copy()
    monkeypatch.setenv("LJMD_AP_CHUNK", "7")
    (Rb, Vb), tb = sim.run((R, V), 45, sample_every=10, energy_every=5)
    eb = sim.last_energies.numpy().copy()
    monkeypatch.delenv("LJMD_AP_CHUNK")
    assert np.array_equal(Ra.numpy(), Rb.numpy()) and np.array_equal(Va.numpy(), Vb.numpy())
    assert np.array_equal(ta.numpy(), tb.numpy()) and np.array_equal(ea, eb)


def test_cluster_kernel_matches_grid_kernel(monkeypatch):
    """The single-cluster kernel (predicate-free pair loop, slices summed by a shuffle butterfly) and the
    grid kernel evaluate the same ordered pairs: forces to rounding, 100-step trajectory before chaos."""
    N = 400
    R, V, box = lattice_jitter(N, seed=4)
    clu = _sim(N, rc=None, dt=0.001)
    monkeypatch.setenv("LJMD_AP_CLUSTER_NMAX", "0")
    grid = _sim(N, rc=None, dt=0.001)
    monkeypatch.delenv("LJMD_AP_CLUSTER_NMAX")
    assert clu.allpairs_mode() == 4 and grid.allpairs_mode() == 1
    Fc, pc = clu.force_and_energy(R)
    Fg, pg = grid.force_and_energy(R)
    assert np.abs(Fc.numpy() - Fg.numpy()).max() <= 2e-6 * np.abs(Fg.numpy()).max()
    assert abs(float(pc) - float(pg)) <= 1e-6 * abs(float(pg))
    (Rc, Vc), _ = clu.run((R, V), 100, energy_every=10)
    ec = clu.last_energies.numpy().copy()
    (Rg, Vg), _ = grid.run((R, V), 100, energy_every=10)
    eg = grid.last_energies.numpy()
    d = np.abs(Rc.numpy() - Rg.numpy()); d = np.minimum(d, float(box) - d)
    assert d.max() <= 1e-4
    assert np.abs(ec.sum(1) - eg.sum(1)).max() <= 2e-6 * np.abs(eg.sum(1)).max()


def test_momentum_and_energy_conservation(oracle):
    """SURVEY §8c KAT (6) + drift no worse than the reference algorithm's (App. B.5)."""
    N, dt, rc = 400, 0.001, None
    R, V, box = lattice_jitter(N, seed=0)
    sim = _sim(N, rc=rc, dt=dt)
    (R1, V1), _ = sim.run((R, V), 1000, energy_every=100)
    e = sim.last_energies.numpy().astype(np.float64).sum(axis=1)
    p0, p1 = V.astype(np.float64).sum(axis=0), V1.numpy().astype(np.float64).sum(axis=0)
    assert np.abs(p1 - p0).max() <= 2e-3
    _, _, _, ke_pe_c = oracle.c_run(R, V, box, dt, 1000, rc=rc, energy_every=100)
    e_c = ke_pe_c.sum(axis=1)
    drift = abs(e[-1] - e[0]) / abs(e[0])
    drift_c = abs(e_c[-1] - e_c[0]) / abs(e_c[0])
    assert drift <= max(2.0 * drift_c, 2e-5)
    pos = R1.numpy()
    assert pos.min() >= 0.0 and pos.max() <= float(box)      # closed interval [0, box], MD:72


def test_thermostat_rescale(oracle):
    N, dt, rc = 400, 0.002, 2.5
    R, V, box = lattice_jitter(N, seed=0)
    sim = _sim(N, rc=rc, dt=dt, thermostat_kT=0.7, thermostat_every=10)
    (R1, V1), _ = sim.run((R, V), 30)
    kT = 0.5 * float((V1.numpy().astype(np.float64) ** 2).sum()) / N
    assert abs(kT - 0.7) < 1e-4                              # rescaled on the last step
    Rc, Vc, _, _ = oracle.c_run(R, V, box, dt, 30, rc=rc, thermostat_kT=0.7, thermostat_every=10)
    assert np.abs(V1.numpy() - Vc).max() <= 1e-4 * np.abs(Vc).max()


def test_large_all_pairs_rows_vs_oracle(oracle):
    """N = 65,536 (config 3): a random subset of i rows against all j."""
    N, rc = 65536, 2.5
    R, V, box = lattice_jitter(N, seed=0)
    sim = _sim(N, rc=rc)
    F, pe = sim.force_and_energy(R)
    F = F.numpy()
    rng = np.random.default_rng(0)
    worst = 0.0
    fmax = np.abs(F).max()
    for i0 in rng.integers(0, N - 64, size=6):
        Fc, _ = oracle.c_forces(R, box, rc=rc, rows=(int(i0), int(i0) + 64))
        worst = max(worst, np.abs(F[i0:i0 + 64] - Fc).max() / fmax)
    assert worst <= FORCE_TOL
    _, pe_c = oracle.c_forces_cells(R, box, rc)
    assert abs(float(pe) - pe_c) <= ENERGY_TOL * abs(pe_c)


def test_gr_histogram_vs_oracle(oracle):
    """calculate_g_r (MD:108-129): integer pair counts bit-exact against the CPU recount."""
    N = 400
    R, V, box = lattice_jitter(N, seed=0)
    sim = _sim(N, rc=2.5, dt=0.005)
    (_, _), traj = sim.run((R, V), 40, sample_every=10)
    r_max = np.float32(box / np.float32(2.0))
    nbins = int(r_max / 0.05)
    centers, g = sim.calculate_g_r(traj, N, box, nbins, r_max)
    counts = sim.last_gr_counts.numpy()
    th = traj.numpy()
    for s in range(th.shape[0]):
        assert np.array_equal(counts[s], oracle.c_gr_hist(th[s], box, nbins, r_max))
    c_ref, g_ref, hists, _ = oracle.g_r(torch.from_numpy(th), N, box, nbins, r_max)
    assert np.array_equal(hists, counts)
    assert np.allclose(g.numpy(), g_ref, rtol=1e-6, atol=0)
    assert np.array_equal(centers.numpy(), c_ref)
    assert counts.sum() < th.shape[0] * N * (N - 1) // 2     # corner pairs beyond box/2 dropped
