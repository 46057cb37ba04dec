"""Device-side run status (ljmd_check), input wrapping and the documented small-system deviation.

The persistent kernels run asynchronously, so a Verlet list / slab that overflows or a barrier that
times out can only be reported where the caller synchronises: ``block_until_ready`` / ``numpy()`` of
any returned array (-> ``LJSimulation.check`` -> ``ljmd_check``).  None of these tests calls
``last_run_ms()``.
"""
import os

import numpy as np
import pytest
import torch

from jax_tpus_benchmark_physics_simulation_b200 import lattice_jitter
from jax_tpus_benchmark_physics_simulation_b200.ic import box_size

pytestmark = pytest.mark.gpu


def _sim(N, **kw):
    from jax_tpus_benchmark_physics_simulation_b200.md import LJSimulation
    return LJSimulation(N, **kw)


def _dense(N, rho, seed=0):
    n = int(round(np.sqrt(N)))
    box = box_size(N, rho)
    a = float(box) / n
    rng = np.random.default_rng(seed)
    g = (np.stack(np.meshgrid(np.arange(n), np.arange(n), indexing="ij"), -1).reshape(-1, 2) + 0.5) * a
    R = np.mod(g + rng.uniform(-0.05, 0.05, g.shape) * a, float(box)).astype(np.float32)
    V = rng.standard_normal((N, 2)).astype(np.float32)
    return R, V, box


def test_list_overflow_raises_from_run():
    """rho = 2.0, rc = 3.5 (+ skin 0.3): ~90 neighbours per particle on average, more than a Verlet
    list holds -> the run must not hand back silently truncated physics."""
    from jax_tpus_benchmark_physics_simulation_b200._lib import LjmdError, LJMD_E_OVERFLOW
    N = 16384
    R, V, box = _dense(N, 2.0)
    sim = _sim(N, rho=2.0, rc=3.5, dt=1e-4, path="cells")
    (R1, V1), _ = sim.run((R, V), 3)
    with pytest.raises(LjmdError) as ei:
        R1.block_until_ready()
    assert ei.value.code == LJMD_E_OVERFLOW and "dense" in str(ei.value)
    with pytest.raises(LjmdError):
        V1.numpy()
    with pytest.raises(LjmdError):
        sim.force_fn(R).numpy()
    # the flag belongs to a call, not to the handle: a system that fits runs clean afterwards
    ok = _sim(N, rho=0.8, rc=2.5, dt=0.005, path="cells")
    Rk, Vk, _ = lattice_jitter(N, seed=0)
    (R2, _), _ = ok.run((Rk, Vk), 3)
    R2.block_until_ready()
    ok.check()


def test_status_is_per_call():
    """The overflow flag belongs to a call, not to the handle: a clumped configuration (all particles
    squeezed into 16 % of the box: ~120 neighbours within rc + skin) overflows the lists, the regular
    configuration on the SAME handle then runs clean."""
    from jax_tpus_benchmark_physics_simulation_b200._lib import LjmdError
    N = 16384
    R, V, box = lattice_jitter(N, seed=0)
    sim = _sim(N, rc=2.5, dt=0.005, path="cells")
    clumped = (R * np.float32(0.4)).astype(np.float32)
    Fc = sim.force_fn(clumped)
    with pytest.raises(LjmdError):
        Fc.block_until_ready()
    F = sim.force_fn(R)
    F.block_until_ready()                                       # the flags were cleared when the call started
    (R1, _), _ = sim.run((R, V), 3)
    R1.block_until_ready()


@pytest.mark.parametrize("path,N", [("allpairs", 400), ("allpairs", 4096), ("cells", 16384)])
def test_positions_outside_the_box_are_wrapped(oracle, path, N):
    """The reference closures are periodic in R (MD:46-48); ljmd.h: coordinates outside [0, box] are
    wrapped with jnp.mod semantics on load, coordinates inside pass bit for bit."""
    R, V, box = lattice_jitter(N, seed=0)
    sim = _sim(N, rc=2.5, dt=0.005, path=path)
    F0 = sim.force_fn(R).numpy()
    shift = np.zeros_like(R)
    shift[::3, 0] = float(box)
    shift[1::3, 1] = -2.0 * float(box)
    Rs = (R.astype(np.float64) + shift).astype(np.float32)
    F1 = sim.force_fn(Rs).numpy()
    Fo, _ = oracle.c_forces(np.remainder(Rs, box).astype(np.float32), box, rc=2.5)
    scale = np.abs(Fo).max()
    assert np.abs(F1 - Fo).max() / scale <= 1e-5
    # (R + box) - box is not R bit for bit in fp32: the shifted coordinates lose up to 2 ulp(2 box) = 6e-5
    # sigma, which the stiffest pair (dF/dr ~ 500) turns into a few 1e-4 of max|F|
    assert np.abs(F1 - F0).max() / scale <= 2e-3
    (R1, _), _ = sim.run((Rs, V), 2)
    pos = R1.numpy()
    assert pos.min() >= 0.0 and pos.max() <= float(box)


def test_periodic_displacement_closure():
    """MD:46-48 as a stand-alone closure: dr - box * round(dr / box), round-half-even, any shape."""
    sim = _sim(400, rc=2.5)
    box = np.float32(sim.box_size)
    rng = np.random.default_rng(0)
    dr = (rng.uniform(-1.0, 1.0, (7, 5, 2)) * float(box)).astype(np.float32)
    dr[0, 0, 0] = box / np.float32(2.0)                     # the half-box tie rounds to even (0)
    dr[0, 0, 1] = -box / np.float32(2.0)
    got = sim.periodic_displacement(dr).numpy()
    want = dr - box * np.round(dr / box)                    # numpy: same half-even rule, fp32
    assert got.shape == dr.shape and np.array_equal(got, want.astype(np.float32))
    assert got[0, 0, 0] == dr[0, 0, 0] and got[0, 0, 1] == dr[0, 0, 1]


def test_cluster_kernel_r2min_clamp_documented(oracle, monkeypatch):
    """the single-cluster kernel (N <= 640, force-only steps) clamps r^2 from below at
    r2min ~ 1.1e-5 instead of masking i == j.  For two DISTINCT particles 0.002 sigma apart the
    reference's own force overflows fp32 (inf, then NaN everywhere); the cluster kernel returns a
    finite capped force on that pair.  Every other particle agrees with the oracle, and the energy
    variant (exact index mask) behaves exactly like the reference."""
    N, rc = 400, 2.5
    R, V, box = lattice_jitter(N, seed=0)
    R[1] = R[0] + np.array([0.002, 0.0], dtype=np.float32)
    clu = _sim(N, rc=rc, path="allpairs")
    assert clu.allpairs_mode() == 4
    Fc = clu.force_fn(R).numpy()
    Fo, _ = oracle.c_forces(R, box, rc=rc)
    assert not np.isfinite(Fo[:2]).all()                    # the reference formulation overflows here
    assert np.isfinite(Fc).all() and abs(Fc[0, 0]) > 1e30 and Fc[0, 0] * Fc[1, 0] < 0.0
    rest = np.abs(Fc[2:] - Fo[2:]).max() / np.abs(Fo[2:]).max()
    assert rest <= 1e-5
    Fe, pe = clu.force_and_energy(R)
    # energy steps keep the exact index mask: the close pair's force overflows like the reference's
    # (MD:56-64) and its finite pair energy 4 r^-12 ~ 1e33 dominates PE exactly as in the oracle
    _, pe_o = oracle.c_forces(R, box, rc=rc)
    assert not np.isfinite(Fe.numpy()[:2]).all()
    assert abs(float(pe) - pe_o) <= 1e-5 * abs(pe_o)
    # the grid kernel (what N > 640 runs) keeps the index mask on every step
    monkeypatch.setenv("LJMD_AP_IPT", "1")
    grid = _sim(N, rc=rc, path="allpairs")
    assert grid.allpairs_mode() == 1
    Fg = grid.force_fn(R).numpy()
    assert not np.isfinite(Fg[:2]).all()
    assert np.abs(Fg[2:] - Fo[2:]).max() / np.abs(Fo[2:]).max() <= 1e-5
