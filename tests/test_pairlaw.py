"""The other dense pairwise kernels of the reference repo: gravity r^-3.
CPU: the C restatement against the closures as written (numpy fp32) and analytic two-body answers.
GPU: ljmd_pair_accel against the C restatement."""
import numpy as np
import pytest


def _bodies(n, seed=0, spread=10.0):
    rng = np.random.default_rng(seed)
    pos = (rng.uniform(-spread, spread, (n, 2))).astype(np.float32)
    mass = rng.uniform(0.5, 30.0, n).astype(np.float32)
    return pos, mass


@pytest.mark.parametrize("n", [2, 3, 5, 17])
def test_oracle_gravity_matches_the_closures_as_written(oracle, n):
    pos, mass = _bodies(n, seed=n)
    G = 1.0
    a_c = oracle.c_gravity(pos, mass, G, "nbody")
    a_py = oracle.gravity_nbody_loops(pos, mass, G)
    assert np.abs(a_c - a_py).max() <= 2e-6 * np.abs(a_py).max()
    e_c = oracle.c_gravity(pos, mass, G, "em3")
    e_py = oracle.gravity_em3_broadcast(pos, mass, G)
    assert np.abs(e_c - e_py).max() <= 2e-6 * np.abs(e_py).max()
    # away from the guards the two laws are the same physics
    assert np.abs(a_c - e_c).max() <= 1e-5 * np.abs(a_c).max()


def test_oracle_gravity_two_body_known_answer(oracle):
    pos = np.array([[0.0, 0.0], [3.0, 4.0]], dtype=np.float32)
    mass = np.array([2.0, 7.0], dtype=np.float32)
    G = 0.5
    for law in ("nbody", "em3"):
        a = oracle.c_gravity(pos, mass, G, law)
        # |a_0| = G m_1 / r^2 towards body 1, r = 5
        assert np.allclose(a[0], G * 7.0 / 25.0 * np.array([0.6, 0.8]), rtol=1e-6)
        assert np.allclose(a[1], -G * 2.0 / 25.0 * np.array([0.6, 0.8]), rtol=1e-6)
    # the guards: coincident bodies contribute nothing (NBODY:63) / a clamped, still zero term (EM3:30)
    pos2 = np.array([[1.0, 1.0], [1.0, 1.0], [2.0, 1.0]], dtype=np.float32)
    m3 = np.ones(3, dtype=np.float32)
    for law in ("nbody", "em3"):
        a = oracle.c_gravity(pos2, m3, 1.0, law)
        assert np.isfinite(a).all() and np.allclose(a[0], [1.0, 0.0]) and np.allclose(a[1], [1.0, 0.0])


@pytest.mark.gpu
@pytest.mark.parametrize("n", [2, 5, 129, 1000, 4096])
def test_gpu_pair_accel_vs_oracle(oracle, n):
    from jax_tpus_benchmark_physics_simulation_b200.pairwise import pairwise_forces, gravity_acceleration
    pos, mass = _bodies(n, seed=n, spread=10.0 * np.sqrt(n))
    if n >= 5:
        pos[3] = pos[1]                       # coincident pair: exercises both guards
    G = 1.5
    a = pairwise_forces(pos, mass, G).numpy()
    a_c = oracle.c_gravity(pos, mass, G, "nbody")
    scale = np.abs(a_c).max()
    # same operations, each rounded once, summed in the same j order: bit-equal up to the sqrt / divide
    # implementations (both correctly rounded on this path) -> a few ulp at most
    assert np.isfinite(a).all() and np.abs(a - a_c).max() <= 2e-6 * scale
    e = gravity_acceleration(pos, mass, G).numpy()
    e_c = oracle.c_gravity(pos, mass, G, "em3")
    assert np.isfinite(e).all() and np.abs(e - e_c).max() <= 2e-6 * np.abs(e_c).max()
    if n <= 5:                                # the reference's own sizes: exact summation order
        assert np.array_equal(a, a_c)


@pytest.mark.gpu
def test_gpu_pair_accel_rejects_bad_arguments():
    from jax_tpus_benchmark_physics_simulation_b200 import _lib
    lib = _lib.load()
    assert lib.ljmd_pair_accel(0, None, None, 4, 1.0, None, None) != 0
    assert lib.ljmd_pair_accel(7, None, None, 4, 1.0, None, None) != 0


# ---- vectors from the reference's own function source (tests/golden/make_pairlaw_golden.py) -----------
import os as _os

_GOLD = _os.path.join(_os.path.dirname(_os.path.abspath(__file__)), "golden", "ref_pairlaw.npz")
_CASES = ["em3_ic", "n2", "n5", "n5_coincident"]


@pytest.mark.parametrize("name", _CASES)
def test_oracle_gravity_vs_reference_function_source(oracle, name):
    g = np.load(_GOLD)
    pos, mass, G = g[f"{name}_pos"], g[f"{name}_mass"], float(g["G"])
    for law in ("nbody", "em3"):
        ref = g[f"{name}_{law}"]
        got = oracle.c_gravity(pos, mass, G, law)
        assert np.isfinite(ref).all()
        assert np.abs(got - ref).max() <= 2e-6 * max(np.abs(ref).max(), 1e-30)
    assert np.abs(oracle.gravity_nbody_loops(pos, mass, G) - g[f"{name}_nbody"]).max() <= 2e-6 * np.abs(g[f"{name}_nbody"]).max()


@pytest.mark.gpu
@pytest.mark.parametrize("name", _CASES)
def test_gpu_pair_accel_vs_reference_function_source(name):
    from jax_tpus_benchmark_physics_simulation_b200.pairwise import pairwise_forces, gravity_acceleration
    g = np.load(_GOLD)
    pos, mass, G = g[f"{name}_pos"], g[f"{name}_mass"], float(g["G"])
    a = pairwise_forces(pos, mass, G).numpy()
    e = gravity_acceleration(pos, mass, G).numpy()
    assert np.abs(a - g[f"{name}_nbody"]).max() <= 2e-6 * np.abs(g[f"{name}_nbody"]).max()
    assert np.abs(e - g[f"{name}_em3"]).max() <= 2e-6 * np.abs(g[f"{name}_em3"]).max()
