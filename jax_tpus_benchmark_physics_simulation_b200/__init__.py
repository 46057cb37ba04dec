"""B200-native drop-in for the hot path of molecular_dynamics_jax_single-host_workload.py.

Only what the path needs: ``csrc/`` (hand-written sm_100a kernels + the C ABI of
include/ljmd.h, built in-tree as ``libljmd.so``), ``_lib`` (ctypes binding), ``md``
(host-side mirror of the reference's closures), ``ic`` (initial conditions) and ``driver``
(the reference's CLI).
"""
from .ic import box_size, lattice_jitter, reference_style_uniform  # noqa: F401

__all__ = ["box_size", "lattice_jitter", "reference_style_uniform", "LJSimulation", "DeviceArray"]


def __getattr__(name):
    if name in ("LJSimulation", "DeviceArray", "fp32_peak_probe"):
        from . import md
        return getattr(md, name)
    raise AttributeError(name)
