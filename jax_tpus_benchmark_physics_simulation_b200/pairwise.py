"""Host-side mirror of the other dense pairwise closures of the reference repo,
served by the same C ABI (``ljmd_pair_accel``, csrc/pairlaw.cu: the all-pairs tiling with the pair law
as a template functor):

* ``pairwise_forces(positions, masses)`` — nbody_bh_merger_sim_single-host_workload.py NBODY:54-67
* ``gravity_acceleration(pos, masses)``  — the gravity term of ``acceleration`` in
  three_particles_em_nonuni_single-host_workload.py EM3:25-38

No CPU fallback: without the built library / a B200 these raise.
"""
from __future__ import annotations

import numpy as np
import torch

from . import _lib
from .md import DeviceArray


def _accel(law: int, positions, masses, G: float) -> DeviceArray:
    if not torch.cuda.is_available():
        raise _lib.LjmdError("no CUDA device: the pairwise kernels have no CPU fallback")
    lib = _lib.load()

    def dev(x):
        if isinstance(x, DeviceArray):
            x = x.tensor
        if not isinstance(x, torch.Tensor):
            x = torch.from_numpy(np.ascontiguousarray(x, dtype=np.float32))
        return x.to("cuda", dtype=torch.float32).contiguous()

    pos, m = dev(positions), dev(masses)
    if pos.dim() != 2 or pos.shape[1] != 2 or m.shape != (pos.shape[0],):
        raise ValueError(f"expected positions (n,2) and masses (n,), got {tuple(pos.shape)} / {tuple(m.shape)}")
    acc = torch.empty_like(pos)
    stream = torch.cuda.current_stream(pos.device).cuda_stream
    _lib.check(lib.ljmd_pair_accel(law, pos.data_ptr(), m.data_ptr(), pos.shape[0], float(G),
                                   acc.data_ptr(), stream), "ljmd_pair_accel")
    return DeviceArray(acc)


def pairwise_forces(positions, masses, G: float = 1.0) -> DeviceArray:
    """acc[i] = sum_{j != i} where(r >= 1e-6, G m_j / r^3, 0) (pos_j - pos_i)   (NBODY:54-67)."""
    return _accel(_lib.LJMD_LAW_GRAVITY_NBODY, positions, masses, G)


def gravity_acceleration(pos, masses, G: float = 1.0) -> DeviceArray:
    """acc[i] = sum_j G m_j (pos_j - pos_i) max(r^2 + [i == j], 1e-12)^(-3/2)   (EM3:25-38)."""
    return _accel(_lib.LJMD_LAW_GRAVITY_EM3, pos, masses, G)
