"""TEST INFRASTRUCTURE ONLY — torch-CPU restatement of the reference hot path.

Op-for-op restatement of
    /root/reference/molecular_dynamics_jax_single-host_workload.py   (cited as MD:<line>)
using torch CPU tensors: ``torch.round`` is round-half-to-even like ``jnp.round``,
``torch.remainder`` takes the divisor's sign like ``jnp.mod`` and ``torch.func.grad`` gives a
zero gradient through ``round`` and respects the double-``where``.  The
dense autodiff force below is the closest available analogue of ``jit(grad(total_energy_fn))``.

PARITY PIN: the reference has no tests / golden vectors and JAX / XLA cannot be installed in this
image, so neither this file nor oracle/lj_oracle.c can be checked against outputs of the reference
running on JAX.  What they ARE pinned to (tests/test_reference_golden.py): vectors produced by the
reference's OWN SOURCE FILE executed unmodified on a torch facade of the jax API
(tests/golden/jax_facade.py, make_reference_golden.py -> tests/golden/ref_md_*.npz) - box size and
periodic_displacement bit for bit, energy, autodiff forces, one step, equilibration, production with
its sampling rule and dropped sample, g(r) - plus each other and analytic known answers
(tests/test_oracle.py).  Not covered by that pin: XLA's own reduction order and pow lowering.

Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs may import this module;
the product package never does.
"""
from __future__ import annotations

import ctypes
import math
import os
import subprocess

import numpy as np
import torch

_HERE = os.path.dirname(os.path.abspath(__file__))


# --------------------------------------------------------------------------- scalars
def box_size(N: int, rho: float) -> np.float32:
    """MD:30  box_size = jnp.sqrt(N / rho): python-float quotient, fp32 sqrt."""
    return np.sqrt(np.float32(N / rho), dtype=np.float32)


# --------------------------------------------------------------------------- energy / force
def periodic_displacement(dr: torch.Tensor, box) -> torch.Tensor:
    """MD:46-48."""
    return dr - box * torch.round(dr / box)


def total_energy(R: torch.Tensor, box, sigma=1.0, epsilon=1.0, rc=None) -> torch.Tensor:
    """MD:50-62 verbatim (dense N x N).  rc is NOT in the reference: plain truncation mask."""
    N = R.shape[0]
    box = torch.as_tensor(box, dtype=R.dtype)
    dr = R[:, None, :] - R[None, :, :]                       # MD:51
    dr = periodic_displacement(dr, box)                      # MD:52
    r_sq = torch.sum(dr ** 2, dim=-1)                        # MD:53
    mask = ~torch.eye(N, dtype=torch.bool)                   # MD:54
    r_sq_safe = torch.where(mask, r_sq, torch.ones((), dtype=R.dtype))   # MD:55
    s2 = (sigma ** 2) / r_sq_safe                            # MD:56
    s6 = s2 ** 3                                             # MD:57
    s12 = s6 ** 2                                            # MD:58
    e = 4.0 * epsilon * (s12 - s6)                           # MD:59
    if rc is not None and math.isfinite(rc):
        rc2 = torch.as_tensor(np.float32(rc) * np.float32(rc), dtype=R.dtype)
        mask = mask & (r_sq < rc2)
    e = torch.where(mask, e, torch.zeros((), dtype=R.dtype))  # MD:60
    return 0.5 * torch.sum(e)                                # MD:61


def force_autodiff(R: torch.Tensor, box, sigma=1.0, epsilon=1.0, rc=None) -> torch.Tensor:
    """MD:64  grad(lambda R: -total_energy_fn(R))."""
    return torch.func.grad(lambda r: -total_energy(r, box, sigma, epsilon, rc))(R)


def _pair_blocks(Ri, Rj, box, sigma, epsilon, rc, i_off, j_off):
    dr = Ri[:, None, :] - Rj[None, :, :]
    dr = periodic_displacement(dr, box)
    r2 = torch.sum(dr ** 2, dim=-1)
    ii = torch.arange(Ri.shape[0])[:, None] + i_off
    jj = torch.arange(Rj.shape[0])[None, :] + j_off
    mask = ii != jj
    if rc is not None and math.isfinite(rc):
        rc2 = torch.as_tensor(np.float32(rc) * np.float32(rc), dtype=Ri.dtype)
        mask = mask & (r2 < rc2)
    r2s = torch.where(mask, r2, torch.ones((), dtype=Ri.dtype))
    s2 = (sigma ** 2) / r2s
    s6 = s2 ** 3
    s12 = s6 ** 2
    zero = torch.zeros((), dtype=Ri.dtype)
    e = torch.where(mask, 4.0 * epsilon * (s12 - s6), zero)
    fs = torch.where(mask, 24.0 * epsilon * (2.0 * s12 - s6) * s2 / (sigma ** 2), zero)
    return dr, e, fs


def force_analytic(R: torch.Tensor, box, sigma=1.0, epsilon=1.0, rc=None, rows=None,
                   chunk=1024):
    """Closed-form force, row-chunked so N x N never materialises.
    Returns (F[rows], PE) with PE summed in float64 over all visited rows (PE is the full
    total only when rows is None)."""
    N = R.shape[0]
    box = torch.as_tensor(box, dtype=R.dtype)
    idx = torch.arange(N) if rows is None else torch.as_tensor(rows)
    F = torch.zeros((idx.shape[0], 2), dtype=R.dtype)
    pe = 0.0
    contiguous = rows is None
    for a in range(0, idx.shape[0], chunk):
        sel = idx[a:a + chunk]
        Ri = R[sel]
        facc = torch.zeros((sel.shape[0], 2), dtype=torch.float64)
        for b in range(0, N, 4096):
            Rj = R[b:b + 4096]
            if contiguous:
                dr, e, fs = _pair_blocks(Ri, Rj, box, sigma, epsilon, rc, a, b)
            else:
                dr = periodic_displacement(Ri[:, None, :] - Rj[None, :, :], box)
                r2 = torch.sum(dr ** 2, dim=-1)
                mask = sel[:, None] != (torch.arange(Rj.shape[0])[None, :] + b)
                if rc is not None and math.isfinite(rc):
                    rc2 = torch.as_tensor(np.float32(rc) * np.float32(rc), dtype=R.dtype)
                    mask = mask & (r2 < rc2)
                r2s = torch.where(mask, r2, torch.ones((), dtype=R.dtype))
                s2 = (sigma ** 2) / r2s
                s6 = s2 ** 3
                s12 = s6 ** 2
                zero = torch.zeros((), dtype=R.dtype)
                e = torch.where(mask, 4.0 * epsilon * (s12 - s6), zero)
                fs = torch.where(mask, 24.0 * epsilon * (2.0 * s12 - s6) * s2 / (sigma ** 2), zero)
            facc += torch.sum((fs[:, :, None] * dr).to(torch.float64), dim=1)
            pe += float(torch.sum(e.to(torch.float64)))
        F[a:a + chunk] = facc.to(R.dtype)
    return F, 0.5 * pe


def kinetic_energy(V: torch.Tensor) -> float:
    """New quantity: KE = 0.5 * sum |v|^2, unit mass (MD:70-74)."""
    return 0.5 * float(torch.sum(V.to(torch.float64) ** 2))


# --------------------------------------------------------------------------- dynamics
def verlet_step(state, box, dt, force_fn):
    """MD:66-75 verbatim: two force evaluations, F not carried."""
    R, V = state
    dt = torch.as_tensor(dt, dtype=R.dtype)
    box_t = torch.as_tensor(box, dtype=R.dtype)
    F = force_fn(R)
    V_half = V + 0.5 * F * dt
    R_new = R + V_half * dt
    R_new = torch.remainder(R_new, box_t)
    F_new = force_fn(R_new)
    V_new = V_half + 0.5 * F_new * dt
    return R_new, V_new


def run(state, box, dt, nsteps, sample_every=0, rc=None, sigma=1.0, epsilon=1.0,
        autodiff=True, energy_every=0):
    """equilibrate_fn MD:77-83 (sample_every=0) / production_fn MD:85-106.
    Carries F between steps (bit-identical to recomputing it)."""
    R, V = state
    ff = (lambda r: force_autodiff(r, box, sigma, epsilon, rc)) if autodiff else \
         (lambda r: force_analytic(r, box, sigma, epsilon, rc)[0])
    dt_t = torch.as_tensor(dt, dtype=R.dtype)
    box_t = torch.as_tensor(box, dtype=R.dtype)
    S = nsteps // sample_every if sample_every else 0
    traj = torch.zeros((S, R.shape[0], 2), dtype=R.dtype)          # MD:89
    energies = []
    F = ff(R)
    for i in range(nsteps):
        V_half = V + 0.5 * F * dt_t
        R = torch.remainder(R + V_half * dt_t, box_t)
        F = ff(R)
        V = V_half + 0.5 * F * dt_t
        if sample_every and i % sample_every == 0 and i // sample_every < S:   # MD:93-100
            traj[i // sample_every] = R
        if energy_every and i % energy_every == 0:
            pe = float(total_energy(R.to(torch.float64), float(box), sigma, epsilon, rc)) \
                if R.shape[0] <= 4096 else force_analytic(R, box, sigma, epsilon, rc)[1]
            energies.append((kinetic_energy(V), pe))
    return (R, V), traj, energies


# --------------------------------------------------------------------------- g(r)
def g_r(R_history: torch.Tensor, N: int, box, nbins: int, r_max):
    """_calculate_g_r_internal MD:108-129 (numpy restatement; histogram in float32 like jnp)."""
    box32 = np.float32(box)
    r_max32 = np.float32(r_max)
    r_bins = np.linspace(0, r_max32, nbins + 1, dtype=np.float32)           # MD:110
    centers = (r_bins[:-1] + r_bins[1:]) / np.float32(2.0)                   # MD:111
    shell = np.float32(np.pi) * (r_bins[1:] ** 2 - r_bins[:-1] ** 2)         # MD:112
    rho_pairs = np.float32(N * (N - 1) / 2.0) / (box32 ** 2)                 # MD:113
    ideal = rho_pairs * shell                                                # MD:115
    hists = []
    iu = np.triu_indices(N, k=1)
    for R in R_history.numpy():
        dr = R[:, None, :] - R[None, :, :]
        dr = dr - box32 * np.round(dr / box32)
        r2 = np.sum(dr ** 2, axis=-1, dtype=np.float32)
        r = np.sqrt(r2[iu])
        h, _ = np.histogram(r, bins=r_bins)
        hists.append(h)
    hists = np.stack(hists) if hists else np.zeros((0, nbins), dtype=np.int64)
    avg = hists.astype(np.float32).mean(axis=0) if len(hists) else np.zeros(nbins, np.float32)
    return centers, (avg / ideal).astype(np.float32), hists, r_bins


# --------------------------------------------------------------------------- C restatement
class _C:
    lib = None


def c_lib_path() -> str:
    return os.path.join(_HERE, "liblj_oracle.so")


def build_c(force: bool = False) -> str:
    """Compile oracle/lj_oracle.c (the recipe is oracle/Makefile)."""
    so = c_lib_path()
    src = os.path.join(_HERE, "lj_oracle.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-s", "-C", _HERE, "liblj_oracle.so"])
    return so


def c_lib():
    if _C.lib is None:
        lib = ctypes.CDLL(build_c())
        f32p = ctypes.POINTER(ctypes.c_float)
        f64p = ctypes.POINTER(ctypes.c_double)
        i32p = ctypes.POINTER(ctypes.c_int32)
        i64p = ctypes.POINTER(ctypes.c_int64)
        lib.orc_periodic_displacement.restype = ctypes.c_float
        lib.orc_periodic_displacement.argtypes = [ctypes.c_float, ctypes.c_float]
        lib.orc_mod.restype = ctypes.c_float
        lib.orc_mod.argtypes = [ctypes.c_float, ctypes.c_float]
        lib.orc_forces_rows.argtypes = [ctypes.c_int64, f32p, ctypes.c_float, ctypes.c_float,
                                        ctypes.c_float, ctypes.c_float, ctypes.c_int64,
                                        ctypes.c_int64, f32p, f64p, ctypes.c_int]
        lib.orc_forces.argtypes = [ctypes.c_int64, f32p, ctypes.c_float, ctypes.c_float,
                                   ctypes.c_float, ctypes.c_float, f32p, f64p, ctypes.c_int]
        lib.orc_kinetic.restype = ctypes.c_double
        lib.orc_kinetic.argtypes = [ctypes.c_int64, f32p]
        lib.orc_run.argtypes = [ctypes.c_int64, f32p, f32p, ctypes.c_float, ctypes.c_float,
                                ctypes.c_float, ctypes.c_float, ctypes.c_float, ctypes.c_int64,
                                ctypes.c_int64, f32p, ctypes.c_int64, f64p, ctypes.c_float,
                                ctypes.c_int64, ctypes.c_int]
        lib.orc_cell_assign.argtypes = [ctypes.c_int64, f32p, ctypes.c_int32, ctypes.c_int32,
                                        ctypes.c_float, ctypes.c_float, i32p, i32p]
        lib.orc_neighbor_count.argtypes = [ctypes.c_int64, f32p, ctypes.c_float, ctypes.c_float,
                                           i32p]
        lib.orc_forces_cells.argtypes = [ctypes.c_int64, f32p, ctypes.c_float, ctypes.c_float,
                                         ctypes.c_float, ctypes.c_float, f32p, f64p]
        lib.orc_gr_hist.argtypes = [ctypes.c_int64, f32p, ctypes.c_float, ctypes.c_int32, f32p,
                                    i64p]
        lib.orc_gravity_nbody.argtypes = [ctypes.c_int64, f32p, f32p, ctypes.c_float, f32p]
        lib.orc_gravity_em3.argtypes = [ctypes.c_int64, f32p, f32p, ctypes.c_float, f32p]
        _C.lib = lib
    return _C.lib


def _f32(a):
    a = np.ascontiguousarray(a, dtype=np.float32)
    return a, a.ctypes.data_as(ctypes.POINTER(ctypes.c_float))


def _rc2(rc):
    if rc is None or not math.isfinite(rc) or rc <= 0:
        return float("inf")
    return float(np.float32(rc) * np.float32(rc))


def c_forces(R, box, rc=None, sigma=1.0, epsilon=1.0, acc_double=True, rows=None):
    """C restatement of force_fn / total_energy_fn.  Returns (F, PE)."""
    R, Rp = _f32(R)
    N = R.shape[0]
    i0, i1 = (0, N) if rows is None else rows
    F = np.empty((i1 - i0, 2), dtype=np.float32)
    pe = ctypes.c_double(0.0)
    c_lib().orc_forces_rows(N, Rp, float(box), sigma, epsilon, _rc2(rc), i0, i1,
                            F.ctypes.data_as(ctypes.POINTER(ctypes.c_float)),
                            ctypes.byref(pe), int(acc_double))
    return F, pe.value


def c_forces_cells(R, box, rc, sigma=1.0, epsilon=1.0):
    R, Rp = _f32(R)
    N = R.shape[0]
    F = np.empty((N, 2), dtype=np.float32)
    pe = ctypes.c_double(0.0)
    c_lib().orc_forces_cells(N, Rp, float(box), sigma, epsilon, float(rc),
                             F.ctypes.data_as(ctypes.POINTER(ctypes.c_float)), ctypes.byref(pe))
    return F, pe.value


def c_run(R, V, box, dt, nsteps, sample_every=0, rc=None, energy_every=0, thermostat_kT=0.0,
          thermostat_every=0, sigma=1.0, epsilon=1.0, acc_double=True):
    """C restatement of equilibrate_fn / production_fn.  Returns (R, V, traj, ke_pe)."""
    R = np.array(R, dtype=np.float32, order="C", copy=True)
    V = np.array(V, dtype=np.float32, order="C", copy=True)
    N = R.shape[0]
    S = nsteps // sample_every if sample_every else 0
    traj = np.zeros((S, N, 2), dtype=np.float32)
    ne = -(-nsteps // energy_every) if energy_every else 0
    ke_pe = np.zeros((ne, 2), dtype=np.float64)
    f32p = ctypes.POINTER(ctypes.c_float)
    c_lib().orc_run(N, R.ctypes.data_as(f32p), V.ctypes.data_as(f32p), float(box), sigma, epsilon,
                    _rc2(rc), float(dt), nsteps, sample_every,
                    traj.ctypes.data_as(f32p) if S else None, energy_every,
                    ke_pe.ctypes.data_as(ctypes.POINTER(ctypes.c_double)) if ne else None,
                    float(thermostat_kT), thermostat_every, int(acc_double))
    return R, V, traj, ke_pe


def c_cell_assign(R, nrows, nbx, inv_hy, inv_wx):
    """Strip-cell recount: cell = row * nbx + bin (see orc_cell_assign)."""
    R, Rp = _f32(R)
    N = R.shape[0]
    cid = np.empty(N, dtype=np.int32)
    cnt = np.empty(nrows * nbx, dtype=np.int32)
    i32p = ctypes.POINTER(ctypes.c_int32)
    c_lib().orc_cell_assign(N, Rp, nrows, nbx, float(inv_hy), float(inv_wx),
                            cid.ctypes.data_as(i32p), cnt.ctypes.data_as(i32p))
    return cid, cnt


def c_neighbor_count(R, box, radius):
    R, Rp = _f32(R)
    N = R.shape[0]
    out = np.empty(N, dtype=np.int32)
    c_lib().orc_neighbor_count(N, Rp, float(box), float(radius),
                               out.ctypes.data_as(ctypes.POINTER(ctypes.c_int32)))
    return out


def c_gr_hist(R, box, nbins, r_max):
    R, Rp = _f32(R)
    edges = np.linspace(0, np.float32(r_max), nbins + 1, dtype=np.float32)
    counts = np.zeros(nbins, dtype=np.int64)
    c_lib().orc_gr_hist(R.shape[0], Rp, float(box), nbins,
                        edges.ctypes.data_as(ctypes.POINTER(ctypes.c_float)),
                        counts.ctypes.data_as(ctypes.POINTER(ctypes.c_int64)))
    return counts


# --------------------------------------------------------------------------- gravity pair laws (8f rank 4)
def c_gravity(pos, mass, G, law="nbody"):
    """C restatement of pairwise_forces (NBODY:54-67, law="nbody") / the gravity term of acceleration
    (EM3:25-38, law="em3").  Returns acc (n,2) float32."""
    pos, pp = _f32(pos)
    mass = np.ascontiguousarray(mass, dtype=np.float32)
    acc = np.empty_like(pos)
    f32p = ctypes.POINTER(ctypes.c_float)
    fn = c_lib().orc_gravity_nbody if law == "nbody" else c_lib().orc_gravity_em3
    fn(pos.shape[0], pp, mass.ctypes.data_as(f32p), float(G), acc.ctypes.data_as(f32p))
    return acc


def gravity_nbody_loops(pos, mass, G):
    """pairwise_forces exactly as written (NBODY:54-67): Python double loop, fp32 numpy scalars."""
    pos = np.asarray(pos, dtype=np.float32)
    mass = np.asarray(mass, dtype=np.float32)
    G = np.float32(G)
    n = len(pos)
    acc = np.zeros_like(pos)
    for i in range(n):
        for j in range(n):
            if i == j:
                continue
            r_vec = pos[j] - pos[i]
            r_norm = np.sqrt(np.sum(r_vec * r_vec, dtype=np.float32), dtype=np.float32)
            acc_mag = (G * mass[j] / (r_norm * r_norm * r_norm)) if r_norm >= np.float32(1e-6) else np.float32(0.0)
            acc[i] = acc[i] + acc_mag * r_vec
    return acc


def gravity_em3_broadcast(pos, mass, G):
    """Gravity term of acceleration() as written (EM3:25-38): broadcast N x N x 2 arrays, fp32."""
    pos = np.asarray(pos, dtype=np.float32)
    mass = np.asarray(mass, dtype=np.float32)
    r_diff = pos[None, :, :] - pos[:, None, :]
    r2 = np.sum(r_diff ** 2, axis=-1, dtype=np.float32) + np.eye(len(pos), dtype=np.float32)
    r2 = np.where(r2 < np.float32(1e-12), np.float32(1e-12), r2)
    inv3 = r2 ** np.float32(-1.5)
    pairs = np.float32(G) * mass[None, :, None] * r_diff * inv3[..., None]
    return np.sum(pairs, axis=1, dtype=np.float32)
