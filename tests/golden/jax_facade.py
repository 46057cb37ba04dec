"""A torch-CPU facade of exactly the slice of the jax API that the reference script
`molecular_dynamics_jax_single-host_workload.py` touches (MD:1-3 imports; jnp.* MD:30-135;
jit / grad / vmap MD:46-131; lax.fori_loop / lax.cond MD:82,95-103; random.* MD:29,133-135).

TEST INFRASTRUCTURE ONLY.  JAX / XLA are not installable in this image, so the
reference cannot run as it is.  With this facade installed under the names `jax`, `jax.numpy`,
... the reference's OWN SOURCE FILE is executed unmodified (tests/golden/make_reference_golden.py):
its closures (periodic_displacement, total_energy_fn through force_fn, verlet_step, equilibrate_fn,
production_fn, calculate_g_r) run their own expression sequences in IEEE fp32 on torch, and `grad`
differentiates the reference's own energy function (torch.func.grad).  What this is NOT: XLA's
reduction orders, its pow lowering, or the threefry PRNG (initial conditions are injected by the
generator, never drawn here).  Element-wise fp32 add / sub / mul / div / round-half-even /
where / remainder are the same IEEE operations in both.

Arrays are plain torch tensors; the two jax.Array methods the script uses on them
(`.block_until_ready()`, `.at[idx].set(v)`) are attached to torch.Tensor while the facade is
installed.
"""
from __future__ import annotations

import math
import sys
import types

import numpy as np
import torch
from torch import func as tfunc

F32 = torch.float32


def _t(x):
    """jnp weak typing: python scalars become fp32 (x64 is never enabled in the reference)."""
    if isinstance(x, torch.Tensor):
        return x
    if isinstance(x, bool):
        return torch.tensor(x)
    if isinstance(x, (int, float)):
        return torch.tensor(x, dtype=F32)
    return torch.as_tensor(np.asarray(x))


class Registry:
    """What the script passes through jit / grad, in definition order, and what it blocks on."""

    def __init__(self):
        self.jitted = []          # (name, fn)
        self.grad_sources = []    # the functions handed to grad (MD:64: lambda R: -total_energy_fn(R))
        self.blocked = []         # tensors on which block_until_ready() was called (MD:145,152,163)
        self.uniform = None       # arrays returned by random.uniform / random.normal (injected ICs)
        self.normal = None

    def fn(self, name):
        for n, f in self.jitted:
            if n == name:
                return f
        raise KeyError(name)


REG = Registry()


# ------------------------------------------------------------------ jax.numpy
def _make_jnp():
    m = types.ModuleType("jax.numpy")
    m.pi = math.pi
    m.float32 = F32
    def sqrt(x):
        # IEEE correctly rounded (numpy -> sqrtss/sqrtps), as XLA's CPU lowering is; torch's vectorised CPU
        # sqrt is off by one ulp for some fp32 inputs (found on a g(r) bin edge at N = 4096).  None of the
        # reference's sqrt calls (MD:30,123,135) is differentiated.
        x = _t(x)
        return torch.from_numpy(np.asarray(np.sqrt(x.detach().numpy()), dtype=x.detach().numpy().dtype))
    m.sqrt = sqrt
    m.round = lambda x: torch.round(_t(x))                      # half to even, like jnp.round
    m.sum = lambda x, axis=None: torch.sum(x) if axis is None else torch.sum(x, dim=axis)
    m.mean = lambda x, axis=None: (torch.mean(x.to(F32)) if axis is None else torch.mean(x.to(F32), dim=axis))
    m.logical_not = torch.logical_not
    m.where = lambda c, a, b: torch.where(c, _t(a).to(F32) if not isinstance(a, torch.Tensor) else a,
                                          _t(b).to(F32) if not isinstance(b, torch.Tensor) else b)
    m.mod = lambda a, b: torch.remainder(_t(a), _t(b))          # sign of the divisor, like jnp.mod
    m.zeros = lambda shape, dtype=None: torch.zeros(shape, dtype=F32 if dtype is None else dtype)

    def eye(n, dtype=None):
        return torch.eye(n, dtype=torch.bool if dtype is bool else (F32 if dtype is None else dtype))
    m.eye = eye

    def linspace(a, b, n):
        # jnp.linspace(0, r_max, n) in fp32: a + step * arange, end point exact
        return torch.from_numpy(np.linspace(float(a), np.float32(float(b)), int(n), dtype=np.float32))
    m.linspace = linspace

    def triu_indices(n, k=0):
        i = torch.triu_indices(n, n, offset=k)
        return (i[0], i[1])
    m.triu_indices = triu_indices

    def histogram(x, bins):
        # numpy.histogram semantics (which jnp.histogram follows): last bin right-closed
        c, e = np.histogram(x.detach().numpy().astype(np.float32), bins=bins.detach().numpy().astype(np.float32))
        return torch.from_numpy(c.astype(np.int64)), bins
    m.histogram = histogram
    def array(x, dtype=None):
        if isinstance(x, (list, tuple)) and x and isinstance(x[0], torch.Tensor):
            return torch.stack([_t(v) for v in x])
        t = _t(x)
        return t.to(F32) if t.dtype == torch.float64 else t
    m.array = array
    m.zeros_like = torch.zeros_like
    m.arange = lambda n: torch.arange(n)
    m.stack = lambda xs, axis=0: torch.stack([_t(v) for v in xs], dim=axis)
    linalg = types.ModuleType("jax.numpy.linalg")
    # jnp.linalg.norm of a vector: sqrt(sum(|x|^2)), correctly rounded sqrt (see sqrt above)
    linalg.norm = lambda x: m.sqrt(torch.sum(x * x))
    m.linalg = linalg
    return m


# ------------------------------------------------------------------ transforms
def _jit(fun=None, static_argnums=None, **kw):
    if fun is None:
        return lambda f: _jit(f, static_argnums=static_argnums)
    REG.jitted.append((getattr(fun, "__name__", "?"), fun))
    return fun


def _grad(fun):
    REG.grad_sources.append(fun)
    g = tfunc.grad(fun)
    g.__name__ = "grad_" + getattr(fun, "__name__", "fn")
    return g


def _vmap(fun, in_axes=0, out_axes=0):
    def wrapped(*args):
        ia = in_axes
        if isinstance(ia, (tuple, list)):
            # scalars mapped with None may be python numbers / 0-d tensors
            args2 = tuple(_t(a) if d is None else a for a, d in zip(args, ia))
            return tfunc.vmap(fun, in_dims=tuple(ia), out_dims=out_axes)(*args2)
        # vmap over histogram (MD:126): get_histogram leaves torch for numpy, so map in Python
        try:
            return tfunc.vmap(fun, in_dims=ia, out_dims=out_axes)(*args)
        except Exception:
            outs = [fun(*(a[i] for a in args)) for i in range(args[0].shape[0])]
            return torch.stack(outs, dim=out_axes)
    return wrapped


def _fori_loop(lo, hi, body, init):
    val = init
    for i in range(int(lo), int(hi)):
        val = body(i, val)
    return val


def _cond(pred, true_fn, false_fn, *ops):
    return true_fn(*ops) if bool(pred) else false_fn(*ops)


class _At:
    def __init__(self, t):
        self.t = t

    def __getitem__(self, idx):
        t = self.t

        class _Set:
            def add(self, v):                      # acc.at[i].add(x)  (NBODY:64)
                out = t.clone()
                out[idx] = out[idx] + v
                return out

            def set(self, v):
                # out-of-range index: JAX drops the update (MD:95-100 relies on it)
                n = t.shape[0]
                if isinstance(idx, int) and not (-n <= idx < n):
                    return t
                out = t.clone()
                out[idx] = v
                return out
        return _Set()


def _block_until_ready(self):
    REG.blocked.append(self)
    return self


# ------------------------------------------------------------------ jax.random (ICs are injected)
def _make_random():
    m = types.ModuleType("jax.random")
    m.PRNGKey = lambda seed: ("key", int(seed))
    m.split = lambda key, n=2: tuple(("key", key[1], i) for i in range(n))

    def uniform(key, shape):
        if REG.uniform is None:
            raise RuntimeError("inject REG.uniform (unit-box positions) before running the reference")
        assert tuple(REG.uniform.shape) == tuple(shape)
        return REG.uniform
    m.uniform = uniform

    def normal(key, shape):
        if REG.normal is None:
            raise RuntimeError("inject REG.normal before running the reference")
        assert tuple(REG.normal.shape) == tuple(shape)
        return REG.normal
    m.normal = normal
    return m


class _Absorb:
    """matplotlib.pyplot stand-in: every attribute is a callable that returns another stand-in."""

    def __getattr__(self, name):
        return _Absorb()

    def __call__(self, *a, **k):
        return _Absorb()


def install():
    """Put the facade into sys.modules (jax, jax.numpy, jax.random, jax.lax, matplotlib[.pyplot])."""
    REG.__init__()
    jax = types.ModuleType("jax")
    jnp = _make_jnp()
    rnd = _make_random()
    lax = types.ModuleType("jax.lax")
    lax.fori_loop = _fori_loop
    lax.cond = _cond
    jax.numpy, jax.random, jax.lax = jnp, rnd, lax
    jax.jit, jax.grad, jax.vmap = _jit, _grad, _vmap
    jax.default_backend = lambda: "torch-cpu facade (XLA absent)"
    mpl = types.ModuleType("matplotlib")
    plt = types.ModuleType("matplotlib.pyplot")
    def _plt_attr(name):
        if name.startswith("__"):
            raise AttributeError(name)
        return _Absorb()
    plt.__getattr__ = _plt_attr
    mpl.pyplot = plt
    mods = {"jax": jax, "jax.numpy": jnp, "jax.random": rnd, "jax.lax": lax,
            "matplotlib": mpl, "matplotlib.pyplot": plt}
    saved = {k: sys.modules.get(k) for k in mods}
    sys.modules.update(mods)
    torch.Tensor.block_until_ready = _block_until_ready
    torch.Tensor.at = property(lambda self: _At(self))
    return saved


def uninstall(saved):
    for k, v in saved.items():
        if v is None:
            sys.modules.pop(k, None)
        else:
            sys.modules[k] = v
    for attr in ("block_until_ready", "at"):
        if hasattr(torch.Tensor, attr):
            delattr(torch.Tensor, attr)
