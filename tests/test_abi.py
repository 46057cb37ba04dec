"""CPU-side checks of the drop-in boundary: the C-ABI library loads and exports every symbol
include/ljmd.h declares (no compute calls without a GPU), and fails loudly without a device."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    import __graft_entry__ as g
    g.build()
    from jax_tpus_benchmark_physics_simulation_b200 import _lib
    return _lib.load()


def _declared_symbols():
    src = open(os.path.join(ROOT, "include", "ljmd.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(ljmd_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_all_exported(lib):
    from jax_tpus_benchmark_physics_simulation_b200 import _lib
    names = _declared_symbols()
    assert len(names) >= 15
    assert sorted(_lib.EXPORTS) == names
    for n in names:
        assert getattr(lib, n) is not None


def test_abi_version_and_error_string(lib):
    assert lib.ljmd_abi_version() == 1
    assert isinstance(lib.ljmd_last_error(), bytes)


def test_params_struct_layout():
    from jax_tpus_benchmark_physics_simulation_b200._lib import LjmdParams
    # int64 + 6 floats + 2 int32 + pointer -> 8 + 24 + 8 + 8
    assert ctypes.sizeof(LjmdParams) == 48
    assert LjmdParams.stream.offset == 40


def test_create_fails_loudly_without_gpu(lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from jax_tpus_benchmark_physics_simulation_b200 import _lib
    p = _lib.LjmdParams(N=400, box=22.36, sigma=1.0, epsilon=1.0, rc=2.5, dt=1e-3, skin=0.3,
                        path=0, device=0, stream=None)
    h = ctypes.c_void_p()
    code = lib.ljmd_create(ctypes.byref(h), ctypes.byref(p))
    assert code != 0 and not h.value
    assert len(lib.ljmd_last_error()) > 0
    from jax_tpus_benchmark_physics_simulation_b200.md import LJSimulation
    with pytest.raises(_lib.LjmdError):
        LJSimulation(400)


def test_bad_arguments_rejected(lib):
    h = ctypes.c_void_p()
    assert lib.ljmd_create(ctypes.byref(h), None) != 0
    assert lib.ljmd_energy(None, None, None) != 0
    assert lib.ljmd_run(None, None, None, None, None, 1, 0, None, 0, None, 0.0, 0) != 0


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "jax_tpus_benchmark_physics_simulation_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert "lj_oracle" not in text and "from oracle" not in text and "import oracle" not in text, f
