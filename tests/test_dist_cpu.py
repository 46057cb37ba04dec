"""CPU (gloo, world_size 2) tests of the host-side multi-GPU logic: slab partition, unique-id
broadcast, and bench.py's reference arm under a multi-rank launch (rank 0 works, others exit 0)."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = r'''
import os, sys
sys.path.insert(0, os.environ["LJMD_ROOT"])
import torch.distributed as dist
from jax_tpus_benchmark_physics_simulation_b200.md import broadcast_unique_id, slab_range, row_slab_range
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
uid = broadcast_unique_id(rank, get_uid=lambda: bytes(range(128)))
assert uid == bytes(range(128)), uid
lo, hi = slab_range(65536, rank, world)
got = [None] * world
dist.all_gather_object(got, (lo, hi))
assert got[0][0] == 0 and got[-1][1] == 65536
for a, b in zip(got[:-1], got[1:]):
    assert a[1] == b[0]
assert all(h - l == 65536 // world for l, h in got)
# row slabs of the cell-list path: contiguous, cover all rows, sizes differ by at most one
g0, nloc = row_slab_range(1635, rank, world)
rows = [None] * world
dist.all_gather_object(rows, (g0, nloc))
assert rows[0][0] == 0 and rows[-1][0] + rows[-1][1] == 1635
for a, b in zip(rows[:-1], rows[1:]):
    assert a[0] + a[1] == b[0]
assert max(n for _, n in rows) - min(n for _, n in rows) <= 1
dist.barrier()
dist.destroy_process_group()
print("OK", rank)
'''


def _torchrun(args, env=None, timeout=300):
    e = dict(os.environ)
    e["LJMD_ROOT"] = ROOT
    e.update(env or {})
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
           "--master-addr", "127.0.0.1", "--master-port", "29617"] + args
    return subprocess.run(cmd, capture_output=True, text=True, timeout=timeout, env=e, cwd=ROOT)


def test_slab_range_partition():
    from jax_tpus_benchmark_physics_simulation_b200.md import slab_range
    for P in (1, 2, 4, 8):
        edges = [slab_range(65536, r, P) for r in range(P)]
        assert edges[0][0] == 0 and edges[-1][1] == 65536
        assert all(a[1] == b[0] for a, b in zip(edges[:-1], edges[1:]))
    with pytest.raises(ValueError):
        slab_range(1000, 0, 3)
    with pytest.raises(ValueError):
        slab_range(1024, 2, 2)


def test_uid_broadcast_and_partition_gloo_world2(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    r = _torchrun([str(script)])
    assert r.returncode == 0, r.stdout + r.stderr
    assert r.stdout.count("OK") == 2


def test_reference_arm_runs_on_rank0_only():
    r = _torchrun(["bench.py", "--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "0",
                   "--workload", "ap400"])
    assert r.returncode == 0, r.stdout + r.stderr
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["n_gpus"] == 2 and d["value"] > 0
    assert d["cpu_baseline"]["kind"] == "port" and d["e2e"]["h2d_bytes_per_step"] == 0
