"""Phase clocks of the slab-decomposed cell-list path on rank 0 (run under torchrun):
    python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 scripts/dist_cells_prof.py [N steps skin]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
if rank == 0:
    os.environ["LJMD_CELLS_PROF"] = "1"
import torch, torch.distributed as dist
from jax_tpus_benchmark_physics_simulation_b200 import lattice_jitter
from jax_tpus_benchmark_physics_simulation_b200.md import LJSimulation, make_dist_arg
torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
N = int(sys.argv[1]) if len(sys.argv) > 1 else 16777216
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 300
skin = float(sys.argv[3]) if len(sys.argv) > 3 else 0.5
R, V, box = lattice_jitter(N, seed=0)
sim = LJSimulation(N, rc=2.5, dt=0.005, path="cells", skin=skin, device=lr, dist=make_dist_arg(rank, world) if world > 1 else None)
state, _ = sim.run((R, V), 500)
state = (state[0].tensor, state[1].tensor)
for _ in range(2):
    torch.cuda.synchronize(); dist.barrier()
    if rank == 0: print(f"---- {steps}-step call on {world} GPUs", file=sys.stderr, flush=True)
    sim.run(state, steps)
    ms = sim.last_run_ms()
    if rank == 0:
        print(f"N={N} P={world} {1e3 * ms / steps:.2f} us/step rebuilds {sim.last_rebuilds()}", flush=True)
sim.check()
dist.destroy_process_group()
