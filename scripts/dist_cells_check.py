"""Multi-GPU check of the slab-decomposed cell-list path (run under torchrun, one rank per GPU):
sharded result == single-GPU result on identical inputs, then a timing line.
    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 scripts/dist_cells_check.py [N steps skin]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch, torch.distributed as dist
from jax_tpus_benchmark_physics_simulation_b200 import lattice_jitter
from jax_tpus_benchmark_physics_simulation_b200.md import LJSimulation, make_dist_arg
rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
N = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 60
skin = float(sys.argv[3]) if len(sys.argv) > 3 else 0.3
R, V, box = lattice_jitter(N, seed=0)
sim = LJSimulation(N, rc=2.5, dt=0.005, path="cells", skin=skin, device=lr, dist=make_dist_arg(rank, world))
one = LJSimulation(N, rc=2.5, dt=0.005, path="cells", skin=skin, device=lr)
F, pe = sim.force_and_energy(R)
F1, pe1 = one.force_and_energy(R)
ferr = float(np.abs(F.numpy() - F1.numpy()).max() / np.abs(F1.numpy()).max())
se = max(1, min(10, steps // 2))
(Rs, Vs), traj = sim.run((R, V), steps, sample_every=se, energy_every=se)
es = sim.last_energies.numpy()
rb = sim.last_rebuilds()
(Ro, Vo), traj1 = one.run((R, V), steps, sample_every=se, energy_every=se)
eo = one.last_energies.numpy()
d = np.abs(Rs.numpy() - Ro.numpy()); d = np.minimum(d, float(box) - d)
terr = np.abs(traj.numpy() - traj1.numpy()); terr = np.minimum(terr, float(box) - terr)
# a few ulp(box): the summation order of the edge warps differs with P.  The system is chaotic: measured
# max|dR| between the two runs at N = 2^20 is 1.5e-5 / 1.2e-4 / 7.3e-4 after 60 / 120 / 200 steps (same
# rebuild steps, energies equal to 1e-7), so the bound doubles every 25 steps past 60
ptol = max(1e-4, 4.0 * float(np.spacing(np.float32(box)))) * 2.0 ** (max(0, steps - 60) / 25.0)
ok = ferr < 2e-6 and d.max() < ptol and terr.max() < ptol and abs(float(pe) - float(pe1)) < 1e-6 * abs(float(pe1)) \
    and np.abs(es.sum(1) - eo.sum(1)).max() < 2e-6 * np.abs(eo.sum(1)).max() and np.abs(Vs.numpy() - Vo.numpy()).max() < max(1e-3, 20.0 * ptol)
print(f"[rank {rank}] N={N} P={world} rebuilds {rb}/{one.last_rebuilds()} force err {ferr:.2e} max|dR| {d.max():.2e} "
      f"traj {terr.max():.2e} E {es[-1].sum():.4f} vs {eo[-1].sum():.4f} -> {'OK' if ok else 'MISMATCH'}", flush=True)
dist.barrier()
Rd, Vd = torch.from_numpy(R).cuda(), torch.from_numpy(V).cuda()
# per-call costs (selecting the slab from the replicated input, the first sort, the replicated output:
# zero-fill + NCCL all-reduce of R and V) are paid once per call whatever its length, so the per-step
# rate of a long production call is the MARGINAL time between a 3x longer and a 1x call
for s_ in (sim, one):
    t = {}
    for n_ in (steps, 3 * steps):
        torch.cuda.synchronize(); dist.barrier()
        s_.run((Rd, Vd), n_)
        torch.cuda.synchronize(); dist.barrier()
        s_.run((Rd, Vd), n_)
        tt = torch.tensor([s_.last_run_ms()], dtype=torch.float64, device="cuda")
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        t[n_] = float(tt.item())
    marg = (t[3 * steps] - t[steps]) / (2 * steps)
    if rank == 0:
        print(f"  {'sharded' if s_ is sim else 'single '} {steps}-step call {1e3 * t[steps] / steps:.1f} us/step, "
              f"{3 * steps}-step call {1e3 * t[3 * steps] / (3 * steps):.1f} us/step, marginal {1e3 * marg:.1f} us/step "
              f"= {N / marg / 1e6:.3f}e9 particle-steps/s", flush=True)
sim.last_run_ms()   # surfaces device error flags
dist.destroy_process_group()
sys.exit(0 if ok else 1)
