#!/usr/bin/env python
"""bench.py — headline benchmark of the B200-native LJ-MD hot path.

    python bench.py --gpus N --steps K --warmup W [--impl reference] [--workload NAME]

Metric (BASELINE.json): particle-steps/s (and pair-interactions/s).  One bench "step" is one
pass of the hot path over one batch: ``md_steps_per_step`` velocity-Verlet steps of the named
system issued as ONE device dispatch (like equilibrate_fn, MD:77-83).

Workloads (BASELINE.json configs):
  ap4096   all-pairs N=4,096  rc=2.5 dt=0.005   (configs[1]; the default.  A step is a few us of
                                                  work: it does not shard -> --gpus N runs N
                                                  independent replicas, "weak" scaling)
  ap65536  all-pairs N=65,536 rc=2.5 dt=0.005   (configs[2]; --gpus N shards it: atom decomposition)
  cells4m  cell-list N=4,194,304 rho=0.8 rc=2.5 (configs[3])
  cells16m cell-list N=16,777,216               (configs[4]; --gpus N shards it: row slabs + halos)

The reference arm (--impl reference) times the CPU restatement of the reference's own
verlet_step (two dense autodiff force evaluations per step, torch CPU, all host threads) on the
same config — JAX itself is not installable in this image, so kind = "port".
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

# measured DRAM bytes per force evaluation (ncu --set full captures under profiles/)
AP_TRAFFIC = {
    "ap4096": (256512.0 / 1001.0, "profiles/r1_ap4096_n3l_final.ncu-rep: 256.5 KB per 1001-evaluation launch"),
    "ap65536": (127447552.0 / 4.0, "profiles/r1_ap65536_n3l_v4.ncu-rep: 127.4 MB per 4-evaluation launch "
                                   "(partial force vectors)"),
}

WORKLOADS = {
    "ap400":    dict(N=400, rc=None, dt=1e-3, path="allpairs", md_steps=2000,
                     desc="default 2D LJ run of the reference script (N=400, no cutoff, dt=1e-3)"),
    "ap4096":   dict(N=4096, rc=2.5, dt=0.005, path="allpairs", md_steps=1000,
                     desc="2D LJ all-pairs N=4096 rc=2.5 dt=0.005"),
    "ap65536":  dict(N=65536, rc=2.5, dt=0.005, path="allpairs", md_steps=20,
                     desc="2D LJ all-pairs N=65536 rc=2.5 dt=0.005"),
    "cells4m":  dict(N=4194304, rc=2.5, dt=0.005, path="cells", md_steps=1000, skin=0.5,
                     desc="2D LJ cell-list N=4194304 rho=0.8 rc=2.5 dt=0.005 (skin 0.5)"),
    "cells16m": dict(N=16777216, rc=2.5, dt=0.005, path="cells", md_steps=1000, skin=0.5,
                     desc="2D LJ cell-list N=16777216 rho=0.8 rc=2.5 dt=0.005 (skin 0.5)"),
}
FLOP_PER_PAIR_FORCE = 25.0      # SURVEY.md §8d (fixed for builder and judge): one ORDERED pair
FLOP_PER_UNORDERED_N3L = 33.0   # Newton's-third-law tiles: one evaluation (25) + the reaction on j (4 FMA)
BYTES_PER_PARTICLE_STEP = 32.0  # read+write R,V as float2


def load_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return json.load(f), "measured"
    except Exception:
        return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "sm_max_mhz": 1965.0}, "fallback"


class ClockSampler:
    """Samples nvidia-smi clocks / throttle reasons during the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index = index
        self.rows = []
        self.proc = None
        self.thread = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                 "-i", str(self.index), "-lms", "100"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.proc = None
            return
        def pump():
            for line in self.proc.stdout:
                self.rows.append([c.strip() for c in line.split(",")])
        self.thread = threading.Thread(target=pump, daemon=True)
        self.thread.start()

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, smax, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                smax.append(float(r[1]))
                for n, v in zip(names, r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
            except Exception:
                continue
        return {"sm_mhz": statistics.median(sm) if sm else None,
                "sm_max_mhz": max(smax) if smax else None,
                "samples": len(sm), "reasons": sorted(reasons)}


# ------------------------------------------------------------------------------------------
def cpu_reference_step_time(N, rc, dt, n_md_steps, seed=0):
    """Times the reference's verlet_step as the reference defines it (MD:66-75): two dense
    autodiff force evaluations per step (F is not carried), torch CPU fp32, all host threads.
    Returns (seconds per MD step, threads)."""
    import torch
    from oracle import lj_oracle as O
    from jax_tpus_benchmark_physics_simulation_b200 import lattice_jitter
    R, V, box = lattice_jitter(N, seed=seed)
    state = (torch.from_numpy(R), torch.from_numpy(V))
    if N <= 8192:
        ff = lambda r: O.force_autodiff(r, float(box), rc=rc)
    else:   # N x N does not fit: row-chunked restatement of the same formulas
        ff = lambda r: O.force_analytic(r, float(box), rc=rc)[0]
    state = O.verlet_step(state, float(box), dt, ff)            # warm-up call
    t0 = time.perf_counter()
    for _ in range(n_md_steps):
        state = O.verlet_step(state, float(box), dt, ff)
    return (time.perf_counter() - t0) / n_md_steps, torch.get_num_threads()


def run_reference(args, wl_name, wl):
    """--impl reference: the CPU restatement timed on this box's host cores (rank 0 only)."""
    if int(os.environ.get("RANK", "0")) != 0:
        return
    N, rc, dt = wl["N"], wl["rc"], wl["dt"]
    if wl["path"] == "cells":
        # no all-pairs CPU baseline exists at N >= 4M (1.8e13 pairs / evaluation): time the C
        # restatement's cell-grid force evaluation (one per step, F carried) instead.
        from oracle import lj_oracle as O
        from jax_tpus_benchmark_physics_simulation_b200 import lattice_jitter
        R, V, box = lattice_jitter(N, seed=0)
        O.c_forces_cells(R, box, rc)
        times = []
        for i in range(args.warmup + args.steps):
            t0 = time.perf_counter()
            O.c_forces_cells(R, box, rc)
            if i >= args.warmup:
                times.append(time.perf_counter() - t0)
        t = sum(times)
        cores = os.cpu_count()
        sample = "1 cell-grid force evaluation (C restatement, OpenMP) per step"
        md_per_step = 1
    else:
        import torch
        from oracle import lj_oracle as O
        from jax_tpus_benchmark_physics_simulation_b200 import lattice_jitter
        md_per_step = 1
        R, V, box = lattice_jitter(N, seed=0)
        state = (torch.from_numpy(R), torch.from_numpy(V))
        if N <= 8192:
            ff = lambda r: O.force_autodiff(r, float(box), rc=rc)
        else:   # N x N does not fit in memory: row-chunked restatement of the same formulas
            ff = lambda r: O.force_analytic(r, float(box), rc=rc)[0]
        times = []
        for i in range(args.warmup + args.steps):
            t0 = time.perf_counter()
            state = O.verlet_step(state, float(box), dt, ff)       # MD:66-75, 2 force evaluations
            if i >= args.warmup:
                times.append(time.perf_counter() - t0)
        t = sum(times)
        cores = torch.get_num_threads()
        sample = (f"{md_per_step} verlet_step (2 dense autodiff force evaluations, torch CPU fp32) "
                  "per bench step; restatement, JAX unavailable")
    value = N * md_per_step * args.steps / t
    line = {
        "impl": "reference", "metric": "particle-steps/sec", "value": value,
        "unit": "particle-steps/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * t / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic lattice+jitter (seed 0)",
        "config": {"workload": wl_name, "desc": wl["desc"], "md_steps_per_step": md_per_step},
        "cpu_baseline": {"value": value, "unit": "particle-steps/s", "cores": cores,
                         "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "particle-steps/s", "h2d_bytes_per_step": 0,
                "d2h_bytes_per_step": 0},
        "pair_interactions_per_s": value * (N - 1) if wl["path"] == "allpairs" else None,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default=None, choices=list(WORKLOADS))
    ap.add_argument("--md-steps", type=int, default=None, help="MD steps per bench step")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup

    wl_name = args.workload or "ap4096"
    wl = dict(WORKLOADS[wl_name])
    if args.md_steps:
        wl["md_steps"] = args.md_steps
    if args.impl == "reference":
        run_reference(args, wl_name, wl)
        return

    import numpy as np
    import torch
    import torch.distributed as dist
    from jax_tpus_benchmark_physics_simulation_b200 import lattice_jitter
    from jax_tpus_benchmark_physics_simulation_b200.md import LJSimulation, fp32_peak_probe

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a GPU (no CPU fallback on the product path)")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    peaks, peaks_src = load_peaks()

    N, rc, dt, md_steps = wl["N"], wl["rc"], wl["dt"], wl["md_steps"]
    R, V, box = lattice_jitter(N, seed=0)
    dist_arg = None
    parallelism = "single GPU"
    if world > 1:
        if N >= 16384:
            # atom decomposition (all-pairs) / row slabs with halo exchange (cell list) inside the library
            from jax_tpus_benchmark_physics_simulation_b200.md import make_dist_arg
            dist_arg = make_dist_arg(rank, world)
            parallelism = f"sharded x{world}"
        else:
            # a step of this system is a few microseconds: it does not shard
            parallelism = f"replicas x{world}"
    sim = LJSimulation(N, rc=rc, dt=dt, path=wl["path"], device=local_rank, dist=dist_arg,
                       skin=wl.get("skin", 0.3))

    # device-resident inputs for `value`
    Rd = torch.from_numpy(R).cuda()
    Vd = torch.from_numpy(V).cuda()
    # pinned host buffers for `e2e`
    Rh = torch.from_numpy(R).pin_memory()
    Vh = torch.from_numpy(V).pin_memory()
    Rh_out = torch.empty_like(Rh).pin_memory()
    Vh_out = torch.empty_like(Vh).pin_memory()
    E_out = torch.empty((1, 2), dtype=torch.float32).pin_memory()
    flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device="cuda")  # > 126 MB L2

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    state = (Rd, Vd)
    for _ in range(args.warmup):
        state, _ = sim.run(state, md_steps)
    barrier()

    # ---- timed region: K steps, device-resident state, L2 flushed between steps ----------------
    sampler = ClockSampler(local_rank)
    sampler.start()
    launches0 = sim.launch_count()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
          for _ in range(args.steps)]
    kernel_ms = []
    barrier()
    wall0 = time.perf_counter()
    for k in range(args.steps):
        flush.zero_()                                   # evict L2 between timed steps
        ev[k][0].record()
        state, _ = sim.run(state, md_steps)
        ev[k][1].record()
        kernel_ms.append(None)
    barrier()
    wall = time.perf_counter() - wall0
    launches = sim.launch_count() - launches0
    step_ms = [a.elapsed_time(b) for a, b in ev]
    t_dev = sum(step_ms) * 1e-3
    clocks = sampler.stop()
    if world > 1:
        tt = torch.tensor([t_dev], dtype=torch.float64, device="cuda")
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        t_dev = float(tt.item())
    sharded = parallelism.startswith("sharded")
    total_particles = N if (world == 1 or sharded) else N * world
    value = total_particles * md_steps * args.steps / t_dev

    # ---- e2e: host buffers in, host buffers out, copies inside the timed region -----------------
    barrier()
    e0 = time.perf_counter()
    for k in range(args.steps):
        Rk = Rh.to("cuda", non_blocking=True)
        Vk = Vh.to("cuda", non_blocking=True)
        (Ro, Vo), _ = sim.run((Rk, Vk), md_steps, energy_every=md_steps)
        Rh_out.copy_(Ro.tensor, non_blocking=True)
        Vh_out.copy_(Vo.tensor, non_blocking=True)
        E_out.copy_(sim.last_energies.tensor, non_blocking=True)
        torch.cuda.synchronize()
    barrier()
    t_e2e = time.perf_counter() - e0
    if world > 1:
        tt = torch.tensor([t_e2e], dtype=torch.float64, device="cuda")
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        t_e2e = float(tt.item())
    e2e_value = total_particles * md_steps * args.steps / t_e2e
    h2d = Rh.numel() * 4 + Vh.numel() * 4
    d2h = Rh_out.numel() * 4 + Vh_out.numel() * 4 + E_out.numel() * 4

    # one more run on EVERY rank (a sharded run is collective) for the per-launch kernel time
    barrier()
    sim.run(state, md_steps)
    launch_ms = sim.last_run_ms()
    rebuilds = sim.last_rebuilds() if wl["path"] == "cells" else 0
    barrier()

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel ----------------------------------------------------------
    sm_max = float(peaks.get("sm_max_mhz", 1965.0))
    if wl["path"] == "allpairs":
        # persistent kernel: one launch = md_steps steps + the prologue force evaluation
        evals = md_steps + 1
        pairs = float(N) * float(N - 1) * evals / (world if sharded else 1)      # ORDERED pairs
        ap_mode = sim.allpairs_mode()
        n3l = ap_mode == 3
        # executed work: the N3L kernel evaluates each unordered pair once (report that count
        # for roofline.achieved; the headline pair rate keeps the ordered-pair denominator)
        executed_flop = (FLOP_PER_UNORDERED_N3L * pairs / 2.0) if n3l else (FLOP_PER_PAIR_FORCE * pairs)
        achieved = executed_flop / (launch_ms * 1e-3) / 1e12
        peak = 148 * 128 * 2 * sm_max * 1e6 / 1e12      # whole GPU (the cluster kernel uses 16 of the 148 SMs)
        try:
            probe = {"ffma_tflops": fp32_peak_probe(local_rank, False),
                     "ffma2_tflops": fp32_peak_probe(local_rank, True)}
        except Exception as e:  # pragma: no cover
            probe = {"error": str(e)}
        roofline = {"bound": "fp32", "achieved": achieved, "peak": peak, "unit": "TFLOP/s",
                    "frac": achieved / peak, "traffic": None,
                    "peak_source": f"148 SM x 128 lanes x 2 x sm_max_mhz={sm_max:.0f} from "
                                   f"MEASURED_PEAKS.json ({peaks_src}); CUDA-core FP32, no tensor cores",
                    "measured_fp32_probe": probe,
                    "kernel": "ap_cluster_kernel" if ap_mode == 4 else "ap_persistent_kernel",
                    "mode": "newton3 tiles (each unordered pair once)" if n3l
                            else ("ordered pairs, one 16-CTA cluster, state in distributed shared memory"
                                  if ap_mode == 4 else "ordered pairs"),
                    "flop_per_evaluation": FLOP_PER_UNORDERED_N3L if n3l else FLOP_PER_PAIR_FORCE,
                    "ordered_pairs_per_launch": pairs,
                    "ordered_pair_equivalent_tflops": FLOP_PER_PAIR_FORCE * pairs / (launch_ms * 1e-3) / 1e12,
                    "launch_ms": launch_ms}
        if n3l and world == 1 and wl_name in AP_TRAFFIC:
            # dram__bytes_read.sum + dram__bytes_write.sum of the committed ncu --set full capture, per force
            # evaluation, scaled to this launch: the state and the partial vectors are L2 resident
            per_eval, src = AP_TRAFFIC[wl_name]
            roofline["traffic"] = per_eval * evals
            roofline["traffic_source"] = src
    else:
        bytes_ = BYTES_PER_PARTICLE_STEP * N * md_steps
        achieved = bytes_ / (launch_ms * 1e-3) / 1e9
        peak = float(peaks["hbm_gbs"]) * (world if sharded else 1)     # aggregate over the sharded GPUs
        roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                    "frac": achieved / peak, "traffic": None,
                    "peak_source": f"MEASURED_PEAKS.json hbm_gbs ({peaks_src})",
                    "bytes_per_particle_step": BYTES_PER_PARTICLE_STEP,
                    "kernel": "cells_persistent_kernel", "run_ms": launch_ms, "rebuilds": rebuilds}
        if world == 1:
            # the HBM measure is the contract, but the pair loop is FP32-issue bound; report the
            # second roofline and the logically gathered bytes next to it (neighbour statistics of the
            # initial configuration, counted on the device by the library's own recount entry point)
            skin = wl.get("skin", 0.3)
            n_cut = float(sim.neighbor_count(Rd, rc).float().mean().item())
            n_list = float(sim.neighbor_count(Rd, rc + skin).float().mean().item())
            fp32_peak = 148 * 128 * 2 * sm_max * 1e6 / 1e12
            per_s = N * md_steps / (launch_ms * 1e-3)
            roofline["fp32"] = {
                "neighbours_within_rc": n_cut, "neighbours_in_list": n_list,
                "useful_tflops": FLOP_PER_PAIR_FORCE * n_cut * per_s / 1e12,
                "evaluated_tflops": FLOP_PER_PAIR_FORCE * n_list * per_s / 1e12,
                "peak": fp32_peak, "frac_evaluated": FLOP_PER_PAIR_FORCE * n_list * per_s / 1e12 / fp32_peak,
                "binding": "FP32 issue: ~12.5 instructions per listed neighbour (ncu: issue slots 60 % busy, "
                           "DRAM 24 %), see DESIGN.md 4.2"}
            # state + one list byte + one 8-byte position per listed neighbour (positions come from the
            # warp's shared-memory windows, staged once per 32 particles)
            roofline["gathered_bytes_per_particle_step"] = BYTES_PER_PARTICLE_STEP + n_list * (1.0 + 8.0)
            roofline["gathered_GBps"] = roofline["gathered_bytes_per_particle_step"] * per_s / 1e9
            if wl_name == "cells4m":
                # dram__bytes_read.sum + dram__bytes_write.sum of profiles/r1_cells4m_final.ncu-rep
                # (same N, skin; 30 steps incl. the first sort): 10.25 GB / (30 x 4,194,304) particle-steps
                roofline["traffic"] = 81.5 * N * md_steps
                roofline["traffic_source"] = ("81.5 B per particle-step measured by ncu --set full on a 30-step "
                                              "launch (profiles/r1_cells4m_final.ncu-rep), scaled to this launch")

    # ---- CPU baseline: bounded sample of the same workload on this box's host cores ------------
    cpu_baseline = None
    if not args.no_cpu_baseline and world == 1:
        try:
            if wl["path"] == "allpairs":
                nsteps_cpu = 4 if N <= 4096 else 1
                per, cores = cpu_reference_step_time(N, rc, dt, nsteps_cpu)
                cpu_baseline = {"value": N / per, "unit": "particle-steps/s", "cores": cores,
                                "kind": "port",
                                "sample": f"{nsteps_cpu} verlet_step of the torch-CPU fp32 restatement "
                                          "(2 dense autodiff force evaluations per step, MD:66-75), "
                                          "after 1 warm-up step; restatement, JAX unavailable"}
            else:
                from oracle import lj_oracle as O
                t0 = time.perf_counter()
                O.c_forces_cells(R, box, rc)
                per = time.perf_counter() - t0
                cpu_baseline = {"value": N / per, "unit": "particle-steps/s", "cores": os.cpu_count(),
                                "kind": "port",
                                "sample": "1 cell-grid force evaluation of the C restatement (OpenMP); "
                                          "no all-pairs CPU baseline exists at this N"}
        except Exception as e:  # pragma: no cover
            cpu_baseline = {"error": str(e)}

    line = {
        "metric": "particle-steps/sec", "value": value, "unit": "particle-steps/s",
        "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * t_dev / args.steps, "higher_is_better": True,
        "scaling": "strong" if sharded else "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic lattice+jitter (seed 0, jitter 0.05, kT 1.0)",
        "config": {"workload": wl_name, "desc": wl["desc"], "N": N, "rc": rc, "dt": dt,
                   "md_steps_per_step": md_steps, "path": wl["path"], "parallelism": parallelism,
                   "l2": "flushed (256 MiB write) between timed steps; state is L2/SMEM resident by design"},
        "pair_interactions_per_s": (value * (N - 1)) if wl["path"] == "allpairs" else None,
        "us_per_md_step": 1e6 * t_dev / (args.steps * md_steps),
        "wall_s": wall, "clocks": clocks,
        "e2e": {"value": e2e_value, "unit": "particle-steps/s", "h2d_bytes_per_step": h2d,
                "d2h_bytes_per_step": d2h},
        "gpu_launches": int(launches),
        "roofline": roofline, "cpu_baseline": cpu_baseline,
    }
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
