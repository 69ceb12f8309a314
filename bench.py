#!/usr/bin/env python
"""Benchmark of the dynamics-guided diffusion sampling path (BASELINE.json metric: guided designs/sec).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path, N GPUs of one node
    python bench.py --impl reference --steps K --warmup W    # the reference algorithm on the host cores

One "step" = one complete guided-sampling pass over the workload: the full 5-step DDIM schedule (denoiser
forward, dynamics fwd + input-gradient over every candidate x pose row, guided update), the forward-only
scoring pass and best-of-N.  N = 1 runs BASELINE.json configs[1] ("C2": 2D, 64 synthetic objects x 256 candidates
x 36 orientations x 5x5 positions = 900 pose rows per candidate, fp32-grade arithmetic).  N > 1 is weak scaling:
every rank runs 64 objects of a 64*N-object set, no per-step traffic, one NCCL all-gather of scores / best designs
per pass.

The JSON line carries, besides the driver's contract keys:
  roofline     tensor-pipe roofline of the dominant kernel (fused tcgen05 trunk): algorithmic FLOPs
               (1 838 080 per guidance row) / CUDA-event time of that kernel inside the timed region.
               In the fp32-grade mode every product is computed with 3 bf16 MMAs, so the peak it is held against is
               the measured sustained bf16 peak / 3 (stated in `peak_basis`).
  cpu_baseline the CPU oracle port of the reference sampler (oracle/dgdm_oracle.py) on a bounded sample of the
               same workload, on this box's host cores.
  e2e          the same metric through the public Python API with HOST (pinned) inputs and outputs, H2D and D2H
               copies inside the timed region.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time

REPO = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, REPO)

FLOP_PER_ROW_2D = 2 * 2 * (7 * 256 * 256 + 256 * 3)          # 1 838 080, SURVEY.md §8d
FLOP_PER_ROW_3D = 2 * 2 * (512 * 256 + 6 * 256 * 256 + 256 * 3)   # 2 100 224
N_OBJ, N_CAND, GRID, NPOS, P = 64, 256, 36, 5, 14
MODE = "point"
T_TRAIN, T_INF = 15, 5
OBJECTIVE = "rotate_clockwise"
WORKLOAD = "C2 (BASELINE.json configs[1])"


def select_workload(name):
    """c2 (default, the metric's configuration) or c3 (BASELINE.json configs[2]: 3D point clouds, 128 candidates,
    stock 3D pose grid 45 x 5 x 5; object count unspecified there -> 64)."""
    global N_OBJ, N_CAND, GRID, NPOS, P, MODE, WORKLOAD
    if name == "c3":
        N_OBJ, N_CAND, GRID, NPOS, P, MODE = 64, 128, 45, 5, 42, "point_3d"
        WORKLOAD = "C3 (BASELINE.json configs[2])"
    if name == "c2g36":   # the literal "x36 orientations" reading of configs[1]: num_pos = 1 => G = 36 (second point)
        N_OBJ, N_CAND, GRID, NPOS, P, MODE = 64, 256, 36, 1, 14, "point"
        WORKLOAD = "C2 literal reading (num_pos=1, 36 pose rows/candidate; denoiser ~45 % of FLOPs)"
    if name == "c5":   # 8 x B200 sweep: 1024 objects x 512 candidates => 128 objects per GPU (weak scaling below 8 GPUs)
        N_OBJ, N_CAND, GRID, NPOS, P, MODE = 128, 512, 45, 5, 42, "point_3d"
        WORKLOAD = "C5 (BASELINE.json configs[4], per-GPU shard of 1024 objects x 512 candidates)"


def peaks():
    path = os.path.join(REPO, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        d = json.load(open(path))
        return {"hbm_gbs": d["hbm_gbs"], "bf16_burst": d["bf16_tflops"], "bf16_sustained": d["bf16_tflops_sustained"],
                "src": "measured"}
    return {"hbm_gbs": 6650.0, "bf16_burst": 1590.0, "bf16_sustained": 1400.0, "src": "fallback"}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(gpu_index), f"--query-gpu={self.Q}",
                                       "--format=csv,noheader,nounits", "-lms", "200"], stdout=self.f,
                                      stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.p is None:
            return out
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        rows = [l.strip().split(", ") for l in open(self.f.name) if l.strip()]
        os.unlink(self.f.name)
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2]))
                for nm, v in zip(names, r[5:9]):
                    if v.strip().lower() == "active":
                        reasons.add(nm)
            except Exception:
                pass
        if sm:
            out.update(sm_mhz=statistics.median(sm), sm_max_mhz=max(mx), reasons=sorted(reasons), samples=len(sm))
        return out


def oracle_sampler(n_obj):
    sys.path.insert(0, os.path.join(REPO, "oracle"))
    import dgdm_oracle as orc
    from dgdm_b200 import synthetic as syn
    if MODE == "point_3d":
        return orc.OracleSampler("point_3d", syn.unet1d_state_dict(0), syn.dynamics3d_state_dict(0), syn.objects_3d(n_obj),
                                 GRID, NPOS, T_TRAIN, T_INF, sub_batch_size=512, fps_start=syn.fps_starts(n_obj))
    return orc.OracleSampler("point", syn.unet1d_state_dict(0), syn.dynamics2d_state_dict(0), syn.objects_2d(n_obj),
                             GRID, NPOS, T_TRAIN, T_INF)


def cpu_pass(samp, noise):
    """One full pass of the reference algorithm (oracle port) on the CPU: 5 guided steps + scoring + argmax."""
    import torch
    import dgdm_oracle as orc
    designs = samp.guided_sample(noise, OBJECTIVE)
    scores = torch.stack([samp.score(designs[o], o, OBJECTIVE) for o in range(designs.shape[0])])
    return designs, scores, orc.best_of_n(scores)


def run_reference(args):
    """--impl reference: the reference's own algorithm for this path on the host cores.  The reference is pure
    Python/torch and /root/reference does not exist on the GPU box, so this times the oracle port
    (oracle/dgdm_oracle.py: un-hoisted, autograd backward, full layer-1 GEMM at B*G rows), multi-threaded."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import torch
    from dgdm_b200 import synthetic as syn
    torch.set_num_threads(os.cpu_count() or 1)
    select_workload(args.workload)
    n_obj_s, b_s = 1, (32 if MODE == "point" else 8)      # bounded sample: 1 object x 32 (8) candidates x G rows
    samp = oracle_sampler(n_obj_s)
    noise = syn.initial_noise(b_s, P)
    for _ in range(max(0, args.warmup)):
        cpu_pass(samp, noise)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        cpu_pass(samp, noise)
    dt = time.perf_counter() - t0
    val = n_obj_s * b_s * args.steps / dt
    sample = f"{n_obj_s} object x {b_s} candidates x {GRID * NPOS * NPOS} pose rows x {T_INF} steps + scoring, per step"
    line = {"impl": "reference", "metric": "guided designs/sec", "value": val, "unit": "designs/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": dict(workload_config(args.gpus, args.precision),
                           reference_arm="CPU oracle port of the reference sampler, host cores only; each step is the "
                                         "bounded sample in cpu_baseline.sample"),
            "cpu_baseline": {"value": val, "unit": "designs/s", "cores": torch.get_num_threads(), "kind": "port",
                             "sample": sample},
            "e2e": {"value": val, "unit": "designs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


def workload_config(n_gpus, precision):
    return {"workload": f"{WORKLOAD}: {'3D' if MODE == 'point_3d' else '2D'} guided sampling, {N_OBJ} objects/GPU x {N_CAND} candidates x "
                        f"{GRID} orientations x {NPOS}x{NPOS} positions = {GRID * NPOS * NPOS} pose rows/candidate, "
                        f"{T_INF} DDIM steps of {T_TRAIN}, objective {OBJECTIVE}, + scoring pass + best-of-N",
            "objects_per_gpu": N_OBJ, "candidates": N_CAND, "pose_rows": GRID * NPOS * NPOS, "ddim_steps": T_INF,
            "guidance_rows_per_pass": N_OBJ * N_CAND * GRID * NPOS * NPOS * T_INF, "precision": precision,
            "sharding": f"objects x{n_gpus}, no per-step traffic, 1 all-gather/pass",
            "l2": "256 MiB memset between timed passes (L2 flush), inside the bracket (<0.1% of a pass)"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--precision", default="fp32", choices=["fp32", "fp16x3", "bf16", "fp16", "fp32_simt"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--workload", default="c2", choices=["c2", "c2g36", "c3", "c5"])
    args = ap.parse_args()
    select_workload(args.workload)
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist
    from dgdm_b200 import _lib, synthetic as syn
    from dgdm_b200 import distributed as D
    from dgdm_b200.diffusion import Diffusion
    from dgdm_b200.scheduler import DDIMScheduler
    import ctypes as C

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    lib = _lib.lib()
    if args.warmup < 3:
        args.warmup = 3

    # ---- workload: this rank's shard of the 64*world objects -------------------------------------------------
    n_obj_global = N_OBJ * world
    lo, hi = D.shard_range(n_obj_global, world, rank)
    is3d = MODE == "point_3d"
    objs_host = (syn.objects_3d(n_obj_global) if is3d else syn.objects_2d(n_obj_global))[lo:hi].contiguous().pin_memory()
    fps_host = syn.fps_starts(n_obj_global)[lo:hi].contiguous() if is3d else None
    noise_host = syn.initial_noise(N_CAND, P).pin_memory()
    dm = Diffusion(syn.unet1d_state_dict(0), DDIMScheduler(T_TRAIN), T_INF, mode=MODE, num_points=P,
                   classifier_model=syn.dynamics3d_state_dict(0) if is3d else syn.dynamics2d_state_dict(0),
                   grid_size=GRID, num_pos=NPOS, object_vertices=objs_host, object_ids=list(range(lo, hi)),
                   fps_starts=fps_host, precision=args.precision, device=dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    noise_dev = noise_host.to(dev)

    def one_pass_resident():
        local = dm.guided_sample(0, N_CAND, noise_dev, opt_obj=OBJECTIVE)
        return D.gather_per_object_results(local, n_obj_global, gather_designs=False)

    out_host = {}

    def one_pass_e2e():
        # host -> device: this pass's inputs from pinned memory; device -> host: designs, scores, best ids
        nz = noise_host.to(dev, non_blocking=True)
        dm.set_objects(objs_host.to(dev, non_blocking=True), fps_host)      # 3D: PointNet++ (K5) runs here, every pass
        local = dm.guided_sample(0, N_CAND, nz, opt_obj=OBJECTIVE)
        g = D.gather_per_object_results(local, n_obj_global, gather_designs=False)
        for k in ("scores", "best_ids", "best_scores"):
            out_host[k] = g[k].to("cpu", non_blocking=True)
        out_host["designs"] = local["designs"].to("cpu", non_blocking=True)
        return g

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps, with_kernel_timing=False):
        for _ in range(args.warmup):
            fn()
            flush.zero_()
        barrier()
        if with_kernel_timing:
            lib.dgdm_trunk_timing(1)
        l0 = lib.dgdm_launch_count()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
            flush.zero_()
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        launches = lib.dgdm_launch_count() - l0
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms, launches

    clocks = ClockSampler(local_rank) if rank == 0 else None
    ms, launches = timed(one_pass_resident, args.steps, with_kernel_timing=True)
    tot = C.c_double(); nl = C.c_int64(); nrows = C.c_int64()
    lib.dgdm_trunk_timing_read(C.byref(tot), C.byref(nl), C.byref(nrows))
    lib.dgdm_trunk_timing(0)
    clk = clocks.stop() if clocks else {}
    ms_e2e, _ = timed(one_pass_e2e, args.steps)

    designs_per_pass = n_obj_global * N_CAND
    value = designs_per_pass * args.steps / (ms / 1e3)
    e2e_value = designs_per_pass * args.steps / (ms_e2e / 1e3)

    if rank == 0:
        pk = peaks()
        x3 = args.precision in ("fp32", "fp16x3")
        # dominant kernel: all launches inside the timed region (5 backward + 1 forward-only per pass)
        G = GRID * NPOS * NPOS
        bwd_rows = (hi - lo) * N_CAND * G * T_INF * args.steps
        fwd_rows = int(nrows.value) - bwd_rows                      # scoring-pass rows (forward only: half the FLOPs)
        fpr = FLOP_PER_ROW_3D if is3d else FLOP_PER_ROW_2D
        flops = bwd_rows * fpr + max(0, fwd_rows) * (fpr // 2)
        achieved = flops / (tot.value / 1e3) / 1e12 if tot.value > 0 else 0.0
        peak = pk["bf16_sustained"] / (3.0 if x3 else 1.0)
        traffic = None                       # DRAM bytes per launch of this kernel from the committed ncu --set full capture
        tpath = os.path.join(REPO, "profiles", "tc_trunk_traffic.json")
        if os.path.exists(tpath) and args.workload == "c2":   # captured on C2 only
            traffic = json.load(open(tpath)).get("fp32" if x3 else "bf16", {}).get("dram_bytes_per_launch")
        roofline = {"bound": "tensor", "achieved": achieved, "peak": peak, "unit": "TFLOP/s",
                    "frac": achieved / peak if peak else None, "traffic": traffic,
                    "kernel": "tc_trunk_kernel (fused tcgen05 trunk fwd+dgrad+reduction)",
                    "peak_basis": f"bf16_tflops_sustained ({pk['bf16_sustained']}) of {pk['src']}"
                                  + (" / 3: fp32-grade mode issues 3 bf16 MMAs per product" if x3 else ""),
                    "launches": int(nl.value), "avg_launch_ms": tot.value / max(1, nl.value),
                    "kernel_share_of_step": tot.value / ms,
                    "executed_tflops": achieved * (3.0 if x3 else 1.0),
                    "flop_per_row": fpr}
        cpu = None
        if not args.no_cpu_baseline and args.gpus == 1:
            torch.set_num_threads(os.cpu_count() or 1)
            samp = oracle_sampler(1)
            nb = 32 if not is3d else 8
            nz = syn.initial_noise(nb, P)
            t0 = time.perf_counter()
            reps = 0
            while True:
                cpu_pass(samp, nz)
                reps += 1
                if time.perf_counter() - t0 > 10.0 or reps >= 8:
                    break
            dt = time.perf_counter() - t0
            cpu = {"value": nb * reps / dt, "unit": "designs/s", "cores": torch.get_num_threads(), "kind": "port",
                   "sample": f"{reps} x (1 object x {nb} candidates x {G} pose rows x {T_INF} steps + scoring) of the same workload"
                             + (" (PointNet++ hoisted to once per object, unlike the as-written reference)" if is3d else "")}
        h2d = noise_host.numel() * 4 + objs_host.numel() * 4
        d2h = sum(v.numel() * v.element_size() for v in out_host.values())
        line = {"metric": "guided designs/sec", "value": value, "unit": "designs/s", "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "bf16x3 (fp32-grade)" if x3 else args.precision, "data": "synthetic",
                "config": workload_config(world, args.precision),
                "denoise_steps_per_sec": value * T_INF,
                "roofline": roofline, "cpu_baseline": cpu,
                "e2e": {"value": e2e_value, "unit": "designs/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                        "ms_per_step": ms_e2e / args.steps},
                "gpu_launches": int(launches), "clocks": clk}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
