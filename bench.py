#!/usr/bin/env python
"""Benchmark of the dynamics-guided diffusion sampling path (BASELINE.json metric: guided designs/sec).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path, N GPUs of one node
    python bench.py --impl reference --steps K --warmup W    # the reference's own sampler on the host cores

One "step" = one complete guided-sampling pass over the workload: the full 5-step DDIM schedule (denoiser
forward, dynamics fwd + input-gradient over every candidate x pose row, guided update), the forward-only
scoring pass and best-of-N.

Headline line (`value`, `e2e`, `roofline`): BASELINE.json configs[1] ("C2": 2D, 64 synthetic objects per GPU x 256
candidates x 36 orientations x 5x5 positions = 900 pose rows per candidate, fp32-grade arithmetic), weak scaling over
N GPUs (64 objects per rank, no per-step traffic, one NCCL all-gather of scores / best ids / designs per pass).

`extra` (same process, same timing rules, fewer steps): the 3D / single-pass configurations --
  c3_bf16, c3_fp16       BASELINE.json configs[2]: 3D, 64 objects per GPU x 128 candidates x 45 x 5x5 = 1125 pose rows
  c5_strong_bf16         BASELINE.json configs[4]: the FIXED 1024 objects x 512 candidates 3D sweep sharded over the N
                         ranks (strong scaling: N = 8 is the north-star target configuration; N = 1 runs all of it on
                         one GPU in object chunks)
each with its own `value`, `e2e` (host buffers, designs gathered inside the bracket) and `roofline`.
`--workload X --precision Y` runs one configuration alone as the headline instead (no `extra`).

Keys beside the driver's contract:
  roofline     tensor-pipe roofline of the dominant kernel (fused tcgen05 trunk, guidance launches = fwd + dgrad):
               algorithmic FLOPs (1 838 080 per 2D guidance row, 2 100 224 per 3D row) / CUDA-event time
               of those launches inside the timed region; the forward-only scoring launches are reported separately
               (`scoring`).  In the fp32-grade modes every product is 3 MMAs, so the peak is the measured sustained
               bf16 peak / 3 (`peak_basis`).
  cpu_baseline the REFERENCE's own `Diffusion.guided_sample` (its modules byte-compiled under oracle/_ref, oracle/ref_arm.py)
               on a bounded slice of the same workload on this box's host cores; `port` = the oracle restatement
               (hoist-free but without the reference's Python tiling) on the same slice; `c1_as_is` = the reference on
               BASELINE.json configs[0] (1 object x 16 candidates x 360 x 5x5 = 9000 pose rows), one pass.
  e2e          the same metric through the public Python API with HOST (pinned) inputs and outputs: H2D of noise and
               objects, (3D: PointNet++,) the pass, the all-gather of scores / best ids / designs, and the D2H of all
               of them inside the timed region.
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
T_TRAIN, T_INF = 15, 5
OBJECTIVE = "rotate_clockwise"

# name -> objects per GPU (weak) or in total (strong), candidates, grid_size, num_pos, control points, mode
WORKLOADS = {
    "c2": dict(n_obj=64, n_cand=256, grid=36, npos=5, P=14, mode="point", strong=False,
               label="C2 (BASELINE.json configs[1])"),
    # the literal "x36 orientations" reading of configs[1]: num_pos = 1 => G = 36 (second point)
    "c2g36": dict(n_obj=64, n_cand=256, grid=36, npos=1, P=14, mode="point", strong=False,
                  label="C2 literal reading (num_pos=1, 36 pose rows/candidate; denoiser ~45 % of FLOPs)"),
    # configs[2]: 3D point clouds, 128 candidates, stock 3D pose grid; object count unspecified there -> 64
    "c3": dict(n_obj=64, n_cand=128, grid=45, npos=5, P=42, mode="point_3d", strong=False,
               label="C3 (BASELINE.json configs[2])"),
    # configs[4] as weak scaling: the per-GPU shard of the 8-GPU sweep on every rank
    "c5": dict(n_obj=128, n_cand=512, grid=45, npos=5, P=42, mode="point_3d", strong=False,
               label="C5 per-GPU shard (BASELINE.json configs[4]: 128 of 1024 objects x 512 candidates per GPU)"),
    # configs[4] as written: the fixed problem over however many GPUs there are
    "c5_strong": dict(n_obj=1024, n_cand=512, grid=45, npos=5, P=42, mode="point_3d", strong=True,
                      label="C5 (BASELINE.json configs[4]): fixed 1024 objects x 512 candidates sharded over the ranks"),
}


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


# ------------------------------------------------------------------------------------------------ CPU arms
def _synthetic_objects(wl, n):
    from dgdm_b200 import synthetic as syn
    return syn.objects_3d(n) if wl["mode"] == "point_3d" else syn.objects_2d(n)


def oracle_sampler(wl, n_obj):
    sys.path.insert(0, os.path.join(REPO, "oracle"))
    import dgdm_oracle as orc
    from dgdm_b200 import synthetic as syn
    if wl["mode"] == "point_3d":
        return orc.OracleSampler("point_3d", syn.unet1d_state_dict(0), syn.dynamics3d_state_dict(0), syn.objects_3d(n_obj),
                                 wl["grid"], wl["npos"], T_TRAIN, T_INF, sub_batch_size=512, fps_start=syn.fps_starts(n_obj))
    return orc.OracleSampler("point", syn.unet1d_state_dict(0), syn.dynamics2d_state_dict(0), syn.objects_2d(n_obj),
                             wl["grid"], wl["npos"], T_TRAIN, T_INF)


def port_pass(samp, noise):
    """One full pass of the oracle port on the CPU: 5 guided steps + scoring + argmax."""
    import torch
    import dgdm_oracle as orc
    designs = samp.guided_sample(noise, OBJECTIVE)
    scores = torch.stack([samp.score(designs[o], o, OBJECTIVE) for o in range(designs.shape[0])])
    return designs, scores, orc.best_of_n(scores)


def reference_available():
    sys.path.insert(0, os.path.join(REPO, "oracle"))
    try:
        import ref_arm
        return ref_arm.available()
    except Exception:
        return False


def cpu_slice(wl):
    """The bounded slice of a workload a CPU arm runs per step: 1 object x b candidates x the workload's full pose grid
    x the full 5-step schedule.  3D: the reference as written runs PointNet++ on every guidance row (1.1 GFLOP/row,
    SURVEY.md §8a-7), so its slice is 1 candidate."""
    return 1, (32 if wl["mode"] == "point" else 1)


class CpuArm:
    """The reference's own guided_sample (kind 'reference') or, when oracle/_ref did not travel, the oracle port."""

    def __init__(self, wl, prefer_reference=True):
        import torch
        from dgdm_b200 import synthetic as syn
        torch.set_num_threads(os.cpu_count() or 1)
        self.cores = torch.get_num_threads()
        self.wl = wl
        self.n_obj, self.b = cpu_slice(wl)
        self.G = wl["grid"] * wl["npos"] ** 2
        # 3D: the reference as written re-runs PointNet++ on every guidance row (~3500 s per object per pass on 8 threads,
        # SURVEY.md §8d) -- not a bounded sample; the 3D CPU arm is the port (PointNet++ once per object)
        self.kind = "reference" if prefer_reference and reference_available() and wl["mode"] == "point" else "port"
        self.noise = syn.initial_noise(self.b, wl["P"])
        if self.kind == "reference":
            import ref_arm
            self.dm = ref_arm.build(wl["mode"], _synthetic_objects(wl, self.n_obj), wl["grid"], wl["npos"], sub_batch_size=512)
            self._run = lambda: ref_arm.guided_sample(self.dm, self.noise, OBJECTIVE)
        else:
            samp = oracle_sampler(wl, self.n_obj)
            self._run = lambda: port_pass(samp, self.noise)

    def step(self):
        self._run()

    def describe(self):
        what = ("the reference's own Diffusion.guided_sample (generator/diffusion.py:541-576, its unmodified modules, byte-compiled under "
                "oracle/_ref: Python tiling, autograd backward, per-row encoders; DDIM stub for diffusers, MuJoCo call "
                "replaced by a recorder, no predicted-score pass)") if self.kind == "reference" else \
               ("CPU oracle port (oracle/dgdm_oracle.py: reference algorithm with repeat-based tiling and PointNet++ "
                "hoisted to once per object) + scoring pass")
        return f"{self.n_obj} object x {self.b} candidates x {self.G} pose rows x {T_INF} steps per step; {what}"

    def designs_per_step(self):
        return self.n_obj * self.b


def timed_cpu(arm, min_s=8.0, max_reps=6):
    t0 = time.perf_counter()
    reps = 0
    while True:
        arm.step()
        reps += 1
        if time.perf_counter() - t0 > min_s or reps >= max_reps:
            break
    dt = time.perf_counter() - t0
    return arm.designs_per_step() * reps / dt, reps, dt


def run_reference(args):
    """--impl reference: the reference's own sampler code on the host cores (oracle/_ref through oracle/ref_arm.py; the
    oracle port only if those files did not travel), all host threads, each step the bounded slice of the headline
    workload described in cpu_baseline.sample."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    wl = WORKLOADS[args.workload or "c2"]
    arm = CpuArm(wl)
    for _ in range(max(0, args.warmup)):
        arm.step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        arm.step()
    dt = time.perf_counter() - t0
    val = arm.designs_per_step() * args.steps / dt
    line = {"impl": "reference", "metric": "guided designs/sec", "value": val, "unit": "designs/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": dict(workload_config(wl, args.gpus, args.precision or "fp32", wl["n_obj"] * args.gpus),
                           reference_arm="host cores only; each step is the bounded slice in cpu_baseline.sample"),
            "cpu_baseline": {"value": val, "unit": "designs/s", "cores": arm.cores, "kind": arm.kind,
                             "sample": arm.describe()},
            "e2e": {"value": val, "unit": "designs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


def workload_config(wl, n_gpus, precision, n_obj_global):
    G = wl["grid"] * wl["npos"] ** 2
    per_gpu = f"{n_obj_global} objects over {n_gpus} GPU(s)" if wl["strong"] else f"{wl['n_obj']} objects/GPU"
    return {"workload": f"{wl['label']}: {'3D' if wl['mode'] == 'point_3d' else '2D'} guided sampling, {per_gpu} x "
                        f"{wl['n_cand']} candidates x {wl['grid']} orientations x {wl['npos']}x{wl['npos']} positions = "
                        f"{G} pose rows/candidate, {T_INF} DDIM steps of {T_TRAIN}, objective {OBJECTIVE}, "
                        f"+ scoring pass + best-of-N",
            "objects_total": n_obj_global, "candidates": wl["n_cand"], "pose_rows": G, "ddim_steps": T_INF,
            "guidance_rows_per_pass": n_obj_global * wl["n_cand"] * G * T_INF, "precision": precision,
            "sharding": f"objects x{n_gpus}, no per-step traffic, 1 all-gather of scores / best ids"
                        f" (e2e: + designs) per pass",
            "l2": "256 MiB memset between timed passes (L2 flush), inside the bracket (<0.1% of a pass)"}


# ------------------------------------------------------------------------------------------------ GPU arm
class Ctx:
    pass


def run_workload(ctx, wl, precision, steps, warmup, want_clocks=False):
    """Time one configuration on this rank's GPU (all ranks call this together).  Returns the result dict on rank 0."""
    import ctypes as C
    import torch
    import torch.distributed as dist
    from dgdm_b200 import distributed as D, synthetic as syn
    from dgdm_b200.diffusion import Diffusion
    from dgdm_b200.scheduler import DDIMScheduler
    lib, dev, world, rank = ctx.lib, ctx.dev, ctx.world, ctx.rank
    wl_name = next(k for k, v in WORKLOADS.items() if v is wl)
    n_cand, P = wl["n_cand"], wl["P"]
    is3d = wl["mode"] == "point_3d"
    n_obj_global = wl["n_obj"] if wl["strong"] else wl["n_obj"] * world
    lo, hi = D.shard_range(n_obj_global, world, rank)
    objs_host = _synthetic_objects(wl, n_obj_global)[lo:hi].contiguous().pin_memory()
    fps_host = syn.fps_starts(n_obj_global)[lo:hi].contiguous() if is3d else None
    noise_host = syn.initial_noise(n_cand, P).pin_memory()
    dm = Diffusion(syn.unet1d_state_dict(0), DDIMScheduler(T_TRAIN), T_INF, mode=wl["mode"], num_points=P,
                   classifier_model=syn.dynamics3d_state_dict(0) if is3d else syn.dynamics2d_state_dict(0),
                   grid_size=wl["grid"], num_pos=wl["npos"], object_vertices=objs_host, object_ids=list(range(lo, hi)),
                   fps_starts=fps_host, precision=precision, device=dev)
    noise_dev = noise_host.to(dev)

    def one_pass_resident():
        local = dm.guided_sample(0, n_cand, noise_dev, opt_obj=OBJECTIVE)
        return D.gather_per_object_results(local, n_obj_global, gather_designs=False)

    out_host = {}

    def one_pass_e2e():
        # host -> device: this pass's inputs from pinned memory; device -> host: global scores, best ids and designs
        nz = noise_host.to(dev, non_blocking=True)
        dm.set_objects(objs_host.to(dev, non_blocking=True), fps_host)      # 3D: PointNet++ (K5) runs here, every pass
        local = dm.guided_sample(0, n_cand, nz, opt_obj=OBJECTIVE)
        g = D.gather_per_object_results(local, n_obj_global, gather_designs=True)
        for k in ("scores", "best_ids", "best_scores", "designs"):
            out_host[k] = g[k].to("cpu", non_blocking=True)
        return g

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, with_kernel_timing=False):
        for _ in range(warmup):
            fn()
            ctx.flush.zero_()
        barrier()
        if with_kernel_timing:
            lib.dgdm_trunk_timing(1)
        l0 = lib.dgdm_launch_count()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
            ctx.flush.zero_()
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        launches = lib.dgdm_launch_count() - l0
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms, launches

    clocks = ClockSampler(ctx.local_rank) if (rank == 0 and want_clocks) else None
    ms, launches = timed(one_pass_resident, with_kernel_timing=True)
    k_ms = (C.c_double * 2)(); k_n = (C.c_int64 * 2)(); k_rows = (C.c_int64 * 2)()
    lib.dgdm_trunk_timing_read_split(k_ms, k_n, k_rows)
    lib.dgdm_trunk_timing(0)
    clk = clocks.stop() if clocks else None
    ms_e2e, _ = timed(one_pass_e2e)
    del dm
    torch.cuda.empty_cache()
    if rank != 0:
        return None

    designs_per_pass = n_obj_global * n_cand
    pk = peaks()
    x3 = precision in ("fp32", "fp16x3")
    fpr = FLOP_PER_ROW_3D if is3d else FLOP_PER_ROW_2D
    g_ms, s_ms = float(k_ms[0]), float(k_ms[1])
    g_flops, s_flops = int(k_rows[0]) * fpr, int(k_rows[1]) * (fpr // 2)
    achieved = g_flops / (g_ms / 1e3) / 1e12 if g_ms > 0 else 0.0
    peak = pk["bf16_sustained"] / (3.0 if x3 else 1.0)
    traffic = None                       # DRAM bytes per launch of this kernel from the committed ncu --set full capture
    tpath = os.path.join(REPO, "profiles", "tc_trunk_traffic.json")
    if os.path.exists(tpath):                                # captured per (workload, mode): see the file's note
        key = {("c2", True): "fp32", ("c2", False): "bf16"}.get((wl_name, x3)) if wl_name == "c2" else \
            (f"c3_{precision}" if wl_name == "c3" else None)
        traffic = json.load(open(tpath)).get(key, {}).get("dram_bytes_per_launch") if key else None
    roofline = {"bound": "tensor", "achieved": achieved, "peak": peak, "unit": "TFLOP/s",
                "frac": achieved / peak if peak else None, "traffic": traffic,
                "kernel": "tc_trunk_kernel, guidance launches (fused tcgen05 trunk fwd + dgrad + per-pair reduction)",
                "peak_basis": f"bf16_tflops_sustained ({pk['bf16_sustained']}) of {pk['src']}"
                              + (" / 3: fp32-grade mode issues 3 MMAs per product" if x3 else ""),
                "launches": int(k_n[0]), "avg_launch_ms": g_ms / max(1, int(k_n[0])),
                "rows_per_launch": int(k_rows[0]) // max(1, int(k_n[0])),
                "flop_per_row": fpr, "executed_tflops": achieved * (3.0 if x3 else 1.0),
                "kernel_share_of_step": (g_ms + s_ms) / ms,
                "scoring": {"launches": int(k_n[1]), "avg_launch_ms": s_ms / max(1, int(k_n[1])),
                            "flop_per_row": fpr // 2,
                            "achieved": s_flops / (s_ms / 1e3) / 1e12 if s_ms > 0 else 0.0,
                            "note": "forward-only launches of the same kernel (G = grid_size profile rows per design)"}}
    per_rank_h2d = noise_host.numel() * 4 + objs_host.numel() * 4
    per_rank_d2h = sum(v.numel() * v.element_size() for v in out_host.values())
    res = {"value": designs_per_pass * steps / (ms / 1e3), "unit": "designs/s", "steps": steps, "warmup": warmup,
           "ms_per_step": ms / steps, "scaling": "strong" if wl["strong"] else "weak",
           "dtype": {"fp32": "bf16x3 (fp32-grade)", "fp16x3": "fp16x3 (fp32-grade)"}.get(precision, precision),
           "config": workload_config(wl, world, precision, n_obj_global),
           "denoise_steps_per_sec": designs_per_pass * steps / (ms / 1e3) * T_INF,
           "roofline": roofline,
           "e2e": {"value": designs_per_pass * steps / (ms_e2e / 1e3), "unit": "designs/s",
                   "h2d_bytes_per_step": per_rank_h2d * world, "d2h_bytes_per_step": per_rank_d2h * world,
                   "ms_per_step": ms_e2e / steps,
                   "note": "bytes are whole-job (every rank copies its inputs in and the gathered global tables out)"},
           "gpu_launches": int(launches)}
    if clk is not None:
        res["clocks"] = clk
    return res


def cpu_baseline_block(wl, with_c1):
    """Rank 0, N = 1: the reference itself (and the port) on a bounded slice of the headline workload."""
    arm = CpuArm(wl)
    arm.step()                                                    # warm-up (imports, first-touch)
    val, reps, dt = timed_cpu(arm)
    cpu = {"value": val, "unit": "designs/s", "cores": arm.cores, "kind": arm.kind,
           "sample": f"{reps} x ({arm.describe()}) in {dt:.1f} s"}
    if arm.kind == "reference":
        port = CpuArm(wl, prefer_reference=False)
        port.step()
        pv, preps, pdt = timed_cpu(port, min_s=4.0)
        cpu["port"] = {"value": pv, "unit": "designs/s", "kind": "port", "sample": f"{preps} x ({port.describe()}) in {pdt:.1f} s"}
        if with_c1:
            c1 = dict(WORKLOADS["c2"], grid=360, npos=5, label="C1 (BASELINE.json configs[0])")
            a1 = CpuArm(c1)
            a1.b = 16
            from dgdm_b200 import synthetic as syn
            a1.noise = syn.initial_noise(16, 14)
            t0 = time.perf_counter()
            a1.step()
            d1 = time.perf_counter() - t0
            cpu["c1_as_is"] = {"value": 16 / d1, "unit": "designs/s", "seconds": d1, "kind": "reference",
                               "sample": "1 pass of BASELINE.json configs[0] as is: 1 object x 16 candidates x 360 x 5x5 = "
                                         "9000 pose rows x 5 steps (generator/guided_sample_2d.sh arguments)"}
    return cpu


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--precision", default=None, choices=["fp32", "fp16x3", "bf16", "fp16", "fp32_simt"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extra", action="store_true")
    ap.add_argument("--workload", default=None, choices=sorted(WORKLOADS))
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist
    from dgdm_b200 import _lib

    ctx = Ctx()
    ctx.world = int(os.environ.get("WORLD_SIZE", "1"))
    ctx.rank = int(os.environ.get("RANK", "0"))
    ctx.local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product path has no CPU fallback")
    torch.cuda.set_device(ctx.local_rank)
    ctx.dev = torch.device("cuda", ctx.local_rank)
    if ctx.world > 1:
        dist.init_process_group("nccl", device_id=ctx.dev)
    ctx.lib = _lib.lib()
    ctx.flush = torch.empty(256 << 20, dtype=torch.uint8, device=ctx.dev)
    if args.warmup < 3:
        args.warmup = 3

    single = args.workload is not None or args.precision is not None
    wl = WORKLOADS[args.workload or "c2"]
    precision = args.precision or "fp32"
    head = run_workload(ctx, wl, precision, args.steps, args.warmup, want_clocks=True)
    extra = {}
    if not single and not args.no_extra:
        # fewer steps than the headline (same rules: >= 3 warm-up passes, L2 flush between passes, device events, max over
        # ranks); the strong-scaling sweep is 8 s per pass on one GPU, so it gets 1 warm-up + 2 timed passes there
        plan = [("c3_bf16", "c3", "bf16", min(args.steps, 5), 3), ("c3_fp16", "c3", "fp16", min(args.steps, 5), 3)]
        n_c5 = (2, 1) if ctx.world == 1 else (3, 2) if ctx.world == 2 else (min(args.steps, 5), 3)
        plan.append(("c5_strong_bf16", "c5_strong", "bf16", n_c5[0], n_c5[1]))
        for key, wname, prec, st, wu in plan:
            r = run_workload(ctx, WORKLOADS[wname], prec, st, wu)
            if r is not None:
                extra[key] = r

    if ctx.rank == 0:
        cpu = None
        if not args.no_cpu_baseline and ctx.world == 1:
            cpu = cpu_baseline_block(wl, with_c1=wl is WORKLOADS["c2"])
        line = {"metric": "guided designs/sec", "value": head["value"], "unit": "designs/s", "n_gpus": ctx.world,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": head["ms_per_step"], "higher_is_better": True,
                "scaling": head["scaling"], "vs_baseline": None, "dtype": head["dtype"], "data": "synthetic",
                "config": head["config"], "denoise_steps_per_sec": head["denoise_steps_per_sec"],
                "roofline": head["roofline"], "cpu_baseline": cpu, "e2e": head["e2e"],
                "gpu_launches": head["gpu_launches"], "clocks": head.get("clocks")}
        if extra:
            line["extra"] = extra
        print(json.dumps(line))
    if ctx.world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
