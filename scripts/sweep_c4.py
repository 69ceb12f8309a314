#!/usr/bin/env python
"""BASELINE.json configs[3] ("C4"): dynamics-network forward and forward+input-gradient throughput on explicit rows
(the ProfileForward{2,3}DModel.forward signature: a distinct finger, pose, time and object per row, so nothing is
hoisted), N = 1k .. 64k rows, both networks, fp32-grade, fp16 and bf16.  Prints one JSON line per point.
3D rows take pre-computed PointNet++ codes (one K5 call per distinct object), as the sampler does.

    python scripts/sweep_c4.py [--cpu]      # --cpu also times the oracle port on the host cores at N = 4096
"""
import argparse
import json
import os
import sys
import time

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
sys.path.insert(0, os.path.join(REPO, "oracle"))

import numpy as np
import torch

from dgdm_b200 import synthetic as syn
from dgdm_b200.diffusion import Diffusion
from dgdm_b200.scheduler import DDIMScheduler

# as-written FLOPs per row of the paired mode: 2D fwd 1.894 MFLOP, minimal dgrad 1.189 MFLOP
FWD_2D, BWD_2D = 1.894e6, 1.189e6


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--cpu", action="store_true")
    args = ap.parse_args()
    dev = "cuda:0"
    for mode, P in (("point", 14), ("point_3d", 42)):
        is3d = mode == "point_3d"
        n_obj = 16
        objs = syn.objects_3d(n_obj) if is3d else syn.objects_2d(n_obj)
        for precision in ("fp32", "fp16", "bf16"):
            dm = Diffusion(syn.unet1d_state_dict(0), DDIMScheduler(15), 5, mode=mode, num_points=P,
                           classifier_model=syn.dynamics3d_state_dict(0) if is3d else syn.dynamics2d_state_dict(0),
                           grid_size=4, num_pos=1, object_vertices=objs, object_ids=list(range(n_obj)),
                           fps_starts=syn.fps_starts(n_obj) if is3d else None, precision=precision, device=dev)
            for n in (1024, 2048, 4096, 8192, 16384, 32768, 65536):
                rs = np.random.RandomState(n)
                x = torch.from_numpy(rs.randn(n, P).astype(np.float32)).to(dev)
                ori = torch.from_numpy(rs.uniform(-1, 1, (n, 1)).astype(np.float32)).to(dev)
                pos = torch.from_numpy(rs.uniform(-1, 1, (n, 2)).astype(np.float32)).to(dev)
                t = torch.from_numpy((rs.randint(0, 15, n) / 15.0).astype(np.float32)).to(dev)
                pick = torch.from_numpy(rs.randint(0, n_obj, n)).to(dev)
                kw = dict(object_codes=dm._obj_dev[pick].contiguous()) if is3d else \
                    dict(object_vertices=dm._obj_dev[pick].contiguous())
                for grad in (False, True):
                    call = lambda: dm.classifier_model(x, ori, pos, t, opt_obj="rotate", return_grad=grad, **kw)
                    for _ in range(3):
                        call()
                    torch.cuda.synchronize()
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    reps = 10
                    e0.record()
                    for _ in range(reps):
                        call()
                    e1.record()
                    torch.cuda.synchronize()
                    ms = e0.elapsed_time(e1) / reps
                    line = {"net": "3d" if is3d else "2d", "precision": precision, "rows": n,
                            "pass": "fwd+dgrad" if grad else "fwd", "ms": round(ms, 4), "rows_per_s": round(n / ms * 1e3)}
                    if not is3d:
                        line["as_written_tflops"] = round(n * (FWD_2D + (BWD_2D if grad else 0)) / ms / 1e9, 2)
                    print(json.dumps(line), flush=True)
    if args.cpu:
        import dgdm_oracle as orc
        torch.set_num_threads(os.cpu_count() or 1)
        n = 4096
        rs = np.random.RandomState(n)
        sd = orc.strip_prefix(syn.dynamics2d_state_dict(0))
        x = torch.from_numpy(rs.randn(n, 14).astype(np.float32)).requires_grad_(True)
        ori = torch.from_numpy(rs.uniform(-1, 1, (n, 1)).astype(np.float32))
        pos = torch.from_numpy(rs.uniform(-1, 1, (n, 2)).astype(np.float32))
        t = torch.from_numpy((rs.randint(0, 15, n) / 15.0).astype(np.float32))
        ov = syn.objects_2d(16)[rs.randint(0, 16, n)].reshape(n, -1)
        for grad in (False, True):
            t0 = time.perf_counter()
            reps = 5
            for _ in range(reps):
                lg = orc.dynamics2d_forward(sd, x, ori, pos, t, ov)
                if grad:
                    torch.autograd.grad(orc.deltas_to_objective(lg, "rotate").sum(), x)
            dt = (time.perf_counter() - t0) / reps
            print(json.dumps({"net": "2d", "precision": "cpu-fp32 oracle", "cores": torch.get_num_threads(), "rows": n,
                              "pass": "fwd+dgrad" if grad else "fwd", "ms": round(dt * 1e3, 2), "rows_per_s": round(n / dt)}))


if __name__ == "__main__":
    main()
