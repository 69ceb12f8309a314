#!/usr/bin/env python
"""Hardware check that an N-rank sharded guided-sampling run returns the 1-rank answer (VERDICT r01 item 4/5;
reference contract: objects are independent, generator/diffusion.py:561-576).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \\
        scripts/check_sharding.py [--objects 16] [--candidates 128] [--precision fp32]

Every rank runs its shard of a 3D object set through ``distributed.sharded_guided_sample`` (objects sharded; with
``--objects`` < N, candidates sharded), rank 0 also runs the whole set alone, and the two are compared: object shards
must match BIT FOR BIT (scores, best ids, designs).  Candidate shards change the 128-row tile alignment of every pair,
i.e. the fp32 summation order of its pose rows (~1e-6 per step): scores / designs must agree to 1e-4 in 2D (guidance
scale 1e-3; measured 8e-7 / 1.2e-5 at 8 ranks) and to 2e-3 / 2e-1 in 3D, where the five-step loop (SCALE_3D = 0.5,
random-init weights) amplifies any reassociation through ReLU sign flips (measured 8e-4 / 1e-1 at 8 ranks; the CPU
oracle's own response to a 2.5e-4 per-step perturbation is 2e-3 / 2.5e-2, tests/test_gpu_baseline_shape.py), and the
selected design must be within that score tolerance of the 1-rank optimum (near-tied candidates may swap).
Prints one JSON line on rank 0; exit code 1 on a mismatch."""
import argparse
import json
import os
import sys

import torch
import torch.distributed as dist

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
from dgdm_b200 import distributed as D, synthetic as syn  # noqa: E402
from dgdm_b200.diffusion import Diffusion  # noqa: E402
from dgdm_b200.scheduler import DDIMScheduler  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--objects", type=int, default=16)
    ap.add_argument("--candidates", type=int, default=128)
    ap.add_argument("--grid", type=int, default=45)
    ap.add_argument("--num-pos", type=int, default=5)
    ap.add_argument("--precision", default="fp32")
    ap.add_argument("--mode", default="point_3d", choices=["point", "point_3d"])
    a = ap.parse_args()
    world, rank, lr = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(lr)
    dev = torch.device("cuda", lr)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    is3d = a.mode == "point_3d"
    P = 42 if is3d else 14
    objs = syn.objects_3d(a.objects) if is3d else syn.objects_2d(a.objects)
    fps = syn.fps_starts(a.objects) if is3d else None
    noise = syn.initial_noise(a.candidates, P)

    def build(o, f, ids):
        return Diffusion(syn.unet1d_state_dict(0), DDIMScheduler(15), 5, mode=a.mode, num_points=P,
                         classifier_model=syn.dynamics3d_state_dict(0) if is3d else syn.dynamics2d_state_dict(0),
                         grid_size=a.grid, num_pos=a.num_pos, object_vertices=o, object_ids=ids, fps_starts=f,
                         precision=a.precision, device=dev)

    plan = D.plan_per_object(a.objects, world)
    got = D.sharded_guided_sample(build, objs, fps, noise, "rotate_clockwise", top_k=4)
    ok, report = True, {"world": world, "plan": plan, "objects": a.objects, "candidates": a.candidates, "precision": a.precision}
    if rank == 0:
        want = build(objs, fps, list(range(a.objects))).guided_sample(0, a.candidates, noise, opt_obj="rotate_clockwise", top_k=4)
        for k in ("scores", "designs", "best_ids", "best_scores"):
            eq = torch.equal(got[k], want[k])
            err = float((got[k].double() - want[k].double()).abs().max())
            report[k] = {"bit_exact": eq, "max_abs_diff": err}
            if plan == "objects":
                ok = ok and eq
            elif k == "best_ids":
                # candidate shards: the selected design must be (one of) the best under the 1-rank scores, to within the
                # score tolerance -- with near-tied candidates a reassociation-level difference may pick the other one
                tol_s = 1e-4 if not is3d else 2e-3
                picked = want["scores"].gather(1, got["best_ids"][:, :1].long())[:, 0]
                regret = float((want["best_scores"][:, 0] - picked).max())
                report[k]["regret_of_top1"] = regret
                ok = ok and regret <= tol_s
            else:
                tol = 1e-4 if not is3d else (2e-1 if k == "designs" else 2e-3)
                ok = ok and err <= tol * max(1.0, float(want[k].abs().max()))
        report["ok"] = ok
        print(json.dumps(report))
    if world > 1:
        flag = torch.tensor([1 if ok else 0], device=dev)
        dist.broadcast(flag, 0)
        ok = bool(flag.item())
        dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
