"""``guided_sample_2d`` / ``guided_sample_3d`` entry points with the reference's sampler flags.

Mirrors ``python generator/train.py --mode=test --classifier_guidance ...`` (generator/guided_sample_2d.sh:1-4,
guided_sample_3d.sh:1-4; flags from dynamics/parser.py:5-39) for the part of that command that is on the hot
path: build the schedule, load both checkpoints in the reference formats, normalise objects, run guided
sampling for the requested objectives and report predicted scores / best designs.  The Lightning trainer,
wandb tables and the MuJoCo evaluation the reference runs afterwards are out of scope.

    python -m dgdm_b200.cli --mode=test --classifier_guidance --ctrlpts_dim=14 --grid_size=360 --num_pos=5 \
        --object_max_num_vertices=100 --num_train_timesteps=15 --num_inference_steps=5 --batch_size=16 --seed=0 \
        [--checkpoint_path dynamics_2d.pt --diffusion_checkpoint_path diffusion_2d.pt --object_dir objects.npy]

Without checkpoint paths, seeded synthetic weights are used (dgdm_b200.synthetic).  ``--object_dir`` takes a
``.npy``/``.npz`` of already-extracted object vertices ``(n_obj, V, 2)`` / ``(n_obj, 512, 3)`` (contour extraction and
mesh sampling need cv2/open3d assets code that is out of scope); default: synthetic objects.  Vertices in the
simulator's metres (what ``extract_contours`` / ``sample_pts_from_mesh`` return) are min-max normalised to [-1, 1]
exactly as generator/train.py:93-99,107-109 (3D) and :111-114,122-123 (2D) do when ``--objects_in_metres`` is given;
without it the array is taken as already normalised.
"""
from __future__ import annotations

import argparse
import json
import sys

import numpy as np
import torch


# workspace bounds the reference normalises object vertices with (generator/train.py:93-99 3D, :111-114 2D), metres
OBJECT_BOUNDS_3D = ((-0.1, 0.1), (-0.1, 0.1), (0.0, 0.12))
OBJECT_BOUNDS_2D = ((-0.05, 0.05), (-0.05, 0.05))


def normalise_objects(vertices: torch.Tensor, fingers_3d: bool) -> torch.Tensor:
    """Min-max the object vertices from metres to [-1, 1] per axis: ``(v - min) / (max - min) * 2.0 - 1.0`` in fp32,
    the expression and operation order of generator/train.py:107-109 (3D) / :122-123 (2D)."""
    out = vertices.clone().float()
    for ax, (lo, hi) in enumerate(OBJECT_BOUNDS_3D if fingers_3d else OBJECT_BOUNDS_2D):
        out[..., ax] = (out[..., ax] - lo) / (hi - lo) * 2.0 - 1.0
    return out


def parse(argv=None):
    p = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    # the reference's flags that matter for sampling (dynamics/parser.py), same names and defaults
    p.add_argument("--batch_size", type=int, default=1024)
    p.add_argument("--sub_bs", type=int, default=1024, help="accepted for compatibility; rows never hit HBM here")
    p.add_argument("--num_fingers", type=int, default=1000)
    p.add_argument("--ctrlpts_dim", type=int, default=14)
    p.add_argument("--ctrlpts_x_dim", type=int, default=7)
    p.add_argument("--ctrlpts_z_dim", type=int, default=3)
    p.add_argument("--checkpoint_path", type=str, default=None)
    p.add_argument("--diffusion_checkpoint_path", type=str, default=None)
    p.add_argument("--save_dir", type=str, default="")
    p.add_argument("--object_dir", type=str, default="")
    p.add_argument("--mode", type=str, default="test")
    p.add_argument("--grid_size", type=int, default=360)
    p.add_argument("--num_pos", type=int, default=9)
    p.add_argument("--num_train_timesteps", type=int, default=1000)
    p.add_argument("--num_inference_steps", type=int, default=100)
    p.add_argument("--object_max_num_vertices", type=int, default=10)
    p.add_argument("--classifier_guidance", action="store_true")
    p.add_argument("--num_cpus", type=int, default=4)
    p.add_argument("--fingers_3d", action="store_true")
    p.add_argument("--render_video", action="store_true")
    p.add_argument("--seed", type=int, default=0)
    p.add_argument("--num_workers", type=int, default=4)
    p.add_argument("--ema_power", type=float, default=0.75)
    # the reference's remaining flags (training / logging side of dynamics/parser.py): accepted so that any command line
    # written for generator/train.py parses here too; they have no effect on sampling
    p.add_argument("--use_sub_batch", action="store_true")
    p.add_argument("--num_epochs", type=int, default=1000)
    p.add_argument("--learning_rate", type=float, default=1e-4)
    p.add_argument("--lr_warmup_steps", type=int, default=100)
    p.add_argument("--weight_decay", type=float, default=0)
    p.add_argument("--patience", type=int, default=500)
    p.add_argument("--wandb_id", type=str, default=None)
    p.add_argument("--data_dir", type=str, default="")
    p.add_argument("--test_data_dir", type=str, default="")
    p.add_argument("--save_ckpt_step", type=int, default=10)
    p.add_argument("--val_step", type=int, default=100)
    p.add_argument("--num_timesteps_per_batch", type=int, default=1)
    # additions
    p.add_argument("--objectives", type=str, default="rotate_clockwise", help="comma-separated opt_obj names")
    p.add_argument("--num_objects", type=int, default=8, help="synthetic objects when --object_dir is not given")
    p.add_argument("--precision", type=str, default="fp32", choices=["fp32", "fp16x3", "fp16", "bf16", "fp32_simt"])
    p.add_argument("--objects_in_metres", action="store_true",
                   help="--object_dir holds raw vertices in metres: normalise them as generator/train.py does")
    p.add_argument("--multi_object", action="store_true", help="also run guided_sample_multi_object")
    p.add_argument("--device", type=str, default="cuda:0")
    p.add_argument("--top_k", type=int, default=1)
    return p.parse_args(argv)


def main(argv=None) -> int:
    args = parse(argv)
    if args.mode != "test":
        raise SystemExit("only --mode=test (guided sampling) is on the B200 path; training is out of scope")
    if not args.classifier_guidance:
        raise SystemExit("--classifier_guidance is required (this is the guided sampler)")
    from . import synthetic as syn
    from .diffusion import Diffusion
    from .scheduler import DDIMScheduler

    mode = "point_3d" if args.fingers_3d else "point"
    P = args.ctrlpts_dim
    unet = args.diffusion_checkpoint_path or syn.unet1d_state_dict(args.seed)
    if args.checkpoint_path:
        dyn = args.checkpoint_path
    else:
        dyn = syn.dynamics3d_state_dict(args.seed, params_ch=P) if args.fingers_3d else \
            syn.dynamics2d_state_dict(args.seed, params_ch=P, object_ch=2 * args.object_max_num_vertices)
    fps = None
    if args.object_dir:
        arr = np.load(args.object_dir, allow_pickle=True)
        arr = arr[arr.files[0]] if hasattr(arr, "files") else arr
        objs = torch.from_numpy(np.asarray(arr, dtype=np.float32))
        if args.objects_in_metres:
            objs = normalise_objects(objs, args.fingers_3d)
    else:
        objs = syn.objects_3d(args.num_objects, args.object_max_num_vertices) if args.fingers_3d else \
            syn.objects_2d(args.num_objects, args.object_max_num_vertices)
    if args.fingers_3d:
        fps = syn.fps_starts(objs.shape[0], objs.shape[1])
    sched = DDIMScheduler(num_train_timesteps=args.num_train_timesteps)
    dm = Diffusion(unet, sched, args.num_inference_steps, mode=mode, num_points=P, class_cond=True,
                   classifier_model=dyn, grid_size=args.grid_size, num_pos=args.num_pos, object_vertices=objs,
                   object_ids=list(range(objs.shape[0])), sub_batch_size=args.sub_bs, pts_x_dim=args.ctrlpts_x_dim,
                   pts_z_dim=args.ctrlpts_z_dim, seed=args.seed, fps_starts=fps, precision=args.precision,
                   device=args.device)
    noise = syn.initial_noise(args.batch_size, P, args.seed)        # generator/diffusion.py:182-183
    report = {}
    for name in args.objectives.split(","):
        kw = {}
        if name == "convergence":
            kw["unguided_sample"] = dm.unguided_sample(noise)
        out = dm.guided_sample(0, args.batch_size, noise, args.save_dir, opt_obj=name, top_k=args.top_k, **kw)
        report[name] = {"best_ids": out["best_ids"].cpu().tolist(), "best_scores": out["best_scores"].cpu().tolist()}
        if args.save_dir:
            np.savez(f"{args.save_dir}/guided_{name}.npz", designs=out["designs"].cpu().numpy(),
                     scores=out["scores"].cpu().numpy(), best_ids=out["best_ids"].cpu().numpy())
        if args.multi_object and name != "convergence":
            m = dm.guided_sample_multi_object(0, args.batch_size, noise, args.save_dir, opt_obj=name, top_k=args.top_k)
            report[name + "/allobj"] = {"best_ids": m["best_ids"].cpu().tolist(),
                                        "best_scores": m["best_scores"].cpu().tolist()}
    json.dump(report, sys.stdout)
    print()
    return 0


def guided_sample_2d(argv=None) -> int:
    """generator/guided_sample_2d.sh with its stock arguments as defaults."""
    base = ["--mode=test", "--classifier_guidance", "--ctrlpts_dim=14", "--num_fingers=16", "--grid_size=360",
            "--num_pos=5", "--object_max_num_vertices=100", "--num_train_timesteps=15", "--num_inference_steps=5",
            "--batch_size=16", "--seed=0"]
    return main(base + list(argv or sys.argv[1:]))


def guided_sample_3d(argv=None) -> int:
    """generator/guided_sample_3d.sh with its stock arguments as defaults."""
    base = ["--mode=test", "--classifier_guidance", "--fingers_3d", "--ctrlpts_dim=42", "--ctrlpts_x_dim=7",
            "--ctrlpts_z_dim=3", "--num_fingers=16", "--grid_size=45", "--num_pos=5", "--object_max_num_vertices=512",
            "--num_train_timesteps=15", "--num_inference_steps=5", "--batch_size=16", "--sub_bs=512", "--seed=0"]
    return main(base + list(argv or sys.argv[1:]))


if __name__ == "__main__":
    raise SystemExit(main())
