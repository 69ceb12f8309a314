"""Latency of one guided-sampling pass at the reference's stock 2D shape (BASELINE.json configs[0]: 1 object, 16
candidates, grid 360 x 5 x 5 = 9000 pose rows per candidate, 5 DDIM steps), eager launches vs one CUDA-graph replay.

    python scripts/latency_small.py [--precision fp32|bf16]
"""
import argparse
import os
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--precision", default="fp32", choices=["fp32", "bf16"])
    ap.add_argument("--iters", type=int, default=20)
    args = ap.parse_args()
    import torch
    from dgdm_b200 import synthetic as syn
    from dgdm_b200.diffusion import Diffusion
    from dgdm_b200.scheduler import DDIMScheduler

    dm = Diffusion(syn.unet1d_state_dict(0), DDIMScheduler(15), 5, mode="point", num_points=14,
                   classifier_model=syn.dynamics2d_state_dict(0), grid_size=360, num_pos=5,
                   object_vertices=syn.objects_2d(1), object_ids=[0], precision=args.precision)
    noise = syn.initial_noise(16, 14).cuda()

    def timed(graph):
        for _ in range(3):
            dm.guided_sample(0, 16, noise, opt_obj="rotate_clockwise", cuda_graph=graph)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.iters):
            dm.guided_sample(0, 16, noise, opt_obj="rotate_clockwise", cuda_graph=graph)
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / args.iters

    ms_e, ms_g = timed(False), timed(True)
    print(f"C1 shape (1 object x 16 candidates x 9000 pose rows, {args.precision}): eager {ms_e:.3f} ms/pass "
          f"({16 / ms_e * 1e3:.0f} designs/s), CUDA graph {ms_g:.3f} ms/pass ({16 / ms_g * 1e3:.0f} designs/s)")


if __name__ == "__main__":
    main()
