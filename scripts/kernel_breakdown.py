"""Per-kernel device-time breakdown of one guided-sampling pass (CUPTI via torch.profiler: no replay, no serialisation).

    python scripts/kernel_breakdown.py --workload c3 --precision bf16 > profiles/breakdown_r01_c3_bf16.txt

Times are warm (one untimed pass first). Used to decide what to optimise after the trunk kernel.
"""
import argparse
import collections
import os
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="c2", choices=["c2", "c2g36", "c3", "c5"])
    ap.add_argument("--precision", default="fp32", choices=["fp32", "fp16x3", "bf16", "fp16", "fp32_simt"])
    args = ap.parse_args()

    import torch
    import bench
    from dgdm_b200 import synthetic as syn
    from dgdm_b200.diffusion import Diffusion
    from dgdm_b200.scheduler import DDIMScheduler

    wl = bench.WORKLOADS[args.workload]
    n_obj, n_cand, P = wl["n_obj"], wl["n_cand"], wl["P"]
    dev = torch.device("cuda", 0)
    is3d = wl["mode"] == "point_3d"
    objs = (syn.objects_3d(n_obj) if is3d else syn.objects_2d(n_obj)).contiguous()
    fps = syn.fps_starts(n_obj).contiguous() if is3d else None
    dm = Diffusion(syn.unet1d_state_dict(0), DDIMScheduler(bench.T_TRAIN), bench.T_INF, mode=wl["mode"], num_points=P,
                   classifier_model=syn.dynamics3d_state_dict(0) if is3d else syn.dynamics2d_state_dict(0),
                   grid_size=wl["grid"], num_pos=wl["npos"], object_vertices=objs, object_ids=list(range(n_obj)),
                   fps_starts=fps, precision=args.precision, device=dev)
    noise = syn.initial_noise(n_cand, P).to(dev)
    dm.guided_sample(0, n_cand, noise, opt_obj=bench.OBJECTIVE)
    torch.cuda.synchronize()
    from torch.profiler import profile, ProfilerActivity
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        dm.guided_sample(0, n_cand, noise, opt_obj=bench.OBJECTIVE)
        torch.cuda.synchronize()
    tot = collections.defaultdict(lambda: [0.0, 0])
    for ev in prof.events():
        if ev.device_type == torch.autograd.DeviceType.CUDA:
            name = ev.name.replace("(anonymous namespace)::", "").replace("void ", "").split("(")[0]
            tot[name][0] += ev.device_time if hasattr(ev, "device_time") else ev.cuda_time
            tot[name][1] += 1
    total = sum(v[0] for v in tot.values())
    print(f"# workload {args.workload} precision {args.precision}: one warm pass, device time {total / 1e3:.2f} ms over "
          f"{sum(v[1] for v in tot.values())} launches")
    for name, (us, n) in sorted(tot.items(), key=lambda kv: -kv[1][0]):
        print(f"{us / 1e3:10.3f} ms {100 * us / total:6.2f} % {n:6d} x  {name[:110]}")


if __name__ == "__main__":
    main()
