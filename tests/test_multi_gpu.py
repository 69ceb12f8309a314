"""Multi-GPU hardware test: an N-rank sharded run equals the 1-rank run (scripts/check_sharding.py under torchrun).
Needs >= 2 GPUs on the box (``gpurun --gpus 2 -- python -m pytest tests/test_multi_gpu.py -m gpu``); skipped otherwise."""
import json
import os
import socket
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _run(world, *args):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr",
           "127.0.0.1", "--master-port", str(_free_port()), os.path.join(REPO, "scripts", "check_sharding.py"), *args]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert lines, r.stdout[-2000:] + r.stderr[-2000:]
    rep = json.loads(lines[-1])
    print(rep)
    assert r.returncode == 0 and rep["ok"], rep
    return rep


def _world():
    n = torch.cuda.device_count() if torch.cuda.is_available() else 0
    return 8 if n >= 8 else 4 if n >= 4 else 2 if n >= 2 else 0


@pytest.mark.parametrize("precision", ["fp32", "fp16"])
def test_object_shards_reproduce_one_rank_bit_for_bit(precision):
    w = _world()
    if w < 2:
        pytest.skip("needs >= 2 GPUs")
    rep = _run(w, "--objects", "16", "--candidates", "128", "--precision", precision)
    assert rep["plan"] == "objects"
    assert all(rep[k]["bit_exact"] for k in ("scores", "designs", "best_ids", "best_scores"))


@pytest.mark.parametrize("mode", ["point_3d", "point"])
def test_candidate_shards_when_fewer_objects_than_ranks(mode):
    w = _world()
    if w < 2:
        pytest.skip("needs >= 2 GPUs")
    # the stock 3D set is 5 objects (assets/object_names_test.txt); with 1 object even 2 ranks shard candidates
    rep = _run(w, "--objects", "1" if w == 2 else "5", "--candidates", "64", "--precision", "fp32", "--mode", mode,
               *(["--grid", "36"] if mode == "point" else []))
    assert rep["plan"] == "candidates"


def test_diffusion_on_a_non_current_device():
    """every launch must run in the context and on the stream of the device that owns the tensors.  Build the
    sampler on cuda:1 while cuda:0 is the process-wide current device and compare with the same run on cuda:0."""
    if not torch.cuda.is_available() or torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs")
    sys.path.insert(0, REPO)
    from dgdm_b200 import synthetic as syn
    from dgdm_b200.diffusion import Diffusion
    from dgdm_b200.scheduler import DDIMScheduler
    torch.cuda.set_device(0)
    objs, noise = syn.objects_2d(3), syn.initial_noise(8, 14)
    outs = []
    for dev in ("cuda:0", "cuda:1"):
        dm = Diffusion(syn.unet1d_state_dict(0), DDIMScheduler(15), 5, mode="point", num_points=14,
                       classifier_model=syn.dynamics2d_state_dict(0), grid_size=12, num_pos=2, object_vertices=objs,
                       object_ids=[0, 1, 2], precision="fp32", device=dev)
        assert torch.cuda.current_device() == 0
        out = dm.guided_sample(0, 8, noise, opt_obj="rotate_clockwise", top_k=2)
        assert out["designs"].device == torch.device(dev)
        data = dm.unguided_sample(noise)                        # scheduler.guided_step on `dev`
        metrics = dm.denoise_from_data(data, seed=1)            # the add_noise launch on `dev`
        res = {k: v.cpu() for k, v in out.items()}
        res["unguided"] = data.cpu()
        res["metrics"] = torch.tensor([metrics[k] for k in sorted(metrics)])
        outs.append(res)
        torch.cuda.synchronize(dev)
    for k in outs[0]:
        assert torch.equal(outs[0][k], outs[1][k]), k
