"""CPU-only tests of the host logic: checkpoint ingestion, BatchNorm folding / layer-1 split / conv re-layout
(dgdm_b200.pack) checked by re-evaluating the *hoisted* formulation in plain torch against the oracle, the
scheduler mirror, and that the C-ABI library loads and exports every symbol include/dgdm_b200.h declares.
No GPU compute is issued here."""
import os
import re

import numpy as np
import pytest
import torch
import torch.nn.functional as F

import dgdm_oracle as orc
from dgdm_b200 import pack, synthetic as syn

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def rel(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30)


def hoisted_logits(h, x, obj_code_in, t_frac, ori, pos):
    """The restructured computation the CUDA path performs, in torch: layer-1 = Cst + U[b] + V[g]."""
    g = F.linear(F.relu(F.linear(x, h["ge_w0"], h["ge_b0"])), h["ge_w1"], h["ge_b1"])
    U = F.linear(g, h["w1_ctrl"])
    pose = torch.cat([orc.fourier_embed(ori), orc.fourier_embed(pos)], dim=1)
    V = F.linear(pose, h["w1_pose"])
    if h["is_3d"]:
        oc = obj_code_in
        te = orc.timestep_embedding(torch.tensor([t_frac]), 256)
    else:
        oc = F.linear(F.relu(F.linear(obj_code_in, h["oe_w0"], h["oe_b0"])), h["oe_w1"], h["oe_b1"])
        te = F.linear(F.silu(F.linear(orc.timestep_embedding(torch.tensor([t_frac]), 128), h["te_w0"], h["te_b0"])),
                      h["te_w1"], h["te_b1"])
    Cst = F.linear(oc, h["w1_obj"]) + F.linear(te, h["w1_time"], h["b1"])
    a = F.relu(Cst[:, None, :] + U[:, None, :] + V[None, :, :])          # (B,G,H1)
    for i in range(7):
        a = F.relu(F.linear(a, h[f"wl{i}"], h[f"bl{i}"]))
    return F.linear(a, h["w_out"], h["b_out"])


def test_fold_dynamics_2d_matches_oracle():
    sd = syn.dynamics2d_state_dict(0)
    h = pack.fold_dynamics(sd)
    assert (h["P"], h["H1"], h["is_3d"], h["obj_dim"]) == (14, 256, 0, 200)
    B, grid, npos = 3, 4, 2
    x = syn.initial_noise(B, 14)[..., 0]
    obj = syn.objects_2d(1)[0].reshape(1, -1)
    ori, pos = orc.pose_grid(1, grid, npos, (-1.0, 1.0))
    got = hoisted_logits(h, x, obj, 9 / 15, ori, pos)                    # (B,G,3)
    G = grid * npos ** 2
    osd = orc.strip_prefix(sd)
    want = orc.dynamics2d_forward(osd, x.repeat_interleave(G, 0), ori.repeat(B, 1), pos.repeat(B, 1),
                                  torch.full((B * G,), 9 / 15), obj.expand(B * G, -1)).reshape(B, G, 3)
    assert rel(got.detach().numpy(), want.detach().numpy()) < 2e-6


def test_fold_dynamics_3d_and_pointnet():
    sd = syn.dynamics3d_state_dict(0)
    h = pack.fold_dynamics(sd)
    assert (h["P"], h["H1"], h["is_3d"], h["obj_dim"]) == (42, 512, 1, 256)
    assert tuple(h["wl0"].shape) == (256, 512) and tuple(h["wl_t0"].shape) == (512, 256)
    osd = orc.strip_prefix(sd)
    objs, st = syn.objects_3d(1), syn.fps_starts(1)
    with torch.no_grad():
        code = orc.pointnet2_encode(osd, objs.permute(0, 2, 1).contiguous(), st)
    B, grid, npos = 2, 3, 2
    x = syn.initial_noise(B, 42)[..., 0]
    ori, pos = orc.pose_grid(1, grid, npos, (-1.0, 1.0))
    G = grid * npos ** 2
    got = hoisted_logits(h, x, code, 3 / 15, ori, pos)
    pts = torch.zeros(B, 3, 42); pts[:, 1, :] = x
    want = orc.dynamics3d_forward(osd, pts.repeat_interleave(G, 0), ori.repeat(B, 1), pos.repeat(B, 1),
                                  torch.full((B * G,), 3 / 15), object_code=code.expand(B * G, -1)).reshape(B, G, 3)
    assert rel(got.detach().numpy(), want.detach().numpy()) < 2e-6
    # PointNet++ conv+BN folding on random grouped features
    pn = pack.fold_pointnet2(sd)
    xg = torch.randn(2, 131, 5, 7)
    ref = xg
    for l in range(2):
        ref = F.relu(F.batch_norm(F.conv2d(ref, osd[f"object_encoder.sa2.mlp_convs.{l}.weight"],
                                           osd[f"object_encoder.sa2.mlp_convs.{l}.bias"]),
                                  osd[f"object_encoder.sa2.mlp_bns.{l}.running_mean"],
                                  osd[f"object_encoder.sa2.mlp_bns.{l}.running_var"],
                                  osd[f"object_encoder.sa2.mlp_bns.{l}.weight"],
                                  osd[f"object_encoder.sa2.mlp_bns.{l}.bias"], training=False, eps=1e-5))
    y = xg.permute(0, 2, 3, 1)
    y = F.relu(F.linear(F.relu(F.linear(y, pn["w2"], pn["b2"])), pn["w3"], pn["b3"]))
    assert rel(y.permute(0, 3, 1, 2).numpy(), ref.numpy()) < 2e-6


def test_checkpoint_formats():
    dsd = syn.dynamics2d_state_dict(0, data_parallel_prefix=True)
    usd = syn.unet1d_state_dict(0)
    a = pack.load_dynamics_state_dict(dsd)
    b = pack.load_dynamics_state_dict(syn.dynamics2d_state_dict(0, data_parallel_prefix=False))
    assert set(a) == set(b) and all(torch.equal(a[k], b[k]) for k in a)
    ck = syn.lightning_diffusion_checkpoint(usd, dsd)
    ck["state_dict"] = {k.replace("noise_pred_net.", "noise_pred_net._orig_mod.") if k.startswith("ema_nets") else k: v
                        for k, v in ck["state_dict"].items()}            # torch.compile'd save (diffusion.py:731-743)
    u = pack.load_unet_state_dict(ck)
    assert set(u) == set(usd) and all(torch.equal(u[k], usd[k]) for k in usd)
    c = pack.load_dynamics_state_dict(ck)                                # classifier embedded in a Lightning ckpt
    assert set(c) == set(a)
    with pytest.raises(ValueError):
        bad = dict(usd); bad["diffusion_step_encoder.1.weight"] = torch.zeros(64, 16)
        pack.fold_unet(bad)


def test_fold_unet_conv_layouts():
    sd = syn.unet1d_state_dict(0)
    u = pack.fold_unet(sd)
    assert [(b["cin"], b["cout"]) for b in u["blocks"]] == [(1, 128), (128, 128), (128, 256), (256, 256), (256, 256),
                                                            (256, 256), (512, 128), (128, 128)]
    x = torch.randn(2, 128, 8)
    xp = F.pad(x.permute(0, 2, 1), (0, 0, 2, 2))                          # channels-last, 2 zero rows either side
    # k=5 conv as the implicit GEMM the kernel runs: row l = 5*C contiguous floats starting at padded row l
    w = u["blocks"][1]["conv0_w"].reshape(128, -1)
    rows = torch.stack([xp[:, l:l + 5, :].reshape(2, -1) for l in range(8)], dim=1)
    got = rows @ w.t() + u["blocks"][1]["conv0_b"]
    want = F.conv1d(x, sd["down_modules.0.1.blocks.0.block.0.weight"], sd["down_modules.0.1.blocks.0.block.0.bias"], padding=2)
    assert rel(got.permute(0, 2, 1).numpy(), want.numpy()) < 1e-5
    # stride-2 downsample: padded rows 2j+1 .. 2j+3
    w = u["down_w"].reshape(128, -1)
    rows = torch.stack([xp[:, 2 * j + 1:2 * j + 4, :].reshape(2, -1) for j in range(4)], dim=1)
    got = rows @ w.t() + u["down_b"]
    want = F.conv1d(x, sd["down_modules.0.2.conv.weight"], sd["down_modules.0.2.conv.bias"], stride=2, padding=1)
    assert rel(got.permute(0, 2, 1).numpy(), want.numpy()) < 1e-5
    # transposed conv as two 2-tap phases: even rows start at padded row m+1, odd at m+2
    ev = torch.stack([xp[:, m + 1:m + 3, :].reshape(2, -1) for m in range(8)], dim=1) @ u["up_w"][0].reshape(128, -1).t()
    od = torch.stack([xp[:, m + 2:m + 4, :].reshape(2, -1) for m in range(8)], dim=1) @ u["up_w"][1].reshape(128, -1).t()
    got = torch.stack([ev, od], dim=2).reshape(2, 16, 128) + u["up_b"]
    want = F.conv_transpose1d(x, sd["up_modules.0.2.conv.weight"], sd["up_modules.0.2.conv.bias"], stride=2, padding=1)
    assert rel(got.permute(0, 2, 1).numpy(), want.numpy()) < 1e-5


def test_abi_library_exports_every_declared_symbol():
    """Loads libdgdm_b200.so (built by __graft_entry__.build / python -m dgdm_b200.build) and checks that every
    function declared in include/dgdm_b200.h is exported and bound.  No kernel is launched."""
    from dgdm_b200 import _lib
    hdr = open(os.path.join(REPO, "include", "dgdm_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(dgdm_[a-z0-9_]+)\s*\(", hdr))
    assert declared, "no declarations parsed"
    l = _lib.lib()
    for name in declared:
        assert hasattr(l, name), f"{name} declared in the header but not exported"
    assert declared == set(_lib.EXPORTS), (declared ^ set(_lib.EXPORTS))
    assert l.dgdm_abi_version() == 1
    assert l.dgdm_launch_count() == 0
    # argument validation happens before any CUDA call
    assert l.dgdm_ddim_guided_update(None, None, None, None, 4, 0.1, 0.9, 0.9, 0.1, 1.0, 1, None) == -1
    assert b"null pointer" in l.dgdm_last_error()


def test_struct_layouts_match_header_sizes():
    import ctypes as C
    from dgdm_b200 import _lib
    assert C.sizeof(_lib.DynWeights) == 16 + 20 * 8 + 21 * 8 + 3 * 8
    assert C.sizeof(_lib.PoseGrid) == 20 and C.sizeof(_lib.Objective) == 24
    assert C.sizeof(_lib.UnetResBlock) == 8 + 12 * 8
    assert C.sizeof(_lib.UnetWeights) == 4 * 8 + 8 * C.sizeof(_lib.UnetResBlock) + 11 * 8
    assert C.sizeof(_lib.PointNet2Weights) == 80


def test_scheduler_mirror_matches_oracle():
    from dgdm_b200.scheduler import DDIMScheduler
    s = DDIMScheduler(num_train_timesteps=15)
    s.set_timesteps(5)
    assert torch.equal(s.alphas_cumprod, orc.ddim_alphas_cumprod(15))
    assert s.timesteps.tolist() == orc.ddim_timesteps(15, 5).tolist()
    a = s.alphas_cumprod
    for t in (12, 9, 6, 3, 0):
        c = s.coefficients(t)
        ap = a[t - 3] if t >= 3 else torch.tensor(1.0)
        assert c == (float((1 - a[t]).sqrt()), float(a[t] ** 0.5), float(ap ** 0.5), float((1 - ap) ** 0.5))
    with pytest.raises(ValueError):
        DDIMScheduler(beta_schedule="linear")


def test_convergence_host_helpers_match_oracle():
    from dgdm_b200.diffusion import _convergence_center, _slicer_idx
    rs = np.random.RandomState(0)
    for _ in range(200):
        n = rs.randint(3, 40)
        cls = rs.choice([0.0, 1.0, 2.0], size=n, p=[0.4, 0.2, 0.4])
        lens, centers = orc.convergence_mode_three_class(torch.from_numpy(cls))
        assert _convergence_center(cls) == int(centers[torch.argmax(lens)])
        lo = rs.randint(-n + 1, n); up = lo + rs.randint(1, n)
        if lo >= 0 and up > n and lo > n:
            continue
        assert _slicer_idx(n, lo, up).tolist() == orc.slicer(torch.arange(n), lo, up).tolist()


# ---------------------------------------------------------------------------------------------- CLI
# every flag of the reference's parser (dynamics/parser.py:5-39) with its default
_REF_FLAGS = {
    "batch_size": 1024, "use_sub_batch": False, "sub_bs": 1024, "num_epochs": 1000, "num_fingers": 1000, "ctrlpts_dim": 14,
    "ctrlpts_x_dim": 7, "ctrlpts_z_dim": 3, "learning_rate": 1e-4, "lr_warmup_steps": 100, "weight_decay": 0,
    "patience": 500, "checkpoint_path": None, "wandb_id": None, "data_dir": "", "test_data_dir": "", "object_dir": "",
    "num_workers": 4, "grid_size": 360, "num_pos": 9, "save_ckpt_step": 10, "val_step": 100,
    "num_train_timesteps": 1000, "num_timesteps_per_batch": 1, "num_inference_steps": 100, "ema_power": 0.75,
    "object_max_num_vertices": 10, "diffusion_checkpoint_path": None, "classifier_guidance": False, "num_cpus": 4,
    "fingers_3d": False, "render_video": False, "seed": 0,
}


def test_cli_accepts_the_reference_command_lines():
    """The sampler CLI parses every flag of dynamics/parser.py with the reference's defaults, and the two stock
    command lines (generator/guided_sample_2d.sh, guided_sample_3d.sh) verbatim."""
    from dgdm_b200 import cli
    a = cli.parse([])
    for k, v in _REF_FLAGS.items():
        assert getattr(a, k) == v, k
    cmd2d = ("--mode=test --checkpoint_path=ckpts/dynamics_2d.pt --classifier_guidance "
             "--diffusion_checkpoint_path=ckpts/diffusion_2d.pt --object_dir=icons/Icons-50.npy --save_dir= "
             "--ctrlpts_dim=14 --num_fingers=16 --grid_size=360 --num_pos=5 --object_max_num_vertices=100 --num_workers=0 "
             "--num_train_timesteps=15 --num_inference_steps=5 --ema_power=0.85 --batch_size=16 --num_cpus=32 --seed=0").split()
    a = cli.parse(cmd2d)
    assert (a.mode, a.classifier_guidance, a.fingers_3d, a.ctrlpts_dim, a.grid_size, a.num_pos) == ("test", True, False, 14, 360, 5)
    assert (a.num_train_timesteps, a.num_inference_steps, a.batch_size, a.object_max_num_vertices) == (15, 5, 16, 100)
    cmd3d = ("--mode=test --checkpoint_path=ckpts/dynamics_3d.pt --diffusion_checkpoint_path=ckpts/diffusion_3d.ckpt "
             "--object_dir= --save_dir= --classifier_guidance --num_fingers=16 --grid_size=45 --num_pos=5 --fingers_3d "
             "--object_max_num_vertices=512 --ctrlpts_dim=42 --ctrlpts_x_dim=7 --ctrlpts_z_dim=3 --num_workers=0 "
             "--num_train_timesteps=15 --num_inference_steps=5 --ema_power=0.85 --batch_size=16 --sub_bs=512 --num_cpus=32 "
             "--seed=0").split()
    a = cli.parse(cmd3d)
    assert (a.fingers_3d, a.ctrlpts_dim, a.ctrlpts_x_dim, a.ctrlpts_z_dim, a.grid_size, a.sub_bs) == (True, 42, 7, 3, 45, 512)
    # training is not on this path
    with pytest.raises(SystemExit):
        cli.main(["--mode=train", "--classifier_guidance"])
    with pytest.raises(SystemExit):
        cli.main(["--mode=test"])


def test_cli_normalises_objects_as_the_reference():
    """``--objects_in_metres``: the min-max to [-1, 1] of generator/train.py:107-109 (3D) and :122-123 (2D), same
    expression and fp32 operation order (restated here independently from the reference's lines)."""
    from dgdm_b200 import cli
    rng = np.random.RandomState(5)
    v3 = torch.from_numpy(rng.uniform([-0.1, -0.1, 0.0], [0.1, 0.1, 0.12], size=(3, 512, 3)).astype(np.float32))
    want = v3.clone()
    want[..., 0] = (want[..., 0] - (-0.1)) / (0.1 - (-0.1)) * 2.0 - 1.0
    want[..., 1] = (want[..., 1] - (-0.1)) / (0.1 - (-0.1)) * 2.0 - 1.0
    want[..., 2] = (want[..., 2] - 0.0) / (0.12 - 0.0) * 2.0 - 1.0
    got = cli.normalise_objects(v3, True)
    assert torch.equal(got, want) and float(got.abs().max()) <= 1.0 + 1e-6
    v2 = torch.from_numpy(rng.uniform(-0.05, 0.05, size=(4, 100, 2)).astype(np.float32))
    want2 = v2.clone()
    want2[..., 0] = (want2[..., 0] - (-0.05)) / (0.05 - (-0.05)) * 2.0 - 1.0
    want2[..., 1] = (want2[..., 1] - (-0.05)) / (0.05 - (-0.05)) * 2.0 - 1.0
    assert torch.equal(cli.normalise_objects(v2, False), want2)
    assert torch.equal(v2, torch.from_numpy(np.asarray(v2)))        # input untouched


def test_torch_library_ops_are_registered_with_fake_kernels():
    """the C ABI is also exposed as torch.library custom ops.  On CPU: the schemas exist, the fake (meta)
    kernels propagate shapes, and a call with CPU tensors fails loudly (no CPU implementation)."""
    import dgdm_b200.ops  # noqa: F401
    from torch._subclasses.fake_tensor import FakeTensorMode
    ns = torch.ops.dgdm_b200
    for name in ("unet1d_forward", "dyn_guidance", "ddim_guided_update", "best_of_n"):
        assert hasattr(ns, name), name
    with FakeTensorMode():
        x = torch.empty(6, 14)
        assert ns.unet1d_forward(x, 0, 9, 1).shape == (6, 14)
        assert ns.dyn_guidance(x, torch.empty(2, 200), 0, 0.4, 36, 5, -1.0, 1.0, -1.0, 0.0, 0.0, 0.0, 1.0, 1).shape == (6, 14)
        assert ns.ddim_guided_update(x, x, x, 0.9, 0.4, 0.5, 0.8, 1e-3, True).shape == (6, 14)
        idx, best = ns.best_of_n(torch.empty(3, 16), 4)
        assert idx.shape == (3, 4) and idx.dtype == torch.int64 and best.shape == (3, 4)
    with pytest.raises(RuntimeError):
        ns.ddim_guided_update(torch.zeros(2, 14), torch.zeros(2, 14), torch.zeros(2, 14), 0.9, 0.4, 0.5, 0.8, 1e-3, True)


def test_bench_reference_arm_prints_the_contract_line():
    """``bench.py --impl reference`` (CPU only): one JSON line with the contract keys, the reference's own sampler as the
    thing timed when oracle/_ref is present (kind "reference"), else the oracle port (kind "port")."""
    import json
    import subprocess
    import sys
    repo = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, os.path.join(repo, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                       capture_output=True, text=True, timeout=600, env=dict(os.environ, RANK="0", WORLD_SIZE="1"))
    assert r.returncode == 0, r.stderr[-2000:]
    line = json.loads([l for l in r.stdout.splitlines() if l.startswith("{")][-1])
    for k in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
              "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert k in line, k
    assert line["impl"] == "reference" and line["value"] > 0 and line["e2e"]["h2d_bytes_per_step"] == 0
    have_ref = os.path.exists(os.path.join(repo, "oracle", "_ref", "reference", "generator", "diffusion.refbin"))
    assert line["cpu_baseline"]["kind"] == ("reference" if have_ref else "port")
    # other ranks of a torchrun launch exit 0 without work
    r = subprocess.run([sys.executable, os.path.join(repo, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                       capture_output=True, text=True, timeout=120, env=dict(os.environ, RANK="1", WORLD_SIZE="2"))
    assert r.returncode == 0 and r.stdout.strip() == ""
