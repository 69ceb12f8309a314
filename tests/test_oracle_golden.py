"""The oracle (oracle/dgdm_oracle.py) against fixtures produced by the REAL reference code
(tests/golden/make_golden.py).  CPU only.  Tolerances: the oracle and the reference use the same
torch-CPU kernels on the same numbers, so agreement is at fp32 round-off (different tiling shapes
can change GEMM blocking, hence not exactly 0)."""
import os

import numpy as np
import pytest
import torch

import dgdm_oracle as orc
from dgdm_b200 import synthetic as syn


def rel(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30)


@pytest.fixture(scope="module")
def g2(golden_dir):
    return dict(np.load(os.path.join(golden_dir, "golden_2d.npz")))


@pytest.fixture(scope="module")
def g3(golden_dir):
    return dict(np.load(os.path.join(golden_dir, "golden_3d.npz")))


@pytest.fixture(scope="module")
def samp2(g2):
    return orc.OracleSampler("point", syn.unet1d_state_dict(0), syn.dynamics2d_state_dict(0),
                             torch.from_numpy(g2["objects"]), int(g2["grid_size"]), int(g2["num_pos"]))


@pytest.fixture(scope="module")
def samp3(g3):
    return orc.OracleSampler("point_3d", syn.unet1d_state_dict(0), syn.dynamics3d_state_dict(0),
                             torch.from_numpy(g3["objects"]), int(g3["grid_size"]), int(g3["num_pos"]),
                             sub_batch_size=int(g3["sub_bs"]), fps_start=torch.from_numpy(g3["fps_starts"]))


def test_synthetic_inputs_reproducible(g2, g3):
    assert np.array_equal(syn.initial_noise(4, 14).numpy(), g2["noise"])
    assert np.array_equal(syn.objects_2d(2).numpy(), g2["objects"])
    assert np.array_equal(syn.objects_3d(2).numpy(), g3["objects"])
    assert np.array_equal(syn.fps_starts(2).numpy(), g3["fps_starts"])


def test_ddim_schedule_closed_form(g2):
    a = orc.ddim_alphas_cumprod(15)
    assert np.array_equal(a.numpy(), g2["alphas_cumprod"])
    # SURVEY.md §8c table, derived from the formula
    want = [0.986676, 0.952420, 0.898706, 0.827844, 0.742884, 0.647478, 0.545732, 0.442022, 0.340810,
            0.246448, 0.162997, 0.094046, 0.042560, 0.010756, 0.000011]
    assert np.allclose(a.numpy(), want, atol=2e-6)
    assert orc.ddim_timesteps(15, 5).tolist() == [12, 9, 6, 3, 0] == g2["timesteps"].tolist()
    # closed form of one eta=0 step incl. clipping and the alpha_prev := 1 boundary
    x = torch.tensor([[0.3], [-2.0]]); e = torch.tensor([[0.5], [0.25]])
    for t in (12, 0):
        got = orc.ddim_step(e, t, x, a, 15, 5)
        at = float(a[t]); ap = float(a[t - 3]) if t >= 3 else 1.0
        x0 = np.clip((x.numpy() - np.sqrt(1 - at) * e.numpy()) / np.sqrt(at), -1, 1)
        assert np.allclose(got.numpy(), np.sqrt(ap) * x0 + np.sqrt(1 - ap) * e.numpy(), atol=1e-6)


def test_dynamics2d_forward(g2):
    sd = orc.strip_prefix(syn.dynamics2d_state_dict(0))
    with torch.no_grad():
        out = orc.dynamics2d_forward(sd, *(torch.from_numpy(g2[k]) for k in
                                           ("fwd_ctrl", "fwd_ori", "fwd_pos", "fwd_t", "fwd_obj")))
    assert rel(out.numpy(), g2["fwd_logits"]) < 1e-6


def test_unet(g2, g3):
    sd = syn.unet1d_state_dict(0)
    for g, P in ((g2, 14), (g3, 42)):
        x = torch.from_numpy(g["noise"])
        for t in (12, 0):
            with torch.no_grad():
                out = orc.unet1d_forward(sd, x, torch.full((x.shape[0],), t, dtype=torch.int64))
            assert out.shape == (x.shape[0], P, 1)
            assert rel(out.numpy(), g[f"unet_t{t}"]) < 1e-5


def test_lightning_checkpoint_roundtrip():
    sd = syn.unet1d_state_dict(0)
    ck = syn.lightning_diffusion_checkpoint(sd, syn.dynamics2d_state_dict(0))
    back = orc.unet_from_lightning(ck)
    assert set(back) == set(sd) and all(torch.equal(back[k], sd[k]) for k in sd)


@pytest.mark.parametrize("name", ["rotate", "rotate_clockwise", "clockwise_up", "shift_left", "counterclockwise_right"])
def test_cond_fn_2d(g2, samp2, name):
    x = torch.from_numpy(g2["noise"])
    for oi in range(2):
        for t in (12, 3):
            got = samp2.cond_fn(x, t, name, oi)
            assert rel(got.numpy(), g2[f"grad_o{oi}_t{t}_{name}"]) < 2e-5


def test_cond_fn_2d_ori_range(g2, samp2):
    x = torch.from_numpy(g2["noise"])
    for oi in range(2):
        got = samp2.cond_fn(x, 6, "rotate", oi, ori_range=(-0.5, 0.25))
        assert rel(got.numpy(), g2[f"grad_o{oi}_t6_rotate_narrow"]) < 2e-5


@pytest.mark.parametrize("name", ["rotate_clockwise", "rotate"])
def test_guided_loop_2d(g2, samp2, name):
    trace = []
    out = samp2.guided_sample(torch.from_numpy(g2["noise"]), name, trace=trace)
    for rec, i in zip(trace, [s for _ in range(2) for s in range(5)]):
        oi = rec["obj"]
        assert rel(rec["eps"].numpy(), g2[f"loop_{name}_eps_o{oi}_s{i}"]) < 1e-4
        assert rel(rec["grad"].numpy(), g2[f"loop_{name}_grad_o{oi}_s{i}"]) < 1e-4
        assert rel(rec["sample"].numpy(), g2[f"loop_{name}_sample_o{oi}_s{i}"]) < 1e-4
    assert out.shape == (2, 4, 14, 1)


def test_multi_object_loop_2d(g2, samp2):
    trace = []
    samp2.guided_sample_multi_object(torch.from_numpy(g2["noise"]), "shift_up", trace=trace)
    for i, rec in enumerate(trace):
        for k in ("eps", "grad", "sample"):
            assert rel(rec[k].numpy(), g2[f"multi_shift_up_{k}_s{i}"]) < 1e-4


def test_convergence_2d(g2, samp2):
    noise = torch.from_numpy(g2["noise"])
    ung = samp2.unguided_sample(noise)
    assert rel(ung.numpy(), g2["unguided"]) < 1e-4
    ung = torch.from_numpy(g2["unguided"])
    for oi in range(2):
        c = samp2.convergence_centers(ung, oi)
        assert np.array_equal(c.numpy(), g2[f"conv_centers_o{oi}"])
        got = samp2.cond_fn(noise, 9, "convergence", oi, convergence_centers=c)
        assert rel(got.numpy(), g2[f"grad_o{oi}_t9_convergence"]) < 2e-5


def test_profile_pass_2d(g2, samp2):
    final = torch.from_numpy(g2["loop_rotate_clockwise_sample_o0_s4"])
    lg = samp2.profile_logits(final, 0)
    assert rel(lg.reshape(-1, 3).numpy(), g2["profile_logits_o0"]) < 1e-5
    s = samp2.score(final, 0, "rotate_clockwise")
    want = -g2["profile_logits_o0"].reshape(int(g2["grid_size"]), 4, 3)[..., 0].mean(0)
    assert np.allclose(s.numpy(), want, atol=1e-5)
    assert int(orc.best_of_n(s)) == int(np.argmax(want))


def test_pointnet2(g3):
    sd = orc.strip_prefix(syn.dynamics3d_state_dict(0))
    objs = torch.from_numpy(g3["objects"]); st = torch.from_numpy(g3["fps_starts"])
    with torch.no_grad():
        for oi in range(2):
            code = orc.pointnet2_encode(sd, objs[oi].t()[None], st[oi:oi + 1])
            assert rel(code.numpy(), g3[f"code_o{oi}"]) < 1e-5
        # batched over objects == one at a time
        both = orc.pointnet2_encode(sd, objs.permute(0, 2, 1).contiguous(), st)
        assert rel(both[1:2].numpy(), g3["code_o1"]) < 1e-5


def test_dynamics3d_forward_per_row_clouds(g3):
    sd = orc.strip_prefix(syn.dynamics3d_state_dict(0))
    n = g3["fwd_ctrl"].shape[0]
    cloud = torch.from_numpy(g3["objects"][1]).t()[None].repeat(n, 1, 1)
    st = torch.from_numpy(g3["fps_starts"][1:2]).expand(n, -1)
    with torch.no_grad():
        out = orc.dynamics3d_forward(sd, *(torch.from_numpy(g3[k]) for k in ("fwd_ctrl", "fwd_ori", "fwd_pos", "fwd_t")),
                                     object_vertices=cloud, fps_start=st)
    assert rel(out.numpy(), g3["fwd_logits_o1"]) < 1e-5


@pytest.mark.parametrize("t,name", [(12, "rotate_clockwise"), (6, "rotate"), (0, "counterclockwise_down")])
def test_cond_fn_3d(g3, samp3, t, name):
    x = torch.from_numpy(g3["noise"])
    for oi in range(2):
        got = samp3.cond_fn(x, t, name, oi)
        assert rel(got.numpy(), g3[f"grad_o{oi}_t{t}_{name}"]) < 5e-5


def test_guided_loop_3d(g3, samp3):
    one = orc.OracleSampler("point_3d", samp3.unet_sd, samp3.dyn_sd, samp3.object_vertices[:1], samp3.grid_size,
                            samp3.num_pos, sub_batch_size=samp3.sub_batch_size, fps_start=samp3.fps_start[:1])
    trace = []
    one.guided_sample(torch.from_numpy(g3["noise"]), "rotate_clockwise", trace=trace)
    for i, rec in enumerate(trace):
        for k in ("eps", "grad", "sample"):
            assert rel(rec[k].numpy(), g3[f"loop_{k}_s{i}"]) < 2e-4
