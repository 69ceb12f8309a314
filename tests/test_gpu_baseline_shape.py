"""Oracle-direct parity at BASELINE shape, unscaled tolerances.

One object of BASELINE.json configs[1] (2D: 256 candidates x 36 orientations x 5x5 positions = 900 pose rows) and of
configs[2] (3D: 128 candidates x 45 x 5x5 = 1125 pose rows): one ``cond_fn`` step and one full 5-step guided run +
scores + top-4, every arithmetic mode of the CUDA path against ``OracleSampler`` (the CPU restatement pinned to the
real reference by tests/test_oracle_golden.py).  No sqrt(900/G) scaling, no set-aside candidates:

    per-step guidance gradient and denoiser output  rel-err  <= 1e-3 (fp32, fp32_simt)   <= 2e-2 (bf16, fp16)
        (every step of the oracle's own trajectory: the step's input state is the oracle's, "teacher forced")
    final predicted scores of the free-running run  abs-err  <= 1e-3                      (bf16: see SCORE_TOL)
    top-4 indices                                   exact

The free-running run is a 5-step feedback loop through a piecewise-linear network.  2D (guidance scale 1e-3): final
scores and designs are held to 1e-3 as is (single-pass bf16: 2e-3 / 2e-2).  The 3D slice (SCALE_3D = 0.5, random-init
weights, clipping at +-1) amplifies any per-step difference -- the exact-order fp32 CUDA-core path itself lands 1.2e-3
from the CPU in designs -- so its bound is CALIBRATED WITH THE ORACLE: the oracle's own loop is re-run with every
step's gradient perturbed by random noise of relative size equal to the north-star per-step tolerance of the mode
(1e-3 / 2e-2), and a CUDA mode, whose per-step error is inside that tolerance, must land within twice the oracle's
own response (never tighter than 1e-3).  The measured responses are printed.

The worst candidate and the rank-1/rank-2 score margin are printed (pytest -s) for DESIGN.md.
"""
import numpy as np
import pytest
import torch

import dgdm_oracle as orc
from dgdm_b200 import synthetic as syn
from test_gpu_parity import make2d, make3d, rel

pytestmark = pytest.mark.gpu

MODES = ["fp32_simt", "fp32", "fp16x3", "fp16", "bf16"]
GRAD_TOL = {"fp32_simt": 1e-3, "fp32": 1e-3, "fp16x3": 1e-3, "fp16": 2e-2, "bf16": 2e-2}
EXACT_TOPK = {"fp32_simt", "fp32", "fp16x3", "fp16"}
# Free-running run (5 guided steps feeding back): final scores (north-star: 1e-3) and designs.  3D: see the module docstring
# (bounds calibrated per fixture by _response below).
SCORE_TOL = {("2d", "fp32_simt"): 1e-3, ("2d", "fp32"): 1e-3, ("2d", "fp16x3"): 1e-3, ("2d", "fp16"): 1e-3, ("2d", "bf16"): 2e-3}
DESIGN_TOL = {("2d", "fp32_simt"): 1e-3, ("2d", "fp32"): 1e-3, ("2d", "fp16x3"): 1e-3, ("2d", "fp16"): 1e-3, ("2d", "bf16"): 2e-2}
OBJ = "rotate_clockwise"


def _per_candidate(got, want):
    a = got.detach().cpu().double().reshape(want.shape[0], -1)
    b = want.double().reshape(want.shape[0], -1)
    return ((a - b).norm(dim=1) / b.norm(dim=1).clamp_min(1e-30)).numpy()


def _response(samp, noise, designs, scores, rel_step, seed=7):
    """The oracle's own sensitivity: its guided loop (OracleSampler.guided_sample) with every step's guidance gradient
    perturbed, per candidate, by white noise of relative L2 size ``rel_step``.  Returns (designs rel-err, scores max
    abs-err) of the perturbed run against the unperturbed one."""
    g = torch.Generator().manual_seed(seed)
    scale = samp.classifier_scale(OBJ)
    sample = noise.clone()
    for t in samp.timesteps:
        grad = samp.cond_fn(sample, int(t), OBJ, 0)
        n = torch.randn(grad.shape, generator=g)
        n = n * (rel_step * grad.flatten(1).norm(dim=1) / n.flatten(1).norm(dim=1)).view(-1, 1, 1)
        sample, _ = samp._step(sample, int(t), grad + n, scale)
    return rel(sample[None], designs), float((samp.score(sample, 0, OBJ) - scores).abs().max())


@pytest.fixture(scope="module")
def oracle2d():
    B, grid, npos = 256, 36, 5
    objs = syn.objects_2d(1)
    samp = orc.OracleSampler("point", syn.unet1d_state_dict(0), syn.dynamics2d_state_dict(0), objs, grid, npos)
    noise = syn.initial_noise(B, 14)
    out = {"objs": objs, "noise": noise, "B": B, "grid": grid, "npos": npos}
    out["grad_lin"] = samp.cond_fn(noise, 6, OBJ, 0)
    out["grad_sq"] = samp.cond_fn(noise, 6, "rotate", 0)
    out["trace"] = []
    designs = samp.guided_sample(noise, OBJ, trace=out["trace"])    # (1, B, P, 1)
    out["designs"] = designs
    out["scores"] = samp.score(designs[0], 0, OBJ)
    return out


@pytest.fixture(scope="module")
def oracle3d():
    B, grid, npos = 128, 45, 5
    objs, st = syn.objects_3d(1), syn.fps_starts(1)
    samp = orc.OracleSampler("point_3d", syn.unet1d_state_dict(0), syn.dynamics3d_state_dict(0), objs, grid, npos,
                             sub_batch_size=512, fps_start=st)
    noise = syn.initial_noise(B, 42)
    out = {"objs": objs, "st": st, "noise": noise, "B": B, "grid": grid, "npos": npos}
    out["grad_lin"] = samp.cond_fn(noise, 6, OBJ, 0)
    out["trace"] = []
    designs = samp.guided_sample(noise, OBJ, trace=out["trace"])
    out["designs"] = designs
    out["scores"] = samp.score(designs[0], 0, OBJ)
    out["response"] = {tol: _response(samp, noise, designs, out["scores"], tol) for tol in (1e-3, 2e-2)}
    print("[3d oracle] response of the final (designs rel-err, scores abs-err) to a per-step gradient perturbation of "
          + ", ".join(f"{tol:g}: ({d:.3e}, {s:.3e})" for tol, (d, s) in out["response"].items()))
    return out


def _check_step(dm, o, precision, tag):
    for name, key in ((OBJ, "grad_lin"), ("rotate", "grad_sq")):
        if key not in o:
            continue
        g = dm.cond_fn(o["noise"].cuda(), 6, opt_obj=name, object_vertices=dm.object_vertices[0])
        r = rel(g, o[key])
        per = _per_candidate(g, o[key])
        print(f"[{tag} {precision} {name}] cond_fn vs oracle: aggregate rel-err {r:.3e}, worst candidate "
              f"{per.max():.3e} (#{int(per.argmax())}), median {np.median(per):.3e}")
        assert r <= GRAD_TOL[precision], f"{tag} {precision} {name}: {r:.3e} > {GRAD_TOL[precision]}"


def _check_run(dm, o, precision, tag):
    B = o["B"]
    # every step of the oracle's trajectory, from the oracle's own input state
    state = o["noise"]
    for st in o["trace"]:
        xin = state.cuda()
        eps = dm.noise_pred_net(xin, st["t"])
        g = dm.cond_fn(xin, st["t"], opt_obj=OBJ, object_vertices=dm.object_vertices[0])
        e_err, g_err = rel(eps, st["eps"]), rel(g, st["grad"])
        per = _per_candidate(g, st["grad"])
        print(f"[{tag} {precision}] step t={st['t']:2d}: denoiser rel-err {e_err:.3e}, guidance rel-err {g_err:.3e} "
              f"(worst candidate {per.max():.3e})")
        assert e_err <= GRAD_TOL[precision] and g_err <= GRAD_TOL[precision], (st["t"], e_err, g_err)
        state = st["sample"]
    out = dm.guided_sample(0, B, o["noise"], opt_obj=OBJ, top_k=4)
    d_err = rel(out["designs"], o["designs"])
    s_err = float((out["scores"].cpu()[0] - o["scores"]).abs().max())
    want_top = orc.top_k(o["scores"][None], 4)[0].tolist()
    got_top = out["best_ids"][0].cpu().tolist()
    srt = torch.sort(o["scores"], descending=True).values
    margins = (srt[:4] - srt[1:5]).tolist()
    print(f"[{tag} {precision}] full run vs oracle: designs rel-err {d_err:.3e}, scores max abs-err {s_err:.3e}, "
          f"top-4 {got_top} (oracle {want_top}), oracle rank margins {['%.2e' % m for m in margins]}")
    if (tag, precision) in SCORE_TOL:
        s_tol, d_tol = SCORE_TOL[(tag, precision)], DESIGN_TOL[(tag, precision)]
    else:                                             # 3D: twice the oracle's own response to an in-tolerance per-step error
        d_resp, s_resp = o["response"][GRAD_TOL[precision]]
        s_tol, d_tol = max(1e-3, 2 * s_resp), max(1e-3, 2 * d_resp)
        print(f"[{tag} {precision}] calibrated bounds: scores {s_tol:.3e}, designs {d_tol:.3e}")
    assert s_err <= s_tol, (s_err, s_tol)
    assert d_err <= d_tol, (d_err, d_tol)
    # Selection: every rank whose oracle score is separated from both neighbours by more than twice THIS run's score
    # error must be the same candidate; closer ranks are a tie at this precision (the 3D slice has a 5e-6 margin
    # between its two best candidates -- 200x below even the exact-order fp32 path's score error).
    for k in range(4):
        lo_m = margins[k]
        hi_m = margins[k - 1] if k > 0 else float("inf")
        if min(lo_m, hi_m) > 2 * s_err:
            assert got_top[k] == want_top[k], (k, got_top, want_top, margins, s_err)
    assert sorted(got_top[:2]) == sorted(want_top[:2]) or margins[1] <= 2 * s_err
    if tag == "2d" and precision in EXACT_TOPK:
        assert got_top == want_top


@pytest.mark.parametrize("precision", MODES)
def test_cond_fn_baseline_shape_2d_vs_oracle(precision, oracle2d):
    o = oracle2d
    _check_step(make2d(precision, o["objs"], o["grid"], o["npos"]), o, precision, "2d")


@pytest.mark.parametrize("precision", MODES)
def test_full_run_baseline_shape_2d_vs_oracle(precision, oracle2d):
    o = oracle2d
    _check_run(make2d(precision, o["objs"], o["grid"], o["npos"]), o, precision, "2d")


@pytest.mark.parametrize("precision", MODES)
def test_cond_fn_baseline_shape_3d_vs_oracle(precision, oracle3d):
    o = oracle3d
    _check_step(make3d(precision, o["objs"], o["st"], o["grid"], o["npos"]), o, precision, "3d")


@pytest.mark.parametrize("precision", MODES)
def test_full_run_baseline_shape_3d_vs_oracle(precision, oracle3d):
    o = oracle3d
    _check_run(make3d(precision, o["objs"], o["st"], o["grid"], o["npos"]), o, precision, "3d")


# ---------------------------------------------------------------------------------------------- the real reference
@pytest.fixture(scope="module")
def gbig(golden_dir):
    import os
    return dict(np.load(os.path.join(golden_dir, "golden_bigG.npz")))


@pytest.mark.parametrize("precision", MODES)
def test_cond_fn_reference_fixture_at_baseline_pose_grid(precision, gbig):
    """``cond_fn`` of the REAL reference (tests/golden/make_golden.py bigG: generator/diffusion.py:473-504 imported
    unmodified) at the BASELINE pose grids -- 2D 36 x 5 x 5 = 900 rows per candidate, 3D 45 x 5 x 5 = 1125 with the stock
    512-row sub-batches -- 16 / 8 candidates.  Unscaled bounds, no candidate set aside."""
    g = gbig
    objs2 = torch.from_numpy(g["objects_2d"])
    dm = make2d(precision, objs2, 36, 5)
    x = torch.from_numpy(g["noise_2d"]).cuda()
    for t, name in ((6, "rotate_clockwise"), (12, "rotate"), (0, "counterclockwise_left")):
        got = dm.cond_fn(x, t, opt_obj=name, object_vertices=dm.object_vertices[1])
        want = torch.from_numpy(g[f"grad_2d_t{t}_{name}"])
        r, per = rel(got, want), _per_candidate(got, want)
        print(f"[ref 2d G=900 {precision} t={t} {name}] rel-err {r:.3e}, worst candidate {per.max():.3e}")
        assert r <= GRAD_TOL[precision]
    objs3, st = torch.from_numpy(g["objects_3d"]), torch.from_numpy(g["fps_starts"])
    dm3 = make3d(precision, objs3, st, 45, 5)
    x3 = torch.from_numpy(g["noise_3d"]).cuda()
    for t, name in ((6, "rotate_clockwise"), (3, "rotate")):
        got = dm3.cond_fn(x3, t, opt_obj=name, object_vertices=dm3.object_vertices[1])
        want = torch.from_numpy(g[f"grad_3d_t{t}_{name}"])
        r, per = rel(got, want), _per_candidate(got, want)
        print(f"[ref 3d G=1125 {precision} t={t} {name}] rel-err {r:.3e}, worst candidate {per.max():.3e}")
        assert r <= GRAD_TOL[precision]
