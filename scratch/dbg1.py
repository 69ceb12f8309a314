import os, sys, numpy as np, torch
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/oracle"); sys.path.insert(0, "/root/repo/tests")
from dgdm_b200 import synthetic as syn
from test_gpu_parity import make2d, rel
g2 = dict(np.load("tests/golden/golden_2d.npz"))
dm = make2d("fp32_simt", torch.from_numpy(g2["objects"]), 6, 2)
noise = torch.from_numpy(g2["noise"])
for name in ("rotate_clockwise", "rotate"):
    trace = []
    dm.guided_sample(0, 4, noise, opt_obj=name, trace=trace)
    for i, rec in enumerate(trace):
        for oi in range(2):
            print(name, "step", i, "t", rec["t"], "obj", oi, "eps %.2e grad %.2e sample %.2e" % (
                rel(rec["eps"][oi], g2[f"loop_{name}_eps_o{oi}_s{i}"]), rel(rec["grad"][oi], g2[f"loop_{name}_grad_o{oi}_s{i}"]),
                rel(rec["sample"][oi], g2[f"loop_{name}_sample_o{oi}_s{i}"])))
    # guidance evaluated on the golden inputs of each step
    ts = [12, 9, 6, 3, 0]
    for oi in range(2):
        for i in range(1, 5):
            xin = torch.from_numpy(g2[f"loop_{name}_sample_o{oi}_s{i-1}"]).cuda()
            g = dm.cond_fn(xin, ts[i], opt_obj=name, object_vertices=dm.object_vertices[oi])
            d = (g.cpu().numpy() - g2[f"loop_{name}_grad_o{oi}_s{i}"]).reshape(4, -1)
            print("  on golden input: obj", oi, "step", i, "rel %.2e" % rel(g, g2[f"loop_{name}_grad_o{oi}_s{i}"]), "per-cand", np.abs(d).max(1))
