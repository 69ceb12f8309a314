import sys, numpy as np, torch
import torch.nn.functional as F
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/oracle")
import dgdm_oracle as orc
from dgdm_b200 import synthetic as syn
g2 = dict(np.load("tests/golden/golden_2d.npz"))
objs = torch.from_numpy(g2["objects"])
sd = orc.strip_prefix(syn.dynamics2d_state_dict(0))
x = torch.from_numpy(g2["loop_rotate_sample_o1_s2"])[..., 0]   # (4,14)
B, G = 4, 24
ori, pos = orc.pose_grid(B, 6, 2, (-1., 1.))
pts = x.repeat(G, 1)
tf = torch.full((B*G,), 3.0)/15
obj = objs[1].reshape(1, -1).expand(B*G, -1)
def pre_acts(sd, dt):
    c = lambda t: t.to(dt)
    g = orc._mlp2({k: c(v) for k, v in sd.items() if v.is_floating_point()}, "gripper_encoder", c(pts), F.relu)
    sdd = {k: (c(v) if v.is_floating_point() else v) for k, v in sd.items()}
    pose = torch.cat([orc.fourier_embed(ori), orc.fourier_embed(pos)], 1)
    o = orc._mlp2(sdd, "object_encoder", c(obj), F.relu)
    te = orc._mlp2(sdd, "time_encoder", c(orc.timestep_embedding(tf, 128)), F.silu)
    h = torch.cat([o, g, c(pose), te], 1)
    zs = []
    for i in range(8):
        h = F.linear(h, sdd[f"linears.{3*i}.weight"], sdd[f"linears.{3*i}.bias"])
        p = f"linears.{3*i+1}"
        h = F.batch_norm(h, sdd[p+".running_mean"], sdd[p+".running_var"], sdd[p+".weight"], sdd[p+".bias"], training=False, eps=1e-5)
        zs.append(h.clone())
        h = F.relu(h)
    return zs
z32 = pre_acts(sd, torch.float32)
z64 = pre_acts(sd, torch.float64)
for l in range(8):
    a = z64[l].abs()
    rows = torch.arange(B*G) % B      # candidate index of reference row r = g*B + b
    m = a[rows == 2]
    flips = ((z32[l] > 0) != (z64[l] > 0)).sum().item()
    print("layer", l+1, "min|z| cand2 %.3e" % m.min().item(), "scale %.3f" % a.mean().item(), "fp32-vs-fp64 sign flips:", flips,
          "max |z32-z64| %.2e" % (z32[l].double()-z64[l]).abs().max().item())
# also gripper encoder relu
