"""bench.py's reference arm: the reference's own ``Diffusion.guided_sample`` (generator/diffusion.py:541-576), run
unmodified from the byte-compiled modules under oracle/_ref/reference (oracle/build_ref.py) on the host cores.

What runs is the reference's method itself -- its objects loop, its ``cond_fn`` with the Python list-comprehension
tiling (diffusion.py:478-486, a third of its time), its ConditionalUnet1D, its DataParallel wrapper -- with random-init
synthetic weights (dgdm_b200.synthetic, loaded ``strict=True``).  Two things are substituted, both outside the path:
the DDIM scheduler (diffusers is not installable offline: oracle/ref_stubs.StubDDIM) and the MuJoCo evaluator the
method calls once per object AFTER the guided loop (diffusion.py:577-580), replaced by a recorder that keeps the final
designs and returns no metrics -- the method then skips its logging for that object (``if len(metrics) == 0:
continue``) and raises IndexError at its summary line (:611), which ``guided_sample`` below swallows.
TEST INFRASTRUCTURE (see dgdm_oracle.py).
"""
import importlib.abc
import importlib.util
import marshal
import os
import sys
import tempfile

import torch
import torch.nn as nn

HERE = os.path.dirname(os.path.abspath(__file__))
REF_ROOT = os.path.join(HERE, "_ref", "reference")
REPO = os.path.dirname(HERE)
for p in (REPO, HERE):
    if p not in sys.path:
        sys.path.insert(0, p)

_RECORDED = []


class _RefbinFinder(importlib.abc.MetaPathFinder, importlib.abc.Loader):
    """Imports ``<package dir>/<module>.refbin`` (the bytes of a .pyc written by oracle/build_ref.py) as ``module``."""

    def find_spec(self, fullname, path, target=None):
        name = fullname.rpartition(".")[2]
        for entry in (path or [REF_ROOT]):
            f = os.path.join(entry, name + ".refbin")
            if os.path.isfile(f) and os.path.abspath(f).startswith(os.path.abspath(REF_ROOT)):
                return importlib.util.spec_from_loader(fullname, self, origin=f)
        return None

    def exec_module(self, module):
        with open(module.__spec__.origin, "rb") as fh:
            data = fh.read()
        if data[:4] != importlib.util.MAGIC_NUMBER:
            raise ImportError(f"{module.__spec__.origin}: bytecode of another CPython (rebuild with oracle/build_ref.py)")
        module.__file__ = module.__spec__.origin[:-len(".refbin")] + ".py"      # the reference modules read __file__
        exec(marshal.loads(data[16:]), module.__dict__)


def _install_finder():
    if not any(isinstance(f, _RefbinFinder) for f in sys.meta_path):
        sys.meta_path.append(_RefbinFinder())


def available() -> bool:
    return os.path.exists(os.path.join(REF_ROOT, "generator", "diffusion.refbin"))


def _record(sample, object_ids, save_dir, **kwargs):
    _RECORDED.append(torch.as_tensor(sample).clone())
    return [], [], [], [], [], [], [], []          # gripper_imgs, metrics, profiles, profiles_x, profiles_y, finals, videos, dirs


def build(mode: str, objects: torch.Tensor, grid_size: int, num_pos: int, sub_batch_size: int = 512, seed: int = 0):
    """The reference's Diffusion LightningModule around its real networks (eval mode, parameters frozen as
    generator/train.py:91-92 does)."""
    import ref_stubs
    _install_finder()
    ref_stubs.install_stubs(REF_ROOT, sim_test_batch=_record, sim_test_batch_3d=_record)
    import warnings
    warnings.filterwarnings("ignore")
    from generator.diffusion import Diffusion
    from generator.diffusion_utils import ConditionalUnet1D
    from dynamics.profile_forward_2d import ProfileForward2DModel
    from dynamics.profile_forward_3d import ProfileForward3DModel
    from dgdm_b200 import synthetic as syn
    P = 14 if mode == "point" else 42
    unet = ConditionalUnet1D(input_dim=1, global_cond_dim=0, down_dims=[128, 256], diffusion_step_embed_dim=32)
    unet.load_state_dict(syn.unet1d_state_dict(seed), strict=True)
    if mode == "point":
        net = nn.DataParallel(ProfileForward2DModel(output_ch=3, params_ch=P, object_ch=200))
        net.load_state_dict(syn.dynamics2d_state_dict(seed), strict=True)
    else:
        net = nn.DataParallel(ProfileForward3DModel(output_ch=3, params_ch=P))
        net.load_state_dict(syn.dynamics3d_state_dict(seed), strict=True)
    for p in net.parameters():
        p.requires_grad = False
    dm = Diffusion(noise_pred_net=unet, noise_scheduler=ref_stubs.StubDDIM(num_train_timesteps=15),
                   num_inference_steps=5, mode=mode, input_dim=1, num_points=P, class_cond=True, classifier_model=net,
                   grid_size=grid_size, num_pos=num_pos, object_vertices=objects, object_ids=list(range(len(objects))),
                   sub_batch_size=sub_batch_size, seed=seed)
    dm.eval()
    return dm


def guided_sample(dm, noise: torch.Tensor, opt_obj: str = "rotate_clockwise") -> torch.Tensor:
    """Run the reference's ``guided_sample`` as is; returns the final designs (n_obj, B, P, 1) its simulator call received."""
    del _RECORDED[:]
    with tempfile.TemporaryDirectory() as tmp, torch.no_grad():       # validation_step runs under no_grad (train.py:152)
        try:
            dm.guided_sample(0, noise.shape[0], noise, tmp, opt_obj=opt_obj, ori_range=[-1.0, 1.0])
        except IndexError:
            pass                                                       # the summary line after the objects loop (no metrics)
    return torch.stack(_RECORDED)
