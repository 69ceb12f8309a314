#!/usr/bin/env python
"""Recipe for oracle/_ref/: the reference's OWN sampler sources, for bench.py's ``--impl reference`` arm.

The reference is pure Python, so "building" it is copying the seven files of the path that import with torch + numpy
alone, from where they lie under /root/reference, into oracle/_ref/reference/ (git-ignored build output: it travels
to the GPU box with the snapshot like a built .so, and never enters the history).  Nothing is modified.  Run by
``__graft_entry__.build()`` whenever /root/reference is present (the authoring container); on the GPU box the copies
that travelled are used.  TEST INFRASTRUCTURE: only bench.py's reference arm (through oracle/ref_arm.py) reads it.
"""
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("DGDM_REFERENCE_ROOT", "/root/reference")
DST = os.path.join(HERE, "_ref", "reference")
FILES = ["generator/diffusion.py", "generator/diffusion_utils.py", "dynamics/profile_forward_2d.py",
         "dynamics/profile_forward_3d.py", "dynamics/models/pointnet2.py", "dynamics/models/pointnet2_utils.py",
         "dynamics/metrics.py"]


def build(verbose: bool = True) -> bool:
    if not os.path.isdir(REF):
        if verbose:
            print(f"oracle/_ref: {REF} not present; keeping whatever travelled with the tree")
        return os.path.isdir(DST)
    for rel in FILES:
        src, dst = os.path.join(REF, rel), os.path.join(DST, rel)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        shutil.copyfile(src, dst)
    with open(os.path.join(DST, "PROVENANCE.txt"), "w") as f:
        f.write("Unmodified copies of real-stanford/dgdm files made by oracle/build_ref.py from " + REF + ":\n"
                + "\n".join(FILES) + "\n")
    if verbose:
        print(f"oracle/_ref: {len(FILES)} reference files -> {DST}")
    return True


if __name__ == "__main__":
    sys.exit(0 if build() else 1)
