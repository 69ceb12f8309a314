#!/usr/bin/env python
"""Recipe for oracle/_ref/: the reference's OWN sampler, compiled, for bench.py's ``--impl reference`` arm.

The reference is pure Python, so "building" it is byte-compiling the seven files of the path that import with torch +
numpy alone, from the sources where they lie under /root/reference, into CPython bytecode (``<module>.refbin``: a .pyc under a
name the snapshot tools do not filter out; oracle/ref_arm.py has the importer) under oracle/_ref/reference/ (git-ignored build output: it travels to the GPU box with the snapshot like a built .so, and never
enters the history; no reference source is copied into this repository).  Run by ``__graft_entry__.build()`` whenever
/root/reference is present (the authoring container); on the GPU box the modules that travelled are used (same image,
same CPython).  TEST INFRASTRUCTURE: only bench.py's reference arm (through oracle/ref_arm.py) imports it.
"""
import os
import py_compile
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
    if os.path.isdir(DST):
        shutil.rmtree(DST)                        # never leave stale modules (or sources of an older recipe) behind
    for rel in FILES:
        src, dst = os.path.join(REF, rel), os.path.join(DST, rel[:-3] + ".refbin")   # module.py -> module.refbin (pyc bytes)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        py_compile.compile(src, cfile=dst, dfile=rel, doraise=True, optimize=0)
    with open(os.path.join(DST, "PROVENANCE.txt"), "w") as f:
        f.write("CPython %d.%d bytecode of real-stanford/dgdm files, compiled unmodified by oracle/build_ref.py from %s:\n"
                % (sys.version_info[0], sys.version_info[1], REF) + "\n".join(FILES) + "\n")
    if verbose:
        print(f"oracle/_ref: {len(FILES)} reference modules byte-compiled -> {DST}")
    return True


if __name__ == "__main__":
    sys.exit(0 if build() else 1)
