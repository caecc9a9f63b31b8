"""Stage the UNMODIFIED reference (corenel/pytorch-glow, pure Python, no packaging) under baseline/_ref/ so that
`bench.py --impl reference` / `--impl reference-gpu` can run the real thing on the GPU box.

The reference has no setup.py / pyproject (`pip install` is not applicable), so "installing" it
is copying its importable packages.  baseline/_ref/ is git-ignored (never part of the history) but travels with
gpurun.  Run in the build container, where /root/reference exists (called from __graft_entry__.build()):

    python baseline/stage_reference.py
"""
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("GLOW_REFERENCE", "/root/reference")
DST = os.path.join(HERE, "_ref")
SHIMS = os.path.join(os.path.dirname(HERE), "tests", "golden", "_shims")


def stage(verbose=True):
    if not os.path.isdir(os.path.join(REF, "network")):
        if verbose:
            print("stage_reference: %s not found, nothing staged" % REF)
        return False
    os.makedirs(DST, exist_ok=True)
    for pkg in ("network", "misc", "profile"):
        dst = os.path.join(DST, pkg)
        if os.path.isdir(dst):
            shutil.rmtree(dst)
        shutil.copytree(os.path.join(REF, pkg), dst, ignore=shutil.ignore_patterns("*.png", "*.pyc", "__pycache__"))
    # the two import shims for packages the image lacks (easydict, tensorboardX)
    dst = os.path.join(DST, "_shims")
    if os.path.isdir(dst):
        shutil.rmtree(dst)
    shutil.copytree(SHIMS, dst, ignore=shutil.ignore_patterns("*.pyc", "__pycache__"))
    if verbose:
        print("stage_reference: staged network/ misc/ profile/ + shims under", DST)
    return True


if __name__ == "__main__":
    sys.exit(0 if stage() else 1)
