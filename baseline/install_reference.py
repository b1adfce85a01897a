#!/usr/bin/env python
"""Stage the UNMODIFIED reference tree under ``baseline/_ref/`` (git-ignored, shipped to the GPU box by gpurun).

The reference (GUOShuxuan/kd-6d-pose-adlp) is a flat script tree without setup.py / pyproject.toml, so the generic
``pip install --target baseline/_ref /root/reference`` recipe does not apply; its Python sources are staged as they lie
(nothing is edited, nothing is committed).  ``tests/test_dropin_reference_gpu.py`` then applies the import swap of
INTEGRATION.md section 2 to the reference's own ``models/model_kd.py`` ON THE GPU BOX and runs its ``PoseModuleKD.forward``
-- teacher branch through our ``PostProcessorKD``, student branch through our ``KDPoseLoss`` on the reference's real
``PoseLossDzi`` / ``prepare_targets``.  ``__graft_entry__.build()`` calls this whenever ``/root/reference`` is present.
"""
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
DST = os.path.join(HERE, "_ref")
KEEP_DIRS = ("arguments", "backbone", "configs", "libs", "losses", "models", "postprocess", "tools")


def install(src="/root/reference", quiet=False):
    if not os.path.isdir(os.path.join(src, "losses")):
        return False
    if os.path.isdir(DST):
        shutil.rmtree(DST)
    os.makedirs(DST)
    n = 0
    for d in KEEP_DIRS:
        for root, _dirs, files in os.walk(os.path.join(src, d)):
            for f in files:
                if f.endswith((".py", ".yaml", ".json", ".txt")):
                    rel = os.path.relpath(os.path.join(root, f), src)
                    os.makedirs(os.path.dirname(os.path.join(DST, rel)), exist_ok=True)
                    shutil.copyfile(os.path.join(root, f), os.path.join(DST, rel))
                    n += 1
    for f in os.listdir(src):
        if f.endswith((".py", ".txt", ".md", ".sh")) and os.path.isfile(os.path.join(src, f)):
            shutil.copyfile(os.path.join(src, f), os.path.join(DST, f))
            n += 1
    if not quiet:
        print(f"[baseline] staged {n} reference files under {DST}")
    return True


if __name__ == "__main__":
    sys.exit(0 if install(*(sys.argv[1:2])) else 1)
