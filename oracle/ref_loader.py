"""Import the reference's OWN in-tree modules from ``/root/reference`` (authoring container only).

TEST INFRASTRUCTURE ONLY.  ``/root/reference`` does not exist on the GPU box, so this loader is
used exclusively (a) by ``tests/golden/make_golden.py`` to generate the committed fixtures and
(b) by CPU tests that are skipped when the tree is absent.  Nothing is copied: the modules are
imported from where they lie.

Work-arounds:
* stub modules for packages the reference imports but this image lacks
  (``trimesh, transforms3d, pyrender, psutil?, matplotlib, tensorboardX``) -- none is on the path;
* ``geomloss`` -> ``oracle.geomloss_ref`` (the restatement; geomloss 0.2.4 is not installable here);
* ``np.float`` alias (``models/model.py:292,300`` use the removed numpy alias);
* ``Tensor.cuda`` -> identity for CPU runs (``models/model.py:128``, ``losses/kd_loss.py:103``,
  ``postprocess/postprocess_kd.py:93,96,202``);
* the two visualiser functions called by ``losses/kd_loss.py:88-97`` are no-ops.
"""
from __future__ import annotations

import importlib
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("KDOT_REFERENCE_ROOT", "/root/reference")


def available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "losses"))


def _stub(name, **attrs):
    if name in sys.modules:
        return sys.modules[name]
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


_loaded = {}


def load():
    """Returns a namespace with the reference's classes/functions on the hot path."""
    if _loaded:
        return types.SimpleNamespace(**_loaded)
    if not available():
        raise RuntimeError("reference tree not present at %s" % REFERENCE_ROOT)
    import numpy as np
    import torch

    from oracle import geomloss_ref

    for name in ("trimesh", "transforms3d", "pyrender", "tensorboardX"):
        try:
            importlib.import_module(name)
        except Exception:
            _stub(name)
    try:
        importlib.import_module("matplotlib")
    except Exception:
        mpl = _stub("matplotlib", use=lambda *a, **k: None)
        mpl.pyplot = _stub("matplotlib.pyplot")
        mpl.cm = _stub("matplotlib.cm")
        mpl.patches = _stub("matplotlib.patches")
    try:
        importlib.import_module("psutil")
    except Exception:
        _stub("psutil")
    _stub("geomloss", SamplesLoss=geomloss_ref.SamplesLoss)
    if not hasattr(np, "float"):
        np.float = float  # noqa: removed alias used by models/model.py
    if not torch.cuda.is_available():
        torch.Tensor.cuda = lambda self, *a, **k: self  # CPU run of code that hard-codes .cuda()

    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)

    loss_libs = importlib.import_module("losses.loss_libs")
    loss_base = importlib.import_module("losses.loss")
    kd_loss = importlib.import_module("losses.kd_loss")
    kd_loss.vis_pxpy_post_train = lambda *a, **k: None
    kd_loss.vis_pxpy_post_train_weight = lambda *a, **k: None
    model = importlib.import_module("models.model")
    pp_kd = importlib.import_module("postprocess.postprocess_kd")
    poses = importlib.import_module("libs.poses")
    boxlist = importlib.import_module("libs.boxlist")

    _loaded.update(
        kd_loss_2d=loss_libs.kd_loss_2d,
        KDPoseLoss=kd_loss.KDPoseLoss,
        PoseLossDzi=loss_base.PoseLossDzi,
        SigmoidFocalLoss=loss_base.SigmoidFocalLoss,
        concat_box_prediction_layers=loss_base.concat_box_prediction_layers,
        TargetCoder=model.TargetCoder,
        AnchorGenerator=model.AnchorGenerator,
        PostProcessorKD=pp_kd.PostProcessorKD,
        PoseAnnot=poses.PoseAnnot,
        BoxList=boxlist.BoxList,
        modules=dict(loss_libs=loss_libs, loss=loss_base, kd_loss=kd_loss, model=model, pp_kd=pp_kd),
    )
    return types.SimpleNamespace(**_loaded)
