#!/usr/bin/env python
"""Generates the committed golden fixtures by running the REFERENCE'S OWN in-tree code in the authoring
container (``/root/reference`` mounted read-only).  Run:  ``python tests/golden/make_golden.py``.

Nothing from the reference is copied: its modules are imported where they lie (``oracle/ref_loader.py``),
driven on seeded synthetic inputs, and only inputs/outputs are stored as small ``.npz`` files:

* ``ot_boundary_*.npz``   reference ``losses.loss_libs.kd_loss_2d`` (the real one) driving the restated
                          geomloss ``SamplesLoss`` in fp32 (+ the fp64 analytic oracle) at the OT boundary;
* ``kd_pose_loss.npz``    reference ``losses.kd_loss.KDPoseLoss.__call__`` end to end (seam B0) on synthetic
                          head outputs / targets: its ``prepare_targets`` outputs, the three losses and the
                          gradients w.r.t. every head-output level;
* ``postprocess_kd.npz``  reference ``postprocess.postprocess_kd.PostProcessorKD`` (with the real
                          ``cv2.solvePnPRansac``) on synthetic teacher head outputs: the teacher dict fed
                          to the loss plus, per image, the selected cells recovered from its outputs.

geomloss itself is NOT available (parity unpinned, see ``oracle/__init__.py``): the ``SamplesLoss`` plugged
into the reference modules is ``oracle/geomloss_ref.py``.
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import ref_loader, sinkhorn_analytic  # noqa: E402
from kd_6d_pose_adlp_b200.synthetic import cu_seqlens, ot_batch  # noqa: E402
from tests.scenario import (ANCHOR_SIZES, ANCHOR_STRIDES, CFG_KD, INTERNAL_K, MESH_DIAMETERS, make_head_outputs,  # noqa: E402
                            make_target_arrays)

ref = ref_loader.load()
torch.backends.cuda.matmul.allow_tf32 = False


def digest(arrays):
    import hashlib

    h = hashlib.sha256()
    for a in arrays:
        h.update(np.ascontiguousarray(a).tobytes())
    return h.hexdigest()


def save(name, **arrays):
    path = os.path.join(HERE, name)
    np.savez_compressed(path, **arrays)
    print("wrote", path, "%.1f KiB" % (os.path.getsize(path) / 1024))


# ---------------------------------------------------------------------------------------------------------
# 1. OT boundary: the reference's kd_loss_2d
# ---------------------------------------------------------------------------------------------------------
def golden_ot_boundary(tag, batch, blur=0.001, reach=0.5, scaling=0.5, weighted=True):
    from oracle import geomloss_ref

    xs = torch.tensor(batch["xs"].reshape(-1, 2), requires_grad=True)
    ws = torch.tensor(batch["ws"], requires_grad=True)
    xt = torch.tensor(batch["xt"].reshape(-1, 2))
    wt = torch.tensor(batch["wt"])
    L = geomloss_ref.SamplesLoss("sinkhorn", p=2.0, blur=blur, scaling=scaling, reach=reach)
    work = xs.clone()
    losses = ref.kd_loss_2d(work, xt, ws if weighted else None, wt if weighted else None, 640, 480, "point", L, dim=2,
                            pos_per_img=batch["pos_per_img"], pos_per_img_t=batch["pos_per_img_t"])
    sum(losses).backward()
    keep = [i for i, (n, m) in enumerate(zip(batch["pos_per_img"], batch["pos_per_img_t"])) if n > 0 and m > 0]
    loss32 = np.zeros(len(batch["pos_per_img"]), np.float32)
    loss32[keep] = torch.stack(losses).detach().numpy()
    o64 = sinkhorn_analytic.kdot_fwd_bwd_f64(batch["xs"], batch["ws"] if weighted else None, batch["xt"],
                                             batch["wt"] if weighted else None, cu_seqlens(batch["pos_per_img"]),
                                             cu_seqlens(batch["pos_per_img_t"]), 8, 2, blur=blur, reach=reach,
                                             scaling=scaling)
    save(f"ot_boundary_{tag}.npz", xs=batch["xs"], ws=batch["ws"], xt=batch["xt"], wt=batch["wt"],
         pos_per_img=np.asarray(batch["pos_per_img"]), pos_per_img_t=np.asarray(batch["pos_per_img_t"]),
         blur=blur, reach=-1.0 if reach is None else reach, scaling=scaling, weighted=weighted,
         ref32_loss=loss32, ref32_grad_xs=xs.grad.numpy().reshape(-1, 8, 2),
         ref32_grad_ws=ws.grad.numpy() if ws.grad is not None else np.zeros_like(batch["ws"]),
         ref32_xs_norm=work.detach().numpy().reshape(-1, 8, 2), ref32_xt_norm=xt.numpy().reshape(-1, 8, 2),
         ref64_loss=o64["loss_per_img"], ref64_grad_xs=o64["grad_xs"], ref64_grad_ws=o64["grad_ws"],
         nits=o64["nits"], valid=o64["valid"])


# ---------------------------------------------------------------------------------------------------------
# shared synthetic scene in the reference's own types
# ---------------------------------------------------------------------------------------------------------
def build_targets(nimg, seed):
    arrs = make_target_arrays(nimg, seed)
    targets = []
    for i in range(nimg):
        t = ref.PoseAnnot(torch.tensor(arrs["keypoints_3d"]), torch.tensor(arrs["K"]), torch.tensor(arrs["mask"][i]),
                          torch.tensor(arrs["class_ids"][i]), torch.tensor(arrs["rotations"][i]),
                          torch.tensor(arrs["translations"][i]), 256, 256,
                          bbox_scale=torch.tensor(1.0), bbox_trans=torch.tensor(arrs["bbox_trans"][i]))
        targets.append(t)
    return targets, arrs


def ref_anchors(nimg, level_hw):
    gen = ref.modules["model"].make_anchor_generator_atss(ANCHOR_SIZES[:len(level_hw)], ANCHOR_STRIDES[:len(level_hw)])

    class _IL:
        sizes = [(256, 256)] * nimg

    feats = [torch.zeros(nimg, 1, h, w) for h, w in level_hw]
    return gen(_IL(), feats)


def sparse(prefix, grads):
    """Gradients w.r.t. head outputs are non-zero on ~10 cells per image: store (flat index, value) per level."""
    out = {}
    for l, g in enumerate(grads):
        idx = np.flatnonzero(g)
        out[f"{prefix}_{l}_idx"] = idx.astype(np.int64)
        out[f"{prefix}_{l}_val"] = g.reshape(-1)[idx]
    return out


# ---------------------------------------------------------------------------------------------------------
# 2 + 3. PostProcessorKD on the teacher, then KDPoseLoss on the student
# ---------------------------------------------------------------------------------------------------------
def golden_postprocess_and_loss(nimg=6, seed=0):
    torch.manual_seed(seed)
    np.random.seed(seed)
    targets, tarr = build_targets(nimg, seed)
    t_hw = [(32, 32), (16, 16), (8, 8), (4, 4), (2, 2)]
    s_hw = t_hw[:4]
    t_cls, t_reg = make_head_outputs(nimg, t_hw, seed + 100, teacher=True)
    s_cls, s_reg = make_head_outputs(nimg, s_hw, seed + 200, teacher=False)

    coder = ref.TargetCoder("POINT", ANCHOR_SIZES, ANCHOR_STRIDES, target_type="3D")
    pp = ref.PostProcessorKD(0.1, coder, 10, 1.0, {})
    anchors_t = ref_anchors(nimg, t_hw)
    with torch.no_grad():
        res = pp([torch.tensor(a) for a in t_cls], [torch.tensor(a) for a in t_reg], targets, anchors_t)
    post_kp_cls = torch.cat(res[0], dim=0)
    post_kp_2d = torch.cat(res[3], dim=0)
    pos_t = [len(r) for r in res[0]]
    print("teacher cells per image:", pos_t)
    # recover which (level, cell) each selected row came from: scores are sqrt(sigmoid(logit)) of a unique cell
    sel_level, sel_loc = [], []
    for i in range(nimg):
        sc = res[0][i][:, 0].numpy() if pos_t[i] else np.zeros(0, np.float32)
        lv_i, loc_i = [], []
        for s in sc:
            found = None
            for lv, a in enumerate(t_cls):
                cand = torch.sqrt(torch.sigmoid(torch.tensor(a[i, 0].reshape(-1)))).numpy()
                hit = np.nonzero(cand == s)[0]
                if len(hit):
                    assert found is None and len(hit) == 1, "ambiguous score"
                    found = (lv, int(hit[0]))
            assert found is not None
            lv_i.append(found[0])
            loc_i.append(found[1])
        sel_level.append(np.asarray(lv_i, np.int32))
        sel_loc.append(np.asarray(loc_i, np.int32))
    # head outputs are NOT stored (7.5 MB): tests regenerate them from tests/scenario.py (seeded numpy) and
    # verify this digest first
    save("postprocess_kd.npz", nimg=nimg, seed=seed, inputs_sha256=digest(t_cls + t_reg),
         post_kp_cls=post_kp_cls.numpy(), post_kp_2d=post_kp_2d.numpy(), post_pos_per_img=np.asarray(pos_t),
         sel_level=np.concatenate(sel_level), sel_loc=np.concatenate(sel_loc),
         anchors_lvl0=anchors_t[0][0].bbox.numpy(), anchors_lvl4=anchors_t[0][4].bbox.numpy(),
         **{k: v for k, v in tarr.items()})

    # ---- student loss through the reference's KDPoseLoss ----
    pred_t = {"post_kp_2d": post_kp_2d.clone(), "post_kp_cls": post_kp_cls.clone(), "post_pos_per_img": pos_t}
    cfg_kd = dict(CFG_KD)
    loss_fn = ref.KDPoseLoss(2.0, 0.25, ANCHOR_SIZES, ANCHOR_STRIDES, "SSC", 10, 1.0, 9, INTERNAL_K, MESH_DIAMETERS,
                             ref.TargetCoder("POINT", ANCHOR_SIZES, ANCHOR_STRIDES, target_type="3D"), cfg_kd)
    loss_fn.step = 0
    loss_fn.vis_dir = "/tmp/kdot_vis"
    anchors_s = ref_anchors(nimg, s_hw)
    # record what the (random, host-side) target assignment produced so the test double can replay it
    torch.manual_seed(seed + 7)
    prep = loss_fn.prepare_targets(targets, anchors_s)
    torch.manual_seed(seed + 7)
    pc = [torch.tensor(a, requires_grad=True) for a in s_cls]
    pr = [torch.tensor(a, requires_grad=True) for a in s_reg]
    cls_loss, reg_loss, kd_loss = loss_fn(pc, pr, targets, anchors_s, pred_t)
    print("reference losses: cls %.6f reg %.6f kd %.6f" % (float(cls_loss), float(reg_loss), float(kd_loss)),
          "student cells per image:", loss_fn.pos_per_img)
    g_kd = torch.autograd.grad(kd_loss, pc + pr, retain_graph=True, allow_unused=True)
    g_all = torch.autograd.grad(cls_loss * 0.1 + reg_loss + 5.0 * kd_loss, pc + pr, allow_unused=True)
    z = lambda g, like: np.zeros_like(like) if g is None else g.numpy()
    save("kd_pose_loss.npz", nimg=nimg, seed=seed, inputs_sha256=digest(s_cls + s_reg),
         labels=torch.cat(prep[0]).numpy(), reg_targets=torch.cat(prep[1]).numpy(),
         aux_raw_boxes=torch.cat(prep[2]).numpy(), aux_3d=torch.cat(prep[3]).numpy(),
         aux_bbox_trans=torch.cat(prep[4]).numpy(), cells_per_img=np.asarray([len(p) for p in prep[0]]),
         post_kp_2d=post_kp_2d.numpy(), post_kp_cls=post_kp_cls.numpy(), post_pos_per_img=np.asarray(pos_t),
         pos_per_img=np.asarray(loss_fn.pos_per_img),
         cls_loss=float(cls_loss), reg_loss=float(reg_loss), kd_loss=float(kd_loss),
         **sparse("gkd_cls", [z(g_kd[l], s_cls[l]) for l in range(4)]),
         **sparse("gkd_reg", [z(g_kd[4 + l], s_reg[l]) for l in range(4)]),
         **sparse("gall_reg", [z(g_all[4 + l], s_reg[l]) for l in range(4)]),
         gall_cls_sum=np.asarray([float(z(g_all[l], s_cls[l]).astype(np.float64).sum()) for l in range(4)]),
         gall_cls_abs=np.asarray([float(np.abs(z(g_all[l], s_cls[l])).astype(np.float64).sum()) for l in range(4)]))


# ---------------------------------------------------------------------------------------------------------
# 4. hard selection scene: every label of PostProcessorKD.pose_infer_ml and the student-eval PostProcessor
# ---------------------------------------------------------------------------------------------------------
def golden_postprocess_hard(nimg=64, seed=5):
    """Reference ``PostProcessorKD`` (all candidate labels, not only the first) and reference ``PostProcessor``
    (``postprocess/postprocess.py``, with its ``clsId in target.class_ids`` filter) on the hard scene of
    ``tests/scenario.py``: batch 64, 2-3 live classes per image, object scales 0.7x..2.6x, exact duplicate logits and
    logits within 2 ulps of the 0.1 threshold.  Stored per (image, label): candidate counts per level, selected
    (level, cell) sequence, scores; per image the student-eval results (class, score, R, T)."""
    import importlib

    from tests.scenario import hard_live_classes, make_hard_scene

    torch.manual_seed(seed)
    np.random.seed(seed)
    t_hw = [(32, 32), (16, 16), (8, 8), (4, 4), (2, 2)]
    tarr, t_cls, t_reg = make_hard_scene(nimg, t_hw, seed)
    targets = []
    for i in range(nimg):
        ids = [c for c in hard_live_classes(i) if c in (0, 3)]
        t = ref.PoseAnnot(torch.tensor(tarr["keypoints_3d"]), torch.tensor(tarr["K"]), torch.tensor(tarr["mask"][i]),
                          torch.tensor(ids, dtype=torch.int64), torch.tensor(np.repeat(tarr["rotations"][i], len(ids), 0)),
                          torch.tensor(np.repeat(tarr["translations"][i], len(ids), 0)), 256, 256,
                          bbox_scale=torch.tensor(1.0), bbox_trans=torch.tensor(tarr["bbox_trans"][i]))
        targets.append(t)
    coder = ref.TargetCoder("POINT", ANCHOR_SIZES, ANCHOR_STRIDES, target_type="3D")
    pp = ref.PostProcessorKD(0.1, coder, 10, 1.0, {})
    anchors_t = ref_anchors(nimg, t_hw)
    cls_t = [torch.tensor(a) for a in t_cls]
    reg_t = [torch.tensor(a) for a in t_reg]
    with torch.no_grad():
        sampled = [pp.forward_for_single_feature_map(o, b, a) for o, b, a in zip(cls_t, reg_t, list(zip(*anchors_t)))]
        per_img = list(zip(*sampled))
        kd_first = pp(cls_t, reg_t, targets, anchors_t)
        pp_eval = importlib.import_module("postprocess.postprocess").PostProcessor(0.1, coder, 10, 1.0, {})
        ev = pp_eval(cls_t, reg_t, targets, anchors_t)

    # decode of EVERY cell of every level for a class, un-cropped, to recover which cells the reference selected
    def all_cells_xy(i, c):
        out = []
        bt = tarr["bbox_trans"][i].astype(np.float64)
        Ainv = np.linalg.inv(bt[:, :2])
        for lv, (h, w) in enumerate(t_hw):
            stride, size = ANCHOR_STRIDES[lv], ANCHOR_SIZES[lv]
            cy, cx = np.meshgrid(np.arange(h) * stride + stride / 2, np.arange(w) * stride + stride / 2, indexing="ij")
            pts = []
            for k in range(8):
                x = t_reg[lv][i, 16 * c + k].astype(np.float64) * size + cx
                y = t_reg[lv][i, 16 * c + 8 + k].astype(np.float64) * size + cy
                pts.append((Ainv @ (np.stack([x.reshape(-1), y.reshape(-1)]) - bt[:, 2:3])).T)
            out.append(np.concatenate(pts, axis=1))  # (cells, 16): x0 y0 x1 y1 ...
        return out

    rec = dict(img=[], label=[], count=[], valid=[], level=[], loc=[], score=[])
    n_dup_ties = 0
    for i in range(nimg):
        preds = per_img[i]
        with torch.no_grad():
            res = pp.pose_infer_ml(preds, targets[i])
        for scores, cls_id, R, T, xy2d in res:
            cells = all_cells_xy(i, cls_id)
            valid = [0 if p is None else int((p[1] == cls_id + 1).sum()) for p in preds]
            lv_seq, loc_seq = [], []
            for row in xy2d.reshape(len(xy2d), 16).numpy().astype(np.float64):
                hits = [(lv, int(k)) for lv in range(len(t_hw)) for k in np.flatnonzero(np.abs(cells[lv] - row).max(axis=1) < 2e-2)]
                assert len(hits) == 1, ("ambiguous / missing cell", i, cls_id, hits)
                lv_seq.append(hits[0][0])
                loc_seq.append(hits[0][1])
            sc = scores[:, 0].numpy()
            n_dup_ties += int(len(np.unique(sc)) < len(sc))
            rec["img"].append(i); rec["label"].append(cls_id); rec["count"].append(len(lv_seq)); rec["valid"].append(valid)
            rec["level"].append(lv_seq); rec["loc"].append(loc_seq); rec["score"].append(sc)
    print("hard scene: %d (image, label) results, %d with tied scores among the selected cells" % (len(rec["img"]), n_dup_ties))
    ev_img, ev_cls, ev_score, ev_R, ev_T = [], [], [], [], []
    for i, lst in enumerate(ev):
        for score, cls_id, R, T, _xy in lst:
            ev_img.append(i); ev_cls.append(cls_id); ev_score.append(score); ev_R.append(R); ev_T.append(T.reshape(3))
    first_cnt = [len(r) for r in kd_first[0]]
    save("postprocess_hard.npz", nimg=nimg, seed=seed, inputs_sha256=digest(t_cls + t_reg),
         bbox_trans=tarr["bbox_trans"], K=tarr["K"], keypoints_3d=tarr["keypoints_3d"],
         img=np.asarray(rec["img"]), label=np.asarray(rec["label"]), count=np.asarray(rec["count"]),
         valid=np.asarray(rec["valid"]), level=np.concatenate([np.asarray(v, np.int32) for v in rec["level"]]),
         loc=np.concatenate([np.asarray(v, np.int32) for v in rec["loc"]]), score=np.concatenate(rec["score"]),
         first_label_count=np.asarray(first_cnt),
         ev_img=np.asarray(ev_img), ev_cls=np.asarray(ev_cls), ev_score=np.asarray(ev_score, np.float64),
         ev_R=np.asarray(ev_R, np.float64), ev_T=np.asarray(ev_T, np.float64))


if __name__ == "__main__":
    if "--hard-only" in sys.argv:
        golden_postprocess_hard()
        raise SystemExit(0)
    golden_ot_boundary("ape_b8", ot_batch(8, seed=0))
    golden_ot_boundary("ape_b8_tight", ot_batch(8, seed=1, sigma=0.005))
    golden_ot_boundary("balanced", ot_batch(4, seed=2), reach=None)
    golden_ot_boundary("unweighted", ot_batch(4, seed=3), weighted=False)
    golden_ot_boundary("mid", ot_batch(2, seed=4, n_range=(40, 60), m_range=(50, 70), p_empty_teacher=0.0), scaling=0.7)
    golden_postprocess_and_loss()
    golden_postprocess_hard()
