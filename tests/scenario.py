"""Seeded synthetic LINEMOD-ape shaped scene at the head-output boundary.

Pure numpy, no reference imports: used by ``tests/golden/make_golden.py`` (which wraps the arrays in the
reference's own ``PoseAnnot`` etc.) and by the GPU tests (which replay the stored target assignment).
One object of class 0 per image, 256x256 dynamic-zoom-in crop of a 640x480 frame; student = 4 FPN levels
(32,16,8,4 -> 1360 cells), teacher = 5 levels (+2x2 -> 1364 cells); 15 classes, 16 regression channels per
class ([dx0..dx7, dy0..dy7] in units of the anchor size).
"""
import numpy as np

ANCHOR_SIZES = [32, 64, 128, 256, 512]          # configs/ape.yaml:3
ANCHOR_STRIDES = [8, 16, 32, 64, 128]           # configs/ape.yaml:4
INTERNAL_K = [572.4114, 0, 325.2611, 0, 573.57043, 242.04899, 0, 0, 1]   # configs/ape.yaml:20
MESH_DIAMETERS = [104.26, 250.85, 167.49, 177.43, 204.83, 154.63, 129.85, 264.12, 110.83, 164.65, 178.35,
                  145.61, 279.04, 287.24, 213.25]
N_CLASS = 15
CFG_KD = dict(GTYPE="sinkhorn", GP=2.0, GBLUR=0.001, SCALING=0.5, REACH=0.5, WEIGHTED_OT=True, DETACH=False,
              GnD=2, GLEVEL="point", LEVEL="pred", LOSS_WEIGHT_KD=5.0)
HALF_EXTENT = np.array([38.0, 39.0, 46.0], np.float32)  # ape-sized box, mm


def _rodrigues(v):
    th = np.linalg.norm(v)
    if th < 1e-12:
        return np.eye(3)
    k = v / th
    K = np.array([[0, -k[2], k[1]], [k[2], 0, -k[0]], [-k[1], k[0], 0]])
    return np.eye(3) + np.sin(th) * K + (1 - np.cos(th)) * K @ K


def make_target_arrays(nimg, seed, scale_range=(1.4, 1.8)):
    rng = np.random.default_rng(seed)
    corners = np.array([[sx, sy, sz] for sx in (-1, 1) for sy in (-1, 1) for sz in (-1, 1)], np.float32) * HALF_EXTENT
    kp3d = np.tile(corners[None], (N_CLASS, 1, 1)).astype(np.float32)
    K = np.asarray(INTERNAL_K, np.float32).reshape(3, 3)
    out = dict(keypoints_3d=kp3d, K=K, mask=[], class_ids=[], rotations=[], translations=[], bbox_trans=[], kp2d_crop=[])
    for _ in range(nimg):
        u, v = rng.uniform(250, 400), rng.uniform(180, 300)
        z = rng.uniform(800, 1000)
        T = (np.linalg.inv(K.astype(np.float64)) @ np.array([u, v, 1.0]) * z).reshape(3, 1)
        R = _rodrigues(rng.normal(0, 0.4, 3))
        s = rng.uniform(*scale_range)
        bt = np.array([[s, 0, 128 - s * u], [0, s, 128 - s * v]], np.float32)
        cam = K.astype(np.float64) @ (R @ corners.T.astype(np.float64) + T)
        uv = cam[:2] / cam[2]
        crop = bt.astype(np.float64) @ np.vstack([uv, np.ones(8)])
        mask = np.zeros((256, 256), np.float32)
        x0, x1 = int(max(crop[0].min(), 0)), int(min(crop[0].max(), 255))
        y0, y1 = int(max(crop[1].min(), 0)), int(min(crop[1].max(), 255))
        mask[y0:y1 + 1, x0:x1 + 1] = 1.0
        out["mask"].append(mask)
        out["class_ids"].append(np.array([0], np.int64))
        out["rotations"].append(R[None].astype(np.float32))
        out["translations"].append(T[None].astype(np.float32))
        out["bbox_trans"].append(bt)
        out["kp2d_crop"].append(crop.astype(np.float32))  # (2, 8)
    for k in ("mask", "class_ids", "rotations", "translations", "bbox_trans", "kp2d_crop"):
        out[k] = np.stack(out[k])
    return out


def make_head_outputs(nimg, level_hw, seed, teacher, target_seed=0, tarr=None):
    """Per level ``(nimg, 15, H, W)`` class logits and ``(nimg, 240, H, W)`` keypoint offsets.

    Class-0 logits are high inside the object mask (so the teacher's 0.1 confidence threshold selects cells
    there: truly random-init heads never pass it, SURVEY.md 8(d) pitfall), other classes ~ -8.  Offsets of
    class 0 are the encoded true keypoints plus noise (teacher: small, student: larger), so that RANSAC-PnP in
    the reference post-processor succeeds and the OT problem looks like real distillation data."""
    rng = np.random.default_rng(seed)
    tarr = make_target_arrays(nimg, target_seed) if tarr is None else tarr
    noise = 0.03 if teacher else 0.12
    cls_l, reg_l = [], []
    for lv, (h, w) in enumerate(level_hw):
        stride, size = ANCHOR_STRIDES[lv], ANCHOR_SIZES[lv]
        cy, cx = np.meshgrid(np.arange(h) * stride + stride / 2, np.arange(w) * stride + stride / 2, indexing="ij")
        cls = np.full((nimg, N_CLASS, h, w), -8.0, np.float32) + rng.normal(0, 0.2, (nimg, N_CLASS, h, w)).astype(np.float32)
        reg = rng.normal(0, 0.3, (nimg, N_CLASS * 16, h, w)).astype(np.float32)
        for i in range(nimg):
            m = tarr["mask"][i][np.clip(cy.astype(int), 0, 255), np.clip(cx.astype(int), 0, 255)] > 0
            cls[i, 0] = np.where(m, rng.normal(0.5, 1.0, (h, w)), rng.normal(-4.0, 1.0, (h, w))).astype(np.float32)
            kp = tarr["kp2d_crop"][i]  # (2, 8) crop pixels
            for k in range(8):
                reg[i, k] = ((kp[0, k] - cx) / size + rng.normal(0, noise, (h, w))).astype(np.float32)
                reg[i, 8 + k] = ((kp[1, k] - cy) / size + rng.normal(0, noise, (h, w))).astype(np.float32)
        cls_l.append(cls)
        reg_l.append(reg)
    return cls_l, reg_l


# ---------------------------------------------------------------------------------------------------------
# hard selection scene (tests/golden/postprocess_hard.npz): several live classes per image, object scales from 0.7x to
# 2.6x (every per-level budget pattern), exact duplicate logits (arg-max / top-k ties) and logits within a few ulps of
# the 0.1 confidence threshold
# ---------------------------------------------------------------------------------------------------------
HARD_EXTRA_CLASSES = (3, 7, 5)
THRESHOLD_LOGIT = np.float32(np.log(0.1 / 0.9))  # sigmoid(x) == 0.1 up to rounding


def hard_live_classes(i):
    """Classes with cells above the threshold in image i (class 0 always; PostProcessor's target.class_ids filter keeps
    only classes 0 and 3)."""
    return [0, HARD_EXTRA_CLASSES[i % 2]] + ([HARD_EXTRA_CLASSES[2]] if i % 3 == 0 else [])


def make_hard_scene(nimg, level_hw, seed):
    tarr = make_target_arrays(nimg, seed, scale_range=(0.7, 2.6))
    cls_l, reg_l = make_head_outputs(nimg, level_hw, seed + 100, teacher=True, tarr=tarr)
    rng = np.random.default_rng(seed + 999)
    for lv, (h, w) in enumerate(level_hw):
        stride, size = ANCHOR_STRIDES[lv], ANCHOR_SIZES[lv]
        cy, cx = np.meshgrid(np.arange(h) * stride + stride / 2, np.arange(w) * stride + stride / 2, indexing="ij")
        for i in range(nimg):
            m = tarr["mask"][i][np.clip(cy.astype(int), 0, 255), np.clip(cx.astype(int), 0, 255)] > 0
            kp = tarr["kp2d_crop"][i]
            for c in hard_live_classes(i)[1:]:
                cls_l[lv][i, c] = np.where(m, rng.normal(0.3, 1.0, (h, w)), rng.normal(-4.0, 1.0, (h, w))).astype(np.float32)
                for k in range(8):
                    reg_l[lv][i, 16 * c + k] = ((kp[0, k] - cx) / size + rng.normal(0, 0.03, (h, w))).astype(np.float32)
                    reg_l[lv][i, 16 * c + 8 + k] = ((kp[1, k] - cy) / size + rng.normal(0, 0.03, (h, w))).astype(np.float32)
            if lv <= 1:
                flat = cls_l[lv][i, 0].reshape(-1)           # view: edits land in cls_l
                inside = np.flatnonzero(m.reshape(-1))
                outside = np.flatnonzero(~m.reshape(-1))
                if len(inside) >= 12:
                    a = inside[np.argmax(flat[inside])]
                    others = inside[inside != a]
                    pick = rng.choice(others, size=11, replace=False)
                    if i % 4 == 1:
                        flat[pick[0]] = flat[a]              # duplicate of the level's maximum (arg-max tie), every 4th image
                    for q in range(5):                        # five more exact duplicates among the candidates (top-k ties)
                        flat[pick[1 + 2 * q]] = flat[pick[2 + 2 * q]]
                if lv == 0 and len(outside) >= 6:
                    edge = rng.choice(outside, size=6, replace=False)
                    v = THRESHOLD_LOGIT
                    vals = [np.nextafter(np.nextafter(v, np.float32(-np.inf)), np.float32(-np.inf)), np.nextafter(v, np.float32(-np.inf)), v,
                            np.nextafter(v, np.float32(np.inf)), np.nextafter(np.nextafter(v, np.float32(np.inf)), np.float32(np.inf)),
                            np.float32(v + 1e-5)]
                    for e, val in zip(edge, vals):
                        flat[e] = val
    return tarr, cls_l, reg_l
