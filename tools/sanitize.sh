#!/bin/bash
# compute-sanitizer passes over every kernel family (run on the GPU box); summaries -> gpurun_out/
R=${1:-r02}
mkdir -p gpurun_out
cat > /tmp/kdot_sanitize_drive.py <<'PY'
import os, sys, numpy as np, torch
sys.path.insert(0, os.getcwd())
from kd_6d_pose_adlp_b200.ops import OTConfig, ot_loss_batched
from kd_6d_pose_adlp_b200.synthetic import ot_batch
from kd_6d_pose_adlp_b200 import SamplesLoss
dev = torch.device("cuda:0")
def run(b, cfg=OTConfig(), **kw):
    t = {k: torch.from_numpy(b[k]).to(dev) for k in ("xs", "ws", "xt", "wt")}
    o = ot_loss_batched(t["xs"], t["ws"], t["xt"], t["wt"], b["pos_per_img"], b["pos_per_img_t"], cfg, **kw)
    torch.cuda.synchronize(); return float(o["loss_per_img"].sum())
print("small fast", run(ot_batch(6, seed=1, p_empty_teacher=0.3)))
print("tiled (33..256 points)", run(ot_batch(3, seed=2, n_range=(20, 30), m_range=(20, 30))))
print("stream", run(ot_batch(3, seed=3, n_range=(140, 190), m_range=(140, 190), p_empty_teacher=0.3)))
print("tiled, empty teachers", run(ot_batch(4, seed=4, n_range=(40, 90), m_range=(40, 90), p_empty_teacher=0.5)))
print("mmd", run(ot_batch(3, seed=5), OTConfig(loss="energy", blur=0.05)))
print("p1", run(ot_batch(2, seed=6), OTConfig(p=1.0, blur=0.01)))
x = torch.sigmoid(torch.randn(1, 60, 16)).to(dev); y = torch.sigmoid(torch.randn(1, 50, 16)).to(dev)
print("stream D=16", float(SamplesLoss("sinkhorn", p=2, blur=0.05)(x, y).sum()))
from kd_6d_pose_adlp_b200.postprocess.postprocess_kd import select_cells
cls = [torch.randn(2, 15, h, h, device=dev) - 2 for h in (32, 16, 8, 4, 2)]
reg = [torch.randn(2, 240, h, h, device=dev) * 0.3 for h in (32, 16, 8, 4, 2)]
s = select_cells(cls, reg, [32, 64, 128, 256, 512], [8, 16, 32, 64, 128], 0.1, 10, 1.0); torch.cuda.synchronize()
print("select", int(s["sel_count"].sum()))
# fused dense losses + device target assignment
import types
from kd_6d_pose_adlp_b200.ops import FocalLossFunction, Reg3dLossFunction
from kd_6d_pose_adlp_b200.targets import ssc_assign
from kd_6d_pose_adlp_b200.target_coder import grid_anchors
from tests import scenario
HW = [(32, 32), (16, 16), (8, 8), (4, 4)]
lab = torch.randint(-1, 3, (2 * 1360,), device=dev)
pc = [torch.randn(2, 15, h, w, device=dev, requires_grad=True) for h, w in HW]
l = FocalLossFunction.apply(lab, 2.0, 0.25, *pc); l.backward(); torch.cuda.synchronize(); print("focal", float(l))
xy = (torch.rand(40, 2, device=dev) * 400).requires_grad_(True)
r = Reg3dLossFunction.apply(xy, torch.randn(5, 8, 3, device=dev) * 30 + torch.tensor([0., 0., 900.], device=dev), torch.full((5,), 100.0, device=dev), np.linalg.inv(np.asarray(scenario.INTERNAL_K).reshape(3, 3)).reshape(-1).tolist())
r.sum().backward(); torch.cuda.synchronize(); print("reg3d", float(r.sum()))
arr = scenario.make_target_arrays(3, 0); tt = lambda a: torch.tensor(a).to(dev)
tg = [types.SimpleNamespace(keypoints_3d=tt(arr["keypoints_3d"]), K=tt(arr["K"]), mask=tt(arr["mask"][i]), class_ids=tt(arr["class_ids"][i]), rotations=tt(arr["rotations"][i]), translations=tt(arr["translations"][i]), bbox_trans=tt(arr["bbox_trans"][i])) for i in range(3)]
an = torch.cat(grid_anchors(HW, scenario.ANCHOR_SIZES, scenario.ANCHOR_STRIDES, device=dev))
for mode in ("parity", "philox"):
    res = ssc_assign(tg, an, [h * w for h, w in HW], scenario.ANCHOR_SIZES, 10, 1.0, mode=mode); torch.cuda.synchronize(); print("ssc", mode, res["npos"].tolist())
PY
for tool in memcheck racecheck synccheck; do
  timeout 900 compute-sanitizer --tool $tool --print-limit 5 python /tmp/kdot_sanitize_drive.py > gpurun_out/${R}_sanitizer_${tool}.txt 2>&1
  echo "== $tool: $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' gpurun_out/${R}_sanitizer_${tool}.txt | tail -1)"
done
