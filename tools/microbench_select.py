"""Selection kernel alone: single-launch event time (queued behind a GPU-side delay) and steady-state time per launch of
a back-to-back train, batch 64 of the ape-shaped teacher head (tests/scenario.py)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from tests import scenario
from kd_6d_pose_adlp_b200.postprocess.postprocess_kd import select_cells

dev = torch.device("cuda:0")
hw = [(32, 32), (16, 16), (8, 8), (4, 4)]
nimg = int(sys.argv[1]) if len(sys.argv) > 1 else 64
t_cls, t_reg = scenario.make_head_outputs(nimg, hw, 200, teacher=True, target_seed=0)
tc = [torch.from_numpy(a).to(dev) for a in t_cls]
tr = [torch.from_numpy(a).to(dev) for a in t_reg]
run = lambda: select_cells(tc, tr, scenario.ANCHOR_SIZES, scenario.ANCHOR_STRIDES, 0.1, 10, 1.0)
run(); torch.cuda.synchronize()
one = []
for _ in range(40):
    torch.cuda._sleep(400000)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); run(); e1.record(); e1.synchronize()
    one.append(e0.elapsed_time(e1) * 1e3)
torch.cuda._sleep(20000000)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(100): run()
e1.record(); e1.synchronize()
print("%s: single launch %.2f us (min %.2f), back-to-back %.2f us per launch" % (os.environ.get("KDOT_LIB", "default")[-20:], np.median(one[5:]), min(one), e0.elapsed_time(e1) * 10))
