import sys, os
sys.path.insert(0, os.getcwd())
import numpy as np, torch
from kd_6d_pose_adlp_b200 import _lib
from kd_6d_pose_adlp_b200.ops import cu_seqlens
from kd_6d_pose_adlp_b200.synthetic import ot_batch
L = _lib.lib(); dev = torch.device("cuda:0")
nimg = 64
b = ot_batch(nimg, seed=1)
t = lambda a: torch.from_numpy(a).to(dev)
xs0, xt0, ws, wt = t(b["xs"]), t(b["xt"]), t(b["ws"]), t(b["wt"])
cn, cm = cu_seqlens(b["pos_per_img"], dev), cu_seqlens(b["pos_per_img_t"], dev)
loss = torch.empty(nimg, device=dev); valid = torch.empty(nimg, dtype=torch.int32, device=dev)
nits = torch.empty(nimg, dtype=torch.int32, device=dev); gx = torch.empty_like(xs0); gw = torch.empty_like(ws)
clk = torch.zeros(nimg + 8, 16, dtype=torch.int64, device=dev)
L.kdot_debug_set_clock_buffer(clk.data_ptr())
for i in range(3):
    xs, xt = xs0.clone(), xt0.clone()
    rc = L.kdot_sinkhorn_fwd_bwd(xs.data_ptr(), ws.data_ptr(), xt.data_ptr(), wt.data_ptr(), cn.data_ptr(), cm.data_ptr(), nimg, 8, 2, 12, 12, 0, 2.0, 0.001, 0.5, 0.5, 640.0, 480.0, 1, loss.data_ptr(), None, valid.data_ptr(), gx.data_ptr(), gw.data_ptr(), nits.data_ptr(), None, 0, torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
c = clk.cpu().numpy()
print("image0 N,M", b["pos_per_img"][0], b["pos_per_img_t"][0], "nits", int(nits[0]))
print("phase stamps img0:", np.diff(c[0, :7]))
r = c[64:66].reshape(-1)[:20]
print("round stamps:", np.diff(np.concatenate([[c[0, 4]], r[r > 0]])))
print("hmag x1000:", c[66:68].reshape(-1)[:16])
