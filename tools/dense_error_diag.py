import sys, os
sys.path.insert(0, os.getcwd())
import numpy as np, torch
from kd_6d_pose_adlp_b200.ops import OTConfig, ot_loss_batched
from kd_6d_pose_adlp_b200.synthetic import cu_seqlens, ot_batch
from oracle import sinkhorn_analytic
dev = torch.device("cuda:0")
NM = int(os.environ.get("NM", "1360"))
b = ot_batch(nimg=1, seed=5, dense=(NM, NM + 4), sigma=0.1)
xs, xt = torch.from_numpy(b["xs"]).to(dev), torch.from_numpy(b["xt"]).to(dev)
ws, wt = torch.from_numpy(b["ws"]).to(dev), torch.from_numpy(b["wt"]).to(dev)
out = ot_loss_batched(xs, ws, xt, wt, b["pos_per_img"], b["pos_per_img_t"], OTConfig())
torch.cuda.synchronize()
refp = "/tmp/ref_dense_%d.npz" % NM
if os.path.exists(refp):
    ref = dict(np.load(refp))
else:
    ref = sinkhorn_analytic.kdot_fwd_bwd_f64(b["xs"], b["ws"], b["xt"], b["wt"], cu_seqlens(b["pos_per_img"]), cu_seqlens(b["pos_per_img_t"]), 8, 2)
    np.savez(refp, **ref)
g = out["grad_xs"].cpu().numpy().astype(np.float64); r = ref["grad_xs"]
den = np.abs(r).max()
err = np.abs(g - r) / den
print(os.environ.get("KDOT_LIB", "default")[-20:], os.environ.get("KDOT_FORCE_PATH", "auto"), "max %.2e" % err.max(), "quantiles 50/90/99/99.9: " + " ".join("%.1e" % np.quantile(err, q) for q in (0.5, 0.9, 0.99, 0.999)),
      "count>1e-4:", int((err > 1e-4).sum()), "of", err.size, "per slot max:", ["%.1e" % err[:, s].max() for s in range(8)])
gw = out["grad_ws"].cpu().numpy(); print("  d/dalpha max rel %.2e" % (np.abs(gw - ref["grad_ws"]).max() / np.abs(ref["grad_ws"]).max()))
i, s, d = np.unravel_index(np.argmax(err), err.shape)
print("  worst cell", i, "slot", s, "dim", d, "got", g[i, s], "ref", r[i, s], "w", b["ws"][i, s])
