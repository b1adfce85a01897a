import sys, os
sys.path.insert(0, os.getcwd())
import numpy as np, torch
from kd_6d_pose_adlp_b200.ops import OTConfig, ot_loss_batched
from kd_6d_pose_adlp_b200.synthetic import cu_seqlens, ot_batch
from oracle import sinkhorn_analytic
dev = torch.device("cuda:0")
rel = lambda a, b: float(np.abs(a - b).max() / np.abs(b).max())
res = []
for n in (300, 600):
    for sigma in (0.02, 0.05, 0.15, 0.3):
        for seed in (2, 7):
            b = ot_batch(nimg=1, seed=seed, n_range=(n, n + 20), m_range=(n - 20, n), p_empty_teacher=0.0, sigma=sigma)
            t = lambda a: torch.from_numpy(a).to(dev)
            out = ot_loss_batched(t(b["xs"]), t(b["ws"]), t(b["xt"]), t(b["wt"]), b["pos_per_img"], b["pos_per_img_t"], OTConfig())
            torch.cuda.synchronize()
            key = "/tmp/ref_%d_%g_%d.npz" % (n, sigma, seed)
            if os.path.exists(key): ref = dict(np.load(key))
            else:
                ref = sinkhorn_analytic.kdot_fwd_bwd_f64(b["xs"], b["ws"], b["xt"], b["wt"], cu_seqlens(b["pos_per_img"]), cu_seqlens(b["pos_per_img_t"]), 8, 2)
                np.savez(key, **ref)
            res.append(rel(out["grad_xs"].cpu().numpy(), ref["grad_xs"]))
print(os.environ.get("KDOT_LIB", "default")[-22:], " ".join("%.1e" % v for v in res), "| max %.1e" % max(res))
