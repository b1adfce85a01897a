"""GPU kernels (through the C ABI) against the committed golden fixtures generated from the reference's own code."""
import glob
import os

import numpy as np
import pytest
import torch

from tests import parity

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


@pytest.mark.parametrize("path", sorted(glob.glob(os.path.join(GOLDEN, "ot_boundary_*.npz"))))
def test_ot_boundary_fixture(path):
    from kd_6d_pose_adlp_b200.ops import OTConfig, ot_loss_batched

    z = np.load(path)
    dev = torch.device("cuda:0")
    weighted = bool(z["weighted"])
    reach = None if float(z["reach"]) < 0 else float(z["reach"])
    xs, xt = torch.from_numpy(z["xs"]).to(dev), torch.from_numpy(z["xt"]).to(dev)
    ws = torch.from_numpy(z["ws"]).to(dev) if weighted else None
    wt = torch.from_numpy(z["wt"]).to(dev) if weighted else None
    out = ot_loss_batched(xs, ws, xt, wt, z["pos_per_img"].tolist(), z["pos_per_img_t"].tolist(),
                          OTConfig(2.0, float(z["blur"]), float(z["scaling"]), reach))
    torch.cuda.synchronize()
    assert np.array_equal(out["valid"].cpu().numpy(), z["valid"])
    assert np.array_equal(out["nits"].cpu().numpy(), z["nits"])
    assert np.array_equal(xs.cpu().numpy(), z["ref32_xs_norm"])       # in-place normalisation, bit-exact
    assert np.array_equal(xt.cpu().numpy(), z["ref32_xt_norm"])
    rows = [parity.report("loss_per_img", out["loss_per_img"].cpu().numpy(), z["ref32_loss"], z["ref64_loss"]),
            parity.report("grad_xs", out["grad_xs"].cpu().numpy(), z["ref32_grad_xs"], z["ref64_grad_xs"])]
    if weighted:
        rows.append(parity.report("grad_ws", out["grad_ws"].cpu().numpy(), z["ref32_grad_ws"], z["ref64_grad_ws"]))
    print("\n" + os.path.basename(path) + "\n" + parity.fmt(rows))
    assert all(r["ok"] for r in rows), parity.fmt(rows)
