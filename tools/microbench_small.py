"""Where does the small kernel's time go?  Sweeps rounds (via scaling), batch size and cache state."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from kd_6d_pose_adlp_b200 import _lib
from kd_6d_pose_adlp_b200.ops import cu_seqlens
from kd_6d_pose_adlp_b200.synthetic import ot_batch

L = _lib.lib()
dev = torch.device("cuda:0")
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

def run(nimg, scaling, cold, reps=50, n_range=(8, 12)):
    b = ot_batch(nimg, seed=1, n_range=n_range, m_range=n_range)
    t = lambda a: torch.from_numpy(a).to(dev)
    xs0, xt0, ws, wt = t(b["xs"]), t(b["xt"]), t(b["ws"]), t(b["wt"])
    xs, xt = xs0.clone(), xt0.clone()
    cn, cm = cu_seqlens(b["pos_per_img"], dev), cu_seqlens(b["pos_per_img_t"], dev)
    loss = torch.empty(nimg, device=dev); valid = torch.empty(nimg, dtype=torch.int32, device=dev)
    nits = torch.empty(nimg, dtype=torch.int32, device=dev); gx = torch.empty_like(xs); gw = torch.empty_like(ws)
    mx, mm = max(b["pos_per_img"]), max(b["pos_per_img_t"])
    def step():
        rc = L.kdot_sinkhorn_fwd_bwd(xs.data_ptr(), ws.data_ptr(), xt.data_ptr(), wt.data_ptr(), cn.data_ptr(), cm.data_ptr(),
                                     nimg, 8, 2, mx, mm, 0, 2.0, 0.001, 0.5, scaling, 640.0, 480.0, 1, loss.data_ptr(), None,
                                     valid.data_ptr(), gx.data_ptr(), gw.data_ptr(), nits.data_ptr(), None, 0,
                                     torch.cuda.current_stream().cuda_stream)
        assert rc == 0
    ts = []
    for i in range(reps + 5):
        xs.copy_(xs0); xt.copy_(xt0)
        if cold: flush.fill_(1)
        torch.cuda._sleep(100000)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); step(); e1.record(); torch.cuda.synchronize()
        if i >= 5: ts.append(e0.elapsed_time(e1) * 1e3)
    return np.median(ts), int(nits.max())

# empty kernel launch floor
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
ts = []
for i in range(30):
    torch.cuda._sleep(100000); e0.record(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1) * 1e3)
print("event pair alone: %.2f us" % np.median(ts))
for nimg in (() if os.environ.get('MB_QUICK') else (1, 64, 148, 512)):
    for scaling in (0.5, 0.05, 0.9):
        for cold in (False, True):
            us, ni = run(nimg, scaling, cold)
            print(f"nimg={nimg:4d} scaling={scaling:<5} nits={ni:3d} cold={cold!s:5}: {us:7.2f} us")
us, ni = run(64, 0.5, False, n_range=(4, 4)); print("nimg=64 N=M=4 warm:", us, ni)
us, ni = run(64, 0.5, False, n_range=(16, 16)); print("nimg=64 N=M=16 warm:", us, ni)

# ---- in-kernel clock stamps (kdot_debug_set_clock_buffer) ----
nimg = 64
clk = torch.zeros(nimg, 16, dtype=torch.int64, device=dev)
L.kdot_debug_set_clock_buffer(clk.data_ptr())
us, ni = run(nimg, 0.5, False, reps=10)
L.kdot_debug_set_clock_buffer(None)
c = clk.cpu().numpy()
d = np.diff(c[:, :7], axis=1)
names = ["loads+logw", "barrier+normalise+bbox", "schedule", "stage smem+round consts", "rounds", "final+sum"]
print("kernel %.2f us; median SM cycles per phase (thread 0 of each CTA):" % us)
for k, n in enumerate(names):
    print("  %-28s %8.0f cycles  (%.2f us @1.965GHz)" % (n, np.median(d[:, k]), np.median(d[:, k]) / 1965))
print("  total in-kernel %.0f cycles = %.2f us" % (np.median(c[:, 6] - c[:, 0]), np.median(c[:, 6] - c[:, 0]) / 1965))
g = c[:, 7]
print("CTA start skew (globaltimer ns): min %d max-min %d ; in-kernel cycles min %d max %d" % (0, g.max() - g.min(), (c[:,6]-c[:,0]).min(), (c[:,6]-c[:,0]).max()))
# back-to-back launches without sleep/flush: steady-state per-launch time
xs = torch.zeros(1, device=dev)
b = ot_batch(64, seed=1)
t = lambda a: torch.from_numpy(a).to(dev)
xs, xt, ws, wt = t(b["xs"]), t(b["xt"]), t(b["ws"]), t(b["wt"])
cn, cm = cu_seqlens(b["pos_per_img"], dev), cu_seqlens(b["pos_per_img_t"], dev)
loss = torch.empty(64, device=dev); valid = torch.empty(64, dtype=torch.int32, device=dev)
nits = torch.empty(64, dtype=torch.int32, device=dev); gx = torch.empty_like(xs); gw = torch.empty_like(ws)
def step(norm):
    L.kdot_sinkhorn_fwd_bwd(xs.data_ptr(), ws.data_ptr(), xt.data_ptr(), wt.data_ptr(), cn.data_ptr(), cm.data_ptr(),
                            64, 8, 2, 12, 12, 0, 2.0, 0.001, 0.5, 0.5, 640.0, 480.0, norm, loss.data_ptr(), None,
                            valid.data_ptr(), gx.data_ptr(), gw.data_ptr(), nits.data_ptr(), None, 0,
                            torch.cuda.current_stream().cuda_stream)
step(1); torch.cuda.synchronize()
for n in (1, 10, 100):
    torch.cuda._sleep(2000000)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): step(0)
    e1.record(); torch.cuda.synchronize()
    print("back-to-back x%d (queued behind a sleep): %.2f us per launch" % (n, e0.elapsed_time(e1) * 1e3 / n))
if c[:, 8:].any():  # built with -DKDOT_SMALL_ROUND_STAMPS: clock at the start of rounds 0..7 (thread 0 of part 0)
    print("cycles per round (7 stamped rounds, median over images):", np.median(np.diff(c[:, 8:16], axis=1), axis=0))
    print("  last stamp -> end of the round phase (stamp 5):", np.median(c[:, 5] - c[:, 15]))
# which CTAs finish last?  (the launch lasts as long as its slowest cluster)
bb = ot_batch(64, seed=1, n_range=(8, 12), m_range=(8, 12))
tot = c[:, 6] - c[:, 0]
order = np.argsort(-tot)[:8]
print("slowest images: (cycles, N, M)", [(int(tot[i]), bb["pos_per_img"][i], bb["pos_per_img_t"][i]) for i in order])
order = np.argsort(tot)[:4]
print("fastest images: (cycles, N, M)", [(int(tot[i]), bb["pos_per_img"][i], bb["pos_per_img_t"][i]) for i in order])
