import sys, os, time
sys.path.insert(0, os.getcwd())
import torch
sys.argv = ["x", "64"]
import tools.time_kd_pose_loss as T
# reuse the setup by running main() pieces: monkeypatch timed() to profile
import types
src = open("tools/time_kd_pose_loss.py").read().replace('    kf, ku = fused(), unfused()', '''    for _ in range(5): fused()
    from torch.profiler import profile, ProfilerActivity
    with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
        for _ in range(5):
            for t in s_cls + s_reg: t.grad = None
            fused()
        torch.cuda.synchronize()
    print(prof.key_averages().table(sort_by="self_cpu_time_total", row_limit=45, max_name_column_width=60))
    return
    kf, ku = fused(), unfused()''')
exec(compile(src, "t", "exec"))
