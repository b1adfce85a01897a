"""Turns gpurun_out/*.ncu-rep + launch csv into small tracked text summaries under profiles/ (run in the authoring
container: `python tools/summarize_ncu.py r01`)."""
import csv
import io
import os
import subprocess
import sys
from collections import defaultdict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KEYS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__cluster_size", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "launch__waves_per_multiprocessor", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.max.pct_of_peak_sustained_active",
    "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "dram__cycles_active.avg.pct_of_peak_sustained_elapsed", "smsp__inst_executed.sum", "sm__cycles_active.avg",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__warps_eligible.avg.per_cycle_active",
]


def raw(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    return rows[0], rows[1], rows[2:]


def summarize_rep(rep, dst):
    hdr, units, kernels = raw(rep)
    with open(dst, "w") as fh:
        fh.write(f"# ncu --set full --clock-control none summary of {os.path.basename(rep)}\n")
        for vals in kernels:
            d = dict(zip(hdr, vals))
            u = dict(zip(hdr, units))
            fh.write(f"\n## {d.get('Kernel Name', '?')}  grid {d.get('Grid Size', '?')} block {d.get('Block Size', '?')}\n")
            for k in KEYS:
                if k in d:
                    fh.write(f"{k:72s} {d[k]:>18s} {u[k]}\n")
            fh.write("# warp stall reasons (cycles per issued instruction)\n")
            st = [(float(d[h].replace(",", "")), h) for h in hdr
                  if h.startswith("smsp__average_warps_issue_stalled") and h.endswith("_per_issue_active.ratio") and d[h]]
            for v, h in sorted(st, reverse=True)[:8]:
                fh.write(f"  {h.split('stalled_')[1].split('_per_issue')[0]:28s} {v:8.3f}\n")
    print("wrote", dst)


def summarize_launches(csv_path, dst):
    rows = [r for r in csv.reader(open(csv_path)) if len(r) > 5]
    hdr = rows[0]
    ik, iv = hdr.index("Kernel Name"), hdr.index("Metric Value")
    tot = defaultdict(float)
    cnt = defaultdict(int)
    for r in rows[1:]:
        tot[r[ik]] += float(r[iv].replace(",", ""))
        cnt[r[ik]] += 1
    all_ns = sum(tot.values())
    with open(dst, "w") as fh:
        fh.write(f"# ncu --metrics gpu__time_duration.sum --clock-control none launch list: {os.path.basename(csv_path)}\n")
        fh.write("# per-launch times are cold-cache and serialised under the profiler: compare SHARES, not absolutes\n")
        fh.write(f"{'kernel':90s} {'launches':>8s} {'total us':>12s} {'avg us':>10s} {'share':>7s}\n")
        for k, v in sorted(tot.items(), key=lambda kv: -kv[1]):
            fh.write(f"{k[:90]:90s} {cnt[k]:8d} {v / 1e3:12.1f} {v / 1e3 / cnt[k]:10.2f} {100 * v / all_ns:6.1f}%\n")
    print("wrote", dst)


if __name__ == "__main__":
    tag = sys.argv[1] if len(sys.argv) > 1 else "r01"
    src, dst = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")
    os.makedirs(dst, exist_ok=True)
    for f in sorted(os.listdir(src)):
        if not f.startswith(tag):
            continue
        p = os.path.join(src, f)
        if f.endswith(".ncu-rep"):
            summarize_rep(p, os.path.join(dst, f.replace(".ncu-rep", "_ncu_summary.txt")))
        elif f.endswith(".csv") and "launches" in f:
            summarize_launches(p, os.path.join(dst, f.replace(".csv", "_summary.txt")))
        elif f.endswith(".json") or f.endswith(".txt") or f.endswith("nvidia_smi.csv"):
            with open(p) as a, open(os.path.join(dst, f), "w") as b:
                b.write(a.read())
