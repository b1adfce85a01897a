"""The numerical argument behind the streaming kernel's compensated row sums (``urow_fold``), as an executable check:
in a float64 solver whose ONLY fp32 quantity is the running row sum, sequential fp32 accumulation over a few hundred
columns already costs d/dx an order of magnitude more than fp32 sums over 32-column sub-tiles folded into a wider total
(``tools/rowsum_study.py``).  CPU only, reduced size (the dense 1360-point run takes a minute)."""
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools"))


def test_subtile_sums_beat_running_fp32_sums():
    import rowsum_study

    res = rowsum_study.study(n=400, seed=5, sigma=0.1)
    seq, sub = res["fp32seq"], res["subtile"]
    print(res)
    assert sub[0] < 5e-6                      # compensated: d/dx within a few 1e-6 of the exact solve
    assert seq[0] > 3.0 * sub[0]              # running fp32 sums: several times worse on d/dx ...
    assert seq[1] > 3.0 * sub[1]              # ... and on the final soft-min offsets
