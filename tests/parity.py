"""Parity bookkeeping shared by the GPU tests ("Parity definition").

For every tensor three relative max-norm numbers are reported:
    new_vs_ref32 = |new - ref32| / |ref32|      (the north-star's 1e-4 bar)
    new_vs_ref64 = |new - ref64| / |ref64|
    ref32_vs_ref64                              (the reference formulation's own fp32 rounding noise)
A tensor passes when  new_vs_ref32 <= TOL  or  new_vs_ref64 <= max(ref32_vs_ref64, TOL_F64):
at blur = 1e-3 (eps = 1e-6) the reference's fp32 evaluation is itself up to 6e-3 away from exact
arithmetic on d/dx, so "within 1e-4 of ref32" is only meaningful where ref32 itself is that accurate.
"""
import numpy as np

TOL = 1e-4       # north-star tolerance vs the fp32 reference formulation
TOL_F64 = 2e-5   # alternatively: at least this close to exact (fp64) arithmetic


def rel(a, b):
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    den = np.abs(b).max()
    return float(np.abs(a - b).max() / den) if den > 0 else float(np.abs(a - b).max())


def report(name, new, ref32, ref64):
    r = dict(name=name, new_vs_ref32=rel(new, ref32), new_vs_ref64=rel(new, ref64), ref32_vs_ref64=rel(ref32, ref64))
    r["ok"] = bool(r["new_vs_ref32"] <= TOL or r["new_vs_ref64"] <= max(r["ref32_vs_ref64"], TOL_F64))
    return r


def fmt(rows):
    out = ["%-14s %12s %12s %14s  ok" % ("tensor", "new-ref32", "new-ref64", "ref32-ref64")]
    for r in rows:
        out.append("%-14s %12.3e %12.3e %14.3e  %s" % (r["name"], r["new_vs_ref32"], r["new_vs_ref64"],
                                                        r["ref32_vs_ref64"], r["ok"]))
    return "\n".join(out)
