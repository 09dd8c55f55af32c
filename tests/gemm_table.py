"""Join the planned-GEMM dump (KMBART_DUMP_GEMMS=1) with an ncu launch list: per shape/epilogue class time and TFLOP/s.
   python tests/gemm_table.py shapes.txt launches.csv [launches_other.csv]"""
import csv, re, sys
from collections import OrderedDict
shapes = [l.split() for l in open(sys.argv[1]) if l.startswith("KMB_GEMM")]
def gemm_times(path):
    rows = [r for r in csv.reader(open(path)) if len(r) > 10 and r[0].isdigit()]
    return [float(r[14]) / 1e3 for r in rows if "gemm_tc05_kernel" in r[4]]
ts = [gemm_times(p) for p in sys.argv[2:]]
# the dump holds fwd plan + bwd plan of the first step(s); the profiled step launches them in the same order
n = len(ts[0])
shapes = shapes[:n] if len(shapes) >= n else shapes
agg = OrderedDict()
for i, s in enumerate(shapes):
    M, N, K = int(s[1]), int(s[2]), int(s[3])
    key = (M, N, K) + tuple(s[4:])
    a = agg.setdefault(key, [0] + [0.0] * len(ts))
    a[0] += 1
    for j, t in enumerate(ts):
        a[1 + j] += t[i]
print(f"{'M':>6} {'N':>6} {'K':>6}  n   " + "  ".join(f"us[{j}] TF/s[{j}]" for j in range(len(ts))) + "   flags")
tot = [0.0] * len(ts)
for key, a in agg.items():
    M, N, K = key[:3]
    fl = 2.0 * M * N * K
    cols = []
    for j in range(len(ts)):
        us = a[1 + j] / a[0]
        tot[j] += a[1 + j]
        cols.append(f"{us:7.1f} {fl / us / 1e6:6.0f}")
    print(f"{M:6d} {N:6d} {K:6d} {a[0]:3d}   " + "   ".join(cols) + "   " + " ".join(key[3:]))
print("total GEMM us:", [round(t, 1) for t in tot])
