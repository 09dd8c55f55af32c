"""Top stall-sample SASS lines of an .ncu-rep (source page): python tests/ncu_hot.py file.ncu-rep [N]"""
import csv, subprocess, sys
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hi]
ix = {h: i for i, h in enumerate(hdr)}
body = [r for r in rows[hi + 1:] if len(r) == len(hdr)]
N = int(sys.argv[2]) if len(sys.argv) > 2 else 25
tot = sum(int(r[ix["# Samples"]] or 0) for r in body)
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
agg = {s: sum(int(r[ix[s]] or 0) for r in body) for s in stalls}
print("total samples", tot, {k: v for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:8]})
order = sorted(range(len(body)), key=lambda i: -int(body[i][ix["# Samples"]] or 0))[:N]
for i in sorted(order):
    r = body[i]
    top = sorted(((int(r[ix[s]] or 0), s) for s in stalls), reverse=True)[:2]
    print(f"{i:5d} {int(r[ix['# Samples']]):6d} {100 * int(r[ix['# Samples']]) / tot:5.1f}%  {r[ix['Source']][:70]:70s} {top[0][1]}:{top[0][0]} {top[1][1]}:{top[1][0]}")
