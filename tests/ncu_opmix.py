"""Executed-instruction mix by opcode from an .ncu-rep source page: python tests/ncu_opmix.py file.ncu-rep [N]"""
import csv, subprocess, sys, re
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hi]; ix = {h: i for i, h in enumerate(hdr)}
body = [r for r in rows[hi + 1:] if len(r) == len(hdr)]
mix = {}
tot = 0
for r in body:
    n = int(r[ix["Instructions Executed"]] or 0)
    src = r[ix["Source"]].strip()
    src = re.sub(r"^@!?U?P\d+\s+", "", src)
    op = src.split()[0].split(".")[0] if src else "?"
    mix[op] = mix.get(op, 0) + n
    tot += n
print("total warp-instructions", tot)
for op, n in sorted(mix.items(), key=lambda kv: -kv[1])[: int(sys.argv[2]) if len(sys.argv) > 2 else 25]:
    print(f"  {op:12s} {n:10d} {100 * n / tot:5.1f}%")
