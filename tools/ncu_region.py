"""Dump SASS lines [a,b) of an ncu report with sample counts and stall reasons. usage: ncu_region.py rep a b"""
import csv, subprocess, sys
rep, a, b = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hi]; ci = {h: i for i, h in enumerate(hdr)}
body = [r for r in rows[hi + 1:] if len(r) == len(hdr)]
stall_cols = [h for h in hdr if h.startswith("stall_")]
for i in range(a, min(b, len(body))):
    r = body[i]
    s = float(r[ci["# Samples"]] or 0)
    top = sorted(((float(r[ci[c]] or 0), c) for c in stall_cols), reverse=True)[:2] if stall_cols else []
    tops = " ".join(f"{c[6:]}={int(v)}" for v, c in top if v > 0)
    print(f"{i:5d} {int(s):6d} {r[ci['Instructions Executed']]:>8} {r[ci['Source']].strip()[:90]:90s} {tops}")
