"""Print the top stall-sample SASS lines of an ncu report's source page.  usage: ncu_top.py rep [n]"""
import csv, subprocess, sys
rep = sys.argv[1]; n = int(sys.argv[2]) if len(sys.argv) > 2 else 25
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hi]; ci = {h: i for i, h in enumerate(hdr)}
body = [r for r in rows[hi + 1:] if len(r) == len(hdr)]
tot = sum(float(r[ci["# Samples"]] or 0) for r in body)
print("total samples", tot)
idx = sorted(range(len(body)), key=lambda i: -float(body[i][ci["# Samples"]] or 0))[:n]
for i in sorted(idx):
    r = body[i]
    print(f"{i:5d} {float(r[ci['# Samples']]):8.0f} {100*float(r[ci['# Samples']])/tot:5.1f}%  {r[ci['Source']].strip()[:100]}")
