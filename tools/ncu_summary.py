"""Text summary of an .ncu-rep (key raw metrics per kernel + top stall lines) for profiles/.
usage: ncu_summary.py report.ncu-rep > profiles/xxx.txt"""
import csv, subprocess, sys
rep = sys.argv[1]
WANT = ["gpu__time_duration.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__t_sector_hit_rate.pct",
        "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "launch__shared_mem_per_block_dynamic", "launch__cluster_size"]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
ci = {h: i for i, h in enumerate(hdr)}
print(f"# ncu summary of {rep} (ncu --set full --clock-control none; cold-cache, serialised launches)")
for r in rows[2:]:
    print("\n## " + r[ci["Kernel Name"]][:110])
    for w in WANT:
        if w in ci:
            print(f"  {w:75s} {r[ci[w]]:>16s} {units[ci[w]]}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
srows = list(csv.reader(src.splitlines()))
try:
    hi = next(i for i, r in enumerate(srows) if r and r[0] == "Address")
    sh = srows[hi]; sci = {h: i for i, h in enumerate(sh)}
    body = []
    for r in srows[hi + 1:]:
        if len(r) != len(sh) or r[0] == "Address":   # a second kernel's table starts: stop
            break
        body.append(r)
    tot = sum(float(r[sci["# Samples"]] or 0) for r in body)
    top = sorted(body, key=lambda r: -float(r[sci["# Samples"]] or 0))[:12]
    print(f"\n## top stall samples (first kernel in report; {int(tot)} samples)")
    for r in top:
        s = float(r[sci["# Samples"]] or 0)
        print(f"  {100 * s / tot:5.1f}%  {r[sci['Source']].strip()[:100]}")
except StopIteration:
    pass
