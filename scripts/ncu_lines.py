"""Per-source-line hot spots of one launch in a .ncu-rep: warp instructions, lanes per instruction, stall samples.
usage: ncu_lines.py <rep> [launch_index] [top_n]"""
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
launch = int(sys.argv[2]) if len(sys.argv) > 2 else 0
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--launch-skip", str(launch), "--launch-count", "1"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
fname = None
hdr = None
lines = []
tot_inst = tot_thr = tot_samp = 0
for r in rows:
    if len(r) >= 2 and r[0] in ("File Name", "File Path"):
        fname = r[1].split("/")[-1]; continue
    if len(r) > 4 and r[0] == "Line No":
        hdr = r; continue
    if hdr is None or len(r) < len(hdr) or r[0] == "":
        continue
    d = dict(zip(hdr[4:], r[4:]))
    try:
        inst = int(d["Instructions Executed"]); thr = int(d["Thread Instructions Executed"]); samp = int(d["# Samples"])
    except ValueError:
        continue
    lines.append((inst, thr, samp, fname, r[0], r[1].strip()[:90], int(d.get("stall_long_sb", 0) or 0)))
    tot_inst += inst; tot_thr += thr; tot_samp += samp
print(f"total warp inst {tot_inst/1e6:.1f} M, thread inst {tot_thr/1e6:.1f} M, lanes/inst {tot_thr/max(tot_inst,1):.2f}, samples {tot_samp}")
by_file = {}
for l in lines:
    a = by_file.setdefault(l[3], [0, 0, 0]); a[0] += l[0]; a[1] += l[1]; a[2] += l[2]
for f, a in sorted(by_file.items(), key=lambda kv: -kv[1][0]):
    print(f"  {f:24s} inst {a[0]/1e6:8.1f} M ({a[0]/tot_inst:5.1%}) lanes {a[1]/max(a[0],1):5.2f} samples {a[2]/max(tot_samp,1):5.1%}")
print(f"{'inst(M)':>8s} {'share':>6s} {'lanes':>5s} {'samp%':>6s} {'longsb':>6s}  where")
for l in sorted(lines, key=lambda l: -l[2])[:top]:
    print(f"{l[0]/1e6:8.2f} {l[0]/tot_inst:6.1%} {l[1]/max(l[0],1):5.1f} {l[2]/max(tot_samp,1):6.1%} {l[6]:6d}  {l[3]}:{l[4]}  {l[5]}")
