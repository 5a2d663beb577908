"""Instruction / stall-sample share per named line range of one source file, from an
`ncu --page source --csv --print-source cuda,sass` dump.  usage: ncu_phases.py dump.csv file.cu name:lo-hi ..."""
import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
fname = sys.argv[2]
phases = []
for a in sys.argv[3:]:
    n, r = a.rsplit(":", 1)
    lo, hi = r.split("-")
    phases.append((n, int(lo), int(hi)))
cur = None; hdr = None
inst = collections.Counter(); smp = collections.Counter()
for r in rows:
    if len(r) >= 2 and r[0] == "File Path": cur = r[1].split("/")[-1]; continue
    if r and r[0] == "Line No": hdr = r; continue
    if hdr is None or len(r) < 8 or not r[0].strip().isdigit(): continue
    try: s = int(r[4]); i = int(r[7])
    except ValueError: continue
    name = "other:" + cur
    if cur == fname:
        name = fname + " (unassigned)"
        for n, lo, hi in phases:
            if lo <= int(r[0]) <= hi: name = n; break
    inst[name] += i; smp[name] += s
ti = sum(inst.values()) or 1; ts = sum(smp.values()) or 1
for n, v in inst.most_common():
    print(f"{v/ti*100:5.1f}% inst {smp[n]/ts*100:5.1f}% smp  {v:>10d}  {n}")
print("total warp-instructions", ti)
