"""Aggregate an `ncu --page source --print-source cuda,sass --csv` dump per CUDA source line:
prints the lines with the most stall samples / executed instructions."""
import csv, sys, collections
path = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
rows = list(csv.reader(open(path)))
cur_file = None
agg = collections.OrderedDict()
hdr = None
for r in rows:
    if len(r) >= 2 and r[0] == "File Path":
        cur_file = r[1].split("/")[-1]; continue
    if len(r) >= 2 and r[0] == "Function Name":
        continue
    if r and r[0] == "Line No":
        hdr = r; continue
    if hdr is None or len(r) < 8:
        continue
    if r[0].strip().isdigit():
        key = (cur_file, int(r[0]))
        try:
            samples = int(r[4]); inst = int(r[7])
        except ValueError:
            continue
        a = agg.setdefault(key, [r[1].strip()[:100], 0, 0])
        a[1] += samples; a[2] += inst
tot_s = sum(a[1] for a in agg.values()) or 1
tot_i = sum(a[2] for a in agg.values()) or 1
print(f"total samples {tot_s}, total warp-instructions {tot_i}")
for (f, l), a in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
    print(f"{a[1]/tot_s*100:5.1f}% smp {a[2]/tot_i*100:5.1f}% inst  {f}:{l:<4d} {a[0]}")
