import csv, sys
rows=[r for r in csv.reader(open(sys.argv[1])) if len(r)>5]
hdr=rows[0]; ik=hdr.index("Kernel Name"); iv=hdr.index("Metric Value")
for r in rows[1:]:
    if "ssdhot" in r[ik]: print(f"{float(r[iv])/1000:9.1f} us  {r[ik][:90]}")
