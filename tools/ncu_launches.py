"""Print kernel name, duration (us) and warp instructions of every launch in an ncu --csv launch list."""
import csv, sys
rows = list(csv.reader(l for l in open(sys.argv[1]) if l.startswith('"')))
hdr = rows[0]
ik, im, iv = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value")
iid = hdr.index("ID")
out = {}
for r in rows[1:]:
    d = out.setdefault(r[iid], {"k": r[ik][:60]})
    d[r[im]] = r[iv]
for i, d in out.items():
    t = float(d.get("gpu__time_duration.sum", "0").replace(",", ""))
    print(f"{i:>4} {d['k']:60s} {t / 1e3 if t > 1e4 else t:10.2f} {d.get('smsp__inst_executed.sum', '')}")
