"""Warp-level instructions per CUDA source line from
   ncu -i rep.ncu-rep --page source --csv --print-source cuda,sass --kernel-name regex:<k>  > view.csv
   python tools/ncu_lines.py view.csv [top N]"""
import csv, os, sys
rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
fname, col, per, by_file = "?", None, [], {}
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        fname = os.path.basename(r[1]); continue
    if r[0] == "Line No":
        col = r.index("Instructions Executed"); continue
    if r[0].isdigit() and col is not None and r[col].isdigit():
        v = int(r[col])
        if v:
            per.append((v, fname, int(r[0]), r[1].strip()))
            by_file[fname] = by_file.get(fname, 0) + v
tot = sum(v for v, *_ in per)
print("total warp-instructions %d;  by file: %s" % (tot, ", ".join("%s %d (%.0f%%)" % (k, v, 100.0 * v / tot) for k, v in by_file.items())))
for v, f, ln, src in sorted(per, reverse=True)[:top]:
    print("%9d %5.1f%%  %s:%d  %s" % (v, 100.0 * v / tot, f, ln, src[:105]))
