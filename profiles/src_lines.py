"""Per-source-line instruction and stall-sample totals of one kernel from an ncu report:
   ncu -i X.ncu-rep --page source --csv --print-source cuda,sass --kernel-name regex:K --launch-count 1 > f.csv
   python profiles/src_lines.py f.csv [top]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
cur_file, hdr, out = None, None, {}
line = None
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        cur_file = r[1].split("/")[-1]
        continue
    if r[0] == "Line No":
        hdr = r
        ii, si = hdr.index("Instructions Executed"), hdr.index("# Samples")
        continue
    if hdr is None or len(r) < len(hdr):
        continue
    if r[0] != "":
        line = (cur_file, int(r[0]), r[1].strip())
        continue
    if line is None:
        continue
    try:
        inst, smp = int(r[ii]), int(r[si])
    except ValueError:
        continue
    a = out.setdefault(line, [0, 0, 0])
    a[0] += inst; a[1] += smp; a[2] += 1
ti = sum(v[0] for v in out.values()); ts = sum(v[1] for v in out.values())
print("total warp instructions %d, samples %d" % (ti, ts))
for k, v in sorted(out.items(), key=lambda kv: -kv[1][0])[:top]:
    print("%5.1f%% inst %5.1f%% smp %4d sass  %s:%d  %s" % (100 * v[0] / ti, 100 * v[1] / max(ts, 1), v[2], k[0], k[1], k[2][:90]))
