"""Aggregate an `ncu --page source --print-source cuda,sass --csv` dump per CUDA source line.
usage: python profiles/ncu_lines.py <csv> [top_n]"""
import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
agg = collections.OrderedDict(); cur_file = ""
hdr = None
for r in rows:
    if not r: continue
    if r[0] == "File Path": cur_file = r[1].split("/")[-1]; continue
    if r[0] == "Line No": hdr = r; continue
    if hdr is None or len(r) < 9: continue
    if r[2] != "-":      # SASS row; only per-line summary rows have Address == '-'
        continue
    try:
        ln = int(r[0]); n = int(r[7]); t = int(r[8]); smp = int(r[6])
    except ValueError:
        continue
    k = (cur_file, ln)
    a = agg.setdefault(k, [0, 0, 0, r[1]])
    a[0] += n; a[1] += t; a[2] += smp
tot = sum(a[0] for a in agg.values()); tots = sum(a[2] for a in agg.values())
print("total warp-instructions %d, samples %d" % (tot, tots))
for (f, ln), a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print("%5.1f%% inst %5.1f%% smp thr/inst %4.1f | %s:%d: %s" % (100.0 * a[0] / tot, 100.0 * a[2] / max(tots, 1), a[1] / max(a[0], 1), f, ln, a[3].strip()[:100]))
