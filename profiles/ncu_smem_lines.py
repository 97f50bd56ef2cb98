"""Shared-memory wavefronts (total / excessive = bank conflicts) per CUDA source line from an `ncu --page source --print-source cuda,sass --csv` dump.
usage: python profiles/ncu_smem_lines.py <csv> [top_n]"""
import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 20
hdr = None; agg = collections.OrderedDict(); cur = ""
for r in rows:
    if not r: continue
    if r[0] == "File Path": cur = r[1].split("/")[-1]; continue
    if r[0] == "Line No": hdr = r; continue
    if hdr is None or len(r) < len(hdr) or r[2] != "-": continue
    try:
        ln = int(r[0]); wf = int(r[hdr.index("L1 Wavefronts Shared")]); ex = int(r[hdr.index("L1 Wavefronts Shared Excessive")])
    except ValueError:
        continue
    a = agg.setdefault((cur, ln), [0, 0, r[1]]); a[0] += wf; a[1] += ex
tw = sum(a[0] for a in agg.values()); te = sum(a[1] for a in agg.values())
print("shared wavefronts %d, excessive %d (%.1f%%)" % (tw, te, 100.0 * te / max(tw, 1)))
for (f, ln), a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print("%5.1f%% wf (%5.1f%% of them excessive) | %s:%d: %s" % (100.0 * a[0] / tw, 100.0 * a[1] / max(a[0], 1), f, ln, a[2].strip()[:90]))
