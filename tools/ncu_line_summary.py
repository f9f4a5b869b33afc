"""Per-CUDA-source-line summary of an `ncu --page source --csv --print-source cuda,sass` export.
Usage: python tools/ncu_line_summary.py file.csv [top_n]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 30
out = []
hdr = None; fname = ""
for r in rows:
    if len(r) == 2 and r[0] == "File Path": fname = r[1].split("/")[-1]; continue
    if r and r[0] == "Line No": hdr = r; si = r.index("# Samples"); ei = r.index("Instructions Executed"); continue
    if hdr and len(r) == len(hdr) and r[0] not in ("", ) and r[0].isdigit() and r[2] == "-":
        try: out.append((int(r[si]), int(r[ei]), fname, int(r[0]), r[1].strip()))
        except ValueError: pass
ts = sum(o[0] for o in out); te = sum(o[1] for o in out)
print("samples %d, instructions %d" % (ts, te))
for s, e, f, ln, src in sorted(out, key=lambda o: -o[0])[:top]:
    print("%5.1f%% smp %5.1f%% ins  %s:%d  %s" % (100.0 * s / max(ts, 1), 100.0 * e / max(te, 1), f, ln, src[:100]))
