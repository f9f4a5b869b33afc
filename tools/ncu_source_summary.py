"""Summarise an `ncu --page source --csv` export: opcode mix weighted by executed count, hottest instructions,
stall samples per reason. Usage: python tools/ncu_source_summary.py file.csv [top_n]"""
import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
body, seen = [], set()
for r in rows[2:]:
    if len(r) == len(hdr) and r[ix["Instructions Executed"]].isdigit() and r[ix["Address"]] not in seen:
        seen.add(r[ix["Address"]])     # a report with several results of one kernel lists its SASS once per result
        body.append(r)
tot = sum(int(r[ix["Instructions Executed"]]) for r in body)
samples = sum(int(r[ix["# Samples"]]) for r in body)
print("instructions executed: %d, samples: %d, SASS lines: %d" % (tot, samples, len(body)))
ops = collections.Counter(); opsamp = collections.Counter()
for r in body:
    s = r[ix["Source"]].strip()
    t = s.split()
    op = t[1] if t[0].startswith("@") else t[0]
    op = ".".join(op.split(".")[:3])
    ops[op] += int(r[ix["Instructions Executed"]]); opsamp[op] += int(r[ix["# Samples"]])
print("-- opcode mix (executed %, sample %)")
for op, n in ops.most_common(top):
    print("  %-28s %6.2f%%  %6.2f%%" % (op, 100.0 * n / tot, 100.0 * opsamp[op] / max(samples, 1)))
print("-- stall reasons (all samples)")
st = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
acc = {h: sum(int(r[ix[h]]) for r in body) for h in st}
for h, n in sorted(acc.items(), key=lambda kv: -kv[1])[:10]:
    print("  %-24s %6.2f%%" % (h, 100.0 * n / max(samples, 1)))
print("-- hottest instructions by samples")
for r in sorted(body, key=lambda r: -int(r[ix["# Samples"]]))[:top]:
    print("  %6d smp %9d exec  %s" % (int(r[ix["# Samples"]]), int(r[ix["Instructions Executed"]]), r[ix["Source"]].strip()[:90]))
