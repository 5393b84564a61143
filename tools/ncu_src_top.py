"""Top SASS lines of an ncu `--page source --csv` dump by stall samples, plus an opcode histogram of executed instructions."""
import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]; idx = {h: i for i, h in enumerate(hdr)}
data = [r for r in rows[2:] if len(r) == len(hdr)]
S, E = idx["# Samples"], idx["Instructions Executed"]
f = lambda r, i: float(r[i] or 0)
tot, tote = sum(f(r, S) for r in data), sum(f(r, E) for r in data)
print("samples %d  warp-instr %d  lines %d" % (tot, tote, len(data)))
ops = collections.Counter()
for r in data:
    src = r[idx["Source"]].split()
    op = src[1] if src and src[0].startswith("@") and len(src) > 1 else (src[0] if src else "")
    ops[op.split(".")[0]] += f(r, E)
print("opcode mix:", ", ".join("%s %.1f%%" % (k, 100 * v / tote) for k, v in ops.most_common(22)))
n = int(sys.argv[2]) if len(sys.argv) > 2 else 40
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
for r in sorted(data, key=lambda r: -f(r, S))[:n]:
    top = sorted(((f(r, idx[s]), s[6:]) for s in stalls), reverse=True)[:2]
    print("%6.0f %5.1f%% ex=%8.0f %-22s| %s" % (f(r, S), 100 * f(r, S) / tot, f(r, E), ",".join("%s:%.0f" % (b, a) for a, b in top), r[idx["Source"]][:100]))
