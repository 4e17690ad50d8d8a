"""Summarise an `ncu --page source --csv --print-source sass` export: hottest SASS instructions and opcode histogram."""
import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
data = rows[2:]
def f(r, k):
    try: return float(r[ix[k]])
    except Exception: return 0.0
tot_s = sum(f(r, "# Samples") for r in data); tot_i = sum(f(r, "Instructions Executed") for r in data)
print("instructions", len(data), "samples", tot_s, "warp-inst", tot_i)
ops = collections.Counter(); ops_s = collections.Counter()
for r in data:
    op = r[ix["Source"]].split()[0] if r[ix["Source"]].split() else "?"
    if op.startswith("@"): op = r[ix["Source"]].split()[1]
    op = op.split(".")[0]
    ops[op] += f(r, "Instructions Executed"); ops_s[op] += f(r, "# Samples")
print("opcode: warp-inst share / sample share")
for op, n in ops.most_common(25):
    print("  %-12s %6.2f%% %6.2f%%" % (op, 100 * n / tot_i, 100 * ops_s[op] / tot_s))
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
tot = {s: sum(f(r, s) for r in data) for s in stalls}
print("stall totals:", {k: int(v) for k, v in sorted(tot.items(), key=lambda kv: -kv[1]) if v > 0})
n = int(sys.argv[2]) if len(sys.argv) > 2 else 40
print("top by samples:")
for r in sorted(data, key=lambda r: -f(r, "# Samples"))[:n]:
    st = sorted(((f(r, s), s) for s in stalls), reverse=True)[:2]
    print("  %s %6d smp %9d ex  %-70s %s" % (r[ix["Address"]][-5:], f(r, "# Samples"), f(r, "Instructions Executed"), r[ix["Source"]][:70], [(s, int(v)) for v, s in st]))
