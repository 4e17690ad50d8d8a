"""Static SASS evidence for profiles/: per kernel of libunmicst_b200.so, the count of the Blackwell-native mnemonics
(UTC*MMA = tcgen05.mma, LDTM/STTM = tcgen05.ld/st, UTMALDG/UTMASTG = TMA, UTCBAR = tcgen05.commit, SYNCS = mbarrier) and of
the legacy tensor path (HMMA) that must NOT appear.  Runs on the build host: cuobjdump -sass, no GPU.
usage: sass_evidence.py [lib.so] > profiles/rN_sass_evidence.txt"""
import collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "unmicst_b200", "csrc", "libunmicst_b200.so")
txt = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
WATCH = ["UTCHMMA", "UTCQMMA", "UTCBAR", "UTCATOMSWS", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UTMAPF", "UBLKCP", "SYNCS", "UCGABAR", "HMMA",
         "FFMA", "DFMA", "DMUL", "DADD", "LDG", "STG", "LDS", "STS", "SHFL", "MUFU", "BAR"]
kern, total = collections.OrderedDict(), collections.Counter()
cur = None
for line in txt.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1)
        kern[cur] = collections.Counter()
        continue
    m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)((?:\.[A-Z0-9_]+)*)", line)
    if m and cur:
        op, suffix = m.group(1), m.group(2)
        kern[cur][op] += 1
        kern[cur]["__n"] += 1
        if op in ("UTCHMMA", "UTMALDG", "UTCBAR"):
            kern[cur][op + suffix] += 1
demangle = subprocess.run(["c++filt"] + list(kern), capture_output=True, text=True).stdout.splitlines()
print(f"# cuobjdump -sass {os.path.relpath(lib, ROOT)}   (sm_100a; {len(kern)} kernels)")
print("# tcgen05.mma -> UTCHMMA[.2CTA], tcgen05.ld -> LDTM, cp.async.bulk.tensor -> UTMALDG, tcgen05.commit -> UTCBAR, mbarrier -> SYNCS; HMMA would be mma.sync")
for (name, c), dm in zip(kern.items(), demangle):
    short = re.sub(r"\(.*", "", dm.replace("umx::(anonymous namespace)::", "").replace("void ", ""))
    for op in WATCH:
        total[op] += c.get(op, 0)
    tc = {k: v for k, v in c.items() if k.startswith(("UTCHMMA", "UTMALDG", "UTCBAR")) and "." in k}
    print(f"{short:70s} {c['__n']:6d} instr  " + "  ".join(f"{op} {c[op]}" for op in WATCH if c.get(op)) + ("   [" + ", ".join(f"{k} {v}" for k, v in sorted(tc.items())) + "]" if tc else ""))
print("# totals: " + "  ".join(f"{op} {total[op]}" for op in WATCH))
assert total["HMMA"] == 0, "legacy mma.sync instructions present"
