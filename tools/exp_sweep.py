"""Timing experiments: which component bounds each tensor-path layer (results are invalid numerically)."""
import json, os, subprocess, sys
prec = sys.argv[1] if len(sys.argv) > 1 else "single"
cfgs = [("base", "0"), ("noA", "1"), ("noB", "2"), ("noAB", "3"), ("1mma", "4"), ("noEpi", "8"), ("noAB+noEpi", "11"), ("1mma+noEpi", "12")]
rows = {}
for pair in ("0", "1"):
    for name, e in cfgs:
        env = dict(os.environ, UMX_TC_PAIR=pair, UMX_TC_EXP=e)
        r = subprocess.run([sys.executable, "bench.py", "--size", "4096", "--steps", "2", "--warmup", "1", "--cpu-budget", "0",
                            "--precision", prec], capture_output=True, text=True, env=env)
        try:
            d = json.loads(r.stdout.strip().splitlines()[-1])
        except Exception:
            print(pair, name, "FAILED", r.stderr[-300:]); continue
        rows[f"p{pair}:{name}"] = {k["name"]: round(k["ms"] / k["launches"], 2) for k in d["roofline"]["kernels"] if k["launches"]}
names = list(next(iter(rows.values())).keys())
print("ms per launch (7396 tiles)")
print("%-14s" % "layer" + "".join("%14s" % k for k in rows))
for n in names:
    print("%-14s" % n + "".join("%14s" % rows[k].get(n) for k in rows))
