"""Per-layer timing of the solo network under different kernel settings (env vars), GPU box."""
import json, os, subprocess, sys
cfgs = [dict(UMX_TC_PAIR=p, UMX_TC_STAGES=s) for p in ("0", "1") for s in ("2", "3", "4", "8")]
prec = sys.argv[1] if len(sys.argv) > 1 else "single"
rows = {}
for c in cfgs:
    env = dict(os.environ, **c)
    r = subprocess.run([sys.executable, "bench.py", "--size", "4096", "--steps", "2", "--warmup", "1", "--cpu-budget", "0",
                        "--precision", prec], capture_output=True, text=True, env=env)
    try:
        d = json.loads(r.stdout.strip().splitlines()[-1])
    except Exception:
        print(c, "FAILED", r.stderr[-500:]); continue
    key = f"pair{c['UMX_TC_PAIR']}_st{c['UMX_TC_STAGES']}"
    rows[key] = {k["name"]: k["tflops"] for k in d["roofline"]["kernels"] if k["launches"]}
    rows[key]["MP/s"] = round(d["value"], 1)
names = list(next(iter(rows.values())).keys())
print("%-16s" % "layer" + "".join("%12s" % k for k in rows))
for n in names:
    print("%-16s" % n + "".join("%12s" % rows[k].get(n) for k in rows))
