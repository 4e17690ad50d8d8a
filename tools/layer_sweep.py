"""Per-layer timing of the solo network under different kernel settings (env vars), GPU box.
usage: layer_sweep.py [precision] [KEY=v1,v2 ...]   e.g.  layer_sweep.py single UMX_TC_HALO=0,1 UMX_TC_PAIR=0,1"""
import itertools, json, os, subprocess, sys
prec = sys.argv[1] if len(sys.argv) > 1 else "single"
axes = [a.split("=") for a in sys.argv[2:]] or [["UMX_TC_PAIR", "0,1"], ["UMX_TC_STAGES", "2,3,4,8"]]
keys = [a[0] for a in axes]
cfgs = [dict(zip(keys, vals)) for vals in itertools.product(*[a[1].split(",") for a in axes])]
extra = os.environ.get("SWEEP_ARGS", "--size 4096").split()       # e.g. SWEEP_ARGS="--workload cyto2tma --size 6144"
rows = {}
for c in cfgs:
    env = dict(os.environ, **c)
    r = subprocess.run([sys.executable, "bench.py", *extra, "--steps", "2", "--warmup", "1", "--cpu-budget", "0",
                        "--configs", "none", "--no-modes", "--no-crop-check", "--precision", prec], capture_output=True, text=True, env=env)
    key = " ".join(f"{k.replace('UMX_TC_', '').replace('UMX_', '')}={v}" for k, v in c.items())
    try:
        d = json.loads(r.stdout.strip().splitlines()[-1])
    except Exception:
        print(key, "FAILED", r.stderr[-500:]); continue
    rows[key] = {k["name"]: k["ms"] for k in d["roofline"]["kernels"] if k["launches"]}
    rows[key]["MP/s"] = round(d["value"], 1)
    rows[key]["max|dp|"] = "%.1e" % d["parity"].get("max_abs_dp", float("nan"))
names = list(next(iter(rows.values())).keys())
print("%-16s" % "layer (ms)" + "".join("%18s" % k for k in rows))
for n in names:
    print("%-16s" % n + "".join("%18s" % rows[k].get(n) for k in rows))
