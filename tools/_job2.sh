mkdir -p gpurun_out/r2/fc
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
for w in duo4k; do
timeout 300 python bench.py --workload $w --steps 3 --warmup 2 --cpu-budget 0 --configs none --no-modes --no-crop-check > gpurun_out/r2/fc/bench3_$w.json 2> gpurun_out/r2/fc/bench3_$w.err
tail -c 300 gpurun_out/r2/fc/bench3_$w.err
done
python - <<'PY'
import json
for w in ("duo4k",):
    d=json.loads(open(f"gpurun_out/r2/fc/bench3_{w}.json").read().strip().splitlines()[-1])
    print(w, d["value"], d["ms_per_step"], d["config"].get("precision"), d["parity"]["max_abs_dp"])
    for k in d["roofline"]["kernels"][:3]: print("  ", k["name"], k["ms"], k["mma_x"])
    print(d["roofline"]["step"])
PY
