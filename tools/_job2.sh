mkdir -p gpurun_out/r2/fc
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
for w in cyto2tma duo4k; do
timeout 300 python bench.py --workload $w --steps 2 --warmup 1 --cpu-budget 0 --configs none --no-modes --no-crop-check > gpurun_out/r2/fc/bench2_$w.json 2> gpurun_out/r2/fc/bench2_$w.err
tail -c 300 gpurun_out/r2/fc/bench2_$w.err
done
python - <<'PY'
import json
for w in ("cyto2tma","duo4k"):
    d=json.loads(open(f"gpurun_out/r2/fc/bench2_{w}.json").read().strip().splitlines()[-1])
    print(w, d["value"], d["ms_per_step"], d["config"].get("precision"), d["parity"]["max_abs_dp"], d["parity"].get("auto",{}).get("op_terms"))
    for k in d["roofline"]["kernels"]: print("  ", k["name"], k["ms"], k["mma_x"])
    print(d["roofline"]["step"])
PY
