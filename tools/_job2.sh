mkdir -p gpurun_out/r2/final
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/r2/final/bench_1gpu.json 2> gpurun_out/r2/final/bench_1gpu.err
tail -c 400 gpurun_out/r2/final/bench_1gpu.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --nvtx --nvtx-include "timed_resident/" -c 400 --csv --log-file gpurun_out/r2/final/launches_4k.csv python bench.py --size 4096 --steps 2 --warmup 1 --cpu-budget 0 --configs none --no-modes --no-crop-check > gpurun_out/r2/final/ncu_bench.log 2>&1
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r2/final/bench_1gpu.json").read().strip().splitlines()[-1])
print(d["value"], d["ms_per_step"], d["e2e"], d["parity"]["max_abs_dp"], d["stitched_u8"], d["modes"])
for c in d.get("configs",[]): print(c.get("workload"), c.get("value"), (c.get("e2e") or {}).get("value"), (c.get("roofline") or {}).get("step",{}) and c["roofline"]["step"].get("frac"))
PY
wc -l gpurun_out/r2/final/launches_4k.csv
