mkdir -p gpurun_out/r2
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
timeout 600 python bench.py --steps 3 --warmup 2 --configs none --cpu-budget 0 > gpurun_out/r2/bench_g.json 2> gpurun_out/r2/bench_g.err; echo "rc=$?"; tail -2 gpurun_out/r2/bench_g.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2/bench_g.json').read().strip().splitlines()[-1])
print('mixed', round(d['value'],1), round(d['ms_per_step'],1), 'e2e', round(d['e2e']['value'],1), {k:round(v['value'],1) for k,v in d['modes'].items()}, d['parity']['max_abs_dp'], d['stitched_u8']['max_abs_u8_diff'], d['stitched_u8_whole_slide']['max_abs_u8_diff'])
print('   '+' '.join('%s %.1f'%(k['name'].split('.conv')[0]+k['name'][-5:],k['ms']) for k in d['roofline']['kernels']))
PY
