mkdir -p gpurun_out/j9
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/j9/pytest.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/j9/pytest.log
python tools/layer_sweep.py single UMX_TC_EXP=0 2>&1 | tee gpurun_out/j9/sweep.txt
python tools/layer_sweep.py split3 UMX_TC_EXP=0 2>&1 | tee gpurun_out/j9/sweep3.txt
env UMX_TC_EXP=64 python bench.py --size 2048 --steps 1 --warmup 1 --cpu-budget 0 --precision single 2>&1 >/dev/null | grep "umx dbg" | tail -13 > gpurun_out/j9/dbg.txt
