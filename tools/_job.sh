mkdir -p gpurun_out/j11
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/j11/pytest.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/j11/pytest.log
python tools/layer_sweep.py single UMX_TC_ASTAGES=2,3 2>&1 | tee gpurun_out/j11/sweep.txt
