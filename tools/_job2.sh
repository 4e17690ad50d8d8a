mkdir -p gpurun_out/r2/final3
timeout 300 ncu --set full --clock-control none --import-source on --nvtx --nvtx-include "timed_resident/" -k regex:"first_conv|stitch" -c 2 -o /tmp/simt python bench.py --size 4096 --steps 1 --warmup 1 --cpu-budget 0 --configs none --no-modes --no-crop-check > gpurun_out/r2/final3/ncu.log 2>&1
ncu -i /tmp/simt.ncu-rep --page raw --csv > gpurun_out/r2/final3/simt_raw.csv 2>/dev/null
python tools/ncu_summary.py gpurun_out/r2/final3/simt_raw.csv gpurun_out/r2/final3/r2_ncu_full_simt_4k.json "ncu --set full --clock-control none --nvtx --nvtx-include timed_resident/ -k regex:first_conv|stitch -c 2 python bench.py --size 4096 --steps 1 --warmup 1 (final build)" "ld0.conv0+gather+taps,stitch_quantize"
tail -3 gpurun_out/r2/final3/ncu.log
