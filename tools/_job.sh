mkdir -p gpurun_out/j15
for cfg in "UMX_TC_EXP=97" "UMX_TC_EXP=225" "UMX_TC_EXP=97 UMX_TC_PAIR=0" "UMX_TC_EXP=97 UMX_TC_HALO=0" "UMX_TC_EXP=225 UMX_TC_HALO=0"; do
  echo "=== $cfg"
  env $cfg python bench.py --size 2048 --steps 1 --warmup 1 --cpu-budget 0 --precision single 2>&1 >/dev/null | grep "umx dbg" | tail -13 | grep "lu0\|ld1\|lu1.conv2\|lb.conv"
done > gpurun_out/j15/dbg.txt 2>&1
