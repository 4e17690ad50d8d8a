mkdir -p gpurun_out/r2/san2
CS=/usr/local/cuda/bin/compute-sanitizer
for tool in memcheck racecheck synccheck; do
  for cfg in "mixed 1" "split3 0" "single 1"; do
    set -- $cfg
    UMX_TC_PAIR=$2 timeout 1200 $CS --tool $tool --error-exitcode 9 --print-limit 5 python tools/sanitize.py $1 > gpurun_out/r2/san2/${tool}_$1_pair$2.log 2>&1
    echo "$tool $1 pair=$2 rc=$? ok=$(grep -c '\]: ok' gpurun_out/r2/san2/${tool}_$1_pair$2.log) $(grep 'ERROR SUMMARY\|RACECHECK SUMMARY' gpurun_out/r2/san2/${tool}_$1_pair$2.log | tail -1) lines=$(grep -o 'kernels_tc.cu:[0-9]*' gpurun_out/r2/san2/${tool}_$1_pair$2.log | sort | uniq -c | tr '\n' ' ')"
  done
done
