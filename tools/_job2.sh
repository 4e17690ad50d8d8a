for rep in 1 2; do
for lib in A B; do
  if [ $lib = B ]; then export UNMICST_B200_LIB=$PWD/tools/_libB.so; else unset UNMICST_B200_LIB; fi
  timeout 200 python bench.py --size 6144 --steps 3 --warmup 2 --cpu-budget 0 --configs none --no-modes --no-crop-check > /tmp/o.json 2>/tmp/o.err
  python - "$lib" <<'PY'
import json,sys
d=json.loads(open("/tmp/o.json").read().strip().splitlines()[-1])
k={x["name"]:x["ms"] for x in d["roofline"]["kernels"]}
print(sys.argv[1], round(d["value"],1), d["config"]["precision"], {n:k[n] for n in ("ld1.conv0","lu1.conv2","lu0.convT","lu0.conv2+lt","lu1.convT")})
PY
done
done
