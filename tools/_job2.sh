timeout 300 python -m pytest tests -m gpu -x -q -k "multi" 2>&1 | tail -3
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --size 8192 --steps 3 --warmup 2 --cpu-budget 0 --configs none --no-modes > /tmp/b2.json 2>/tmp/b2.err
tail -c 300 /tmp/b2.err
python - <<'PY'
import json
d=json.loads(open("/tmp/b2.json").read().strip().splitlines()[-1])
print(d["value"], d["n_gpus"], d["bands_bit_exact"], d["per_rank"], d["stitched_u8"])
PY
