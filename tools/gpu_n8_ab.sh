#!/bin/bash
# 8-GPU A/B: brick list ordered by cost (TG_BRICK_ORDER=1) against list order (=0)
out=gpurun_out
mkdir -p $out
for order in 0 1; do
TG_BRICK_ORDER=$order python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --steps 8 --warmup 10 \
  > $out/s20_n8_order$order.json 2> $out/s20_n8_order$order.err
python - $out/s20_n8_order$order.json <<'PY'
import json, sys
d = json.loads([l for l in open(sys.argv[1]) if l.startswith('{')][-1])
print(sys.argv[1], "ms/step %.3f  e2e %.3f" % (d["ms_per_step"], d["e2e"]["ms_per_step"]), d["config"]["partition"][:90])
for k, v in d["per_rank_ms"].items(): print("   ", k, [round(x, 3) for x in v])
PY
done
