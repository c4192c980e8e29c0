#!/bin/bash
# bench.py at N GPUs of one box, as the driver launches it.   tools/gpu_scale.sh <tag> <N>
tag=${1:-run}; n=${2:-2}
out=gpurun_out
mkdir -p $out
python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $n --steps 8 --warmup 10 \
  > $out/${tag}_bench_n$n.json 2> $out/${tag}_bench_n$n.err
python - $out/${tag}_bench_n$n.json <<'PY'
import json, sys
d = json.loads([l for l in open(sys.argv[1]) if l.startswith('{')][-1])
print(sys.argv[1], "ms/step %.3f  value %.0f  e2e %.3f ms" % (d["ms_per_step"], d["value"], d["e2e"]["ms_per_step"]), d["run"]["partition"][:100])
for k, v in d["per_rank_ms"].items(): print("   ", k, [round(x, 3) for x in v])
PY
tail -2 $out/${tag}_bench_n$n.err
