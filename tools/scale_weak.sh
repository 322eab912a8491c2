#!/bin/bash
# weak-scaling lines of one workload on a multi-GPU box: tools/scale_weak.sh <workload> <max_gpus> <out.jsonl>
w=$1; maxn=$2; out=$3; : > $out
for n in 1 2 4 8; do
  [ $n -le $maxn ] || continue
  if [ $n = 1 ]; then python bench.py --gpus 1 --no-cpu --no-extras --steps 20 --warmup 3 --workload $w 2>/dev/null | tail -1 >> $out
  else python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 1000)) bench.py --gpus $n --no-cpu --no-extras --steps 20 --warmup 3 --workload $w 2>/dev/null | tail -1 >> $out; fi
done
python - $out <<'PY'
import json, sys
for l in open(sys.argv[1]):
    d = json.loads(l)
    print("%s n=%d value %.4g e2e %.4g step %.3f ms e2e %.3f ms" % (d["config"]["workload"], d["n_gpus"], d["value"], d["e2e"]["value"], d["ms_per_step"], d["e2e"]["ms_per_step"]))
PY
