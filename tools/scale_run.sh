#!/bin/bash
# weak / strong scaling lines on one multi-GPU box: tools/scale_run.sh <max_gpus> <out.jsonl>
maxn=$1; out=$2; : > $out
run() {  # n, extra args...
  n=$1; shift
  if [ "$n" = 1 ]; then python bench.py --gpus 1 --no-cpu --no-extras "$@" 2>/dev/null | tail -1 >> $out
  else python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 1000)) bench.py --gpus $n --no-cpu --no-extras "$@" 2>/dev/null | tail -1 >> $out; fi
}
for n in 1 2 4 8; do [ $n -le $maxn ] && run $n --steps 20 --warmup 3; done                                   # weak, 65536 envs / GPU
for tot in 65536 16384 4096 1024 256; do for n in 1 2 4 8; do [ $n -le $maxn ] && run $n --steps 20 --warmup 3 --scaling strong --total-envs $tot; done; done
python - $out <<'PY'
import json, sys
for l in open(sys.argv[1]):
    try: d = json.loads(l)
    except Exception: print("bad line", l[:80]); continue
    print("%-6s n=%d total_envs=%-7d value %.4g  e2e %.4g  step %.3f ms  e2e %.3f ms" % (d["scaling"], d["n_gpus"], d["config"]["total_envs"], d["value"], d["e2e"]["value"], d["ms_per_step"], d["e2e"]["ms_per_step"]))
PY
