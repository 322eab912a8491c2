#!/bin/bash
# A/B kernel builds on the GPU box: tools/ab.sh <workload> <envs> lib1.so lib2.so ...
w=$1; envs=$2; shift 2
for lib in "$@"; do
  PPR_B200_LIB=$PWD/$lib python bench.py --steps 3 --warmup 3 --no-cpu --no-extras --workload $w --envs $envs 2>/dev/null | tail -1 | \
    python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$lib', '$w', d['config']['envs_per_gpu'], 'fwd_ms %.3f bwd_ms %.3f step_ms %.3f value %.4g' % (d['kernels_ms']['rollout_forward'], d['kernels_ms']['rollout_backward'], d['ms_per_step'], d['value']))"
done
