#!/bin/bash
# quick look at bench.py lines on the GPU box: tools/bench_quick.sh [bench args...]
out=$(python bench.py --no-cpu --no-extras "$@" 2>gpurun_out/bench_err.log | tail -1)
echo "$out" | python -c "
import sys, json
try:
    d = json.loads(sys.stdin.read())
except Exception as e:
    print('NO JSON LINE'); print(open('gpurun_out/bench_err.log').read()[-3000:]); sys.exit(0)
print(d['config']['workload'], 'n_gpus', d['n_gpus'], 'envs/gpu', d['config']['envs_per_gpu'], 'value %.4g' % d['value'], 'step %.3f ms (eager %.3f)' % (d['ms_per_step'], d['ms_per_step_eager']),
      'kernels fwd %.3f bwd %.3f' % (d['kernels_ms']['rollout_forward'], d['kernels_ms']['rollout_backward']), 'launches/step', d['gpu_launches_per_step'])
print('   e2e %.4g (%.3f ms, h2d %d B)  e2e_full %.4g (%.3f ms)  frac fwd+bwd %.3f step %.3f  refconv %s' % (d['e2e']['value'], d['e2e']['ms_per_step'], d['e2e']['h2d_bytes_per_step'],
      d['e2e_full_refs_from_host']['value'], d['e2e_full_refs_from_host']['ms_per_step'], d['roofline']['fwd_bwd_combined_frac'], d['roofline']['step_frac'], d.get('reference_convention')))
"
