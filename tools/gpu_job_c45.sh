#!/bin/bash
# configs 4 and 5 at full length on N GPUs: bash tools/gpu_job_c45.sh N "c4 c5"
cd "$GRAFT_REPO_ROOT"
N=${1:-1}
for wl in ${2:-c4 c5}; do
  if [ "$N" = "1" ]; then
    timeout 900 python bench.py --workload $wl --steps 2 --warmup 1 > gpurun_out/bench_${wl}_g$N.log 2>&1
  else
    timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29551 bench.py --gpus $N --workload $wl --steps 2 --warmup 1 --no-python-reference > gpurun_out/bench_${wl}_g$N.log 2>&1
  fi
  grep '^{' gpurun_out/bench_${wl}_g$N.log | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print('$wl', d['n_gpus'], 'value %.3g'%d['value'], 'ms/step %.1f'%d['ms_per_step'], 'e2e %.3g'%d['e2e']['value'], 'parity', d['parity'], d['kernel_ms_per_step'])
" || tail -5 gpurun_out/bench_${wl}_g$N.log
done
