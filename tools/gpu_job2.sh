#!/bin/bash
# several GPUs: one hashgraph sharded over the ranks
cd "$GRAFT_REPO_ROOT"
N=${1:-2}
timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -q -x 2>&1 | tail -15 > gpurun_out/pytest_multi_$N.log
cat gpurun_out/pytest_multi_$N.log
for wl in c4 c5; do
  timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus $N --workload $wl --events 1048576 --steps 3 --warmup 1 --no-python-reference > gpurun_out/bench_${wl}_1M_g$N.log 2>&1
  grep '^{' gpurun_out/bench_${wl}_1M_g$N.log | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print('$wl 1M', d['n_gpus'], 'value %.3g'%d['value'], 'ms/step %.1f'%d['ms_per_step'], 'e2e %.3g'%d['e2e']['value'], 'parity', d['parity'], d['kernel_ms_per_step'])
" || tail -5 gpurun_out/bench_${wl}_1M_g$N.log
done
