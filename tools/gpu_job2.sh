#!/bin/bash
# several GPUs: one hashgraph sharded over the ranks; A/B of the can_see sharding
cd "$GRAFT_REPO_ROOT"
N=${1:-2}
for sh in 1 0; do
for wl in c4 c5; do
  SW_CS_SHARD=$sh timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus $N --workload $wl --events 1048576 --steps 3 --warmup 1 --no-python-reference --no-find-order > gpurun_out/bench_${wl}_1M_g${N}_shard$sh.log 2>&1
  grep '^{' gpurun_out/bench_${wl}_1M_g${N}_shard$sh.log | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print('$wl 1M shard=$sh', d['n_gpus'], 'value %.3g'%d['value'], 'ms/step %.1f'%d['ms_per_step'], 'e2e %.3g'%d['e2e']['value'], 'parity', d['parity'], d['kernel_ms_per_step'])
" || tail -5 gpurun_out/bench_${wl}_1M_g${N}_shard$sh.log
done; done
