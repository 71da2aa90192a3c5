#!/bin/bash
# two (or more) GPUs: one hashgraph sharded over the ranks
cd "$GRAFT_REPO_ROOT"
N=${1:-2}
nvidia-smi -L > gpurun_out/multi_gpus.log 2>&1
timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -q -x 2>&1 | tail -15 > gpurun_out/pytest_multi_$N.log
cat gpurun_out/pytest_multi_$N.log
for ev in 1048576; do
  timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus $N --workload c4 --events $ev --steps 3 --warmup 1 --no-python-reference > gpurun_out/bench_c4_${ev}_g$N.log 2>&1
  tail -c 600 gpurun_out/bench_c4_${ev}_g$N.log
done
