#!/bin/bash
# development job for one gpurun call (edit freely)
cd "$GRAFT_REPO_ROOT"
timeout 1500 python -m pytest tests -m gpu -q -x --timeout=600 --timeout-method=thread 2>&1 | tail -25 > gpurun_out/pytest4.log
for a in "64 1000000 65536 gossip_np 2" "256 1000000 262144 gossip_np 2" "1024 1000000 262144 gossip_np 1"; do timeout 200 python tools/prof_run.py $a; done > gpurun_out/prof4.log 2>&1
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/l4_64.csv python tools/prof_run.py 64 1000000 65536 gossip_np 1 > /dev/null 2>&1
timeout 300 python bench.py --steps 3 --warmup 1 > gpurun_out/bench4_c3.log 2>&1
timeout 300 python bench.py --workload c4 --events 1048576 --steps 2 --warmup 1 > gpurun_out/bench4_c4.log 2>&1
cat gpurun_out/pytest4.log gpurun_out/prof4.log; tail -c 1500 gpurun_out/bench4_c3.log; tail -c 1500 gpurun_out/bench4_c4.log
