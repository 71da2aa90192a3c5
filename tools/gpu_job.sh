#!/bin/bash
cd "$GRAFT_REPO_ROOT"
timeout 900 python -m pytest tests -m gpu -q -x --timeout=600 --timeout-method=thread 2>&1 | tail -8 > gpurun_out/pytest6.log
for a in "64 1000000 65536 gossip_np 2" "256 1000000 262144 gossip_np 2" "1024 1000000 262144 gossip_np 2" "1024 1000000 262144 adversarial_np 2"; do timeout 200 python tools/prof_run.py $a; done > gpurun_out/prof6.log 2>&1
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/l6_1024.csv python tools/prof_run.py 1024 1000000 262144 gossip_np 1 >/dev/null 2>&1
cat gpurun_out/pytest6.log gpurun_out/prof6.log
