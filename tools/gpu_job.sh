#!/bin/bash
cd "$GRAFT_REPO_ROOT"
timeout 900 python -m pytest tests -m gpu -q -x --timeout=600 --timeout-method=thread 2>&1 | tail -8 > gpurun_out/pytest8.log
for a in "64 1000000 65536 gossip_np 2" "64 1000000 65536 adversarial_np 2" "256 1000000 262144 gossip_np 2" "1024 1000000 262144 gossip_np 2" "1024 1000000 262144 adversarial_np 2"; do timeout 200 python tools/prof_run.py $a; done > gpurun_out/prof8.log 2>&1
timeout 600 python bench.py --steps 5 --warmup 3 --views 1,8,32 --views-events 262144 --no-python-reference > gpurun_out/bench8_c3.log 2>&1
cat gpurun_out/pytest8.log gpurun_out/prof8.log; grep -o '"views": \[.*\]' gpurun_out/bench8_c3.log | head -3
