#!/bin/bash
cd "$GRAFT_REPO_ROOT"
(timeout 300 python tools/triage.py quick | cut -c1-200) > gpurun_out/triage5.log 2>&1
for ct in 32 16 8; do SW_CS_CT=$ct timeout 100 python tools/prof_run.py 64 1000000 65536 gossip_np 2 | cut -c1-900; done > gpurun_out/prof5_ct.log 2>&1
for a in "256 1000000 262144 gossip_np 2" "1024 1000000 262144 gossip_np 2"; do timeout 200 python tools/prof_run.py $a; done > gpurun_out/prof5.log 2>&1
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/l5_256.csv python tools/prof_run.py 256 1000000 262144 gossip_np 1 >/dev/null 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/l5_1024.csv python tools/prof_run.py 1024 1000000 262144 gossip_np 1 >/dev/null 2>&1
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -x --timeout=300 --timeout-method=thread -k "find_order or fixture or checkpoint" 2>&1 | tail -3 > gpurun_out/pytest5.log
cat gpurun_out/triage5.log gpurun_out/prof5_ct.log gpurun_out/prof5.log gpurun_out/pytest5.log
