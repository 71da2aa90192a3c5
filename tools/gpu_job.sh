#!/bin/bash
cd "$GRAFT_REPO_ROOT"
timeout 900 python -m pytest tests -m gpu -q -x --timeout=600 --timeout-method=thread 2>&1 | tail -5 > gpurun_out/pytest11.log
for a in "64 1000000 65536 gossip_np 2" "64 1000000 65536 adversarial_np 2" "1024 1000000 262144 adversarial_np 2"; do timeout 200 python tools/prof_run.py $a | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); p=d['passes'][-1]; print(d['M'],d['gen'],{k:p[k] for k in ('ms','events_per_s','ms_can_see','ms_rounds_kernel','ms_decide_fame')})
"; done > gpurun_out/prof11.log 2>&1
cat gpurun_out/pytest11.log gpurun_out/prof11.log
