#!/bin/bash
cd "$GRAFT_REPO_ROOT"
timeout 900 python -m pytest tests -m gpu -q -x --timeout=600 --timeout-method=thread 2>&1 | tail -5 > gpurun_out/pytest10.log
for a in "64 1000000 65536 adversarial_np 2" "1024 1000000 262144 adversarial_np 2"; do timeout 200 python tools/prof_run.py $a | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); p=d['passes'][-1]; print(d['M'],d['gen'],{k:p[k] for k in ('ms','events_per_s','ms_can_see','ms_rounds_kernel','ms_decide_fame')})
"; done > gpurun_out/prof10.log 2>&1
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/l10_1024adv.csv python tools/prof_run.py 1024 1000000 262144 adversarial_np 1 >/dev/null 2>&1
cat gpurun_out/pytest10.log gpurun_out/prof10.log
bash tools/gpu_job_c45.sh 1 "c4 c5"
