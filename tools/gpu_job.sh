#!/bin/bash
cd "$GRAFT_REPO_ROOT"
timeout 300 python bench.py --no-python-reference --no-find-order --steps 3 --warmup 3 --views 1,8,32 --views-events 262144 > gpurun_out/bench15.json 2> gpurun_out/bench15.err
echo "bench rc=$?" >> gpurun_out/bench15.err
timeout 600 python -m pytest tests/test_gpu_parity.py -q -x --timeout=300 --timeout-method=thread -k "batch or views or View" 2>&1 | tail -5 > gpurun_out/pytest15.log
timeout 200 ncu --set full --clock-control none --import-source on -k regex:k_rounds_cluster -s 1 -c 1 -o gpurun_out/rc_full -f python tools/prof_run.py 64 300000 65536 gossip_np 1 > gpurun_out/ncu15.log 2>&1
echo "ncu rc=$?" >> gpurun_out/ncu15.log
tail -3 gpurun_out/bench15.err; python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/bench15.json') if l.startswith('{')][0])
print(d['value'], d['ms_per_step'], d['parity'], d.get('kernel_ms_per_step'))
print(json.dumps(d.get('views')))
PY
cat gpurun_out/pytest15.log; tail -3 gpurun_out/ncu15.log
