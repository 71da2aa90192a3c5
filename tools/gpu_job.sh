#!/bin/bash
cd "$GRAFT_REPO_ROOT"
timeout 200 python bench.py --no-python-reference --no-find-order --steps 3 --warmup 3 --views 1,8,32 --views-events 262144 > gpurun_out/bench17.json 2> gpurun_out/bench17.err
echo "bench rc=$?" >> gpurun_out/bench17.err
timeout 100 python tools/rounds_cycles.py > gpurun_out/cycles17.log 2>&1
SW_RC_MB=0 timeout 100 python tools/rounds_cycles.py > gpurun_out/cycles17_nomb.log 2>&1
timeout 600 python -m pytest tests/test_gpu_parity.py -q -x --timeout=200 --timeout-method=thread 2>&1 | tail -6 > gpurun_out/pytest17.log
tail -3 gpurun_out/bench17.err; python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/bench17.json') if l.startswith('{')][0])
print(d['value'], d['ms_per_step'], d['parity'], d.get('kernel_ms_per_step'))
print(json.dumps(d.get('views')))
PY
cat gpurun_out/cycles17.log gpurun_out/cycles17_nomb.log gpurun_out/pytest17.log
