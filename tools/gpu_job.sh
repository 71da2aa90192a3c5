#!/bin/bash
cd "$GRAFT_REPO_ROOT"
timeout 200 python __graft_entry__.py --smoke > gpurun_out/smoke_final.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke_final.log
timeout 900 python -m pytest tests -m gpu -q -x --timeout=300 --timeout-method=thread 2>&1 | tail -5 > gpurun_out/pytest_final.log
timeout 120 python bench.py --no-python-reference --no-find-order --steps 5 --warmup 3 2>/dev/null | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('c3', d['value'], d['ms_per_step'], d['parity'], d['e2e']['value'])
" > gpurun_out/bench_final_check.log 2>&1
cat gpurun_out/smoke_final.log gpurun_out/pytest_final.log gpurun_out/bench_final_check.log
