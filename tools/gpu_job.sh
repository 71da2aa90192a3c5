#!/bin/bash
cd "$GRAFT_REPO_ROOT"
timeout 200 python bench.py --no-python-reference --no-find-order --steps 5 --warmup 3 --views 1,8,32 --views-events 262144 > gpurun_out/bench19.json 2> gpurun_out/bench19.err
echo "bench rc=$?" >> gpurun_out/bench19.err
timeout 100 python tools/rounds_cycles.py > gpurun_out/cycles19.log 2>&1
timeout 600 python -m pytest tests/test_gpu_parity.py -q -x --timeout=200 --timeout-method=thread 2>&1 | tail -6 > gpurun_out/pytest19.log
run() {  tool=$1; tag=$2; shift 2
  timeout 250 compute-sanitizer --tool $tool --print-limit 10 python tools/triage.py --one "$@" > gpurun_out/sanitize_${tool}_${tag}.log 2>&1
  echo "== $tool $tag rc=$? : $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY|"bad"' gpurun_out/sanitize_${tool}_${tag}.log | tr '\n' ' ' | cut -c1-400)"; }
run memcheck m64c gossip 64 12000 4096 0 > gpurun_out/sanitize19.log 2>&1
run racecheck m64c gossip 64 8000 4096 0 >> gpurun_out/sanitize19.log 2>&1
run synccheck m64c gossip 64 8000 4096 0 >> gpurun_out/sanitize19.log 2>&1
tail -3 gpurun_out/bench19.err; python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/bench19.json') if l.startswith('{')][0])
print(d['value'], d['ms_per_step'], d['parity'], d.get('kernel_ms_per_step'))
print(json.dumps(d.get('views')))
PY
cat gpurun_out/cycles19.log gpurun_out/pytest19.log gpurun_out/sanitize19.log
