#!/bin/bash
cd "$GRAFT_REPO_ROOT"
SW_DEBUG=1 timeout 150 python tools/prof_run.py 64 1000000 65536 gossip_np 2 > gpurun_out/prof13.log 2>&1
echo "prof rc=$?" >> gpurun_out/prof13.log
timeout 120 python tools/rounds_cycles.py > gpurun_out/cycles13.log 2>&1
echo "cycles rc=$?" >> gpurun_out/cycles13.log
timeout 200 python bench.py --no-python-reference --steps 5 --warmup 3 > gpurun_out/bench13.json 2> gpurun_out/bench13.err
echo "bench rc=$?" >> gpurun_out/bench13.err
timeout 900 python -m pytest tests/test_gpu_parity.py -q -x --timeout=300 --timeout-method=thread 2>&1 | tail -15 > gpurun_out/pytest13.log
tail -5 gpurun_out/prof13.log; cat gpurun_out/cycles13.log; tail -3 gpurun_out/bench13.err; head -c 600 gpurun_out/bench13.json; echo; cat gpurun_out/pytest13.log
