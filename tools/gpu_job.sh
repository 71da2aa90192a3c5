#!/bin/bash
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out/final3
timeout 400 python bench.py --stream --views 1,8,32 --views-events 262144 > gpurun_out/final3/bench_c3.json 2> gpurun_out/final3/bench_c3.err
echo "bench rc=$?" >> gpurun_out/final3/bench_c3.err
tail -2 gpurun_out/final3/bench_c3.err; head -c 400 gpurun_out/final3/bench_c3.json
