#!/bin/bash
# final evidence of the round on one B200: the bench line the driver will ask for (+ stream / views legs), the ncu launch
# list of the same command, one ncu --set full capture of the cluster round kernel, the whole GPU test suite
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out/final2
timeout 400 python bench.py --stream --views 1,8,32 --views-events 262144 > gpurun_out/final2/bench_c3.json 2> gpurun_out/final2/bench_c3.err
echo "bench rc=$?" >> gpurun_out/final2/bench_c3.err
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/final2/launches_bench_c3.csv python bench.py --steps 2 --warmup 1 --no-python-reference --no-find-order > gpurun_out/final2/ncu_list.log 2>&1
timeout 200 ncu --set full --clock-control none --import-source on -k regex:k_rounds_cluster -s 1 -c 1 -o gpurun_out/final2/rc_full -f python tools/prof_run.py 64 300000 65536 gossip_np 1 > gpurun_out/final2/ncu_full.log 2>&1
timeout 100 python tools/rounds_cycles.py > gpurun_out/final2/cycles.log 2>&1
timeout 900 python -m pytest tests -m gpu -q -x --timeout=300 --timeout-method=thread 2>&1 | tail -6 > gpurun_out/final2/pytest.log
tail -2 gpurun_out/final2/bench_c3.err; head -c 700 gpurun_out/final2/bench_c3.json; echo; cat gpurun_out/final2/cycles.log gpurun_out/final2/pytest.log; tail -2 gpurun_out/final2/ncu_full.log
run() {  tool=$1; tag=$2; shift 2
  timeout 250 compute-sanitizer --tool $tool --print-limit 10 python tools/triage.py --one "$@" > gpurun_out/final2/sanitize_${tool}_${tag}.log 2>&1
  echo "== $tool $tag rc=$? : $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY|"bad"' gpurun_out/final2/sanitize_${tool}_${tag}.log | tr '\n' ' ' | cut -c1-400)"; }
run memcheck m64c gossip 64 12000 4096 0 > gpurun_out/final2/sanitize.log 2>&1
run racecheck m64c gossip 64 8000 4096 0 >> gpurun_out/final2/sanitize.log 2>&1
run synccheck m64c gossip 64 8000 4096 0 >> gpurun_out/final2/sanitize.log 2>&1
cat gpurun_out/final2/sanitize.log
