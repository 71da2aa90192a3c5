#!/bin/bash
# final single-GPU evidence run: bench lines for c1-c3 (+ stream / views legs), reference arm, launch list of the bench
# command, ncu --set full of the kernels the docs quote, cycle counters of the wide round kernel
cd "$GRAFT_REPO_ROOT"
O=gpurun_out/final; mkdir -p $O
timeout 900 python bench.py --stream --views 1,8,32 --views-events 262144 > $O/bench_c3.log 2>&1
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > $O/bench_c3_reference.log 2>&1
timeout 200 python bench.py --workload c1 --steps 10 --warmup 3 > $O/bench_c1.log 2>&1
timeout 300 python bench.py --workload c2 --steps 10 --warmup 3 > $O/bench_c2.log 2>&1
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file $O/launches_bench_c3.csv python bench.py --steps 2 --warmup 1 --no-python-reference --no-find-order > /dev/null 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_cs_pass -s 1 -c 1 -f -o $O/cs_pass2_1M python tools/prof_run.py 64 1000000 65536 gossip_np 1 > /dev/null 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_rounds_wide -s 1 -c 1 -f -o $O/rw1024 python tools/prof_run.py 1024 524288 262144 gossip_np 1 > /dev/null 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_cs_pass -s 1 -c 1 -f -o $O/cs_pass2_adv1024 python tools/prof_run.py 1024 400000 400000 adversarial_np 1 > /dev/null 2>&1
for a in "256 1000000 262144 gossip_np 2" "1024 1000000 262144 gossip_np 2" "1024 1000000 262144 adversarial_np 2"; do timeout 200 python tools/prof_run.py $a; done > $O/prof_wide.log 2>&1
ls -la $O; grep -c '^{' $O/bench_c3.log $O/bench_c3_reference.log $O/bench_c1.log $O/bench_c2.log
