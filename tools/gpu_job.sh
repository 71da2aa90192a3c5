#!/bin/bash
# development job for one gpurun call (edit freely): triage + timings + launch lists + ncu captures
cd "$GRAFT_REPO_ROOT"
(timeout 400 python tools/triage.py quick | cut -c1-260) > gpurun_out/triage3.log 2>&1
for a in "64 1000000 65536 gossip_np 2" "256 1000000 262144 gossip_np 2" "1024 1000000 262144 gossip_np 1" "1024 1000000 262144 adversarial_np 1" "128 500000 65536 adversarial_np 1"; do timeout 200 python tools/prof_run.py $a; done > gpurun_out/prof3.log 2>&1
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/l3_64.csv python tools/prof_run.py 64 1000000 65536 gossip_np 1 > /dev/null 2>&1
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/l3_256.csv python tools/prof_run.py 256 524288 262144 gossip_np 1 >/dev/null 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_cs_pass -s 1 -c 1 -f -o gpurun_out/cs_pass2 python tools/prof_run.py 64 262144 65536 gossip_np 1 > /dev/null 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_rounds_wide -c 1 -f -o gpurun_out/rw256 python tools/prof_run.py 256 524288 262144 gossip_np 1 > /dev/null 2>&1
cat gpurun_out/triage3.log gpurun_out/prof3.log; ls -la gpurun_out/*.ncu-rep
