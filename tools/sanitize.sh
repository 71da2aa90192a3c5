#!/bin/bash
# compute-sanitizer over the default paths on small traces (logs -> gpurun_out/sanitize_*.log; summaries go to profiles/)
cd "${GRAFT_REPO_ROOT:-.}"
run() {  # tool, tag, triage case...
  tool=$1; tag=$2; shift 2
  timeout 300 compute-sanitizer --tool $tool --print-limit 10 python tools/triage.py --one "$@" > gpurun_out/sanitize_${tool}_${tag}.log 2>&1
  echo "== $tool $tag rc=$? : $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY|"bad"' gpurun_out/sanitize_${tool}_${tag}.log | tr '\n' ' ' | cut -c1-400)"
}
run memcheck c1 gossip 4 2000 50 0
run memcheck m64 gossip 64 12000 4096 0
run memcheck wide96 gossip 96 6000 2000 0
run memcheck stream7 gossip 7 500 1 0
run racecheck c1 gossip 4 2000 50 0
run racecheck m64 gossip 64 8000 4096 0
run racecheck wide96 gossip 96 6000 2000 0
run synccheck m64 gossip 64 8000 4096 0
run synccheck wide96 gossip 96 6000 2000 0
