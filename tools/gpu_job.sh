#!/bin/bash
cd "$GRAFT_REPO_ROOT"
for a in "64 1000000 65536 adversarial_np" "64 1000000 65536 gossip_np" "32 500000 65536 gossip_np" "16 200000 4096 adversarial_np" "64 1000000 4096 gossip_np"; do timeout 100 python tools/rc_handover.py $a; done > gpurun_out/handover.log 2>&1
for t in 2048 512 128; do SW_RC_MIN_N=$t timeout 120 python bench.py --workload c2 --no-python-reference --no-find-order --steps 5 --warmup 3 2>/dev/null | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('c2 min_n $t', d['value'], d['ms_per_step'], d['parity'], d['kernel_ms_per_step'])
"; done > gpurun_out/c2_minn.log 2>&1
cat gpurun_out/handover.log gpurun_out/c2_minn.log
