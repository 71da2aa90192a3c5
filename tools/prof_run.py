"""One pass of divide_rounds + decide_fame over a synthetic trace (for ncu launch lists / quick timings).
    python tools/prof_run.py M N K [gossip_np|adversarial_np|gossip|...] [passes]"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "py-swirld_b200"))
from swirld_b200 import engine, traces  # noqa: E402
from swirld_b200.traces import chunks   # noqa: E402

M, N, K = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
gen = sys.argv[4] if len(sys.argv) > 4 else "gossip_np"
passes = int(sys.argv[5]) if len(sys.argv) > 5 else 1
t0 = time.time()
tr = getattr(traces, gen)(M, N, 1)
t_gen = time.time() - t0
e = engine.Engine(M, N)
e.append_trace(tr)
out = []
for p in range(passes):
    if p:
        e.rewind()
    s0 = e.stats()
    e.debug_counters(clear=True)
    e.record(0)
    for first, cnt in chunks(N, K):
        e.divide_rounds(first, cnt)
        e.decide_fame()
    e.record(1)
    ms = e.elapsed_ms(0, 1)
    s1 = e.stats()
    dbg = e.debug_counters(clear=True).tolist()
    out.append({"dbg": dbg[:11], "ms": round(ms, 3), "events_per_s": round(N / ms * 1e3), **{k: round(s1[k] - s0[k], 3) for k in s1 if k.startswith("ms_")}})
print(json.dumps({"M": M, "N": N, "K": K, "gen": gen, "gen_s": round(t_gen, 1), "max_round": e.max_round, "passes": out}))
