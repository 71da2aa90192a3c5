"""How often does the cluster round kernel hand a chunk back to the grid-wide kernel?  usage: rc_handover.py M N K gen"""
import os, sys
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(R, 'py-swirld_b200')); sys.path.insert(0, R)
from swirld_b200 import engine, traces
from swirld_b200.traces import chunks
M, N, K, gen = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3]), sys.argv[4]
tr = getattr(traces, gen)(M, N, 3)
e = engine.Engine(M, tr.N)
e.append_trace(tr)
for rep in range(2):
    e.rewind(); e.debug_counters()
    for first, cnt in chunks(tr.N, K):
        e.divide_rounds(first, cnt); e.decide_fame()
    c = e.debug_counters(); st = e.stats()
print({"M": M, "N": N, "K": K, "gen": gen, "steps": int(c[6]), "cluster_launches": int(c[7]), "handed_over": int(c[15]),
       "ms_rounds_kernel": round(st["ms_rounds_kernel"], 3), "max_round": e.max_round})
