"""Small driver for `ncu --metrics gpu__time_duration.sum`: three per-chunk can_see scans
(the e2e pattern) and one scan over the whole trace (the resident pattern)."""
import sys, os
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(R, 'py-swirld_b200')); sys.path.insert(0, R)
import bench
from swirld_b200 import engine
from swirld_b200.traces import chunks
wl = bench.WORKLOADS['c3']; tr = bench.make_trace(wl, 1)
e = engine.Engine(64, tr.N)
for first, cnt in list(chunks(tr.N, 65536))[:3]:
    e.append_trace(tr, first, cnt); e.divide_rounds(first, cnt); e.decide_fame()
e.reset()
e.append_trace(tr)
first, cnt = next(iter(chunks(tr.N, 65536)))
e.divide_rounds(first, cnt); e.decide_fame()
e.sync()
