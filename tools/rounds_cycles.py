import sys, os
R=os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0,os.path.join(R,'py-swirld_b200')); sys.path.insert(0,R)
os.environ['SW_DIVIDE_IMPL']='5'
import numpy as np, bench
from swirld_b200 import engine
from swirld_b200.traces import chunks
wl=bench.WORKLOADS['c3']; tr=bench.make_trace(wl,1)
e=engine.Engine(64, tr.N)
e.append_trace(tr)
for rep in range(2):
    e.rewind(); e.debug_counters()
    for first,cnt in chunks(tr.N, 65536):
        e.divide_rounds(first,cnt); e.decide_fame()
    c=e.debug_counters(); st=e.stats()
    for b in (0,8):
        n=max(1,c[b+4])
        print('cta',b//8,'steps',c[b+4],'per step: eval %.0f sync %.0f update %.0f cycles ; evals by this warp %d'%(c[b]/n,c[b+1]/n,c[b+2]/n,c[b+3]))
    print('all warps: max single eval %d cycles, total misses %d, evals %d, mean eval %.0f cycles'%(c[5],c[6],c[13],c[7]/max(1,c[13])))
    print('slow tests (>20k cycles): %d, mean j %.1f, deferred %d, mean misses %.1f'%(c[9], c[10]/max(1,c[9]), c[11], c[12]/max(1,c[9])))
    print('stats', {k:(round(v,2) if isinstance(v,float) else v) for k,v in st.items()})
