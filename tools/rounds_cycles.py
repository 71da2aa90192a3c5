import sys, os
R=os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0,os.path.join(R,'py-swirld_b200')); sys.path.insert(0,R)
import numpy as np, bench
from swirld_b200 import engine
from swirld_b200.traces import chunks
wl=bench.WORKLOADS['c3']; tr=bench.make_trace(wl,1)
e=engine.Engine(64, tr.N)
e.append_trace(tr)
for rep in range(1):
    e.rewind(); e.debug_counters()
    for first,cnt in chunks(tr.N, 65536):
        e.divide_rounds(first,cnt); e.decide_fame()
    c=e.debug_counters(); st=e.stats()
    n=max(1,c[6])
    print('cta 0 warp 0: steps %d; cycles per step: prep %.0f  masks %.0f  sync1 %.0f  test %.0f  sync2 %.0f  update %.0f'%((c[6],)+tuple(c[i]/n for i in range(6))))
    print('  (k_rounds_cluster: prep = control, masks = masks + push, sync1 = deferred work + test phase 1 + wait for the masks, test = test phase 2 + results, sync2 = cluster barrier (0 with st.async); k_rounds_batch: as named)')
    print('all warps: tests %d (deferred %d), mean %.0f cycles, max %d; mask misses %d'%(c[11],c[12],c[10]/max(1,c[11]),c[8],c[9]))
    print('cluster kernel: %d launches, %d handed the rest of their chunk to k_rounds_batch' % (c[7], c[15]))
    print('gather: mean %.0f max %d'%(c[13]/max(1,c[11]), c[14]))
    print('stats', {k:(round(v,2) if isinstance(v,float) else v) for k,v in st.items()})
