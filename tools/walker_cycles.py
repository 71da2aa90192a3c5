import sys, time
import os; R=os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0,os.path.join(R,'py-swirld_b200')); sys.path.insert(0,R)
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
    c=e.debug_counters()
    st=e.stats()
    print('compute: proc %.1fM lbar %.1fM bbar %.1fM  nproc %d nlev %d nbatch %d'%(c[0]/1e6,c[1]/1e6,c[2]/1e6,c[3],c[4],c[5]))
    print('  per level: proc %.0f lbar %.0f ; per batch bbar %.0f ; proc per call %.0f'%(c[0]/c[4], c[1]/c[4], c[2]/c[5], c[0]/max(1,c[3])))
    print('prep per batch: prep %.0f wait %.0f bbar %.0f'%(c[9]/c[5],c[10]/c[5],c[11]/c[5]))
    print('stats', {k: (round(v,2) if isinstance(v,float) else v) for k,v in st.items()})
