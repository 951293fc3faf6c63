import sys, time, json
sys.path.insert(0,'/root/repo')
import numpy as np
import adsel_b200
from adsel_b200 import native
h=native.Handle(0)
F=int(sys.argv[1]) if len(sys.argv)>1 else 2048
sizes=300_000+(np.arange(F)*977)%4001-2000
off=np.concatenate([[0],np.cumsum(sizes)]).astype(np.int64)
h.bev_generate(off,10,1)
bytes_=off[-1]*20
for agg in (True,False):
    for it in range(4):
        h.bev_build(aggregate_atomics=agg)
        ms=h.stage_ms()["bev_histogram"]
    print("agg",agg,"frames",F,"points",off[-1],"ms",ms,"GB/s",bytes_/ms/1e6,"frames/s",F/ms*1e3, "frac of 6441", bytes_/ms/1e6/6441)
