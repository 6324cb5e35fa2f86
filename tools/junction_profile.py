import sys, time; sys.path.insert(0,'/root/repo')
import numpy as np
import opticalflowcontainer_b200 as ofb
from oracle import synth
eng=ofb.FlowEngine(64,64,1,0)
img=synth.synth_net(1080,1920,1)
for _ in range(3): eng.find_junctions(img,200,2.0,6)
t0=time.perf_counter(); r,c=eng.find_junctions(img,200,2.0,6,return_candidates=True); print('total ms',(time.perf_counter()-t0)*1e3,len(r),len(c))
import ctypes as C
from opticalflowcontainer_b200 import _lib
lib=_lib.load()
out=np.empty((len(c),2),np.float32); n=C.c_int()
t0=time.perf_counter(); lib.ofb_cluster_junctions(c.ctypes.data,len(c),6,out.ctypes.data,len(out),C.byref(n)); print('cluster ms',(time.perf_counter()-t0)*1e3)
