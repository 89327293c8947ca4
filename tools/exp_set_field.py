import sys, time
sys.path[:0]=['/root/repo','/root/repo/tests']
import numpy as np
from grasptrajopt_b200 import capi
ctx=capi.GtoContext(0)
for n in (128,256):
    c=(np.random.default_rng(0).random((n,n,n))<0.05).astype(np.float32)
    ctx.set_field(0,c,np.zeros(3),0.01)
    t0=time.perf_counter()
    for _ in range(3): ctx.set_field(0,c,np.zeros(3),0.01)
    print(f"set_field {n}^3: {(time.perf_counter()-t0)/3*1e3:.1f} ms")
ctx.close()
