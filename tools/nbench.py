import sys, importlib, numpy as np
sys.path.insert(0,'.'); sys.path.insert(0,'tests')
import torch
pkg=importlib.import_module("rtl-power-fftw_b200")
for N in (100, 500, 1000, 3000, 8190, 10000, 100000, 131072, 1048576):
    nspec=max(4, (512<<20)//(2*N)); nbytes=2*N*nspec
    dev=torch.empty(nbytes,dtype=torch.uint8,device="cuda")
    pkg.synth_fill_device(dev.data_ptr(), nbytes, 0, seed=1, mode=1); torch.cuda.synchronize()
    with pkg.Context(N) as c:
        ts=[]
        for i in range(4):
            c.reset(nspec); e0=torch.cuda.Event(enable_timing=True); e1=torch.cuda.Event(enable_timing=True)
            s=torch.cuda.Stream(); c.sync()
            import time; t0=time.perf_counter(); c.push_device(dev.data_ptr(), nbytes); c.sync(); ts.append(time.perf_counter()-t0)
        print("N=%8d %-10s spectra=%8d %.3f ms %.2f Gsamples/s" % (N, c.plan_name, nspec, min(ts)*1e3, nspec*N/min(ts)/1e9), flush=True)
    del dev
