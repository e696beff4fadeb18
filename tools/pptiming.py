import ctypes as C, importlib, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
pkg = importlib.import_module("rtl-power-fftw_b200")
L = pkg.lib()
N, nspec = 4096, 262144
dev = torch.empty(2 * N * nspec, dtype=torch.uint8, device="cuda")
pkg.synth_fill_device(dev.data_ptr(), dev.numel(), 0, seed=1, mode=1)
torch.cuda.synchronize()
ctx = pkg.Context(N)
out = (C.c_ulonglong * 4)()
for rep in range(2):
    ctx.reset(nspec); ctx.push_device(dev.data_ptr(), dev.numel()); ctx.sync()
    L.rpf_debug_pp_timing(out, 1)
sections = out[3] / 2     # each PP_ADD(3,0,2) adds 2 per (warp, iteration): phases 1,2 timed
iters = sections
print("per warp-iteration (phases 1+2 only): wait-for-turn %.0f  math %.0f  memory(between) %.0f cycles; iterations %d"
      % (out[0] / iters, out[1] / iters, out[2] / iters, iters))
