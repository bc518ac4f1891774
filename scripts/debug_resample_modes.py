"""Resampling time at 1M: reference-f32 (exact parallel emulation) vs fixed point."""
import os, sys, numpy as np, torch, ctypes as C
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from mcmh_localization_b200 import parallel_utils as pu
n = 1_000_000
c = pu._ctx()
w = torch.from_numpy(np.exp(np.random.RandomState(0).normal(0, 1, n)).astype(np.float32)).to(c.device)
idx = torch.empty(n, dtype=torch.int32, device=c.device)
for mode in (0, 1):
    for _ in range(3): c.h.call("mcl_resample_indices", C.c_void_p(w.data_ptr()), n, n, 3e-7, mode, C.c_void_p(idx.data_ptr()))
    torch.cuda.synchronize(); ts = []
    for _ in range(10):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); c.h.call("mcl_resample_indices", C.c_void_p(w.data_ptr()), n, n, 3e-7, mode, C.c_void_p(idx.data_ptr())); e1.record()
        torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    print("resample mode %d: %.1f us" % (mode, 1e3 * np.median(ts)))
