"""Debug driver for the persistent tail kernel: one raw resampling call (mode, n from argv)."""
import ctypes as C
import sys
import numpy as np
import torch
sys.path.insert(0, ".")
from mcmh_localization_b200 import parallel_utils as pu

mode, n = int(sys.argv[1]), int(sys.argv[2])
c = pu._ctx()
rs = np.random.RandomState(1)
w = rs.uniform(0, 1, n).astype(np.float32)
wd = torch.from_numpy(w).to(c.device)
idx = torch.empty(n, dtype=torch.int32, device=c.device)
c.h.call("mcl_debug_tail_resample", C.c_void_p(wd.data_ptr()), n, 0.3 / n, mode, C.c_void_p(idx.data_ptr()), None)
torch.cuda.synchronize()
err = C.c_int(-1)
c.h.call("mcl_tail_status", C.byref(err))
print("ok mode", mode, "n", n, "err", err.value, "idx[:5]", idx[:5].tolist())
