"""Timing of k_motion alone on the cloud of a running filter, with the attempt histogram (att == 0: stuck)."""
import os, sys, time
import numpy as np, torch, ctypes as C
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from mcmh_localization_b200 import Localizer
from mcmh_localization_b200.params import YAML_PARAMS as P
from mcmh_localization_b200.synth import free_space_particles

n = 1_000_000
gm = bench.load_world()
K = 130
poses = bench.trajectory(K + 1)
scans, angles = bench.make_scans(gm, poses, 360)
loc = Localizer(params=P, mode="MHMCL", seed=1, resample_mode="fixed")
loc.load_map(gm)
loc.set_particles(free_space_particles(gm, n))
loc.stage_scans(scans, angles)
loc.predict(poses[0])
h = loc.h
att = torch.zeros(n, dtype=torch.int32, device="cuda")
ox, oy, ot = [torch.empty(n, dtype=torch.float64, device="cuda") for _ in range(3)]
d3 = (C.c_double * 3)(0.0, 0.02, 0.01)
def run(max_att, with_att=True):
    ts = []
    for _ in range(3):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(); a.record()
        h.call("mcl_predict", *[C.c_void_p(t.data_ptr()) for t in loc.cur], n, d3, 7, 3, 0, None, 0, max_att,
               C.c_void_p(ox.data_ptr()), C.c_void_p(oy.data_ptr()), C.c_void_p(ot.data_ptr()),
               C.c_void_p(att.data_ptr()) if with_att else None)
        b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
    return min(ts)
for k in range(1, K):
    if k in (1, 5, 10, 20, 40, 60, 80, 120):
        t1000 = run(1000); a = att.cpu().numpy()
        t1 = run(1)
        hst = np.bincount(np.minimum(a, 40), minlength=41)
        print("step %3d  k_motion %.3f ms (1 attempt only: %.3f ms)  att==0 %7d  att==1 %7d  2..32 %6d  >32 %6d  max %d" % (
            k, t1000, t1, hst[0], hst[1], hst[2:33].sum(), hst[33:].sum(), a.max()), flush=True)
    loc.step_staged(poses[k], k)
