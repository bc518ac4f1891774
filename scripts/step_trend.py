"""Per-step device time over a benchmark run, with the likelihood / motion kernel times sampled every 20 steps
(does the step get slower as the cloud converges?)."""
import os, sys
import numpy as np, torch, ctypes as C
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from mcmh_localization_b200 import Localizer
from mcmh_localization_b200.params import YAML_PARAMS
from mcmh_localization_b200.synth import free_space_particles

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
K = 210
gm = bench.load_world()
poses = bench.trajectory(K + 1)
scans, angles = bench.make_scans(gm, poses, 360)
loc = Localizer(params=YAML_PARAMS, mode="MHMCL", seed=1, resample_mode="fixed")
loc.load_map(gm)
loc.set_particles(free_space_particles(gm, n))
loc.stage_scans(scans, angles)
loc.predict(poses[0])
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
h = loc.h

def dev_time(fn, reps=3):
    ts = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(); a.record(); fn(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return min(ts)

for k in range(1, K + 1):
    flush.zero_()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); loc.step_staged(poses[k], k); b.record(); torch.cuda.synchronize()
    t = a.elapsed_time(b)
    if k % 20 == 0 or k < 4:
        sc = loc.score_post
        tl = dev_time(lambda: h.call("mcl_likelihood", *[C.c_void_p(x.data_ptr()) for x in loc.cur], n, C.c_void_p(sc.data_ptr())))
        p = loc.particles()
        print("step %3d  %.3f ms   likelihood(1 set) %.3f ms   spread x %.2f y %.2f m  distinct cells %d" % (
            k, t, tl, p[:, 0].std(), p[:, 1].std(),
            len(np.unique((np.floor((p[:, 0] + 10) / 0.05) * 384 + np.floor((p[:, 1] + 10) / 0.05)).astype(np.int64)))), flush=True)
