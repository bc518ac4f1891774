"""BASELINE config 4: 1M chains x 32 MH iterations per scan, timing."""
import os, sys, numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from mcmh_localization_b200 import Localizer
from mcmh_localization_b200.params import YAML_PARAMS as P
from mcmh_localization_b200.synth import free_space_particles
n, iters = 1_000_000, 32
gm = bench.load_world()
poses = bench.trajectory(8); scans, angles = bench.make_scans(gm, poses, 360)
loc = Localizer(params=P, mode="MHMCL", seed=1, resample_mode="fixed")
loc.load_map(gm); loc.set_particles(free_space_particles(gm, n)); loc.stage_scans(scans, angles)
loc.predict(poses[0])
ts = []
for k in range(1, 7):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    loc.predict(poses[k]); loc.h.call("mcl_use_scan", k); loc.update_chain(None, iters=iters); loc.estimate_async(loc.est18); loc.resample()
    e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
valid = int(np.isfinite(scans[3]).sum())
ms = float(np.median(ts[1:]))
print("config 4: %d chains x %d MH iterations: %.2f ms per scan, %.3e likelihood evals/s" % (n, iters, ms, n * valid * (iters + 1) / (ms * 1e-3)))
