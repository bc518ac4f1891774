"""Counters of the motion kernel's rejection loop over the benchmark's steps (MCL_MOTION_STATS=1)."""
import os, sys
os.environ["MCL_MOTION_STATS"] = "1"
import ctypes as C
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from mcmh_localization_b200 import Localizer
from mcmh_localization_b200.params import YAML_PARAMS as P
from mcmh_localization_b200.synth import free_space_particles

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 210
gm = bench.load_world()
poses = bench.trajectory(steps + 1)
scans, angles = bench.make_scans(gm, poses, 360)
loc = Localizer(device=0, params=P, mode="MHMCL", seed=1)
loc.load_map(gm)
loc.set_particles(free_space_particles(gm, n, seed=1))
out = (C.c_ulonglong * 16)()
names = "fail0 stuck retried screen_rounds eval_rounds found hard evaluated warps".split()
print("step " + " ".join("%13s" % s for s in names))
for k in range(steps):
    loc.step(poses[k], scans[k], angles=angles)
    loc.h.call("mcl_debug_motion_stats", C.c_void_p(C.addressof(out)))
    if k < 20 or k % 10 == 0:
        print("%4d " % k + " ".join("%13d" % out[j] for j in range(9)), flush=True)
