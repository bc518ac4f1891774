"""Attempt histogram and timing of the motion kernel on the particle cloud of a running filter."""
import os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from mcmh_localization_b200 import Localizer, parallel_utils as pu
from mcmh_localization_b200.params import YAML_PARAMS as P
from mcmh_localization_b200.synth import free_space_particles

n = 1_000_000
gm = bench.load_world()
K = 16
poses = bench.trajectory(K + 1)
scans, angles = bench.make_scans(gm, poses, 360)
loc = Localizer(params=P, mode="MHMCL", seed=1, resample_mode="fixed")
loc.load_map(gm)
loc.set_particles(free_space_particles(gm, n))
loc.stage_scans(scans, angles)
loc.predict(poses[0])
alpha = np.array([P["alpha1"], P["alpha2"], P["alpha3"], P["alpha4"]], dtype=np.float32)
for k in range(1, K):
    if k in (1, 5, 10, 15):
        parts = loc.particles()
        torch.cuda.synchronize(); t0 = time.perf_counter()
        out, att = pu.apply_motion_model_parallel(parts, loc.delta if k > 1 else (0.0, 0.02, 0.01), alpha, gm.occ.ravel(),
                                                  gm.resolution, gm.origin_x, gm.origin_y, gm.width, gm.height,
                                                  return_attempts=True)
        h = np.bincount(np.minimum(att, 40), minlength=41)
        uniq = len(np.unique(parts.view([('', parts.dtype)] * 3)))
        print("step %2d unique %7d  att==0 %7d  att==1 %7d  2..32 %6d  >32 %6d  max %d" % (
            k, uniq, h[0], h[1], h[2:33].sum(), h[33:].sum(), att.max()))
    loc.step_staged(poses[k], k)
torch.cuda.synchronize()
