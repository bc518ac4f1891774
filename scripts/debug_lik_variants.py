"""Time the G=1 likelihood kernel variants (MCL_LIK_VARIANT is read once per process)."""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
code = r'''
import os, sys, numpy as np, torch, ctypes as C
sys.path.insert(0, %r)
import bench
from mcmh_localization_b200 import Localizer
from mcmh_localization_b200.params import YAML_PARAMS as P
from mcmh_localization_b200.synth import free_space_particles
n = 1_000_000
gm = bench.load_world()
poses = bench.trajectory(3); scans, angles = bench.make_scans(gm, poses, 360)
loc = Localizer(params=P, mode="MHMCL", seed=1, resample_mode="fixed")
loc.load_map(gm); loc.set_particles(free_space_particles(gm, n)); loc.set_scan(scans[0], angles=angles)
h = loc.h; sc = loc.score_post
args = [C.c_void_p(t.data_ptr()) for t in loc.cur] + [n, C.c_void_p(sc.data_ptr())]
for _ in range(5): h.call("mcl_likelihood", *args)
torch.cuda.synchronize()
ts = []
for _ in range(20):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); h.call("mcl_likelihood", *args); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
print("variant", os.environ.get("MCL_LIK_VARIANT"), "median %%.1f us  min %%.1f us  checksum %%.6f" %% (1e3*np.median(ts), 1e3*min(ts), float(sc.double().sum())))
''' % ROOT
for v in sys.argv[1:] or ["0", "1", "2", "3", "4", "5", "6", "7"]:
    env = dict(os.environ, MCL_LIK_VARIANT=v)
    subprocess.run([sys.executable, "-c", code], env=env)
