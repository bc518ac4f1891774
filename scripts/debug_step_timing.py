"""Where does a filter step's time go?  Host enqueue time vs device time per stage (debug aid)."""
import os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from mcmh_localization_b200 import Localizer
from mcmh_localization_b200.params import YAML_PARAMS
from mcmh_localization_b200.synth import free_space_particles

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
gm = bench.load_world()
K = 30
poses = bench.trajectory(K + 1)
scans, angles = bench.make_scans(gm, poses, 360)
loc = Localizer(params=YAML_PARAMS, mode="MHMCL", seed=1, resample_mode="fixed")
loc.load_map(gm)
loc.set_particles(free_space_particles(gm, n))
loc.stage_scans(scans, angles)
loc.predict(poses[0])
for k in range(1, 6):
    loc.step_staged(poses[k], k)
torch.cuda.synchronize()

def timed(fn, reps=10):
    host, dev = [], []
    for _ in range(reps):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter(); e0.record(); fn(); e1.record(); t1 = time.perf_counter()
        torch.cuda.synchronize()
        host.append((t1 - t0) * 1e3); dev.append(e0.elapsed_time(e1))
    return np.median(host), np.median(dev)

k = [6]
def full():
    loc.step_staged(poses[k[0] % K + 1], k[0] % K + 1); k[0] += 1
print("full step        host %.3f ms  device %.3f ms" % timed(full))
print("predict          host %.3f ms  device %.3f ms" % timed(lambda: loc.predict(poses[(k[0] % K) + 1])))
print("update_staged    host %.3f ms  device %.3f ms" % timed(lambda: loc.update_staged(3)))
print("estimate_async   host %.3f ms  device %.3f ms" % timed(lambda: loc.estimate_async(loc.est18)))
print("resample         host %.3f ms  device %.3f ms" % timed(lambda: loc.resample()))
import ctypes as C
h = loc.h
sc = loc.score_post
print("likelihood only  host %.3f ms  device %.3f ms" % timed(lambda: h.call("mcl_likelihood", *[C.c_void_p(t.data_ptr()) for t in loc.cur], n, C.c_void_p(sc.data_ptr()))))
print("softmax only     host %.3f ms  device %.3f ms" % timed(lambda: h.call("mcl_softmax", C.c_void_p(sc.data_ptr()), n, C.c_void_p(loc.w_post.data_ptr()), None, None)))
print("use_scan+lik     host %.3f ms  device %.3f ms" % timed(lambda: (h.call("mcl_use_scan", 4), h.call("mcl_likelihood", *[C.c_void_p(t.data_ptr()) for t in loc.cur], n, C.c_void_p(sc.data_ptr())))))
# back-to-back steps, no per-step sync
torch.cuda.synchronize(); t0 = time.perf_counter()
for j in range(20):
    full()
t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
print("20 steps back-to-back: enqueue %.3f ms/step, total %.3f ms/step" % ((t1 - t0) * 50, (t2 - t0) * 50))
