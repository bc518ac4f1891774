"""One launch each of the coded shared-memory path and the global/L2 path on map_house (for ncu)."""
import os, sys, numpy as np, torch, ctypes as C
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from mcmh_localization_b200 import Localizer
from mcmh_localization_b200.maps import load_npz
from mcmh_localization_b200.params import YAML_PARAMS as P
from mcmh_localization_b200.synth import free_space_particles, raycast_scan
gm = load_npz(os.path.join(ROOT, "tests", "golden", "map_house.npz"))
n = 1_000_000
parts = free_space_particles(gm, n)
scan, angles = raycast_scan(gm, parts[0])
loc = Localizer(params=P, mode="MHMCL", seed=1, resample_mode="fixed")
loc.load_map(gm); loc.set_particles(parts); loc.set_scan(scan, angles=angles)
h = loc.h
args = [C.c_void_p(t.data_ptr()) for t in loc.cur] + [n, C.c_void_p(loc.score_post.data_ptr())]
for path in (0, 1, 0, 1):
    h.call("mcl_set_likelihood_path", path)
    h.call("mcl_likelihood", *args)
torch.cuda.synchronize()
print("ok")
