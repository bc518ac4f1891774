"""Span of the persistent tail kernel per step over the benchmark's trajectory (MCL_TAIL_PROF=1), with the tiles of the
two exact passes that were not CLEAN (kind, verification failures)."""
import ctypes as C
import os
import sys
os.environ["MCL_TAIL_PROF"] = "1"
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import load_world, make_scans, trajectory
from mcmh_localization_b200 import Localizer
from mcmh_localization_b200.params import YAML_PARAMS
from mcmh_localization_b200.synth import free_space_particles

mode = sys.argv[1] if len(sys.argv) > 1 else "reference"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 1_000_000
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 210
gm = load_world()
poses = trajectory(steps + 1)
scans, angles = make_scans(gm, poses, 360)
loc = Localizer(params=YAML_PARAMS, mode="MHMCL", seed=2024, resample_mode=mode)
loc.load_map(gm)
loc.set_particles(free_space_particles(gm, n, seed=1234))
loc.stage_scans(scans, angles)
loc.predict(poses[0])
for k in range(1, steps):
    loc.step_staged(poses[k], k)
    out = (C.c_uint64 * (1024 * 32))()
    g = C.c_int(0)
    loc.h.call("mcl_tail_prof", out, C.byref(g))
    full = np.array(out[: g.value * 32], dtype=np.float64).reshape(g.value, 32)
    a = (full[:, :12] - full[:, 0].min()) / 1e3
    if k % 10 == 0 or k < 5:
        st = a[:, 1:12].max(axis=0)
        txt = "step %3d span %6.1f us  stage ends: %s" % (k, a[:, 11].max(), " ".join("%.0f" % v for v in st))
        if mode == "reference":
            for base, nm in ((16, "p1"), (20, "p2")):
                items = full[:, 24 + (base == 20)].astype(int)
                odd = ["%d%s%s%d" % (v, "CTM"[items[v] // 100000], "F" if (items[v] // 1000) % 100 else "", items[v] % 1000)
                       for v in range(len(items)) if items[v] != 1]
                txt += "  %s[%s]" % (nm, " ".join(odd))
        print(txt, flush=True)
