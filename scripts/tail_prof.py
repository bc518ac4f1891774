"""Stage timeline of the persistent tail kernel (MCL_TAIL_PROF=1): per-stage time, min/median/max over CTAs."""
import ctypes as C
import os
import sys
os.environ["MCL_TAIL_PROF"] = "1"
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import load_world, make_scans, trajectory
from mcmh_localization_b200 import Localizer
from mcmh_localization_b200.params import YAML_PARAMS
from mcmh_localization_b200.synth import free_space_particles

mode = sys.argv[1] if len(sys.argv) > 1 else "reference"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 1_000_000
gm = load_world()
poses = trajectory(40)
scans, angles = make_scans(gm, poses, 360)
loc = Localizer(params=YAML_PARAMS, mode="MHMCL", seed=2024, resample_mode=mode)
loc.load_map(gm)
loc.set_particles(free_space_particles(gm, n, seed=1234))
loc.stage_scans(scans, angles)
loc.predict(poses[0])
names = ["S1 sumexp", "bar1", "S2 weights+MH", "bar2", "S3a moments/prefix", "S3b central", "S3c pass1|tilesum", "bar3",
         "S4 pass2|scan", "bar4", "S5 search+gather"]
acc = []
for k in range(1, 30):
    loc.step_staged(poses[k], k)
    out = (C.c_uint64 * (1024 * 32))()
    g = C.c_int(0)
    loc.h.call("mcl_tail_prof", out, C.byref(g))
    full = np.array(out[: g.value * 32], dtype=np.float64).reshape(g.value, 32)
    a = full[:, :12]
    if k >= 10:
        acc.append(a - a[:, :1].min())
        last_full = full
a = np.mean(acc, axis=0) / 1e3          # us, relative to the earliest CTA start
print("mode", mode, "n", n, "grid", a.shape[0])
print("kernel span: %.1f us" % (a[:, 11].max()))
for j, nm in enumerate(names):
    d = a[:, j + 1] - a[:, j]
    print("%-20s start(max) %7.1f  dur min/med/max %6.1f %6.1f %6.1f" % (nm, a[:, j].max(), d.min(), np.median(d), d.max()))

if mode == "reference":
    t0 = last_full[:, 0].min()
    for base, nm in ((16, "pass1"), (20, "pass2")):
        print(nm, "per tile (us since kernel start): ready | lookback done | walked+published | tile done | fail*1000+items")
        for v in (0, 1, 2, 3, 4, 5, 8, 9, 17, 18, 35, 36, 70, 71, 100, 139):
            if v < last_full.shape[0]:
                r = (last_full[v, base:base + 4] - t0) / 1e3
                x = (last_full[v, (26 if base == 20 else 29):(29 if base == 20 else 32)] - t0) / 1e3
                print("  tile %3d  [loaded %6.1f scanned %6.1f classified %6.1f] %7.1f %7.1f %7.1f %7.1f   %d" % (
                    v, x[0], x[1], x[2], r[0], r[1], r[2], r[3], int(last_full[v, 24 + (base == 20)])))
if mode == "reference":
    for base, nm in ((16, "pass1"), (20, "pass2")):
        lb = (last_full[:, base + 1] - t0) / 1e3
        items = last_full[:, 24 + (base == 20)].astype(int)
        print(nm, "lookback-done per tile (v:us, *kind[C/T/M][F=failed]items):",
              " ".join("%d:%.0f%s" % (v, lb[v], "" if items[v] == 1 else "*%s%s%d" % ("CTM"[items[v] // 100000], "F" if (items[v] // 1000) % 100 else "", items[v] % 1000))
                       for v in range(len(lb))))

if mode == "reference":
    for v in range(last_full.shape[0]):
        if False:
            print("pass1 tile", v, "failed at item", int(last_full[v, 12]), "run e", int(last_full[v, 13]) - 1000, "exponent(c)",
                  int(last_full[v, 14]) - 1000, "Kn", int(last_full[v, 15]), "2^24 =", 1 << 24)
