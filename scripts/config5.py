"""BASELINE config 5: global localisation (kidnapped robot) -- uniformly initialised particles on a
4096 x 4096 occupancy map (map_house tiled 11 x 11 and cropped), sharded over the ranks.
Run under torchrun; argv[1] = particles per rank (default 6.25M -> 50M on 8 GPUs)."""
import os, sys, time
import numpy as np, torch, torch.distributed as dist
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from mcmh_localization_b200 import Localizer
from mcmh_localization_b200.sharded import ShardedLocalizer
from mcmh_localization_b200.maps import load_npz, tiled_map
from mcmh_localization_b200.params import YAML_PARAMS as P
from mcmh_localization_b200.synth import free_space_particles, raycast_scan

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
n = int(sys.argv[1]) if len(sys.argv) > 1 else 6_250_000
side = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
t0 = time.time()
base = load_npz(os.path.join(ROOT, "tests", "golden", "map_house.npz"))
occ = np.tile(base.occ, (11, 11))[:side, :side]
loc = (ShardedLocalizer(device=local, params=P, mode="MHMCL", seed=7) if world > 1
       else Localizer(device=local, params=P, mode="MHMCL", seed=7, resample_mode="fixed"))
loc.load_map(np.ascontiguousarray(occ), base.resolution, (-0.5 * side * base.resolution, -0.5 * side * base.resolution),
             gpu_edt=True)                        # exact EDT on the device (SciPy: ~3 s for 4096 x 4096)
gm = loc.map
t_map = time.time() - t0
loc.init_uniform(n)
pose = free_space_particles(gm, 1, seed=3)[0]
steps = 8
ts = []
for k in range(steps):
    scan, angles = raycast_scan(gm, pose)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t1 = time.perf_counter()
    est = loc.step(pose, scan, angles=angles)
    torch.cuda.synchronize()
    ts.append(time.perf_counter() - t1)
    pose = pose + np.array([0.02 * np.cos(pose[2]), 0.02 * np.sin(pose[2]), 0.01])
valid = int(np.isfinite(scan).sum())
if rank == 0:
    ms = 1e3 * float(np.median(ts[2:]))
    print("config5: %d x %d map (EDT %.1f s), %d ranks x %d particles = %.1fM, step %.2f ms, %.3e evals/s, est (%.2f, %.2f) true (%.2f, %.2f)" % (
        side, side, t_map, world, n, world * n / 1e6, ms, world * n * valid * 2 / (ms * 1e-3), est[0], est[1], pose[0], pose[1]), flush=True)
if world > 1:
    dist.destroy_process_group()
