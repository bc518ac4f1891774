"""Host-side cost of each stage of the sharded step (run under torchrun)."""
import os, sys, time
import numpy as np, torch, torch.distributed as dist
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from mcmh_localization_b200.sharded import ShardedLocalizer
from mcmh_localization_b200.params import YAML_PARAMS as P
from mcmh_localization_b200.synth import free_space_particles
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
n = 1_000_000
gm = bench.load_world()
K = 40
poses = bench.trajectory(K + 1); scans, angles = bench.make_scans(gm, poses, 360)
sh = ShardedLocalizer(device=local, params=P, mode="MHMCL", seed=3)
sh.load_map(gm); sh.set_particles(free_space_particles(gm, n, seed=rank)); sh.stage_scans(scans, angles)
sh.predict(poses[0])
for k in range(1, 6):
    sh.step_staged(poses[k], k)
torch.cuda.synchronize()
acc = {"predict": 0, "update": 0, "estimate": 0, "resample": 0}
dev = {"predict": 0, "update": 0, "estimate": 0, "resample": 0}
t_all0 = time.perf_counter()
for k in range(6, K):
    for name, fn in (("predict", lambda: sh.predict(poses[k])), ("update", lambda: sh.update_staged(k)),
                     ("estimate", lambda: sh.estimate_async(sh.est18)), ("resample", lambda: sh.resample())):
        t0 = time.perf_counter(); fn(); acc[name] += time.perf_counter() - t0
        torch.cuda.synchronize(); dev[name] += time.perf_counter() - t0
t_all = time.perf_counter() - t_all0
if rank == 0:
    m = K - 6
    print("per step (ms): " + "  ".join("%s host %.3f total %.3f" % (k, 1e3 * acc[k] / m, 1e3 * dev[k] / m) for k in acc))
    print("sum %.3f ms/step (with a sync after every stage)" % (1e3 * t_all / m))
dist.destroy_process_group()
