"""Stage stamps of the tail kernel on every rank of a sharded run (torchrun; MCL_TAIL_PROF=1): where the ranks wait for
each other in the reference arithmetic (the exact passes are handed from rank to rank)."""
import ctypes as C
import os
import sys
os.environ["MCL_TAIL_PROF"] = "1"
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from mcmh_localization_b200.sharded import ShardedLocalizer
from mcmh_localization_b200.params import YAML_PARAMS as P
from mcmh_localization_b200.synth import free_space_particles

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
mode = sys.argv[1] if len(sys.argv) > 1 else "reference"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 1_000_000
gm = bench.load_world()
poses = bench.trajectory(30)
scans, angles = bench.make_scans(gm, poses, 360)
sh = ShardedLocalizer(device=local, params=P, mode="MHMCL", seed=99, resample_mode=mode)
sh.load_map(gm)
sh.set_particles(free_space_particles(gm, n, seed=5 + rank))
sh.stage_scans(scans, angles)
sh.predict(poses[0])
names = ["S1", "bar1", "S2", "bar2", "S3a", "S3b", "S3c/pass1", "bar3", "S4/pass2", "bar4", "S5"]
for k in range(1, 20):
    sh.step_staged(poses[k], k)
    out = (C.c_uint64 * (1024 * 32))()
    g = C.c_int(0)
    sh.h.call("mcl_tail_prof", out, C.byref(g))
full = np.array(out[: g.value * 32], dtype=np.float64).reshape(g.value, 32)
a = (full[:, :12] - full[:, 0].min()) / 1e3
ends = a[:, 1:12].max(axis=0)
for r in range(world):
    dist.barrier()
    if r == rank:
        print("rank %d span %.1f us; stage ends (max over CTAs): %s" % (rank, a[:, 11].max(), "  ".join("%s %.0f" % (nm, e) for nm, e in zip(names, ends))), flush=True)
dist.destroy_process_group()
