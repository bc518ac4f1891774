"""Multi-GPU check (run under torchrun, one rank per GPU): a ShardedLocalizer over R ranks must
reproduce a single-GPU Localizer of the same total size -- same Philox streams (keyed by the global
particle index), same softmax statistics, same global resampling -- in the fixed-point arithmetic or
(MCL_RESAMPLE=reference) in the reference's own sequential-float32 arithmetic continued from rank to rank."""
import os, sys
import numpy as np, torch, torch.distributed as dist
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from mcmh_localization_b200 import Localizer
from mcmh_localization_b200.sharded import ShardedLocalizer
from mcmh_localization_b200.params import YAML_PARAMS as P
from mcmh_localization_b200.synth import free_space_particles


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    n_local = int(sys.argv[1]) if len(sys.argv) > 1 else 50000
    n = n_local * world
    gm = bench.load_world()
    steps = 6
    poses = bench.trajectory(steps + 1)
    scans, angles = bench.make_scans(gm, poses, 360)
    parts = free_space_particles(gm, n, seed=5)
    mode = os.environ.get("MCL_EXCHANGE", "native")      # native | push | nccl
    arith = os.environ.get("MCL_RESAMPLE", "fixed")      # fixed | reference (native exchanges only)
    sh = ShardedLocalizer(device=local, params=P, mode="MHMCL", seed=99, peer_push=mode != "nccl",
                          native_comm=mode == "native", resample_mode=arith)
    sh.load_map(gm)
    sh.set_particles(parts[rank * n_local:(rank + 1) * n_local])
    ref = None
    if rank == 0:
        ref = Localizer(device=local, params=P, mode="MHMCL", seed=99, resample_mode=arith)
        ref.load_map(gm)
        ref.set_particles(parts)
    ok = True
    if rank == 0:
        print("exchange:", ("native peer-memory exchanges + peer-push" if sh.native else "NCCL scalars + peer-push")
              if sh.symm is not None else "NCCL all-to-all (%s)" % getattr(sh, "symm_error", "peer push disabled"), flush=True)
    for k in range(steps):
        est = sh.step(poses[k], scans[k], angles=angles)
        allp = sh.gather_particles()
        if rank == 0:
            rest = ref.step(poses[k], scans[k], angles=angles)
            rp = ref.particles()
            same = np.isclose(allp, rp, rtol=0, atol=1e-12).all(axis=1).mean()
            de = max(abs(est[0] - rest[0]), abs(est[1] - rest[1]))
            dc = np.abs(est[3] - rest[3]).max()
            print("step %d: identical particles %.6f  |d mean| %.2e  |d cov| %.2e" % (k, same, de, dc), flush=True)
            ok = ok and same > 0.9999 and de < 1e-9 and dc < 1e-9
    if mode == "native":
        # asymmetric MH and the MH chain (BASELINE config 4) on sharded particles == single GPU, in either arithmetic;
        # the scan ends with finish(): estimate + resampling as one launch of the tail kernel on every rank
        for lm, chain in (("AMHMCL", 0), ("MHMCL", 4)):
            sh2 = ShardedLocalizer(device=local, params=P, mode=lm, seed=7, resample_mode=arith)
            sh2.load_map(gm)
            sh2.set_particles(parts[rank * n_local:(rank + 1) * n_local])
            ref2 = None
            if rank == 0:
                ref2 = Localizer(device=local, params=P, mode=lm, seed=7, resample_mode=arith)
                ref2.load_map(gm)
                ref2.set_particles(parts)
            for k in range(3):
                ests = []
                for loc_ in (sh2, ref2):
                    if loc_ is None:
                        continue
                    loc_.predict(poses[k])
                    if chain:
                        loc_.update_chain(scans[k], angles=angles, iters=chain)
                    else:
                        loc_.update(scans[k], angles=angles)
                    ests.append(loc_.finish())
                allp = sh2.gather_particles()
                if rank == 0:
                    same = np.isclose(allp, ref2.particles(), rtol=0, atol=1e-12).all(axis=1).mean()
                    de = max(abs(ests[0][0] - ests[1][0]), abs(ests[0][1] - ests[1][1]))
                    print("%s%s step %d: identical particles %.6f  |d mean| %.2e" % (lm, " chain x%d" % chain if chain else "", k, same, de), flush=True)
                    ok = ok and same > 0.9999 and de < 1e-9
    flag = torch.tensor([1 if ok else 0], device="cuda")
    dist.broadcast(flag, 0)
    dist.destroy_process_group()
    if rank == 0:
        print("DIST_CHECK", "OK" if ok else "FAILED")
    sys.exit(0 if int(flag.item()) else 1)


if __name__ == "__main__":
    main()
