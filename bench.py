#!/usr/bin/env python3
"""bench.py -- BASELINE.json's metric on BASELINE.json's config.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl native|reference]

metric : particle*beam likelihood evaluations per second over the FULL filter step
         (predict + update[2 likelihood passes + softmax + MH] + estimate + resample), plus the
         step latency (ms_per_step).  Workload at N=1 = BASELINE configs[1]:
         1M particles x 360 beams, likelihood-field model, map_world, amhmcl.yaml parameters.
value  : inputs resident in HBM (scans pre-staged on the device, estimates left on the device).
e2e    : the same steps through the public Localizer API with HOST scan/odom buffers in and the
         host estimate out, wall clock per step (copies and sync inside the timed region).
One process per GPU (torchrun for N > 1); particles are sharded (weak scaling: per-GPU N fixed),
the map is replicated; timing = barrier + synchronize on both sides, max over ranks.
--impl reference times the CPU oracle port of the reference filter (oracle/, all host cores) on a
bounded sample of the same workload.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "particle_beam_likelihood_evals_per_sec"
UNIT = "evals/s"
START_POSE = np.array([-2.0, -0.5, 0.0])          # mcmh_localization.launch:25-27


def trajectory(steps):
    """Per step the robot advances 0.02 m and turns 0.01 rad."""
    poses = [START_POSE.copy()]
    for _ in range(steps):
        p = poses[-1]
        poses.append(np.array([p[0] + 0.02 * np.cos(p[2]), p[1] + 0.02 * np.sin(p[2]), p[2] + 0.01]))
    return poses


def load_world():
    from mcmh_localization_b200.maps import load_npz
    return load_npz(os.path.join(ROOT, "tests", "golden", "map_world.npz"))


def make_scans(gm, poses, beams):
    from mcmh_localization_b200.synth import raycast_scan
    scans, angles = [], None
    for k, p in enumerate(poses):
        r, angles = raycast_scan(gm, p, num_beams=beams, noise_sigma=0.01, seed=4321 + k)
        scans.append(r)
    return np.stack(scans), angles


# ------------------------------------------------------------------------------------------------
class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                 "--format=csv,noheader,nounits", "-lms", "50"], stdout=subprocess.PIPE,
                stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def mark(self):
        """Call at the start of the timed region: only later samples count."""
        self.first = len(self.lines)

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.06)            # let the sample covering the end of the region arrive
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        self.t.join(timeout=2)
        sm, mx, pw, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        lines = self.lines[getattr(self, "first", 0):] or self.lines[-2:]
        for ln in lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1])); pw.append(float(f[2]))
            except ValueError:
                continue
            for nme, v in zip(names, f[3:7]):
                if v == "Active":
                    reasons.add(nme)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None, "samples": len(sm), "reasons": sorted(reasons)}


# ------------------------------------------------------------------------------------------------
def cpu_filter_run(gm, particles, poses, scans, angles, params, min_seconds, min_steps, fixed_steps=None):
    """The oracle port of the reference filter (MHMCL) on the host cores; returns (evals/s, ms/step,
    steps, threads).  Step = node.odom_callback + node.lidar_callback arithmetic."""
    from oracle import clib, node_glue as ng
    mp = ng.load_map(gm.occ, gm.resolution, gm.origin_x, gm.origin_y)
    clib.set_threads(clib.max_threads())
    f = ng.ReferenceFilter(mp, params, particles, mode="MHMCL")
    n = len(particles)
    f.move_particles(poses[0])
    # one untimed warm-up step
    f.move_particles(poses[1], seed=1, step=1)
    f.update(scans[1], angles, seed=1, step=2)
    f.estimate()
    f.resample(0.5 / n)
    evals, t_total, k = 0, 0.0, 2
    while True:
        t0 = time.perf_counter()
        f.move_particles(poses[k], seed=1, step=3 * k)
        f.update(scans[k], angles, seed=1, step=3 * k + 1)
        f.estimate()
        f.resample(0.5 / n)
        t_total += time.perf_counter() - t0
        valid = int(np.sum(np.isfinite(scans[k]) & (scans[k] < params["max_range"])))
        evals += n * valid * 2
        k += 1
        done = k - 2
        if fixed_steps is not None:
            if done >= fixed_steps:
                break
        elif (t_total >= min_seconds and done >= min_steps) or k >= len(poses):
            break
    return evals / t_total, 1e3 * t_total / done, done, clib.max_threads()


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from mcmh_localization_b200.params import YAML_PARAMS
    from mcmh_localization_b200.synth import free_space_particles
    gm = load_world()
    n = args.cpu_sample
    steps, warm = args.steps, args.warmup
    poses = trajectory(steps + 3)
    scans, angles = make_scans(gm, poses, args.beams)
    parts = free_space_particles(gm, n, seed=1234)
    v, ms, done, threads = cpu_filter_run(gm, parts, poses, scans, angles, YAML_PARAMS, 0, 0, fixed_steps=steps)
    sample = "%d of %d particles x %d beams, %d full MHMCL steps, oracle C port (OpenMP) + NumPy glue" % (
        n, args.particles, args.beams, done)
    line = {
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
        "warmup": warm, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic", "config": workload_config(args, gm, n_cpu=n),
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


def workload_config(args, gm, n_cpu=None):
    c = {"workload": "%d particles/GPU x %d beams, likelihood-field model, MHMCL predict+update+estimate+resample "
                     "per step (BASELINE configs[1]), map_world %dx%d @ %.2f m, amhmcl.yaml parameters" % (
                         args.particles, args.beams, gm.width, gm.height, gm.resolution),
         "particles_per_gpu": args.particles, "beams": args.beams, "likelihood_passes_per_step": 2,
         "map": "map_world", "resample_mode": args.resample,
         "l2": "flushed between timed steps (256 MiB memset outside the per-step event pairs)"}
    if n_cpu is not None:
        c["cpu_sample_particles"] = n_cpu
    return c


# ------------------------------------------------------------------------------------------------
def run_native(args):
    # Libraries chat on stdout (NCCL announces its version there): keep the real stdout for the one JSON line and
    # send everything else written to fd 1 to stderr.
    os.environ["NCCL_DEBUG"] = os.environ.get("MCL_NCCL_DEBUG", "WARN")
    sys.stdout.flush()
    json_out = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    import torch
    import torch.distributed as dist
    from mcmh_localization_b200 import Localizer
    from mcmh_localization_b200.params import YAML_PARAMS
    from mcmh_localization_b200.synth import free_space_particles

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product has no CPU fallback (use --impl reference)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()             # nvidia-smi needs ~0.5 s to come up: start it before the set-up work
    gm = load_world()
    n = args.particles
    K, W = args.steps, args.warmup
    poses = trajectory(K + W + 1)
    scans, angles = make_scans(gm, poses, args.beams)
    valid = np.array([int(np.sum(np.isfinite(s) & (s < YAML_PARAMS["max_range"]))) for s in scans])

    if world > 1:
        from mcmh_localization_b200.sharded import ShardedLocalizer
        loc = ShardedLocalizer(device=local, params=YAML_PARAMS, mode="MHMCL", seed=2024,
                               resample_mode=args.resample)
    else:
        loc = Localizer(device=local, params=YAML_PARAMS, mode="MHMCL", seed=2024, resample_mode=args.resample)
    loc.load_map(gm)
    loc.set_particles(free_space_particles(gm, n, seed=1234 + rank))
    loc.stage_scans(scans, angles)
    lib, h = loc.h.lib, loc.h.h

    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    est_buf = torch.zeros((K + W + 1, 18), dtype=torch.float64, device=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    # ---- value leg: device-resident inputs -------------------------------------------------
    loc.predict(poses[0])
    ev0 = [torch.cuda.Event(enable_timing=True) for _ in range(K)]
    ev1 = [torch.cuda.Event(enable_timing=True) for _ in range(K)]
    iters = max(1, args.mh_iters)

    def staged_step(k):
        if iters == 1:
            loc.step_staged(poses[k], k, est_buf[k])
        else:       # config 4: k MH iterations per scan
            loc.predict(poses[k])
            loc.h.call("mcl_use_scan", int(k))
            loc.update_chain(None, iters=iters)
            loc.estimate_async(est_buf[k])
            loc.resample()

    for k in range(1, W + 1):
        staged_step(k)
    barrier()
    sampler.mark()
    launches0 = lib.mcl_launch_count(h)
    lib.mcl_timing_start(h)
    t_wall0 = time.perf_counter()
    for j in range(K):
        k = W + 1 + j
        flush.zero_()
        ev0[j].record()
        staged_step(k)
        ev1[j].record()
    barrier()
    t_wall = time.perf_counter() - t_wall0
    import ctypes as C
    lik_ms, lik_n = C.c_double(0), C.c_int64(0)
    lib.mcl_timing_stop(h, C.byref(lik_ms), C.byref(lik_n))
    lik_sets = int(lib.mcl_timing_sets(h)) or int(lik_n.value)      # particle sets scored by those launches
    launches = lib.mcl_launch_count(h) - launches0
    clocks = sampler.stop() if rank == 0 else None
    step_ms = np.array([a.elapsed_time(b) for a, b in zip(ev0, ev1)])
    total_ms = float(step_ms.sum())
    if world > 1:
        t = torch.tensor([total_ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        total_ms = float(t.item())
        lt = torch.tensor([launches], dtype=torch.int64, device=dev)
        dist.all_reduce(lt)
        launches = int(lt.item())
    passes = 2 if iters == 1 else iters + 1       # likelihood passes per step
    evals = float(np.sum(valid[W + 1:W + 1 + K])) * n * passes * world
    value = evals / (total_ms * 1e-3)

    # ---- e2e leg: public API, host buffers in, host estimate out ----------------------------
    e2e_t, h2d, d2h = 0.0, 0, 0
    for j in range(K):
        k = W + 1 + j
        flush.zero_()
        barrier()
        t0 = time.perf_counter()
        if iters == 1:
            loc.step(poses[k], scans[k], angles=angles)
        else:
            loc.predict(poses[k]); loc.update_chain(scans[k], angles=angles, iters=iters); loc.estimate(); loc.resample()
        barrier()
        e2e_t += time.perf_counter() - t0
        h2d += int(valid[k]) * 16 + 0      # the per-scan beam table (fp64 pairs) is what crosses PCIe
        d2h += 18 * 8
    if world > 1:
        t = torch.tensor([e2e_t], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_t = float(t.item())
    e2e_value = evals / e2e_t

    # ---- the same step with the reference's own resampling arithmetic (sequential-f32 sums, bit-exact) ---
    ref_mode_ms = None
    if world == 1 and args.resample == "fixed" and not args.quick and iters == 1:
        from mcmh_localization_b200 import RESAMPLE_REFERENCE_F32, RESAMPLE_FIXED_POINT
        loc.h.call("mcl_filter_configure", 1, RESAMPLE_REFERENCE_F32, loc.seed, 0, -1)
        ts = []
        for j in range(min(K, 30)):
            k = W + 1 + j
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            flush.zero_()
            a.record(); loc.step_staged(poses[k], k, est_buf[k]); b.record()
            torch.cuda.synchronize(dev)
            ts.append(a.elapsed_time(b))
        ref_mode_ms = float(np.median(ts[3:])) if len(ts) > 3 else float(np.median(ts))
        loc.h.call("mcl_filter_configure", 1, RESAMPLE_FIXED_POINT, loc.seed, 0, -1)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel (likelihood), measured live ------------------------
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except OSError:
        pass
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    lik_launch_ms = lik_ms.value / max(1, lik_n.value)
    mv = float(np.mean(valid[W + 1:W + 1 + K]))
    # SURVEY 8(d) per-unit algorithmic traffic: 4 B gathered from the likelihood table per particle*beam
    # evaluation + (pose read + score write) per particle (here fp64 SoA poses: 24 B + 4 B), times the units
    # one launch processes.  The gathered bytes are served from the shared-memory copy of the table, so the
    # DRAM traffic ncu sees (`traffic`) is only the pose stream.
    sets_per_launch = lik_sets / max(1, lik_n.value)      # 2: particles and particles_prev scored by one launch
    gather_bytes = 4.0 * n * mv * sets_per_launch
    stream_bytes = n * (24 + 4) * sets_per_launch
    # HBM roofline: only the pose stream and the score write must cross HBM (the table is staged in shared memory
    # once per CTA), so that is the algorithmic HBM traffic; the gathered table bytes are reported beside it.
    achieved = stream_bytes / (lik_launch_ms * 1e-3) / 1e9
    roofline = {"bound": "hbm", "achieved": achieved, "peak": hbm_peak, "unit": "GB/s",
                "frac": achieved / hbm_peak, "traffic": int(24107776 * sets_per_launch),
                "traffic_source": "profiles/r1g_summary.txt: dram__bytes_read.sum + dram__bytes_write.sum per k_likelihood_g1 "
                                  "launch (ncu --set full): 48.2 MB for a launch that scores both particle sets",
                "kernel": "k_likelihood_g1", "launch_ms": lik_launch_ms, "launches_timed": int(lik_n.value),
                "particle_sets_per_launch": sets_per_launch,
                "algorithmic_bytes_per_launch": stream_bytes,
                "table_gather_bytes_per_launch": gather_bytes,
                "table_gather_gbs": gather_bytes / (lik_launch_ms * 1e-3) / 1e9,
                "binding": "shared-memory gather rate, see gather_roofline",
                "peak_source": "MEASURED_PEAKS.json hbm_gbs (measured)" if peaks else "fallback 6650 GB/s",
                "note": "algorithmic HBM bytes = 24 B pose read + 4 B score write per particle and set (SURVEY 8d: the per-"
                        "evaluation HBM share is 28/M bytes); the 4 B per evaluation gathered from the likelihood table are "
                        "served by the shared-memory copy (table_gather_*).  The kernel is NOT HBM-bound (frac ~ 0.03, DRAM "
                        "traffic == algorithmic bytes: no re-reads): ncu shows the shared-memory pipe 87 % and the issue "
                        "slots 72 % busy; the binding ceiling and the fraction reached are in gather_roofline"}
    gl = {}
    try:
        if args.quick:
            raise RuntimeError("skipped (--quick)")
        win_bytes = 44 * 1024
        s_rate, g_rate = C.c_double(0), C.c_double(0)
        loc.h.call("mcl_bench_gather", 0, win_bytes, 1 << 32, 3, C.byref(s_rate))
        loc.h.call("mcl_bench_gather", 1, gm.width * gm.height * 4, 1 << 31, 3, C.byref(g_rate))
        lik_rate = n * mv * sets_per_launch / (lik_launch_ms * 1e-3)
        gl = {"bound": "smem_gather", "achieved": lik_rate, "peak": s_rate.value, "unit": "lookups/s",
              "frac": lik_rate / s_rate.value, "l2_gather_peak": g_rate.value,
              "frac_of_l2_gather": lik_rate / g_rate.value,
              "how": "mcl_bench_gather: random 4-byte lookups, table = 44 KiB in shared memory (peak) / "
                     "590 KB in global memory through L1/L2 (l2_gather_peak), 512-thread CTAs at full occupancy"}
    except Exception as e:  # measurement helper only
        gl = {"error": str(e)}

    # ---- CPU baseline: the oracle port on this box's host cores (bounded sample) -------------
    cpu = None
    if world == 1 and not args.no_cpu and not args.quick and iters == 1:
        ns = args.cpu_sample
        cv, cms, cdone, threads = cpu_filter_run(gm, free_space_particles(gm, ns, seed=1234), poses, scans, angles,
                                                 YAML_PARAMS, args.cpu_seconds, 3)
        cpu = {"value": cv, "unit": UNIT, "cores": threads, "kind": "port", "ms_per_step_sample": cms,
               "sample": "%d of %d particles x %d beams, %d full MHMCL steps (%.1f s), oracle C port (OpenMP, all "
                         "host cores) + NumPy glue" % (ns, n, args.beams, cdone, cms * cdone / 1e3)}

    cfg = workload_config(args, gm)
    cfg["likelihood_passes_per_step"] = passes
    if iters > 1:
        cfg["workload"] += "; %d MH iterations per scan (BASELINE configs[3])" % iters
    if world > 1:
        cfg["parallelism"] = "particles sharded over %d ranks (weak scaling), map replicated" % world
        cfg["resample_exchange"] = ("peer-push over NVLink symmetric memory (gather fused with the exchange)"
                                    if getattr(loc, "symm", None) is not None else "NCCL all-to-all")
        cfg["scalar_exchanges"] = ("libmcl kernels over NVLink peer memory (mailbox all-gather)"
                                   if getattr(loc, "native", False) else "NCCL via torch.distributed")
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
        "ms_per_step": total_ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic", "config": cfg,
        "step_ms_median": float(np.median(step_ms)), "step_ms_min": float(step_ms.min()),
        "host_wall_ms_per_step": 1e3 * t_wall / K,
        "step_ms_reference_resampling": ref_mode_ms,
        "valid_beams_mean": mv, "likelihood_kernel_evals_per_s": n * mv * sets_per_launch / (lik_launch_ms * 1e-3),
        "clocks": clocks,
        "e2e": {"value": e2e_value, "unit": UNIT, "ms_per_step": 1e3 * e2e_t / K,
                "h2d_bytes_per_step": h2d // K, "d2h_bytes_per_step": d2h // K,
                "api": "Localizer.step(odom, ranges, angles): host scan + odom in, host estimate out, wall clock"},
        "gpu_launches": int(launches), "roofline": roofline, "gather_roofline": gl, "cpu_baseline": cpu,
    }
    json_out.write(json.dumps(line) + "\n")
    json_out.flush()
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--particles", type=int, default=1_000_000, help="particles per GPU")
    ap.add_argument("--beams", type=int, default=360)
    ap.add_argument("--resample", default="fixed", choices=["fixed", "reference"])
    ap.add_argument("--cpu-sample", type=int, default=262144)
    ap.add_argument("--cpu-seconds", type=float, default=12.0)
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--mh-iters", type=int, default=1,
                    help="MH iterations per scan (BASELINE config 4 uses 32); 1 = the reference's single accept")
    ap.add_argument("--quick", action="store_true", help="skip the gather microbenchmark and the CPU baseline (profiling runs)")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "native":
        args.warmup = 3
    if args.impl == "reference":
        run_reference(args)
    else:
        run_native(args)


if __name__ == "__main__":
    main()
