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
Resampling arithmetic: on one GPU the REFERENCE'S OWN (pu:416-446: sequential float32 sums, reproduced bit for
bit by the persistent tail kernel); the time of the fixed-point arithmetic is reported beside it.  Sharded runs
(N > 1) default to the fixed-point arithmetic (exact sums: cheapest exchanges, results independent of the number of
ranks) and report the time of the reference's arithmetic -- continued from rank to rank inside the same kernel,
bit-identical to one GPU -- beside it (`step_ms_other_resampling`; `--resample reference` makes it the timed mode).
One process per GPU (torchrun for N > 1); particles are sharded (weak scaling: per-GPU N fixed),
the map is replicated; timing = barrier + synchronize on both sides, max over ranks.  For N > 1 the run ends with
a parity check (sharded == single GPU of the same total size) whose result is printed and decides the exit code.
--impl reference times the reference filter on the host cores: its own numba code when the reference tree is on
this box, else the C port (oracle/), on a bounded sample of the same workload.
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


def host_threads():
    """Host cores this process may use (torchrun exports OMP_NUM_THREADS=1: do not trust OpenMP's default)."""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return max(1, os.cpu_count() or 1)


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
def cpu_filter_run(gm, particles, poses, scans, angles, params, min_seconds, min_steps, fixed_steps=None, engine="port"):
    """The reference filter (MHMCL) on the host cores; returns (evals/s, ms/step, steps, threads, description).
    Step = node.odom_callback + node.lidar_callback arithmetic.  engine: "numba" = the reference's own njit
    functions (needs the reference tree on this box), "port" = the C/OpenMP restatement (oracle/)."""
    from oracle import clib, node_glue as ng
    mp = ng.load_map(gm.occ, gm.resolution, gm.origin_x, gm.origin_y)
    threads = host_threads()
    n = len(particles)
    if engine == "numba":
        from oracle import numba_ref
        f = numba_ref.NumbaReferenceFilter(mp, params, particles, mode="MHMCL", threads=threads)
        threads = f.threads
        desc = "the reference's own " + f.describe()
    else:
        clib.set_threads(threads)
        f = ng.ReferenceFilter(mp, params, particles, mode="MHMCL")
        desc = "oracle C port (OpenMP, %d threads) + NumPy glue" % threads
    f.move_particles(poses[0])
    # one untimed warm-up step (numba: JIT compilation excluded)
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
    return evals / t_total, 1e3 * t_total / done, done, threads, desc


def cpu_engine():
    """("numba", None) when the reference's own code can run here, else ("port", why)."""
    try:
        from oracle import numba_ref
        pu, why = numba_ref.available()
        return ("numba", None) if pu is not None else ("port", why)
    except Exception as e:  # measurement helper only
        return "port", str(e)


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
    engine, why = cpu_engine()
    v, ms, done, threads, desc = cpu_filter_run(gm, parts, poses, scans, angles, YAML_PARAMS, 0, 0, fixed_steps=steps,
                                                engine=engine)
    sample = "%d of %d particles x %d beams, %d full MHMCL steps, %s" % (n, args.particles, args.beams, done, desc)
    cpu = {"value": v, "unit": UNIT, "cores": threads, "kind": "reference" if engine == "numba" else "port", "sample": sample}
    if why:
        cpu["why_port"] = why
    line = {
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
        "warmup": warm, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic", "config": workload_config(args, gm, args.resample or "reference", n_cpu=n),
        "cpu_baseline": cpu,
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


def workload_config(args, gm, resample, n_cpu=None):
    c = {"workload": "%d particles/GPU x %d beams, likelihood-field model, MHMCL predict+update+estimate+resample "
                     "per step (BASELINE configs[1]), map_world %dx%d @ %.2f m, amhmcl.yaml parameters" % (
                         args.particles, args.beams, gm.width, gm.height, gm.resolution),
         "particles_per_gpu": args.particles, "beams": args.beams, "likelihood_passes_per_step": 2,
         "map": "map_world", "resample_mode": resample,
         "l2": "flushed between timed steps (256 MiB memset outside the per-step event pairs)"}
    if n_cpu is not None:
        c["cpu_sample_particles"] = n_cpu
    return c


def likelihood_traffic():
    """DRAM bytes per likelihood launch from the committed ncu capture (profiles/likelihood_traffic.json, written by
    scripts/summarize_profiles.py from `ncu --set full`); None when there is no capture."""
    try:
        d = json.load(open(os.path.join(ROOT, "profiles", "likelihood_traffic.json")))
        return d
    except (OSError, ValueError):
        return None


# ------------------------------------------------------------------------------------------------
def run_native(args):
    # Libraries chat on stdout (NCCL announces its version there): keep the real stdout for the one JSON line and
    # send everything else written to fd 1 to stderr.
    os.environ["NCCL_DEBUG"] = os.environ.get("MCL_NCCL_DEBUG", "WARN")
    sys.stdout.flush()
    json_out = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    import ctypes as C
    import torch
    import torch.distributed as dist
    from mcmh_localization_b200 import Localizer, RESAMPLE_REFERENCE_F32, RESAMPLE_FIXED_POINT
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
    # one GPU: the reference's own resampling arithmetic; sharded: fixed point (rank-count independent)
    resample = args.resample or ("reference" if world == 1 else "fixed")

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
        loc = ShardedLocalizer(device=local, params=YAML_PARAMS, mode="MHMCL", seed=2024, resample_mode=resample)
    else:
        loc = Localizer(device=local, params=YAML_PARAMS, mode="MHMCL", seed=2024, resample_mode=resample)
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

    def max_over_ranks(v):
        if world == 1:
            return v
        t = torch.tensor([v], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

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
            loc.finish_async(est_buf[k])

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
    lik_ms, lik_n = C.c_double(0), C.c_int64(0)
    lib.mcl_timing_stop(h, C.byref(lik_ms), C.byref(lik_n))
    lik_sets = int(lib.mcl_timing_sets(h)) or int(lik_n.value)      # particle sets scored by those launches
    launches = lib.mcl_launch_count(h) - launches0
    clocks = sampler.stop() if rank == 0 else None
    step_ms = np.array([a.elapsed_time(b) for a, b in zip(ev0, ev1)])
    total_ms = max_over_ranks(float(step_ms.sum()))
    if world > 1:
        lt = torch.tensor([launches], dtype=torch.int64, device=dev)
        dist.all_reduce(lt)
        launches = int(lt.item())
    passes = 2 if iters == 1 else iters + 1       # likelihood passes per step
    evals = float(np.sum(valid[W + 1:W + 1 + K])) * n * passes * world
    value = evals / (total_ms * 1e-3)

    # ---- e2e leg: public API, host buffers in, host estimate out ----------------------------
    # K steps between two barriers; each step is timed on the host from the call to the moment the device is idle
    # again (the L2 flush between steps is outside the timed intervals); no collective inside the timed region
    # except the ones the step itself needs.
    e2e_t, h2d, d2h = 0.0, 0, 0
    barrier()
    for j in range(K):
        k = W + 1 + j
        flush.zero_()
        torch.cuda.synchronize(dev)
        t0 = time.perf_counter()
        if iters == 1:
            loc.step(poses[k], scans[k], angles=angles)
        else:
            loc.step_chain(poses[k], scans[k], angles=angles, iters=iters)
        torch.cuda.synchronize(dev)
        e2e_t += time.perf_counter() - t0
        h2d += int(valid[k]) * 16 + 0      # the per-scan beam table (fp64 pairs) is what crosses PCIe
        d2h += 18 * 8
    barrier()
    e2e_t = max_over_ranks(e2e_t)
    e2e_value = evals / e2e_t

    def timed_steps(count):
        ts = []
        for j in range(count):
            k = W + 1 + (j % K)
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            flush.zero_()
            a.record(); loc.step_staged(poses[k], k, est_buf[k]); b.record()
            torch.cuda.synchronize(dev)
            ts.append(a.elapsed_time(b))
        return float(np.median(ts[3:])) if len(ts) > 3 else float(np.median(ts))

    # ---- the same step with the OTHER resampling arithmetic, for comparison (one GPU) ------------------------
    other_mode_ms, other_mode = None, None
    if not args.quick and iters == 1:
        other_mode = "fixed" if resample == "reference" else "reference"
        loc.h.call("mcl_filter_configure", 1, RESAMPLE_FIXED_POINT if other_mode == "fixed" else RESAMPLE_REFERENCE_F32,
                   loc.seed, loc.first_index, -1)
        barrier()
        other_mode_ms = max_over_ranks(timed_steps(min(K, 40)))
        loc.h.call("mcl_filter_configure", 1, RESAMPLE_REFERENCE_F32 if resample == "reference" else RESAMPLE_FIXED_POINT,
                   loc.seed, loc.first_index, -1)
        barrier()
    tail_err = C.c_int(0)
    if world == 1:
        loc.h.call("mcl_tail_status", C.byref(tail_err))

    # ---- other BASELINE configurations, short and clearly labelled (not the headline) --------------------------
    extras = {}
    if not args.quick and not args.no_extras and iters == 1:
        extras = run_extras(args, loc, world, rank, local, dev, poses, scans, angles, valid, W, K, flush, barrier, max_over_ranks)

    # ---- N > 1: the sharded run must equal a single-GPU run of the same total size -----------------------------
    parity = None
    if world > 1 and not args.no_parity:
        parity = sharded_parity(world, rank, local, dev, gm)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        if parity is not None and not parity.get("ok", False):
            sys.exit(1)
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
    achieved = stream_bytes / (lik_launch_ms * 1e-3) / 1e9
    tr = likelihood_traffic()
    traffic = None
    if tr and tr.get("particles_per_launch"):
        # the capture's bytes per particle and set, scaled to this launch
        traffic = int(tr["dram_bytes_per_launch"] / (tr["particles_per_launch"] * tr["sets_per_launch"]) * n * sets_per_launch)
    roofline = {"bound": "hbm", "achieved": achieved, "peak": hbm_peak, "unit": "GB/s",
                "frac": achieved / hbm_peak, "traffic": traffic,
                "traffic_source": (tr or {}).get("source", "no ncu capture committed under profiles/"),
                "kernel": "k_likelihood_g1", "launch_ms": lik_launch_ms, "launches_timed": int(lik_n.value),
                "particle_sets_per_launch": sets_per_launch,
                "algorithmic_bytes_per_launch": stream_bytes,
                "table_gather_bytes_per_launch": gather_bytes,
                "table_gather_gbs": gather_bytes / (lik_launch_ms * 1e-3) / 1e9,
                "binding": "shared-memory gather rate, see gather_roofline",
                "peak_source": "MEASURED_PEAKS.json hbm_gbs (measured)" if peaks else "fallback 6650 GB/s",
                "note": "algorithmic HBM bytes = 24 B pose read + 4 B score write per particle and set (SURVEY 8d: the per-"
                        "evaluation HBM share is 28/M bytes); the 4 B per evaluation gathered from the likelihood table are "
                        "served by the shared-memory copy (table_gather_*).  The kernel is NOT HBM-bound (frac ~ 0.03): "
                        "the binding ceiling and the fraction reached are in gather_roofline"}
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
              "how": "mcl_bench_gather, measured in this run: random 4-byte lookups, table = 44 KiB in shared memory "
                     "(peak) / 590 KB in global memory through L1/L2 (l2_gather_peak), 512-thread CTAs at full occupancy"}
    except Exception as e:  # measurement helper only
        gl = {"error": str(e)}

    # The ceiling that actually binds the likelihood kernel: the warp schedulers.  An evaluation costs
    # 10.06 instructions of which 4 are DFMA (11.06 for particles whose beams can leave the window's 256 columns: one
    # more clamp), and a DFMA holds its scheduler for two cycles (scripts/ubench/dmma.cu; the mix is pinned on the
    # built library by tests/test_abi_host.py) = 14.06 issue cycles per warp and evaluation.
    issue = {}
    try:
        props = torch.cuda.get_device_properties(dev)
        sm_hz = float((clocks or {}).get("sm_mhz") or 0.0) * 1e6
        if sm_hz > 0 and iters == 1:
            cyc = 14.06
            peak_issue = props.multi_processor_count * 4 * sm_hz * 32.0 / cyc
            lik_rate_i = n * mv * sets_per_launch / (lik_launch_ms * 1e-3)
            issue = {"bound": "warp-scheduler issue cycles", "achieved": lik_rate_i, "peak": peak_issue, "unit": "evals/s",
                     "frac": lik_rate_i / peak_issue, "issue_cycles_per_evaluation": cyc, "sm_mhz": sm_hz / 1e6,
                     "how": "SMs x 4 schedulers x SM clock under load x 32 lanes / (6.06 single-cycle instructions + 4 DFMA "
                            "x 2 cycles per evaluation: SASS of the beam loop)"}
    except Exception as e:  # explanatory figure only
        issue = {"error": str(e)}

    # ---- CPU baseline: the reference filter on this box's host cores (bounded sample) -------------
    cpu = None
    if world == 1 and not args.no_cpu and not args.quick and iters == 1:
        ns = args.cpu_sample
        sample_parts = free_space_particles(gm, ns, seed=1234)
        engine, why = cpu_engine()
        cv, cms, cdone, threads, desc = cpu_filter_run(gm, sample_parts, poses, scans, angles, YAML_PARAMS, args.cpu_seconds, 3,
                                                       engine=engine)
        cpu = {"value": cv, "unit": UNIT, "cores": threads, "kind": "reference" if engine == "numba" else "port",
               "ms_per_step_sample": cms,
               "sample": "%d of %d particles x %d beams, %d full MHMCL steps (%.1f s), %s" % (
                   ns, n, args.beams, cdone, cms * cdone / 1e3, desc)}
        if why:
            cpu["why_port"] = why
        else:       # the C port beside the reference's own code
            pv, pms, pdone, pth, pdesc = cpu_filter_run(gm, sample_parts, poses, scans, angles, YAML_PARAMS,
                                                        args.cpu_seconds / 2, 3, engine="port")
            cpu["port"] = {"value": pv, "cores": pth, "ms_per_step_sample": pms, "sample": pdesc}

    cfg = workload_config(args, gm, resample)
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
        "step_ms_other_resampling": {"mode": other_mode, "ms_median": other_mode_ms} if other_mode else None,
        "tail_kernel_error": int(tail_err.value),
        "valid_beams_mean": mv, "likelihood_kernel_evals_per_s": n * mv * sets_per_launch / (lik_launch_ms * 1e-3),
        "clocks": clocks,
        "e2e": {"value": e2e_value, "unit": UNIT, "ms_per_step": 1e3 * e2e_t / K,
                "h2d_bytes_per_step": h2d // K, "d2h_bytes_per_step": d2h // K,
                "api": "Localizer.step(odom, ranges, angles): host scan + odom in, host estimate out, wall clock"},
        "gpu_launches": int(launches), "roofline": roofline, "gather_roofline": gl, "issue_roofline": issue,
        "cpu_baseline": cpu,
        "extras": extras, "parity": parity,
    }
    json_out.write(json.dumps(line) + "\n")
    json_out.flush()
    if world > 1:
        dist.destroy_process_group()
    if parity is not None and not parity.get("ok", False):
        sys.exit(1)


def run_extras(args, loc, world, rank, local, dev, poses, scans, angles, valid, W, K, flush, barrier, max_over_ranks):
    """Short runs of the other BASELINE configurations so that the driver's records carry them: configs[3] (32 MH
    iterations per scan), configs[4] (uniform particles on a 4096^2 map, tiled likelihood kernel) and, on 8 GPUs,
    configs[2] (10 M particles).  Each: 2 warm-up + 5 timed steps, device-resident inputs, CUDA events, max over
    ranks.  They are labelled extras, not the headline."""
    import torch
    from mcmh_localization_b200 import Localizer
    from mcmh_localization_b200.params import YAML_PARAMS
    from mcmh_localization_b200.synth import free_space_particles, raycast_scan
    out = {}
    n = args.particles

    def timed(fn, count=5, warm=2):
        for j in range(warm):
            fn(j)
        barrier()
        ts = []
        for j in range(count):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            flush.zero_()
            a.record(); fn(warm + j); b.record()
            torch.cuda.synchronize(dev)
            ts.append(a.elapsed_time(b))
        barrier()
        return max_over_ranks(float(np.mean(ts)))

    try:        # configs[3]: 1 M chains x 32 MH iterations per scan
        def chain_step(j):
            k = W + 1 + (j % K)
            loc.predict(poses[k])
            loc.h.call("mcl_use_scan", int(k))
            loc.update_chain(None, iters=32)
            loc.finish_async(loc.est18)
        ms = timed(chain_step)
        mv = float(np.mean(valid[W + 1:W + 1 + K]))
        out["configs[3] MH x32"] = {"ms_per_scan": ms, "particles_per_gpu": n, "n_gpus": world, "mh_iterations": 32,
                                    "evals_per_s": n * world * mv * 33 / (ms * 1e-3)}
    except Exception as e:
        out["configs[3] MH x32"] = {"error": str(e)}
    if world == 8:
        try:    # configs[2]: 10 M particles over 8 GPUs
            n2 = 1_250_000
            loc.set_particles(free_space_particles(loc.map, n2, seed=99 + rank))
            loc.predict(poses[0])

            def step10m(j):
                k = W + 1 + (j % K)
                loc.step_staged(poses[k], k, None)
            ms = timed(step10m)
            mv = float(np.mean(valid[W + 1:W + 1 + K]))
            out["configs[2] 10M on 8 GPUs"] = {"ms_per_step": ms, "particles_total": n2 * world,
                                               "evals_per_s": n2 * world * mv * 2 / (ms * 1e-3)}
        except Exception as e:
            out["configs[2] 10M on 8 GPUs"] = {"error": str(e)}
    try:        # configs[4]: global localisation on a 4096^2 map, 6.25 M uniform particles per GPU
        from mcmh_localization_b200.maps import load_npz
        side, n5 = 4096, 6_250_000
        base = load_npz(os.path.join(ROOT, "tests", "golden", "map_house.npz"))
        occ = np.ascontiguousarray(np.tile(base.occ, (11, 11))[:side, :side])
        if world > 1:
            from mcmh_localization_b200.sharded import ShardedLocalizer
            big = ShardedLocalizer(device=local, params=YAML_PARAMS, mode="MHMCL", seed=7)
        else:
            big = Localizer(device=local, params=YAML_PARAMS, mode="MHMCL", seed=7, resample_mode="fixed")
        big.load_map(occ, base.resolution, (-0.5 * side * base.resolution, -0.5 * side * base.resolution), gpu_edt=True)
        big.init_uniform(n5)
        pose = free_space_particles(big.map, 1, seed=3)[0]
        scan, ang = raycast_scan(big.map, pose)
        vb = int(np.sum(np.isfinite(scan) & (scan < YAML_PARAMS["max_range"])))
        state = {"pose": pose}

        def big_step(j):
            p = state["pose"]
            big.step(p, scan, angles=ang)
            state["pose"] = p + np.array([0.02 * np.cos(p[2]), 0.02 * np.sin(p[2]), 0.01])
        ms = timed(big_step)
        out["configs[4] 4096^2 map"] = {"ms_per_step": ms, "particles_per_gpu": n5, "n_gpus": world,
                                        "evals_per_s": n5 * world * vb * 2 / (ms * 1e-3),
                                        "note": "uniform particles, tiled likelihood kernel, fixed-point resampling, host API"}
        big.close()
    except Exception as e:
        out["configs[4] 4096^2 map"] = {"error": str(e)}
    return out


def sharded_parity(world, rank, local, dev, gm):
    """ShardedLocalizer over all ranks vs a single-GPU Localizer (rank 0) of the same total size and seed: the
    particle sets must be identical and the estimates equal to rounding (the random draws are keyed by the global
    particle index, the softmax and resampling sums are exact integers).  Returns the dict printed as `parity`."""
    import torch
    import torch.distributed as dist
    from mcmh_localization_b200 import Localizer
    from mcmh_localization_b200.params import YAML_PARAMS
    from mcmh_localization_b200.sharded import ShardedLocalizer
    from mcmh_localization_b200.synth import free_space_particles
    n_local, steps = 320_000, 4            # enough per rank for the one-thread-per-particle likelihood kernel
    n = n_local * world
    poses = trajectory(steps + 1)
    scans, angles = make_scans(gm, poses, 360)
    parts = free_space_particles(gm, n, seed=5)
    sh = ShardedLocalizer(device=local, params=YAML_PARAMS, mode="MHMCL", seed=99)
    sh.load_map(gm)
    sh.set_particles(parts[rank * n_local:(rank + 1) * n_local])
    ref = None
    if rank == 0:
        ref = Localizer(device=local, params=YAML_PARAMS, mode="MHMCL", seed=99, resample_mode="fixed")
        ref.load_map(gm)
        ref.set_particles(parts)
    worst, d_mean, d_cov = 1.0, 0.0, 0.0
    for k in range(steps):
        est = sh.step(poses[k], scans[k], angles=angles)
        allp = sh.gather_particles()
        if rank == 0:
            rest = ref.step(poses[k], scans[k], angles=angles)
            same = float(np.all(allp == ref.particles(), axis=1).mean())
            worst = min(worst, same)
            d_mean = max(d_mean, abs(est[0] - rest[0]), abs(est[1] - rest[1]))
            d_cov = max(d_cov, float(np.abs(est[3] - rest[3]).max()))
    ok = worst == 1.0 and d_mean < 1e-9 and d_cov < 1e-9
    flag = torch.tensor([1 if ok else 0], device=dev)
    dist.broadcast(flag, 0)
    return {"ok": bool(int(flag.item())), "identical_particle_fraction": worst, "max_abs_mean_delta": d_mean,
            "max_abs_cov_delta": d_cov, "particles_total": n, "ranks": world, "steps": steps,
            "what": "ShardedLocalizer over all ranks vs one-GPU Localizer (fixed-point resampling), same seed"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--particles", type=int, default=1_000_000, help="particles per GPU")
    ap.add_argument("--beams", type=int, default=360)
    ap.add_argument("--resample", default=None, choices=["fixed", "reference"],
                    help="resampling arithmetic; default: reference on one GPU, fixed when sharded")
    ap.add_argument("--cpu-sample", type=int, default=262144)
    ap.add_argument("--cpu-seconds", type=float, default=12.0)
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip the short runs of BASELINE configs[2..4]")
    ap.add_argument("--no-parity", action="store_true", help="skip the sharded == single-GPU check (N > 1)")
    ap.add_argument("--mh-iters", type=int, default=1,
                    help="MH iterations per scan (BASELINE config 4 uses 32); 1 = the reference's single accept")
    ap.add_argument("--quick", action="store_true", help="skip the gather microbenchmark, the CPU baseline and the extras (profiling runs)")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "native":
        args.warmup = 3
    if args.impl == "reference":
        run_reference(args)
    else:
        run_native(args)


if __name__ == "__main__":
    main()
