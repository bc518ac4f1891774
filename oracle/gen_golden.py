#!/usr/bin/env python3
"""Generate tests/golden/*.npz by executing the UNMODIFIED reference.

Run in the build container only (needs /root/reference and numba):

    python -m oracle.gen_golden

The reference's ``app/scripts/parallel_utils.py`` is imported read-only via sys.path and called
directly; nothing of it is copied.  Stochastic functions are made reproducible with the recipe of
SURVEY.md Appendix A.3 (one numba thread + seeding inside njit), which makes the compiled code
consume the same MT19937 stream as ``np.random.RandomState(seed)``; fixtures therefore store the
*seed* and the tests regenerate the draws with RandomState (legacy stream, stable across NumPy
versions).  The maps stored under tests/golden/ are the reference's PGM files decoded with
map_server semantics (data, not source).
"""
import os
import sys

import numpy as np

REF = os.environ.get("MCL_REFERENCE", "/root/reference")
sys.path.insert(0, os.path.join(REF, "app", "scripts"))

import numba  # noqa: E402
from numba import njit, prange  # noqa: E402

import parallel_utils as pu  # noqa: E402  (the reference)

from . import node_glue as ng  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(os.path.dirname(HERE), "tests", "golden")

# app/params/amhmcl.yaml values
PARAMS = dict(alpha1=0.002, alpha2=0.03, alpha3=0.08, alpha4=0.002, sigma_hit=0.3, z_hit=0.75,
              z_rand=0.25, max_range=5.0, step=1)

numba.set_num_threads(1)


@njit
def _seed_main(s):
    np.random.seed(s)


@njit(parallel=True)
def _seed_workers(s):
    for _ in prange(1):
        np.random.seed(s)


def seed_reference(s):
    _seed_main(s)
    _seed_workers(s)


def load_ref_map(name):
    img = ng.read_pgm(os.path.join(REF, "app", "maps", name + ".pgm"))
    occ = ng.occupancy_from_pgm(img, 0, 0.65, 0.196)      # app/maps/<name>.yaml
    return occ, 0.05, -10.0, -10.0


def free_particles(mp, n, rs):
    free = np.flatnonzero(mp["map_data"] == 0)
    cells = rs.choice(free, n)
    my, mx = np.divmod(cells, mp["width"])
    x = mp["origin_np"][0] + (mx + rs.uniform(0, 1, n)) * mp["resolution"]
    y = mp["origin_np"][1] + (my + rs.uniform(0, 1, n)) * mp["resolution"]
    th = rs.uniform(-np.pi, np.pi, n)
    return np.column_stack((x, y, th))


def gen_maps():
    for name in ("map_world", "map_house"):
        occ, res, ox, oy = load_ref_map(name)
        np.savez_compressed(os.path.join(OUT, name + ".npz"), occ=occ, resolution=res,
                            origin=np.array([ox, oy]))
        print(name, occ.shape, {int(v): int((occ == v).sum()) for v in np.unique(occ)})


def gen_likelihood():
    for name, start in (("map_world", (-2.0, -0.5, 0.0)), ("map_house", None)):
        occ, res, ox, oy = load_ref_map(name)
        mp = ng.load_map(occ, res, ox, oy)
        rs = np.random.RandomState(1234)
        if start is None:
            start = free_particles(mp, 1, rs)[0]
        scan, angles = ng.synthetic_scan(np.array(start), mp, 360, 3.5,
                                         noise=np.random.RandomState(4321).normal(0, 0.01, 360))
        # special range values: NaN, negative finite, just below / at / above max_range
        scan[3] = np.nan
        scan[10] = -0.1
        scan[20] = 4.9
        scan[30] = 5.0
        scan[40] = 7.0
        scan[50] = 0.0
        pa = free_particles(mp, 1500, rs)
        W, H = mp["width"], mp["height"]
        pb = np.column_stack((rs.uniform(ox - 2, ox + W * res + 2, 400),
                              rs.uniform(oy - 2, oy + H * res + 2, 400),
                              rs.uniform(-4 * np.pi, 4 * np.pi, 400)))
        pc = np.array([
            [ox - 0.5 * res, oy + 3.0, 0.1],          # t in (-1,0): int() -> cell 0, in bounds
            [ox + 3.0, oy - 0.999 * res, -2.0],
            [ox - res, oy - res, 0.0],                # exactly -1 cell
            [ox, oy, 0.0], [ox + W * res, oy + H * res, 3.0],
            [ox + 100 * res, oy + 100 * res, 0.0],    # exactly on a cell corner
            [1e6, -1e6, 0.3], [0.0, 0.0, 0.0],
        ])
        particles = np.ascontiguousarray(np.vstack((pa, pb, pc)))
        outs = {}
        for step in (1, 3):
            outs["scores_step%d" % step] = pu.compute_likelihoods(
                scan, angles, particles, mp["distance_map"], mp["resolution"], mp["origin_np"],
                mp["width"], mp["height"], PARAMS["sigma_hit"], PARAMS["z_hit"], PARAMS["z_rand"],
                PARAMS["max_range"], step)
        # a scan with no valid beam -> -50
        blind = np.full(360, np.inf, np.float32)
        outs["scores_blind"] = pu.compute_likelihoods(
            blind, angles, particles[:16], mp["distance_map"], mp["resolution"], mp["origin_np"],
            mp["width"], mp["height"], PARAMS["sigma_hit"], PARAMS["z_hit"], PARAMS["z_rand"],
            PARAMS["max_range"], 1)
        # other sensor params incl. dist > max_range branch (tiny max_range)
        outs["scores_alt"] = pu.compute_likelihoods(
            scan, angles, particles, mp["distance_map"], mp["resolution"], mp["origin_np"],
            mp["width"], mp["height"], 0.2, 0.8, 0.2, 0.6, 2)
        np.savez_compressed(os.path.join(OUT, "likelihood_%s.npz" % name), map=name, scan=scan,
                            angles=angles, particles=particles, **outs)
        print("likelihood", name, particles.shape, {k: (float(v.min()), float(v.max()))
                                                    for k, v in outs.items()})


def motion_walk(particles, delta, alpha, mp, rs, max_attempts=1000):
    """Sequential-stream restatement of pu:332-363 used ONLY to learn how many attempts each
    particle consumed (so tests can split the regenerated stream per particle).  Its output is
    asserted equal to the reference's before anything is stored."""
    rot1, trans, rot2 = delta
    a1, a2, a3, a4 = [float(a) for a in alpha]
    s1 = a1 * abs(rot1) + a2 * abs(trans)
    s2 = a3 * abs(trans) + a4 * (abs(rot1) + abs(rot2))
    s3 = a1 * abs(rot2) + a2 * abs(trans)
    out = np.empty_like(particles)
    counts = np.zeros(len(particles), np.int32)
    ok = np.zeros(len(particles), np.uint8)
    md, W, H = mp["map_data"], mp["width"], mp["height"]
    res, (ox, oy) = mp["resolution"], mp["origin_np"]
    for i, (x, y, th) in enumerate(particles):
        out[i] = (x, y, th)
        for t in range(max_attempts):
            z = rs.normal(0, 1), rs.normal(0, 1), rs.normal(0, 1)
            counts[i] = t + 1
            r1 = rot1 + (0.0 + s1 * z[0])
            tt = trans + (0.0 + s2 * z[1])
            r2 = rot2 + (0.0 + s3 * z[2])
            xn = x + tt * np.cos(th + r1)
            yn = y + tt * np.sin(th + r1)
            thn = (th + r1 + r2 + np.pi) % (2 * np.pi) - np.pi
            mx = int((xn - ox) / res)
            my = int((yn - oy) / res)
            if 0 <= mx < W and 0 <= my < H and md[my * W + mx] == 0:
                out[i] = (xn, yn, thn)
                ok[i] = 1
                break
    return out, counts, ok


def gen_motion():
    occ, res, ox, oy = load_ref_map("map_world")
    mp = ng.load_map(occ, res, ox, oy)
    alpha = np.array([PARAMS["alpha%d" % k] for k in (1, 2, 3, 4)], dtype=np.float32)
    cases = {}
    rs0 = np.random.RandomState(99)
    particles = free_particles(mp, 384, rs0)
    for tag, delta, seed in (("fwd", (0.0, 0.02, 0.01), 11), ("turn", (0.7, 0.05, -0.3), 12),
                             ("big", (-2.5, 0.4, 1.0), 13)):
        seed_reference(seed)
        ref = pu.apply_motion_model_parallel(particles, delta, alpha, mp["map_data"],
                                             mp["resolution"], mp["origin_np"][0],
                                             mp["origin_np"][1], mp["width"], mp["height"])
        walk, counts, ok = motion_walk(particles, delta, alpha, mp, np.random.RandomState(seed))
        assert np.array_equal(ref, walk), "stream-split walk disagrees with the reference"
        cases["delta_" + tag] = np.array(delta)
        cases["seed_" + tag] = seed
        cases["out_" + tag] = ref
        cases["counts_" + tag] = counts
        cases["ok_" + tag] = ok
        print("motion", tag, "stuck", int((ok == 0).sum()), "max attempts", int(counts.max()),
              "total draws", int(counts.sum()) * 3)
    np.savez_compressed(os.path.join(OUT, "motion_map_world.npz"), particles=particles,
                        alpha=alpha, **cases)


def gen_mh_and_resample():
    occ, res, ox, oy = load_ref_map("map_world")
    mp = ng.load_map(occ, res, ox, oy)
    rs = np.random.RandomState(7)
    n = 2000
    prev = free_particles(mp, n, rs)
    cur = prev + rs.normal(0, 0.12, prev.shape)
    scan, angles = ng.synthetic_scan(np.array([-2.0, -0.5, 0.0]), mp, 360, 3.5)
    lik = lambda p: pu.compute_likelihoods(scan, angles, p, mp["distance_map"], mp["resolution"],
                                           mp["origin_np"], mp["width"], mp["height"],
                                           PARAMS["sigma_hit"], PARAMS["z_hit"], PARAMS["z_rand"],
                                           PARAMS["max_range"], 1)
    s_pre, s_post = lik(prev), lik(cur)
    w_pre, w_post = ng.convert_scores(s_pre), ng.convert_scores(s_post)
    w_pre2 = w_pre.copy()
    w_pre2[:7] = 0.0                                  # p_old == 0 branch
    out = dict(prev=prev, cur=cur, scan=scan, angles=angles, s_pre=s_pre, s_post=s_post,
               w_pre=w_pre, w_post=w_post, w_pre2=w_pre2)
    for tag, wp, seed in (("a", w_pre, 21), ("b", w_pre2, 22)):
        seed_reference(seed)
        newp, neww = pu.mh_resampling(prev, cur, w_post, wp)
        out["mh_seed_" + tag] = seed
        out["mh_particles_" + tag] = newp
        out["mh_weights_" + tag] = neww
    # asymmetric MH + transition densities
    alpha = np.array([PARAMS["alpha%d" % k] for k in (1, 2, 3, 4)], dtype=np.float32)
    delta = np.array([0.1, 0.05, -0.02])
    tf_ = pu.motion_model_odometry_parallel(prev, cur, delta, alpha)
    tb_ = pu.motion_model_odometry_parallel(cur, prev, -delta, alpha)
    seed_reference(23)
    ap, aw = pu.assym_mh_resampling(prev, cur, w_post, w_pre, tf_, tb_)
    out.update(amh_delta=delta, amh_alpha=alpha, amh_tf=tf_, amh_tb=tb_, amh_seed=23,
               amh_particles=ap, amh_weights=aw)
    # estimate inputs are just (particles, weights): restated with numpy itself, nothing to pin.
    # normalize_angle_array
    ang = rs.uniform(-20, 20, 500)
    out["naa_in"] = ang
    out["naa_mean"] = 0.7
    out["naa_out"] = pu.normalize_angle_array(ang, 0.7)
    np.savez_compressed(os.path.join(OUT, "mh_map_world.npz"), **out)
    print("mh accept rate", float((out["mh_weights_a"] == w_post).mean()))

    # low-variance resampling: the node's index form (node:467) -> resampled source indices
    res_out = {}
    cases = {
        "softmax2000": w_post,
        "mhmix2000": out["mh_weights_a"],
        "uniform1000": np.full(1000, 1.0 / 1000, np.float32),
        "peaked500": np.where(np.arange(500) == 123, 0.9, 0.1 / 499).astype(np.float32),
        "zeros_head300": np.concatenate([np.zeros(100), rs.uniform(0, 1, 200)]).astype(np.float32),
        "two": np.array([0.25, 0.75], np.float32),
        "one": np.array([1.0], np.float32),
        "rand100k": rs.uniform(0, 1, 100000).astype(np.float32) ** 8,
        "unnorm5000": (rs.uniform(0, 1, 5000) * 37.0).astype(np.float32),
    }
    for k, (tag, w) in enumerate(cases.items()):
        seed = 100 + k
        n = len(w)
        seed_reference(seed)
        idx, neww = pu.low_variance_resample_numba(np.arange(n), w.astype(np.float32), n)
        res_out["w_" + tag] = w.astype(np.float32)
        res_out["seed_" + tag] = seed
        res_out["idx_" + tag] = idx.astype(np.int32)
        assert neww.dtype == np.float32
    np.savez_compressed(os.path.join(OUT, "resample.npz"), **res_out)
    print("resample cases", list(cases))


def gen_raycast_likelihood():
    """pu:151-201 compute_likelihoods_raycast + pu:4-29 raycast on both maps."""
    for name in ("map_world", "map_house"):
        occ, res, ox, oy = load_ref_map(name)
        mp = ng.load_map(occ, res, ox, oy)
        rs = np.random.RandomState(55)
        grid = (occ != 0).astype(np.float64)
        start = free_particles(mp, 1, rs)[0]
        scan, angles = ng.synthetic_scan(start, mp, 360, 3.5, noise=rs.normal(0, 0.01, 360))
        scan[5] = np.nan; scan[6] = -0.2; scan[7] = 9.99; scan[8] = 10.0; scan[9] = 0.0
        pa = free_particles(mp, 600, rs)
        W, H = mp["width"], mp["height"]
        pb = np.column_stack((rs.uniform(ox - 1, ox + W * res + 1, 150), rs.uniform(oy - 1, oy + H * res + 1, 150),
                              rs.uniform(-4 * np.pi, 4 * np.pi, 150)))
        particles = np.ascontiguousarray(np.vstack((pa, pb)))
        scores = pu.compute_likelihoods_raycast(scan, angles, particles, grid, mp["resolution"], mp["limits"])
        blind = pu.compute_likelihoods_raycast(np.full(360, np.inf, np.float32), angles, particles[:8], grid,
                                               mp["resolution"], mp["limits"])
        rays = np.array([pu.raycast(particles[k, :2], particles[k, 2] + 0.3 * k, 10.0, mp["limits"], mp["resolution"],
                                    grid, W, H) for k in range(200)])
        np.savez_compressed(os.path.join(OUT, "raycast_%s.npz" % name), scan=scan, angles=angles, particles=particles,
                            scores=scores, blind=blind, rays=rays)
        print("raycast", name, float(np.nanmin(scores)), float(np.nanmax(scores)), blind[:2])


def gen_gaussian_init():
    """pu:594-614 initialize_gaussian_parallel (node:183), seeded through NumPy's global legacy generator."""
    occ, res, ox, oy = load_ref_map("map_world")
    mp = ng.load_map(occ, res, ox, oy)
    out = {}
    for k, (mean, cov) in enumerate((([-2.0, -0.5, 0.0], np.diag([0.05, 0.05, 0.1])),      # node:49 initial_cov
                                     ([-9.8, 9.0, 1.0], np.diag([0.5, 0.5, 0.3])))):
        np.random.seed(400 + k)
        p = pu.initialize_gaussian_parallel(np.array(mean), cov, 500, mp["distance_map"].reshape(mp["height"], mp["width"]),
                                            mp["resolution"], mp["origin_np"])
        out["mean_%d" % k] = np.array(mean); out["cov_%d" % k] = cov; out["seed_%d" % k] = 400 + k; out["out_%d" % k] = p
        print("gaussian init", k, p.shape, int((p == 0).all(axis=1).sum()), "zeroed")
    np.savez_compressed(os.path.join(OUT, "init_gaussian.npz"), **out)


def gen_kld():
    """pu:529-591 kld_sampling_amcl on weights from a real update (yaml KLD parameters)."""
    g = np.load(os.path.join(OUT, "mh_map_world.npz"))
    out = {}
    cases = {"spread": (g["mh_particles_a"], g["mh_weights_a"] / np.sum(g["mh_weights_a"]), 2000, 100),
             "minpart": (g["mh_particles_a"], g["mh_weights_a"] / np.sum(g["mh_weights_a"]), 1500, 1200)}
    rs = np.random.RandomState(3)
    tight = np.column_stack((rs.normal(-2.0, 0.03, 3000), rs.normal(-0.5, 0.03, 3000), rs.normal(0.3, 0.02, 3000)))
    wt = rs.uniform(0.5, 1.0, 3000).astype(np.float32)
    cases["tight"] = (tight, (wt / np.sum(wt)).astype(np.float32), 3000, 100)
    for k, (tag, (parts, w, max_samples, min_particles)) in enumerate(cases.items()):
        seed = 300 + k
        seed_reference(seed)
        res = pu.kld_sampling_amcl(np.ascontiguousarray(parts), w.astype(np.float32), 0.20, 0.1745, 0.03, 2,
                                   max_samples, min_particles)
        out["p_" + tag] = parts
        out["w_" + tag] = w.astype(np.float32)
        out["max_" + tag] = max_samples
        out["min_" + tag] = min_particles
        out["seed_" + tag] = seed
        out["out_" + tag] = res
        print("kld", tag, "->", res.shape, res.dtype)
    np.savez_compressed(os.path.join(OUT, "kld.npz"), **out)


def gen_init():
    occ, res, ox, oy = load_ref_map("map_world")
    mp = ng.load_map(occ, res, ox, oy)
    seed_reference(31)
    p = pu.generate_valid_particles(200, mp["map_data"], mp["resolution"], mp["origin_np"][0],
                                    mp["origin_np"][1], mp["width"], mp["height"])
    np.savez_compressed(os.path.join(OUT, "init_map_world.npz"), seed=31, n=200, particles=p)
    print("init", p.shape)


def gen_filter_run():
    """Config 1 of BASELINE.json in miniature: MHMCL on map_world, 1000 particles x 360 beams,
    12 steps (one odom + one scan per step), every function the reference's own, one MT stream."""
    occ, res, ox, oy = load_ref_map("map_world")
    mp = ng.load_map(occ, res, ox, oy)
    alpha = np.array([PARAMS["alpha%d" % k] for k in (1, 2, 3, 4)], dtype=np.float32)
    n, steps, seed = 1000, 12, 4242
    particles = free_particles(mp, n, np.random.RandomState(1234))
    particles0 = particles.copy()
    pose = np.array([-2.0, -0.5, 0.0])
    seed_reference(seed)
    prev = particles.copy()
    last_odom = None
    est, odoms, scans = [], [], []
    angles = None
    for k in range(steps):
        odom = pose.copy()
        odoms.append(odom)
        if last_odom is not None:                                         # node:393-405
            dx, dy = odom[0] - last_odom[0], odom[1] - last_odom[1]
            dth = pu.normalize_angle(odom[2] - last_odom[2])
            rot1 = np.arctan2(dy, dx) - last_odom[2]
            delta = (rot1, np.hypot(dx, dy), dth - rot1)
            prop = pu.apply_motion_model_parallel(particles, delta, alpha, mp["map_data"],
                                                  mp["resolution"], mp["origin_np"][0],
                                                  mp["origin_np"][1], mp["width"], mp["height"])
            prev = particles.copy()
            particles = prop.copy()
        last_odom = odom
        scan, angles = ng.synthetic_scan(pose, mp, 360, 3.5)
        scans.append(scan)
        lik = lambda p: pu.compute_likelihoods(scan, angles, p, mp["distance_map"],
                                               mp["resolution"], mp["origin_np"], mp["width"],
                                               mp["height"], PARAMS["sigma_hit"], PARAMS["z_hit"],
                                               PARAMS["z_rand"], PARAMS["max_range"], 1)
        w_pre = ng.convert_scores(lik(prev))                               # node:254-270
        w_post = ng.convert_scores(lik(particles))
        particles, weights = pu.mh_resampling(prev, particles, w_post, w_pre)   # node:363
        mx, my, mth, cov = ng.estimate(particles, weights)                 # node:586-597
        est.append(np.concatenate(([mx, my, mth], cov.ravel())))
        particles, _ = pu.low_variance_resample_numba(particles, weights, n)    # node:490
        # robot motion for the next step: trans 0.02 m, dtheta 0.01 rad
        pose = np.array([pose[0] + 0.02 * np.cos(pose[2]), pose[1] + 0.02 * np.sin(pose[2]),
                         pose[2] + 0.01])
    np.savez_compressed(os.path.join(OUT, "filter_run_map_world.npz"), seed=seed, n=n,
                        particles0=particles0, odoms=np.array(odoms), scans=np.array(scans),
                        angles=angles, estimates=np.array(est), particles_final=particles,
                        weights_final=weights)
    print("filter run: final estimate", est[-1][:3], "true pose", odoms[-1])


@njit(parallel=True)
def _reinit_draws(n, L):
    """The draws reinitialize_particles_numba consumes (pu:517, 522), exposed: same calls in the same order on the
    same (seeded, single-thread) stream."""
    ci = np.empty(n, np.int64)
    th = np.empty(n, np.float64)
    for i in prange(n):
        ci[i] = np.random.randint(0, L)
        th[i] = np.random.uniform(-np.pi, np.pi)
    return ci, th


def gen_alt():
    """The functions the node imports (node:13) but never reaches from its callbacks:
    compute_valid_indices pu:369-386, parallel_resample_simple pu:467-477, low_variance_resample_amcl pu:486-502,
    reinitialize_particles_numba pu:504-526."""
    occ, res, ox, oy = load_ref_map("map_world")
    mp = ng.load_map(occ, res, ox, oy)
    rs = np.random.RandomState(77)
    out = {}
    # compute_valid_indices: free, occupied, unknown, outside, and (-1, 0)-cell particles
    p = free_particles(mp, 600, rs)
    p[::7, 0] += rs.uniform(-3, 3, len(p[::7]))
    p[5] = [ox - 0.03, oy + 4.0, 0.1]; p[6] = [ox + 4.0, oy - 0.02, 0.2]; p[7] = [ox - 0.07, oy - 0.2, 0.0]
    p[8] = [ox + 384 * res + 0.01, oy + 1.0, 0.0]
    vi = pu.compute_valid_indices(p, mp["map_data"], mp["resolution"], ox, oy, mp["width"], mp["height"])
    out["cvi_particles"] = p; out["cvi_out"] = vi
    print("compute_valid_indices", len(vi), "of", len(p))
    # parallel_resample_simple (multinomial): seeds chosen so that no draw exceeds cum[-1] (Appendix C #7)
    g = np.load(os.path.join(OUT, "mh_map_world.npz"))
    w = (g["mh_weights_a"] / np.sum(g["mh_weights_a"])).astype(np.float32)
    parts = np.ascontiguousarray(g["mh_particles_a"])
    N = len(w)
    cum = np.cumsum(w)
    seed = 500
    while np.random.RandomState(seed).random_sample(N).max() >= cum[-1]:
        seed += 1
    seed_reference(seed)
    res_p = pu.parallel_resample_simple(parts, w, N)
    out["prs_particles"] = parts; out["prs_w"] = w; out["prs_seed"] = seed; out["prs_out"] = res_p
    # low_variance_resample_amcl: target_size below, equal to and above N; weights as given
    for tag, target, ww in (("eq", N, w), ("small", 700, w), ("big", 2 * N + 3, w),
                            ("unnorm", N, (w * np.float32(0.37)).astype(np.float32))):
        seed_reference(520 + len(tag))
        a, b = pu.low_variance_resample_amcl(parts, ww, target)
        out["amcl_w_" + tag] = ww; out["amcl_target_" + tag] = target; out["amcl_seed_" + tag] = 520 + len(tag)
        out["amcl_out_" + tag] = a; out["amcl_wout_" + tag] = b
        print("amcl lvr", tag, a.shape, a.dtype, b.dtype)
    out["amcl_particles"] = parts
    # reinitialize_particles_numba: the draws are read off the same stream by _reinit_draws
    occ2d = mp["map_data"].reshape(mp["height"], mp["width"])
    L = int((occ2d == 0).sum())
    seed_reference(540)
    rp = pu.reinitialize_particles_numba(500, occ2d, res, ox, oy)
    seed_reference(540)
    ci, th = _reinit_draws(500, L)
    cells = np.argwhere(occ2d == 0)
    exp = np.column_stack((cells[ci, 1] * res + ox, cells[ci, 0] * res + oy, th)).astype(np.float32)
    assert np.array_equal(exp, rp), "draw recovery does not reproduce reinitialize_particles_numba"
    out["reinit_choice"] = ci; out["reinit_theta"] = th; out["reinit_out"] = rp
    np.savez_compressed(os.path.join(OUT, "alt_functions.npz"), **out)
    print("alt functions written")



def main():
    os.makedirs(OUT, exist_ok=True)
    if len(sys.argv) > 1 and sys.argv[1] == "alt":        # only the fixtures added in round 2
        gen_alt()
        return
    gen_maps()
    gen_likelihood()
    gen_motion()
    gen_mh_and_resample()
    gen_init()
    gen_gaussian_init()
    gen_raycast_likelihood()
    gen_kld()
    gen_filter_run()
    gen_alt()
    tot = sum(os.path.getsize(os.path.join(OUT, f)) for f in os.listdir(OUT))
    print("golden bytes", tot)


if __name__ == "__main__":
    main()
