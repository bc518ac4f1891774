"""CPU ORACLE (test infrastructure only): ROS-free restatement of the glue arithmetic of the
reference node ``app/scripts/amcmh_localizer.py`` ("node" below), line for line in NumPy, plus a
``ReferenceFilter`` that replays the node's callback order on top of ``oracle.clib``.

The node itself cannot be imported (needs rospy/tf); everything here is plain NumPy/SciPy calling
the same library functions the node calls (np.exp/np.sum/np.average/np.cov,
scipy.ndimage.distance_transform_edt).
"""
import numpy as np
from scipy.ndimage import distance_transform_edt

from . import clib


# --------------------------------------------------------------------------- maps
def read_pgm(path):
    """Binary PGM (P5) reader -> (H, W) uint8/uint16 array, first image row first."""
    with open(path, "rb") as f:
        data = f.read()
    toks, pos = [], 0
    while len(toks) < 4:
        while data[pos:pos + 1].isspace():
            pos += 1
        if data[pos:pos + 1] == b"#":
            while data[pos:pos + 1] != b"\n":
                pos += 1
            continue
        end = pos
        while not data[end:end + 1].isspace():
            end += 1
        toks.append(data[pos:end])
        pos = end
    pos += 1  # single whitespace after maxval
    assert toks[0] == b"P5"
    W, H, maxval = int(toks[1]), int(toks[2]), int(toks[3])
    dt = np.uint8 if maxval < 256 else np.dtype(">u2")
    return np.frombuffer(data, dtype=dt, count=W * H, offset=pos).reshape(H, W)


def occupancy_from_pgm(img, negate=0, occupied_thresh=0.65, free_thresh=0.196):
    """map_server (trinary mode) semantics [SURVEY Appendix B]: occ=(255-px)/255;
    >occ_th -> 100, <free_th -> 0, else -1; image rows flipped so data[0] is the bottom-left cell."""
    px = img.astype(np.float64)
    occ = px / 255.0 if negate else (255.0 - px) / 255.0
    grid = np.full(img.shape, -1, np.int8)
    grid[occ > occupied_thresh] = 100
    grid[occ < free_thresh] = 0
    return np.ascontiguousarray(grid[::-1, :])


def load_map(map_2d, resolution, origin_x, origin_y):
    """node:124-177 load_map given the OccupancyGrid payload as an (H, W) int8 array."""
    map_2d = np.asarray(map_2d, dtype=np.int8)
    height, width = map_2d.shape
    map_data = map_2d.flatten().astype(np.int8)                       # node:150,176
    occupancy_binary = (map_2d != 0).astype(np.uint8)                 # node:153
    dist_2d = distance_transform_edt(occupancy_binary == 0) * resolution   # node:156
    distance_map = dist_2d.flatten().astype(np.float32)               # node:157,177
    origin_np = np.array([origin_x, origin_y])                        # node:147
    limits = np.array([origin_x, origin_x + width * resolution,
                       origin_y, origin_y + height * resolution])      # node:168-173
    return dict(map_data=map_data, distance_map=distance_map, origin_np=origin_np, limits=limits,
                width=width, height=height, resolution=resolution)


# --------------------------------------------------------------------------- glue arithmetic
def convert_scores(scores):
    """node:351-358 softmax (float32 in, float32 out, NumPy pairwise sum)."""
    max_score = np.max(scores)
    weights = np.exp(scores - max_score)
    weights = weights / np.sum(weights)
    return weights


def compute_motion(odom1, odom2):
    """node:410-421 (rot1 is NOT normalised)."""
    dx = odom2[0] - odom1[0]
    dy = odom2[1] - odom1[1]
    dtheta = clib.normalize_angle(odom2[2] - odom1[2])
    rot1 = np.arctan2(dy, dx) - odom1[2]
    trans = np.hypot(dx, dy)
    rot2 = dtheta - rot1
    return rot1, trans, rot2


def transition_probability(particles_prev, particles, delta, alpha):
    """node:424-439 incl. the reference's backward-delta construction."""
    trans_forward = clib.motion_model_odometry_parallel(particles_prev, particles,
                                                        np.array(delta), alpha)
    dx, dy, dtheta = delta
    backward_delta = np.array([
        -dx * np.cos(dtheta) - dy * np.sin(dtheta),
        dx * np.sin(dtheta) - dy * np.cos(dtheta),
        -dtheta,
    ])
    trans_backward = clib.motion_model_odometry_parallel(particles, particles_prev,
                                                         backward_delta, alpha)
    return trans_forward, trans_backward


def estimate(particles, weights):
    """node:584-597 publish_estimate arithmetic -> (mean_x, mean_y, mean_theta, cov 3x3)."""
    particles = np.asarray(particles)
    mean_pose = np.average(particles, axis=0, weights=weights)
    cos_mean = np.sum(np.cos(particles[:, 2]) * weights)
    sin_mean = np.sum(np.sin(particles[:, 2]) * weights)
    mean_theta = np.arctan2(sin_mean, cos_mean)
    diffs = particles.copy()
    diffs[:, 0] -= mean_pose[0]
    diffs[:, 1] -= mean_pose[1]
    diffs[:, 2] = clib.normalize_angle_array(particles[:, 2], mean_theta)
    if len(particles) < 2:
        return mean_pose[0], mean_pose[1], mean_theta, None
    cov = np.cov(diffs.T, aweights=weights)
    return mean_pose[0], mean_pose[1], mean_theta, cov


def get_lidar_angles(angle_min, angle_max, num_ranges):
    """node:346-348"""
    return np.linspace(angle_min, angle_max, num_ranges, dtype=np.float32)


# --------------------------------------------------------------------------- synthetic scans
def synthetic_scan(pose, mp, num_beams=360, sensor_max=3.5, noise=None):
    """ranges from pu:4-29 raycast on (map != 0), returns >= sensor_max -> +inf."""
    angles = np.linspace(0.0, 2 * np.pi - 2 * np.pi / num_beams, num_beams, dtype=np.float32)
    grid = (mp["map_data"].reshape(mp["height"], mp["width"]) != 0).astype(np.float64)
    r = np.empty(num_beams, np.float64)
    for j in range(num_beams):
        r[j] = clib.raycast(pose[:2], pose[2] + float(angles[j]), sensor_max, mp["limits"],
                            mp["resolution"], grid, mp["width"], mp["height"])
    if noise is not None:
        r = r + noise
    r = np.where(r >= sensor_max, np.inf, r)
    return r.astype(np.float32), angles


def mh_near_ties(w_new, w_old, seed, step, first_index=0, rel=1e-5, uniforms=None):
    """Particles whose MH accept decision (pu:229-231: u < min(1, f32(p_new / p_old))) could legitimately differ
    between two implementations whose weights agree to a few 1e-6 relative: the uniform lies within `rel` (relative)
    of alpha.  Tests require every mismatch to be inside this set instead of tolerating a fraction."""
    w_new, w_old = np.asarray(w_new, np.float32), np.asarray(w_old, np.float32)
    with np.errstate(divide="ignore", invalid="ignore"):
        alpha = np.where(w_old > 0, np.minimum(1.0, (w_new / w_old).astype(np.float64)), 1.0)
    if uniforms is None:
        u = np.array([clib.uniform53(seed, step, first_index + i, 0, clib.STREAM_MH) for i in range(len(w_new))])
    else:
        u = np.asarray(uniforms, np.float64)
    return (np.abs(u - alpha) <= rel * np.maximum(alpha, 1e-300)) & (alpha < 1.0)


# --------------------------------------------------------------------------- filter
class ReferenceFilter:
    """ROS-free replay of node.odom_callback (node:379-408) and node.lidar_callback
    (node:294-338) for the fixed-N modes (MCL / MHMCL / AMHMCL); draws are injected.
    """

    def __init__(self, mp, params, particles, mode="MHMCL"):
        self.mp = mp
        self.p = dict(params)
        self.mode = mode
        self.use_mh = "MH" in mode                 # node:19
        self.use_adaptive = "AMCL" in mode         # node:20
        self.assym = "AMH" in mode                 # node:21
        self.w_slow = 1e-3                         # node:86-87
        self.w_fast = 1e-3
        self.num_particles = len(particles)
        self.alpha = np.array([params["alpha1"], params["alpha2"], params["alpha3"],
                               params["alpha4"]], dtype=np.float32)     # node:28-33
        self.particles = np.array(particles, dtype=np.float64)
        self.particles_prev = self.particles.copy()
        n = len(self.particles)
        self.weights = np.ones(n) / n
        self.last_odom = None
        self.delta = (0.0, 0.0, 0.0)

    # predict ------------------------------------------------------------
    def move_particles(self, odom, normals=None, seed=0, step=0):
        current = np.asarray(odom, dtype=np.float64)
        if self.last_odom is not None:
            self.delta = compute_motion(self.last_odom, current)
            mp = self.mp
            prop = clib.apply_motion_model_parallel(
                self.particles, self.delta, self.alpha, mp["map_data"], mp["resolution"],
                mp["origin_np"][0], mp["origin_np"][1], mp["width"], mp["height"],
                normals=normals, seed=seed, step=step)
            self.particles_prev = self.particles.copy()
            self.particles = prop.copy()
        self.last_odom = current

    # update --------------------------------------------------------------
    def likelihood(self, particles, scan, angles):
        mp, p = self.mp, self.p
        return clib.compute_likelihoods(scan, angles, particles, mp["distance_map"],
                                        mp["resolution"], mp["origin_np"], mp["width"],
                                        mp["height"], p["sigma_hit"], p["z_hit"], p["z_rand"],
                                        p["max_range"], p["step"])

    def update(self, scan, angles, uniforms=None, seed=0, step=0):
        scores_pre = self.likelihood(self.particles_prev, scan, angles)      # node:254-259
        weights_pre = convert_scores(scores_pre)
        scores_post = self.likelihood(self.particles, scan, angles)          # node:263-268
        weights_post = convert_scores(scores_post)
        self.scores_pre, self.scores_post = scores_pre, scores_post
        if self.use_mh:
            if not self.assym:
                self.particles, weights = clib.mh_resampling(
                    self.particles_prev, self.particles, weights_post, weights_pre,
                    uniforms=uniforms, seed=seed, step=step)
            else:
                tf_, tb_ = transition_probability(self.particles_prev, self.particles, self.delta,
                                                  self.alpha)
                self.particles, weights = clib.assym_mh_resampling(
                    self.particles_prev, self.particles, weights_post, weights_pre, tf_, tb_,
                    uniforms)
        else:
            weights = weights_post                                           # node:313
        if self.use_adaptive:                                                # node:316-318 -> node:276-286
            self.weights = weights / np.sum(weights)
            w_avg = np.mean(self.weights)
            self.w_slow += self.p["alpha_slow"] * (w_avg - self.w_slow)
            self.w_fast += self.p["alpha_fast"] * (w_avg - self.w_fast)
            return self.weights
        self.weights = weights                                               # node:322
        return weights

    def estimate(self):
        return estimate(self.particles, self.weights)

    def resample_amcl_kld(self, r, normals, random_particles):
        """node:496-527 with injected draws; random_particles: callable(N_random) -> (N_random, 3)
        standing in for generate_valid_particles (node:515-516)."""
        p_random = max(0.0, 1.0 - self.w_fast / (self.w_slow + 1e-9))
        N = self.num_particles
        N_random = int(p_random * N)
        N_resampled = N - N_random
        resampled = clib.kld_sampling_amcl(self.particles, self.weights, self.p["kld_bin_size_xy"],
                                           self.p["kld_bin_size_theta"], self.p["kld_epsilon"], self.p["kld_z"],
                                           N_resampled, self.p["min_particles"], r, normals)
        rnd = np.asarray(random_particles(N_random), dtype=np.float64).reshape(-1, 3)
        self.num_particles = len(self.particles)                             # node:520
        self.particles = np.vstack((rnd, resampled))                         # node:521 (f32 values in an f64 array)
        self.weights = np.full(len(self.particles), 1.0 / len(self.particles))
        return N_random, len(resampled)

    def resample(self, r):
        n = len(self.particles)
        self.particles, _ = clib.low_variance_resample_numba(self.particles, self.weights, n, r)
