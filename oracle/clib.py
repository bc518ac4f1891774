"""ctypes binding of oracle/c/mcl_oracle.c (CPU ORACLE -- test infrastructure only).

Function names and positional signatures follow the reference's
app/scripts/parallel_utils.py so the parity tests read like calls into the
reference; stochastic functions additionally take the injected draws.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libmcl_oracle.so")

STREAM_MOTION, STREAM_MH, STREAM_RESAMPLE, STREAM_INIT, STREAM_KLD = 1, 2, 3, 4, 5


def build(force=False):
    """Compile the C oracle with gcc (oracle/Makefile). Building the checker is not using it."""
    src = os.path.join(_HERE, "c", "mcl_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            build()
        L = C.CDLL(_SO)
        L.orc_normalize_angle.restype = C.c_double
        L.orc_normalize_angle.argtypes = [C.c_double]
        L.orc_raycast.restype = C.c_double
        L.orc_uniform53.restype = C.c_double
        L.orc_uniform53.argtypes = [C.c_uint64, C.c_uint64, C.c_uint64, C.c_uint32, C.c_uint32]
        L.orc_resample_scale.restype = C.c_double
        L.orc_resample_scale.argtypes = [C.c_float, C.c_int64]
        L.orc_max_threads.restype = C.c_int
        _lib = L
    return _lib


def _p(a, t):
    return a.ctypes.data_as(C.POINTER(t))


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def max_threads():
    return int(lib().orc_max_threads())


def set_threads(n):
    lib().orc_set_threads(C.c_int(int(n)))


def normalize_angle(theta):
    return float(lib().orc_normalize_angle(float(theta)))


def normalize_angle_array(angles, mean_angle):
    a = _f64(angles)
    out = np.empty(a.shape[0], np.float32)
    lib().orc_normalize_angle_array(_p(a, C.c_double), C.c_double(float(mean_angle)),
                                    C.c_int64(a.shape[0]), _p(out, C.c_float))
    return out


def compute_likelihoods(scan_ranges, angles, particles, distance_map, map_resolution, map_origin,
                        width, height, sigma_hit=0.35, z_hit=0.9, z_rand=0.1, max_range=10, step=1):
    scan = _f32(scan_ranges)
    ang = _f32(angles)
    p = _f64(particles)
    d = _f32(distance_map)
    N = p.shape[0]
    out = np.zeros(N, np.float32)
    lib().orc_compute_likelihoods(
        _p(scan, C.c_float), _p(ang, C.c_float), C.c_int(scan.shape[0]), _p(p, C.c_double),
        C.c_int64(N), _p(d, C.c_float), C.c_double(float(map_resolution)),
        C.c_double(float(map_origin[0])), C.c_double(float(map_origin[1])), C.c_int(int(width)),
        C.c_int(int(height)), C.c_double(float(sigma_hit)), C.c_double(float(z_hit)),
        C.c_double(float(z_rand)), C.c_double(float(max_range)), C.c_int(int(step)),
        _p(out, C.c_float))
    return out


def raycast(pose, angle, max_range, limits, resolution, grid_map, grid_width, grid_height):
    g = _f64(grid_map)
    lim = _f64(limits)
    return float(lib().orc_raycast(
        C.c_double(float(pose[0])), C.c_double(float(pose[1])), C.c_double(float(angle)),
        C.c_double(float(max_range)), _p(lim, C.c_double), C.c_double(float(resolution)),
        _p(g, C.c_double), C.c_int(int(grid_width)), C.c_int(int(grid_height))))


def compute_likelihoods_raycast(scan_ranges, angles, particles, grid_map, map_resolution, limits):
    """pu:151-201."""
    scan, ang, p = _f32(scan_ranges), _f32(angles), _f64(particles)
    g = _f64(grid_map)
    lim = _f64(limits)
    out = np.zeros(p.shape[0], np.float32)
    lib().orc_compute_likelihoods_raycast(_p(scan, C.c_float), _p(ang, C.c_float), C.c_int(scan.shape[0]),
                                          _p(p, C.c_double), C.c_int64(p.shape[0]), _p(g, C.c_double),
                                          C.c_int(g.shape[1]), C.c_int(g.shape[0]), C.c_double(float(map_resolution)),
                                          _p(lim, C.c_double), _p(out, C.c_float))
    return out


def compute_valid_mask(particles, map_data, width, height, resolution, origin_x, origin_y):
    p = _f64(particles)
    m = np.ascontiguousarray(map_data, np.int8)
    out = np.zeros(p.shape[0], np.uint8)
    lib().orc_compute_valid_mask(_p(p, C.c_double), C.c_int64(p.shape[0]), _p(m, C.c_int8),
                                 C.c_int(int(width)), C.c_int(int(height)),
                                 C.c_double(float(resolution)), C.c_double(float(origin_x)),
                                 C.c_double(float(origin_y)), _p(out, C.c_uint8))
    return out.astype(bool)


def generate_valid_particles(num_particles, map_data, map_resolution, origin_x, origin_y, width,
                             height, ux, uy, ut):
    """pu:450-465 with injected uniforms ux, uy, ut in [0,1) of length max(50*N, 500).
    np.random.uniform(lo, hi) == lo + (hi - lo) * u."""
    max_trials = max(50 * num_particles, 500)
    assert len(ux) == len(uy) == len(ut) == max_trials
    hi_x = origin_x + width * map_resolution
    hi_y = origin_y + height * map_resolution
    x = origin_x + (hi_x - origin_x) * _f64(ux)
    y = origin_y + (hi_y - origin_y) * _f64(uy)
    th = -np.pi + (np.pi - (-np.pi)) * _f64(ut)
    allp = np.column_stack((x, y, th))
    mask = compute_valid_mask(allp, map_data, width, height, map_resolution, origin_x, origin_y)
    return allp[mask][:num_particles]


def apply_motion_model_parallel(particles, delta, alpha, map_data, map_resolution, origin_x,
                                origin_y, width, height, normals=None, seed=0, step=0,
                                first_index=0, max_attempts=1000, return_attempts=False):
    """pu:332-363. ``normals``: injected (N, A, 3) standard normals, else Philox(seed, step)."""
    p = _f64(particles)
    N = p.shape[0]
    d = _f64(np.asarray(delta, dtype=np.float64))
    al = _f32(alpha)
    m = np.ascontiguousarray(map_data, np.int8)
    out = np.empty_like(p)
    att = np.zeros(N, np.int32)
    if normals is not None:
        z = _f64(normals)
        assert z.ndim == 3 and z.shape[0] == N and z.shape[2] == 3
        zp, A = _p(z, C.c_double), z.shape[1]
    else:
        zp, A = None, 0
    lib().orc_apply_motion_model(
        _p(p, C.c_double), C.c_int64(N), _p(d, C.c_double), _p(al, C.c_float), _p(m, C.c_int8),
        C.c_double(float(map_resolution)), C.c_double(float(origin_x)), C.c_double(float(origin_y)),
        C.c_int(int(width)), C.c_int(int(height)), zp, C.c_int(A), C.c_uint64(int(seed)),
        C.c_uint64(int(step)), C.c_uint64(int(first_index)), C.c_int(int(max_attempts)),
        _p(out, C.c_double), _p(att, C.c_int32))
    return (out, att) if return_attempts else out


def mh_resampling(particles, proposed_particles, likelihoods, old_weights, uniforms=None, seed=0,
                  step=0, first_index=0, return_accept=False):
    """pu:208-236 with injected uniforms (or Philox)."""
    p, q = _f64(particles), _f64(proposed_particles)
    lk, ow = _f32(likelihoods), _f32(old_weights)
    N = p.shape[0]
    newp = np.empty_like(p)
    neww = np.empty(N, np.float32)
    acc = np.zeros(N, np.uint8)
    up = _p(_f64(uniforms), C.c_double) if uniforms is not None else None
    if uniforms is not None:
        u = _f64(uniforms)
        up = _p(u, C.c_double)
    lib().orc_mh_resampling(_p(p, C.c_double), _p(q, C.c_double), _p(lk, C.c_float),
                            _p(ow, C.c_float), C.c_int64(N), up, C.c_uint64(int(seed)),
                            C.c_uint64(int(step)), C.c_uint64(int(first_index)),
                            _p(newp, C.c_double), _p(neww, C.c_float), _p(acc, C.c_uint8))
    return (newp, neww, acc) if return_accept else (newp, neww)


def assym_mh_resampling(particles, proposed_particles, likelihoods, old_weights, trans_forward,
                        trans_backward, uniforms, return_accept=False):
    p, q = _f64(particles), _f64(proposed_particles)
    lk, ow = _f32(likelihoods), _f32(old_weights)
    tf, tb, u = _f64(trans_forward), _f64(trans_backward), _f64(uniforms)
    N = p.shape[0]
    newp = np.empty_like(p)
    neww = np.empty(N, np.float32)
    acc = np.zeros(N, np.uint8)
    lib().orc_assym_mh_resampling(_p(p, C.c_double), _p(q, C.c_double), _p(lk, C.c_float),
                                  _p(ow, C.c_float), _p(tf, C.c_double), _p(tb, C.c_double),
                                  C.c_int64(N), _p(u, C.c_double), _p(newp, C.c_double),
                                  _p(neww, C.c_float), _p(acc, C.c_uint8))
    return (newp, neww, acc) if return_accept else (newp, neww)


def motion_model_odometry_parallel(particles_prev, particles_curr, delta, alpha):
    a, b = _f64(particles_prev), _f64(particles_curr)
    d = _f64(np.asarray(delta, dtype=np.float64))
    al = _f32(alpha)
    out = np.empty(a.shape[0], np.float64)
    lib().orc_motion_model_odometry(_p(a, C.c_double), _p(b, C.c_double), C.c_int64(a.shape[0]),
                                    _p(d, C.c_double), _p(al, C.c_float), _p(out, C.c_double))
    return out


def low_variance_resample_indices(weights, N, r):
    """pu:416-446 -> the source index chosen for every output m (r = the one uniform draw)."""
    w = _f32(weights)
    idx = np.empty(int(N), np.int32)
    lib().orc_low_variance_resample(_p(w, C.c_float), C.c_int64(w.shape[0]), C.c_int64(int(N)),
                                    C.c_double(float(r)), _p(idx, C.c_int32))
    return idx


def low_variance_resample_numba(particles, weights, N, r):
    """pu:416-446 drop-in (particles may be (N,3) poses or an (N,) index array, node:467)."""
    idx = low_variance_resample_indices(weights, N, r)
    particles = np.asarray(particles)
    return particles[idx].copy(), np.full(int(N), 1.0 / N, dtype=np.float32)


def resample_scale(wmax, n_global):
    return float(lib().orc_resample_scale(C.c_float(float(wmax)), C.c_int64(int(n_global))))


def systematic_resample_q(weights, N, r, scale=None):
    """Production (fixed-point) systematic resampling restated on the CPU."""
    w = _f32(weights)
    if scale is None:
        scale = resample_scale(float(w.max()), w.shape[0])
    idx = np.empty(int(N), np.int32)
    lib().orc_systematic_resample_q(_p(w, C.c_float), C.c_int64(w.shape[0]), C.c_int64(int(N)),
                                    C.c_double(float(r)), C.c_double(float(scale)),
                                    _p(idx, C.c_int32))
    return idx


def kld_sampling_amcl(particles, weights, bin_size_xy, bin_size_theta, epsilon, z, max_samples, min_particles,
                      r, normals):
    """pu:529-591 with injected draws (r, normals (max_samples, 3)) -> (count, 3) float32."""
    p = _f64(particles)
    w = _f32(weights)
    zz = _f64(normals)
    assert zz.shape[0] >= max_samples and zz.shape[1] == 3
    out = np.zeros((int(max_samples), 3), np.float32)
    L = lib()
    L.orc_kld_sampling.restype = C.c_int64
    cnt = L.orc_kld_sampling(_p(p, C.c_double), _p(w, C.c_float), C.c_int64(p.shape[0]), C.c_double(float(bin_size_xy)),
                             C.c_double(float(bin_size_theta)), C.c_double(float(epsilon)), C.c_double(float(z)),
                             C.c_int64(int(max_samples)), C.c_int64(int(min_particles)), C.c_double(float(r)),
                             _p(zz, C.c_double), _p(out, C.c_float))
    return out[:cnt]


def philox4x32_10(ctr, key):
    c = np.ascontiguousarray(ctr, np.uint32)
    k = np.ascontiguousarray(key, np.uint32)
    out = np.zeros(4, np.uint32)
    lib().orc_philox4x32_10(_p(c, C.c_uint32), _p(k, C.c_uint32), _p(out, C.c_uint32))
    return out


def draw4(seed, step, item, sub, stream):
    out = np.zeros(4, np.uint32)
    lib().orc_draw4(C.c_uint64(int(seed)), C.c_uint64(int(step)), C.c_uint64(int(item)),
                    C.c_uint32(int(sub)), C.c_uint32(int(stream)), _p(out, C.c_uint32))
    return out


def uniform53(seed, step, item, sub, stream):
    return float(lib().orc_uniform53(int(seed), int(step), int(item), int(sub), int(stream)))


def normals3(seed, step, item, attempt):
    z = np.zeros(3, np.float64)
    lib().orc_normals3(C.c_uint64(int(seed)), C.c_uint64(int(step)), C.c_uint64(int(item)),
                       C.c_uint32(int(attempt)), _p(z, C.c_double))
    return z


def normals3_seq(seed, step, item, n):
    """attempts 0 .. n-1 of one particle through the oracle's sequential reader -> (n, 3)"""
    z = np.zeros((int(n), 3), np.float64)
    lib().orc_normals3_seq(C.c_uint64(int(seed)), C.c_uint64(int(step)), C.c_uint64(int(item)), C.c_int(int(n)),
                           _p(z, C.c_double))
    return z


# ---- functions the node imports but its callbacks never reach -------------------------
def compute_valid_indices(particles, map_data, map_resolution, origin_x, origin_y, width, height):
    """pu:369-386 -> int32 indices of the particles on cells with map_data <= 10."""
    p = _f64(particles)
    m = np.ascontiguousarray(map_data, np.int8)
    out = np.empty(p.shape[0], np.int32)
    lib().orc_compute_valid_indices.restype = C.c_int64
    k = lib().orc_compute_valid_indices(_p(p, C.c_double), C.c_int64(p.shape[0]), _p(m, C.c_int8), C.c_int(int(width)),
                                        C.c_int(int(height)), C.c_double(float(map_resolution)),
                                        C.c_double(float(origin_x)), C.c_double(float(origin_y)), _p(out, C.c_int32))
    return out[:k].copy()


def parallel_resample_simple_indices(weights, N, uniforms):
    """pu:467-477 with injected uniforms -> source index per output."""
    w = _f32(weights)
    u = _f64(uniforms)
    idx = np.empty(int(N), np.int32)
    lib().orc_parallel_resample_simple(_p(w, C.c_float), C.c_int64(w.shape[0]), _p(u, C.c_double), C.c_int64(int(N)),
                                       _p(idx, C.c_int32))
    return idx


def low_variance_resample_amcl_indices(weights, target_size, r):
    """pu:486-502 -> source index per output (r = the one uniform draw in [0, 1/target_size))."""
    w = _f32(weights)
    idx = np.empty(int(target_size), np.int32)
    lib().orc_low_variance_resample_amcl(_p(w, C.c_float), C.c_int64(w.shape[0]), C.c_int64(int(target_size)),
                                         C.c_double(float(r)), _p(idx, C.c_int32))
    return idx


def reinitialize_particles(num_new, occupancy_map, res, origin_x, origin_y, choice, theta):
    """pu:504-526 with injected draws (choice into the row-major list of free cells, theta) -> (num_new, 3) float32."""
    occ = np.ascontiguousarray(occupancy_map, np.int8)
    H, W = occ.shape
    ch = np.ascontiguousarray(choice, np.int64)
    th = _f64(theta)
    out = np.empty((int(num_new), 3), np.float32)
    lib().orc_reinitialize_particles.restype = C.c_int64
    lib().orc_reinitialize_particles(C.c_int64(int(num_new)), _p(occ, C.c_int8), C.c_int(W), C.c_int(H),
                                     C.c_double(float(res)), C.c_double(float(origin_x)), C.c_double(float(origin_y)),
                                     _p(ch, C.c_int64), _p(th, C.c_double), _p(out, C.c_float))
    return out


def validate_samples(samples, distance_map, resolution, origin):
    """pu:600-614 -> copy of samples with the invalid ones zeroed (distance_map is (H, W))."""
    s = _f64(samples).copy()
    d = np.ascontiguousarray(distance_map, np.float32)
    H, W = d.shape
    lib().orc_validate_samples(_p(s, C.c_double), C.c_int64(s.shape[0]), _p(d, C.c_float), C.c_int(W), C.c_int(H),
                               C.c_double(float(resolution)), C.c_double(float(origin[0])), C.c_double(float(origin[1])))
    return s
