"""Function shim with the names and positional signatures of the reference's
app/scripts/parallel_utils.py, so that the node's import line (amcmh_localizer.py:13)

    from parallel_utils import compute_likelihoods, mh_resampling, apply_motion_model_parallel, ...

can point at this module unchanged.  Host NumPy arrays in the reference's dtypes go in and come
out ((N,3) float64 particles, float32 scores/weights); every function uploads, runs the libmcl.so
kernel(s) on the GPU and downloads.  This is the compatibility path (it pays PCIe copies per
call); the fast path is mcmh_localization_b200.Localizer, which keeps particles on the device.

Stochastic functions draw from the library's Philox generator (module seed + call counter) unless
the caller injects the draws through the extra keyword arguments (uniforms= / normals= / r=) --
that is how the parity tests reproduce the reference bit-for-bit.
"""
import ctypes as C
import zlib

import numpy as np
import torch

from . import _lib

_pf = C.POINTER(C.c_float)
_state = {"ctx": {}, "seed": 0, "calls": 0}


def seed(s):
    """Seed of the shim's Philox streams (the reference seeds numba's per-thread MT19937)."""
    _state["seed"] = int(s)
    _state["calls"] = 0


def _tick():
    _state["calls"] += 1
    return _state["calls"]


class _Ctx:
    def __init__(self, device):
        if not torch.cuda.is_available():
            raise RuntimeError("mcmh_localization_b200.parallel_utils needs a CUDA device (no CPU fallback)")
        self.device = torch.device("cuda", device)
        self.h = _lib.Handle(device)
        self.map_key = None
        self.sensor_key = None
        self.alpha_key = None

    def bind(self):
        self.h.call("mcl_set_stream", C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream))

    def set_map(self, occ, dist, W, H, res, ox, oy):
        occ = None if occ is None else np.ascontiguousarray(occ, dtype=np.int8).ravel()
        dist = None if dist is None else np.ascontiguousarray(dist, dtype=np.float32).ravel()
        key = (W, H, float(res), float(ox), float(oy),
               None if occ is None else zlib.adler32(occ.view(np.uint8)),
               None if dist is None else zlib.adler32(dist.view(np.uint8)))
        if key != self.map_key:
            self.h.call("mcl_set_map", C.c_void_p(occ.ctypes.data) if occ is not None else None,
                        C.c_void_p(dist.ctypes.data) if dist is not None else None, int(W), int(H),
                        float(res), float(ox), float(oy))
            self.map_key = key
            self.sensor_key = None

    def set_sensor(self, sigma_hit, z_hit, z_rand, max_range, step):
        key = (float(sigma_hit), float(z_hit), float(z_rand), float(max_range), int(step))
        if key != self.sensor_key:
            self.h.call("mcl_set_sensor", *key)
            self.sensor_key = key

    def soa(self, particles):
        p = np.ascontiguousarray(particles, dtype=np.float64)
        n = p.shape[0]
        aos = torch.from_numpy(p).to(self.device)
        x, y, t = (torch.empty(n, dtype=torch.float64, device=self.device) for _ in range(3))
        self.h.call("mcl_aos_to_soa", _p(aos), n, _p(x), _p(y), _p(t))
        return x, y, t

    def aos(self, x, y, t):
        n = x.shape[0]
        out = torch.empty((n, 3), dtype=torch.float64, device=self.device)
        self.h.call("mcl_soa_to_aos", _p(x), _p(y), _p(t), n, _p(out))
        return out.cpu().numpy()


def _p(t):
    return C.c_void_p(t.data_ptr()) if t is not None else None


def _ctx(device=None):
    device = torch.cuda.current_device() if device is None else int(device)
    c = _state["ctx"].get(device)
    if c is None:
        c = _state["ctx"][device] = _Ctx(device)
    c.bind()
    return c


def _dev(c, a, dtype):
    return torch.from_numpy(np.ascontiguousarray(a, dtype=dtype)).to(c.device)


# ------------------------------------------------------------------------------------------------
def normalize_angle(theta):
    """pu:62-67 (scalar, host): (theta + pi) % (2 pi) - pi."""
    return (theta + np.pi) % (2 * np.pi) - np.pi


def normalize_angle_array(angles, mean_angle):
    """pu:69-83 -> float32 array."""
    c = _ctx()
    a = _dev(c, angles, np.float64)
    out = torch.empty(a.shape[0], dtype=torch.float32, device=c.device)
    c.h.call("mcl_normalize_angle_array", _p(a), float(mean_angle), a.shape[0], _p(out))
    return out.cpu().numpy()


def compute_likelihoods(scan_ranges, angles, particles, distance_map, map_resolution, map_origin,
                        width, height, sigma_hit=0.35, z_hit=0.9, z_rand=0.1, max_range=10, step=1):
    """pu:85-149.  Returns (N,) float32 mean log-likelihood per particle."""
    c = _ctx()
    c.set_map(None, distance_map, width, height, map_resolution, map_origin[0], map_origin[1])
    c.set_sensor(sigma_hit, z_hit, z_rand, max_range, step)
    r = np.ascontiguousarray(scan_ranges, dtype=np.float32)
    a = np.ascontiguousarray(angles, dtype=np.float32)
    c.h.call("mcl_set_scan", C.c_void_p(r.ctypes.data), C.c_void_p(a.ctypes.data), len(r))
    n = len(particles)
    if n == 0:
        return np.zeros(0, np.float32)
    x, y, t = c.soa(particles)
    score = torch.empty(n, dtype=torch.float32, device=c.device)
    c.h.call("mcl_likelihood", _p(x), _p(y), _p(t), n, _p(score))
    return score.cpu().numpy()


def compute_likelihoods_raycast(scan_ranges, angles, particles, grid_map, map_resolution, limits):
    """pu:151-201 (ray-marching beam model, the reference's hard-coded parameters)."""
    c = _ctx()
    g = np.asarray(grid_map)
    blocked = np.ascontiguousarray(g > 0.5, dtype=np.uint8)
    key = (g.shape, float(map_resolution), float(limits[0]), float(limits[2]), zlib.adler32(blocked))
    if getattr(c, "rc_key", None) != key:
        c.h.call("mcl_set_raycast_grid", C.c_void_p(blocked.ctypes.data), int(g.shape[1]), int(g.shape[0]),
                 float(map_resolution), float(limits[0]), float(limits[2]))
        c.rc_key = key
    r = np.ascontiguousarray(scan_ranges, dtype=np.float32)
    a = np.ascontiguousarray(angles, dtype=np.float32)
    n = len(particles)
    if n == 0:
        return np.zeros(0, np.float32)
    x, y, t = c.soa(particles)
    score = torch.empty(n, dtype=torch.float32, device=c.device)
    c.h.call("mcl_likelihood_raycast", C.c_void_p(r.ctypes.data), C.c_void_p(a.ctypes.data), len(r), _p(x), _p(y), _p(t),
             n, _p(score))
    return score.cpu().numpy()


def convert_scores(scores):
    """node:351-358 softmax of the scores (float32)."""
    c = _ctx()
    s = _dev(c, scores, np.float32)
    w = torch.empty_like(s)
    c.h.call("mcl_softmax", _p(s), s.shape[0], _p(w), None, None)
    return w.cpu().numpy()


def mh_resampling(particles, proposed_particles, likelihoods, old_weights, uniforms=None,
                  return_accept=False):
    """pu:208-236 -> (new_particles (N,3) f64, new_weights (N,) f32)."""
    c = _ctx()
    n = len(particles)
    x, y, t = c.soa(particles)
    px, py, pt = c.soa(proposed_particles)
    lk, ow = _dev(c, likelihoods, np.float32), _dev(c, old_weights, np.float32)
    u = _dev(c, uniforms, np.float64) if uniforms is not None else None
    xo, yo, to = (torch.empty_like(x) for _ in range(3))
    wo = torch.empty_like(lk)
    acc = torch.empty(n, dtype=torch.uint8, device=c.device)
    c.h.call("mcl_mh_accept", _p(x), _p(y), _p(t), _p(px), _p(py), _p(pt), _p(lk), _p(ow), n, _p(u),
             _state["seed"], _tick(), 0, _p(xo), _p(yo), _p(to), _p(wo), _p(acc))
    out = (c.aos(xo, yo, to), wo.cpu().numpy())
    return out + (acc.cpu().numpy(),) if return_accept else out


def motion_model_odometry_parallel(particles_prev, particles_curr, delta, alpha):
    """pu:282-330 -> (N,) float64 transition densities, normalised by their population sum."""
    c = _ctx()
    al = np.ascontiguousarray(alpha, dtype=np.float32)
    c.h.call("mcl_set_motion", al.ctypes.data_as(_pf))
    px, py, pt = c.soa(particles_prev)
    cx, cy, ct = c.soa(particles_curr)
    n = px.shape[0]
    probs = torch.empty(n, dtype=torch.float64, device=c.device)
    d = (C.c_double * 3)(float(delta[0]), float(delta[1]), float(delta[2]))
    c.h.call("mcl_motion_density", _p(px), _p(py), _p(pt), _p(cx), _p(cy), _p(ct), n, d, _p(probs), None, 1)
    return probs.cpu().numpy()


def assym_mh_resampling(particles, proposed_particles, likelihoods, old_weights, trans_forward, trans_backward,
                        uniforms=None, return_accept=False, corrected=False):
    """pu:238-276 -> (new_particles, new_weights); the reference's always-accept quirk included.
    corrected=True: the Metropolis-Hastings ratio applied unconditionally."""
    c = _ctx()
    n = len(particles)
    x, y, t = c.soa(particles)
    px, py, pt = c.soa(proposed_particles)
    lk, ow = _dev(c, likelihoods, np.float32), _dev(c, old_weights, np.float32)
    tf, tb = _dev(c, trans_forward, np.float64), _dev(c, trans_backward, np.float64)
    u = _dev(c, uniforms, np.float64) if uniforms is not None else None
    xo, yo, to = (torch.empty_like(x) for _ in range(3))
    wo = torch.empty_like(lk)
    acc = torch.empty(n, dtype=torch.uint8, device=c.device)
    c.h.call("mcl_assym_mh_accept_ex", _p(x), _p(y), _p(t), _p(px), _p(py), _p(pt), _p(lk), _p(ow), _p(tf), _p(tb), n,
             _p(u), _state["seed"], _tick(), 0, _p(xo), _p(yo), _p(to), _p(wo), _p(acc), 1 if corrected else 0)
    out = (c.aos(xo, yo, to), wo.cpu().numpy())
    return out + (acc.cpu().numpy(),) if return_accept else out


def apply_motion_model_parallel(particles, delta, alpha, map_data, map_resolution, origin_x, origin_y,
                                width, height, normals=None, return_attempts=False, max_attempts=1000):
    """pu:332-363 -> (N,3) f64 proposed particles.  normals: optional injected (N, A, 3) draws."""
    c = _ctx()
    c.set_map(map_data, None, width, height, map_resolution, origin_x, origin_y)
    al = np.ascontiguousarray(alpha, dtype=np.float32)
    c.h.call("mcl_set_motion", al.ctypes.data_as(_pf))
    n = len(particles)
    x, y, t = c.soa(particles)
    xo, yo, to = (torch.empty_like(x) for _ in range(3))
    att = torch.empty(n, dtype=torch.int32, device=c.device)
    z = _dev(c, normals, np.float64) if normals is not None else None
    d = (C.c_double * 3)(float(delta[0]), float(delta[1]), float(delta[2]))
    c.h.call("mcl_predict", _p(x), _p(y), _p(t), n, d, _state["seed"], _tick(), 0, _p(z),
             int(z.shape[1]) if z is not None else 0, int(max_attempts), _p(xo), _p(yo), _p(to), _p(att))
    out = c.aos(xo, yo, to)
    return (out, att.cpu().numpy()) if return_attempts else out


def low_variance_resample_indices(weights, N, r=None, mode=_lib.RESAMPLE_REFERENCE_F32):
    c = _ctx()
    w = _dev(c, weights, np.float32)
    N = int(N)
    if r is None:
        r = c.h.lib.mcl_resample_offset(_state["seed"], _tick(), N)
    idx = torch.empty(N, dtype=torch.int32, device=c.device)
    c.h.call("mcl_resample_indices", _p(w), w.shape[0], N, float(r), int(mode), _p(idx))
    return idx.cpu().numpy()


def low_variance_resample_numba(particles, weights, N, r=None, mode=_lib.RESAMPLE_REFERENCE_F32):
    """pu:416-446 -> (new_particles, uniform float32 weights 1/N).  particles may be (N,3) poses or
    any array indexed along axis 0 (the node passes np.arange(N) at node:467)."""
    idx = low_variance_resample_indices(weights, N, r, mode)
    particles = np.asarray(particles)
    return particles[idx].copy(), np.full(int(N), 1.0 / N, dtype=np.float32)


def kld_sampling_amcl(particles, weights, bin_size_xy, bin_size_theta, epsilon, z, max_samples, min_particles,
                      r=None, normals=None, mode=_lib.RESAMPLE_REFERENCE_F32):
    """pu:529-591 -> (count, 3) float32 sampled particles.  r / normals: optional injected draws."""
    c = _ctx()
    x, y, t = c.soa(particles)
    w = _dev(c, weights, np.float32)
    ms = int(max_samples)
    if ms == 0:
        return np.zeros((0, 3), np.float32)
    zz = _dev(c, normals, np.float64) if normals is not None else None
    tick = _tick()
    if r is None:
        r = c.h.lib.mcl_resample_offset(_state["seed"], tick, ms)
    xo, yo, to = (torch.empty(ms, dtype=torch.float64, device=c.device) for _ in range(3))
    cnt = C.c_int64(0)
    c.h.call("mcl_kld_resample", _p(x), _p(y), _p(t), _p(w), x.shape[0], ms, int(min_particles), float(bin_size_xy),
             float(bin_size_theta), float(epsilon), float(z), float(r), _p(zz), _state["seed"], tick, int(mode),
             _p(xo), _p(yo), _p(to), C.byref(cnt))
    k = cnt.value
    return c.aos(xo[:k].contiguous(), yo[:k].contiguous(), to[:k].contiguous()).astype(np.float32)


def generate_valid_particles(num_particles, map_data, map_resolution, origin_x, origin_y, width, height,
                             uniforms=None):
    """pu:450-465.  uniforms: optional injected (3, max(50 N, 500)) draws (ux | uy | utheta) for the
    bit-exact restatement; default: per-particle Philox rejection sampling (always returns N)."""
    c = _ctx()
    c.set_map(map_data, None, width, height, map_resolution, origin_x, origin_y)
    n = int(num_particles)
    x, y, t = (torch.empty(n, dtype=torch.float64, device=c.device) for _ in range(3))
    cnt = C.c_int64(0)
    if uniforms is not None:
        u = _dev(c, uniforms, np.float64)
        c.h.call("mcl_init_uniform", n, _p(u), int(u.shape[1]), 0, 0, _p(x), _p(y), _p(t), C.byref(cnt))
    else:
        c.h.call("mcl_init_uniform", n, None, 0, _state["seed"] + _tick(), 0, _p(x), _p(y), _p(t), C.byref(cnt))
    k = cnt.value
    return c.aos(x[:k].contiguous(), y[:k].contiguous(), t[:k].contiguous())


def estimate(particles, weights):
    """node:586-597 publish_estimate arithmetic -> (mean_x, mean_y, mean_theta, cov 3x3)."""
    from .localizer import assemble_estimate
    c = _ctx()
    x, y, t = c.soa(particles)
    w = _dev(c, weights, np.float32)
    out = (C.c_double * 16)()
    c.h.call("mcl_estimate", _p(x), _p(y), _p(t), _p(w), x.shape[0], out)
    return assemble_estimate(list(out))


# ---- names the node imports at node:13 without reaching them from its callbacks --------
def compute_valid_indices(particles, map_data, map_resolution, origin_x, origin_y, width, height):
    """pu:369-386 -> int32 indices (ascending) of the particles on in-map cells with map_data <= 10."""
    c = _ctx()
    c.set_map(map_data, None, width, height, map_resolution, origin_x, origin_y)
    n = len(particles)
    if n == 0:
        return np.zeros(0, np.int32)
    x, y, _ = c.soa(particles)
    idx = torch.empty(n, dtype=torch.int32, device=c.device)
    cnt = C.c_int64(0)
    c.h.call("mcl_compute_valid_indices", _p(x), _p(y), n, _p(idx), C.byref(cnt))
    return idx[:cnt.value].cpu().numpy()


def validate_samples(samples, distance_map, resolution, origin):
    """pu:600-614 -> (N,3) float64 copy of samples, the invalid ones zeroed.  distance_map is (H, W)."""
    c = _ctx()
    d = np.asarray(distance_map)
    c.set_map(None, d, d.shape[1], d.shape[0], resolution, origin[0], origin[1])
    x, y, t = c.soa(samples)
    c.h.call("mcl_validate_samples", _p(x), _p(y), _p(t), x.shape[0])
    return c.aos(x, y, t)


def initialize_gaussian_parallel(mean, cov, num_particles, distance_map, resolution, origin):
    """pu:594-598 (node:183): the samples come from NumPy's global generator exactly like the reference's (it is
    plain Python there); the validation kernel zeroes the invalid ones."""
    samples = np.random.multivariate_normal(mean, cov, size=num_particles)
    return validate_samples(samples, distance_map, resolution, origin)


def parallel_resample_simple(particles, weights, N, uniforms=None):
    """pu:467-477 multinomial resampling -> new particles (same shape/dtype as `particles`).  uniforms: optional
    injected (N,) draws.  Where the reference would read out of bounds (Appendix C #7) the last particle is taken."""
    c = _ctx()
    w = _dev(c, weights, np.float32)
    N = int(N)
    u = _dev(c, uniforms, np.float64) if uniforms is not None else None
    idx = torch.empty(N, dtype=torch.int32, device=c.device)
    c.h.call("mcl_resample_multinomial", _p(w), w.shape[0], N, _p(u), _state["seed"], _tick(), _p(idx))
    particles = np.asarray(particles)
    out = np.empty_like(particles)
    out[:N] = particles[idx.cpu().numpy()]
    return out


def low_variance_resample_amcl(particles, weights, target_size, r=None):
    """pu:486-502 -> ((target_size, 3) float32 particles, float64 weights 1/target_size)."""
    c = _ctx()
    w = _dev(c, weights, np.float32)
    T = int(target_size)
    if r is None:
        r = c.h.lib.mcl_resample_offset(_state["seed"], _tick(), T)
    idx = torch.empty(T, dtype=torch.int32, device=c.device)
    c.h.call("mcl_resample_indices", _p(w), w.shape[0], T, float(r), _lib.RESAMPLE_AMCL_F32, _p(idx))
    return np.asarray(particles)[idx.cpu().numpy()].astype(np.float32), np.full(T, 1.0 / T)


def reinitialize_particles_numba(num_new, occupancy_map, res, origin_x, origin_y, choice=None, theta=None):
    """pu:504-526 -> (num_new, 3) float32 poses on the corners of uniformly chosen free cells.  occupancy_map is
    (H, W); choice / theta: optional injected draws (index into the row-major free-cell list, heading)."""
    c = _ctx()
    occ = np.asarray(occupancy_map)
    c.set_map(occ, None, occ.shape[1], occ.shape[0], res, origin_x, origin_y)
    n = int(num_new)
    x, y, t = (torch.empty(n, dtype=torch.float64, device=c.device) for _ in range(3))
    ch = _dev(c, choice, np.int64) if choice is not None else None
    th = _dev(c, theta, np.float64) if theta is not None else None
    nf = C.c_int64(0)
    c.h.call("mcl_reinitialize_particles", n, _p(ch), _p(th), _state["seed"], _tick(), _p(x), _p(y), _p(t), C.byref(nf))
    return c.aos(x, y, t).astype(np.float32)
