"""Stateful, device-resident localizer: the Python-facing surface of the reference's ROS node
(app/scripts/amcmh_localizer.py, class AMCMHLocalizer) without the ROS message handling.

    node.load_map            (node:124-177)  -> Localizer.load_map
    rosparam reads           (node:18-58)    -> Localizer.set_params (same keys as amhmcl.yaml)
    initialize_particles     (node:179-197)  -> init_uniform / init_gaussian / set_particles
    odom_callback / move_particles (node:379-408) -> predict(odom_xytheta)
    lidar_callback: update_scans + update_weights + update_particles_mh (node:294-322) -> update(...)
    publish_estimate math    (node:584-597)  -> estimate()
    resample_lvr             (node:488-492)  -> resample()

Particles stay on the GPU as SoA fp64 torch tensors; per step the host sends one scan (M floats)
and one odometry pose and reads back the 3+9 numbers of the estimate.  Every arithmetic step is a
libmcl.so kernel (include/mcl.h); torch only allocates buffers.  No CPU fallback.
"""
import ctypes as C
import threading
import warnings

import numpy as np
import torch

from . import _lib
from .maps import GridMap
from .params import DEFAULT_PARAMS, mode_flags

_pd = C.POINTER(C.c_double)


def _ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else None


def _dbl3(v):
    return (C.c_double * 3)(float(v[0]), float(v[1]), float(v[2]))


class Localizer:
    def __init__(self, device=0, params=None, mode=None, seed=0, resample_mode="reference",
                 max_attempts=1000, amh_corrected=False):
        """amh_corrected: in the asymmetric-MH modes use the corrected accept rule and backward increment instead
        of the reference's (see mcl_filter_set_assym in include/mcl.h).  Default: the
        reference's behaviour, quirks included."""
        self.amh_corrected = bool(amh_corrected)
        if not torch.cuda.is_available():
            raise RuntimeError("mcmh_localization_b200.Localizer needs a CUDA device (no CPU fallback)")
        self.device = torch.device("cuda", int(device))
        self.h = _lib.Handle(int(device))
        self._lock = threading.Lock()      # rospy runs the two callbacks on two threads
        self.seed = int(seed)
        self.first_index = 0               # global index of local particle 0 (sharded runs)
        self.max_attempts = int(max_attempts)
        self.resample_mode = {"reference": _lib.RESAMPLE_REFERENCE_F32,
                              "fixed": _lib.RESAMPLE_FIXED_POINT}[resample_mode]
        self.params = dict(DEFAULT_PARAMS)
        self.map = None
        self.n = 0
        self.last_odom = None
        self.delta = (0.0, 0.0, 0.0)
        self.sets = None
        self.w_slow = 1e-3                 # node:86-87
        self.w_fast = 1e-3
        self.num_particles = 0             # node:26 / node:520 (lags one scan behind in the KLD modes)
        self.capacity = 0
        self.set_params(params or {}, mode=mode)

    # ------------------------------------------------------------------ configuration
    def set_params(self, params, mode=None):
        """Accepts the keys of app/params/amhmcl.yaml (unknown keys are kept but unused)."""
        self.params.update(params)
        if mode is not None:
            self.params["localization_mode"] = mode
        p = self.params
        f = mode_flags(p["localization_mode"])
        self.use_mh, self.use_adaptive, self.assym = f["use_mh"], f["use_adaptive"], f["assym"]
        self.alpha = np.array([p["alpha1"], p["alpha2"], p["alpha3"], p["alpha4"]], dtype=np.float32)  # node:28-33
        self._bind_stream()
        self.h.call("mcl_set_sensor", float(p["sigma_hit"]), float(p["z_hit"]), float(p["z_rand"]),
                    float(p["max_range"]), int(p["step"]))
        self.h.call("mcl_set_motion", self.alpha.ctypes.data_as(C.POINTER(C.c_float)))
        if self.n:
            self.h.call("mcl_filter_configure", int(self.use_mh), self.resample_mode, self.seed, self.first_index, -1)
            self.h.call("mcl_filter_set_assym", (2 if self.amh_corrected else 1) if self.assym else 0)

    def load_map(self, occ, resolution=None, origin_xy=None, gpu_edt=False):
        """occ: (H,W) int8 OccupancyGrid payload, or a GridMap (maps.load_map_yaml / map_from_occupancy).
        gpu_edt: compute the distance map (node:153-157) on the device instead of with SciPy (bit-identical)."""
        if isinstance(occ, GridMap):
            gm = occ
        elif gpu_edt:
            o = np.ascontiguousarray(occ, dtype=np.int8)
            dist = np.empty(o.shape, np.float32)
            self._bind_stream()
            self.h.call("mcl_set_map_edt", C.c_void_p(o.ctypes.data), int(o.shape[1]), int(o.shape[0]), float(resolution),
                        float(origin_xy[0]), float(origin_xy[1]), C.c_void_p(dist.ctypes.data))
            self.map = GridMap(o, dist, float(resolution), float(origin_xy[0]), float(origin_xy[1]))
            return
        else:
            from .maps import map_from_occupancy
            gm = map_from_occupancy(occ, resolution, origin_xy[0], origin_xy[1])
        self.map = gm
        self._bind_stream()
        self.h.call("mcl_set_map", C.c_void_p(gm.occ.ctypes.data), C.c_void_p(gm.dist.ctypes.data),
                    gm.width, gm.height, gm.resolution, gm.origin_x, gm.origin_y)

    def _bind_stream(self):
        self.h.call("mcl_set_stream", C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream))

    # ------------------------------------------------------------------ particle buffers
    def _alloc(self, n):
        """Allocate the device buffers (torch tensors) and bind them to the library's filter state."""
        self.sets = self._make_sets(n)                      # three SoA pose sets; roles live in the library
        f32 = lambda: torch.empty(n, dtype=torch.float32, device=self.device)
        self.score_pre, self.score_post, self.w_pre, self.w_post = f32(), f32(), f32(), f32()
        self.wbuf = [torch.full((n,), 1.0 / n, dtype=torch.float32, device=self.device), f32()]   # node:98
        self.idx = torch.empty(n, dtype=torch.int32, device=self.device)
        self.est18 = torch.zeros(18, dtype=torch.float64, device=self.device)
        self.n = n
        self.capacity = n
        self.num_particles = n
        arr = lambda k: (C.c_void_p * 3)(*[self.sets[j][k].data_ptr() for j in range(3)])
        self.h.call("mcl_filter_bind", n, arr(0), arr(1), arr(2), _ptr(self.score_pre), _ptr(self.score_post),
                    _ptr(self.w_pre), _ptr(self.w_post), _ptr(self.wbuf[0]), _ptr(self.wbuf[1]), _ptr(self.idx),
                    int(self.use_mh), self.resample_mode, self.seed, self.first_index, self.max_attempts)
        self.h.call("mcl_filter_set_assym", (2 if self.amh_corrected else 1) if self.assym else 0)

    def _make_sets(self, n):
        mk = lambda: [torch.empty(n, dtype=torch.float64, device=self.device) for _ in range(3)]
        return [mk(), mk(), mk()]

    def _roles(self):
        r = (C.c_int * 4)()
        t = C.c_uint64(0)
        self.h.call("mcl_filter_roles", r, C.byref(t))
        return r[0], r[1], r[2], r[3], t.value

    @property
    def cur(self):
        return self.sets[self._roles()[0]]

    @property
    def prev(self):
        return self.sets[self._roles()[1]]

    @property
    def weights_t(self):
        return self.wbuf[self._roles()[3]]

    @property
    def tick(self):
        return self._roles()[4] if self.n else 0

    def set_particles(self, particles, prev=None, keep_odom=False):
        """(N,3) float64 host array -> device SoA; particles_prev = prev or a copy (node:95-97)."""
        p = np.ascontiguousarray(particles, dtype=np.float64)
        with self._lock:
            self._bind_stream()
            if p.shape[0] != self.n or self.sets is None:
                self._alloc(p.shape[0])
            aos = torch.from_numpy(p).to(self.device)
            self.h.call("mcl_aos_to_soa", _ptr(aos), self.n, *[_ptr(t) for t in self.cur])
            if prev is not None:
                q = torch.from_numpy(np.ascontiguousarray(prev, dtype=np.float64)).to(self.device)
                self.h.call("mcl_aos_to_soa", _ptr(q), self.n, *[_ptr(t) for t in self.prev])
            else:
                for a, b in zip(self.prev, self.cur):
                    a.copy_(b)
            if not keep_odom:
                self.last_odom = None

    def init_uniform(self, n, seed=None, uniforms=None):
        """node:188 generate_valid_particles.  uniforms: optional (3, max(50n,500)) injected draws for
        the bit-exact restatement; default: per-particle Philox rejection sampling."""
        with self._lock:
            self._bind_stream()
            self._alloc(int(n))
            cnt = C.c_int64(0)
            if uniforms is not None:
                u = torch.from_numpy(np.ascontiguousarray(uniforms, dtype=np.float64)).to(self.device)
                self.h.call("mcl_init_uniform", self.n, _ptr(u), int(u.shape[1]), 0, 0,
                            *[_ptr(t) for t in self.cur], C.byref(cnt))
                if cnt.value < self.n:          # pu:462-465 may return fewer than N
                    k = cnt.value
                    kept = [t[:k].clone() for t in self.cur]
                    self._alloc(k)
                    for a, b in zip(self.cur, kept):
                        a.copy_(b)
            else:
                s = self.seed if seed is None else int(seed)
                self.h.call("mcl_init_uniform", self.n, None, 0, s, self.first_index,
                            *[_ptr(t) for t in self.cur], C.byref(cnt))
            for a, b in zip(self.prev, self.cur):
                a.copy_(b)
            self.last_odom = None

    def init_gaussian(self, mean, cov, n, seed=None):
        """node:183 initialize_gaussian_parallel (pu:594-614).  Host-side, one-off."""
        self.set_particles(gaussian_particles(mean, cov, n, self.map, self.seed if seed is None else int(seed)))

    # ------------------------------------------------------------------ predict (odom_callback)
    def predict(self, odom, normals=None):
        """node:384-408 move_particles.  odom = (x, y, yaw).  normals: optional injected draws
        (N, A, 3) (tests); default Philox."""
        with self._lock:
            self._bind_stream()
            cur_odom = np.asarray(odom, dtype=np.float64)
            if self.last_odom is not None:
                self.delta = compute_motion(self.last_odom, cur_odom)
                zp, A = None, 0
                if normals is not None:
                    z = normals if torch.is_tensor(normals) else torch.from_numpy(
                        np.ascontiguousarray(normals, dtype=np.float64))
                    z = z.to(self.device)
                    zp, A = _ptr(z), int(z.shape[1])
                self.h.call("mcl_filter_predict", _dbl3(self.delta), zp, A)
                self._push_transition()
            self.last_odom = cur_odom

    # ------------------------------------------------------------------ update (lidar_callback)
    def set_scan(self, ranges, angle_min=None, angle_max=None, angles=None):
        r = np.ascontiguousarray(ranges, dtype=np.float32)                       # node:343
        if angles is None:
            angles = np.linspace(angle_min, angle_max, len(r), dtype=np.float32)  # node:346-348
        a = np.ascontiguousarray(angles, dtype=np.float32)
        self._bind_stream()
        # (__array_interface__ instead of .ctypes: 4 us less per scan)
        self.h.call("mcl_set_scan", C.c_void_p(r.__array_interface__["data"][0]),
                    C.c_void_p(a.__array_interface__["data"][0]), len(r))

    def update(self, ranges, angle_min=None, angle_max=None, angles=None, uniforms=None):
        """node:296-322: update_scans, update_weights (both particle sets), MH accept by mode."""
        with self._lock:
            self.set_scan(ranges, angle_min, angle_max, angles)
            self._update_core(uniforms)

    def update_chain(self, ranges, angle_min=None, angle_max=None, angles=None, iters=32):
        """update() with `iters` MH iterations per scan (BASELINE config 4); iters=1 == update() in MHMCL."""
        with self._lock:
            if ranges is not None:
                self.set_scan(ranges, angle_min, angle_max, angles)
            else:
                self._bind_stream()
            self.h.call("mcl_filter_update_chain", int(iters))

    def stage_scans(self, ranges_km, angles):
        """Pre-stage K scans on the device (bag replay / device-resident benchmark inputs)."""
        r = np.ascontiguousarray(ranges_km, dtype=np.float32)
        a = np.ascontiguousarray(angles, dtype=np.float32)
        self._bind_stream()
        self.h.call("mcl_set_scan_batch", C.c_void_p(r.ctypes.data), C.c_void_p(a.ctypes.data),
                    int(r.shape[1]), int(r.shape[0]))

    def update_staged(self, k, uniforms=None):
        """update() on pre-staged scan k: no host->device traffic."""
        with self._lock:
            self._bind_stream()
            self.h.call("mcl_use_scan", int(k))
            self._update_core(uniforms)

    def estimate_async(self, out18):
        """Non-blocking estimate into a device tensor of 18 float64 (see include/mcl.h)."""
        with self._lock:
            self._bind_stream()
            self.h.call("mcl_filter_estimate", _ptr(out18), None)

    def _push_transition(self):
        """node:429-434 backward increment with the node's own NumPy calls (bit-exact), for AMH modes."""
        if self.assym and self.amh_corrected:
            self.h.call("mcl_filter_set_transition", _dbl3(self.delta), None)     # the library derives the true inverse
        elif self.assym:
            dx, dy, dth = self.delta
            db = (-dx * np.cos(dth) - dy * np.sin(dth), dx * np.sin(dth) - dy * np.cos(dth), -dth)
            self.h.call("mcl_filter_set_transition", _dbl3(self.delta), _dbl3(db))

    def _update_core(self, uniforms=None):
        up = None
        if uniforms is not None:
            u = uniforms if torch.is_tensor(uniforms) else torch.from_numpy(
                np.ascontiguousarray(uniforms, dtype=np.float64))
            u = u.to(self.device)
            up = _ptr(u)
        self.h.call("mcl_filter_update", up)
        if self.use_adaptive:
            self._update_acml_weights()

    # ------------------------------------------------------------------ KLD-adaptive modes (AMCL family)
    def _update_acml_weights(self):
        """node:276-286: normalise the weights, then the slow / fast running averages of w_avg."""
        out = (C.c_double * 2)()
        self.h.call("mcl_weights_normalize", _ptr(self.weights_t), self.n, out)
        w_avg = float(np.float32(out[1]))                     # np.mean of a float32 array
        self.w_slow += self.params["alpha_slow"] * (w_avg - self.w_slow)
        self.w_fast += self.params["alpha_fast"] * (w_avg - self.w_fast)

    def _resample_amcl_kld(self, r=None, normals=None):
        """node:496-527 resample_amcl_kld: KLD-adaptive systematic resampling (pu:529-591) of
        N - N_random particles plus N_random uniformly re-initialised ones (pu:450-465); N changes."""
        p = self.params
        p_random = max(0.0, 1.0 - self.w_fast / (self.w_slow + 1e-9))            # node:497
        N = self.num_particles
        n_random = min(int(p_random * N), self.capacity)
        n_resampled = N - n_random
        cur, prev, spare, ws, tick = self._roles()
        tick += 1
        if r is None and n_resampled > 0:
            r = self.h.lib.mcl_resample_offset(self.seed, tick, n_resampled)
        zp = None
        if normals is not None:
            z = torch.from_numpy(np.ascontiguousarray(normals, dtype=np.float64)).to(self.device)
            zp = _ptr(z)
        S = self.sets
        off = lambda t, k: C.c_void_p(t.data_ptr() + 8 * k)
        cnt = C.c_int64(0)
        if n_resampled > 0:
            self.h.call("mcl_kld_resample", *[_ptr(t) for t in S[cur]], _ptr(self.wbuf[ws]), self.n, n_resampled,
                        int(p["min_particles"]), float(p["kld_bin_size_xy"]), float(p["kld_bin_size_theta"]),
                        float(p["kld_epsilon"]), float(p["kld_z"]), float(r), zp, self.seed, tick, self.resample_mode,
                        *[off(t, n_random) for t in S[spare]], C.byref(cnt))
        if n_random > 0:                                                         # node:515-516
            c2 = C.c_int64(0)
            self.h.call("mcl_init_uniform", n_random, None, 0, self.seed ^ (tick << 20), self.first_index,
                        *[_ptr(t) for t in S[spare]], C.byref(c2))
        new_n = n_random + cnt.value
        self.num_particles = self.n                                              # node:520 (before the reassignment)
        self.h.call("mcl_filter_set_roles", (C.c_int * 4)(spare, prev, cur, ws), int(tick))
        self.n = new_n
        self.h.call("mcl_filter_set_n", new_n)
        self.wbuf[ws][:new_n].fill_(1.0 / new_n)                                  # node:522
        # particles_prev keeps the old length in the reference until the next odom message (node:404); keep
        # the two sets the same length here so that an update without a predict in between stays defined
        for a, b in zip(S[prev], S[spare]):
            a[:new_n].copy_(b[:new_n])

    # ------------------------------------------------------------------ estimate / resample
    def estimate(self):
        """node:586-597 -> (mean_x, mean_y, mean_theta, cov 3x3) with np.cov(aweights) semantics, or None with
        fewer than two particles (node:594-596 logs a warning and publishes nothing)."""
        if self.n < 2:
            warnings.warn("not enough particles for an estimate (node:594-596)")
            return None
        with self._lock:
            self._bind_stream()
            out = (C.c_double * 16)()
            self.h.call("mcl_filter_estimate", None, out)
        return assemble_estimate(list(out))

    def resample(self, r=None):
        """node:488-492 resample_lvr -> low_variance_resample_numba (pu:416-446)."""
        with self._lock:
            self._bind_stream()
            if self.use_adaptive:                                                 # node:329-331
                self._resample_amcl_kld(r)
                return
            self.h.call("mcl_filter_resample", -1.0 if r is None else float(r))
            # self.weights keeps the pre-resampling values (node:490 discards the uniform weights)

    def finish(self):
        """estimate() followed by resample() -- what lidar_callback does after the weights are updated (node:324-331)
        -- as one library call (fixed-N modes on one GPU: one launch of the tail kernel); returns the estimate."""
        if self.use_adaptive:
            est = self.estimate()
            self.resample()
            return est
        with self._lock:
            self._bind_stream()
            out = (C.c_double * 16)()
            self.h.call("mcl_filter_finish", None, out)
        if self.n < 2:
            return None
        return assemble_estimate(list(out))

    def finish_async(self, out18):
        """finish() with the estimate left in a device tensor of 18 float64 (no host round trip)."""
        with self._lock:
            self._bind_stream()
            if self.use_adaptive:
                self.h.call("mcl_filter_estimate", _ptr(out18), None)
                self._resample_amcl_kld(None)
                return
            self.h.call("mcl_filter_finish", _ptr(out18), None)

    def step_chain(self, odom, ranges, angle_min=None, angle_max=None, angles=None, iters=32):
        """One odom message followed by one scan with `iters` MH iterations per particle (BASELINE config 4):
        predict -> update_chain -> estimate -> resample."""
        self.predict(odom)
        self.update_chain(ranges, angle_min, angle_max, angles, iters=iters)
        return self.finish()

    def step(self, odom, ranges, angle_min=None, angle_max=None, angles=None):
        """One odom message followed by one scan: predict -> update -> estimate -> resample, enqueued by
        ONE library call; returns the host estimate while the resampling kernels are still running."""
        with self._lock:
            cur_odom = np.asarray(odom, dtype=np.float64)
            d = None
            if self.last_odom is not None:
                self.delta = compute_motion(self.last_odom, cur_odom)
                d = _dbl3(self.delta)
            self.last_odom = cur_odom
            if d is not None and not self.use_adaptive and not self.assym:
                # the motion kernels do not need the scan: enqueue them first and build the beam table of the scan
                # on the host while they run (same calls, same results as mcl_filter_step(d, ...))
                self._bind_stream()
                self.h.call("mcl_filter_predict", d, None, 0)
                d = None
            self.set_scan(ranges, angle_min, angle_max, angles)
            if self.use_adaptive:                 # N changes per scan: sequence the stages from the host
                if d is not None:
                    self.h.call("mcl_filter_predict", d, None, 0)
                    self._push_transition()
                self._update_core()
                out = (C.c_double * 16)()
                self.h.call("mcl_filter_estimate", None, out)
                self._resample_amcl_kld()
                return assemble_estimate(list(out))
            out = (C.c_double * 16)()
            if self.assym and d is not None:       # AMH: predict first so the NumPy backward increment can be pushed
                self.h.call("mcl_filter_predict", d, None, 0)
                self._push_transition()
                d = None
            self.h.call("mcl_filter_step", d, -1, None, out)
        if self.n < 2:                                   # node:594-596: no estimate is published
            return None
        return assemble_estimate(list(out))

    def step_staged(self, odom, k, out18=None):
        """step() on pre-staged scan k with the estimate left on the device (no host<->device traffic)."""
        with self._lock:
            self._bind_stream()
            cur_odom = np.asarray(odom, dtype=np.float64)
            d = None
            if self.last_odom is not None:
                self.delta = compute_motion(self.last_odom, cur_odom)
                d = _dbl3(self.delta)
            self.last_odom = cur_odom
            self.h.call("mcl_filter_step", d, int(k), _ptr(out18 if out18 is not None else self.est18), None)

    # ------------------------------------------------------------------ read-back
    def _aos(self, soa):
        out = torch.empty((self.n, 3), dtype=torch.float64, device=self.device)
        self.h.call("mcl_soa_to_aos", *[_ptr(t) for t in soa], self.n, _ptr(out))
        return out.cpu().numpy()

    def particles(self):
        with self._lock:
            self._bind_stream()
            return self._aos(self.cur)

    def particles_sample(self, max_count=2000):
        """Every (n / max_count)-th particle with its weight, sliced on the device: what a visualisation topic needs
        (the node builds one Marker per particle in Python, node:538-581 -- unusable at 1 M particles).
        Returns ((k, 3) float64 poses, (k,) float32 weights)."""
        with self._lock:
            self._bind_stream()
            stride = max(1, -(-self.n // max(1, int(max_count))))
            x, y, t = (c[:self.n:stride] for c in self.cur)
            poses = torch.stack((x, y, t), dim=1).cpu().numpy()
            return poses, self.weights_t[:self.n:stride].cpu().numpy()

    def particles_prev(self):
        with self._lock:
            self._bind_stream()
            return self._aos(self.prev)

    def weights(self):
        with self._lock:
            self._bind_stream()
            return self.weights_t[:self.n].cpu().numpy()

    def set_weights(self, w):
        with self._lock:
            self._bind_stream()
            self.weights_t[:self.n].copy_(torch.from_numpy(np.ascontiguousarray(w, dtype=np.float32)))

    def scores(self):
        with self._lock:
            self._bind_stream()
            return self.score_pre[:self.n].cpu().numpy(), self.score_post[:self.n].cpu().numpy()

    def sync(self):
        self.h.call("mcl_sync")

    def close(self):
        self.h.close()


def gaussian_particles(mean, cov, n, gm, seed):
    """pu:594-614 initialize_gaussian_parallel + validate_samples: N(mean, cov) samples from NumPy's legacy
    generator (np.random.multivariate_normal after np.random.seed(seed)); a sample is kept if its cell
    (int() truncation, no lower-bound quirk: 0 <= mx) lies in the map and distance_map < 1.0, otherwise it is
    replaced by (0, 0, 0) like the reference does."""
    rs = np.random.RandomState(int(seed))
    s = rs.multivariate_normal(np.asarray(mean, float), np.asarray(cov, float), size=int(n))
    mx = np.trunc((s[:, 0] - gm.origin_x) / gm.resolution).astype(np.int64)
    my = np.trunc((s[:, 1] - gm.origin_y) / gm.resolution).astype(np.int64)
    ok = (mx >= 0) & (mx < gm.width) & (my >= 0) & (my < gm.height)
    ok[ok] &= gm.dist[my[ok], mx[ok]] < 1.0
    s[~ok] = 0.0
    return s


def compute_motion(odom1, odom2):
    """node:410-421, host scalars, the node's own NumPy calls (np.arctan2 / np.hypot differ from
    glibc by an ulp now and then, so this stays NumPy; mcl_compute_motion is the C-host variant)."""
    dx = odom2[0] - odom1[0]
    dy = odom2[1] - odom1[1]
    dtheta = (odom2[2] - odom1[2] + np.pi) % (2 * np.pi) - np.pi      # pu:62-67 normalize_angle
    rot1 = np.arctan2(dy, dx) - odom1[2]
    trans = np.hypot(dx, dy)
    rot2 = dtheta - rot1
    return float(rot1), float(trans), float(rot2)


def assemble_estimate(o):
    """np.average / np.cov(aweights) from the 16 numbers of mcl_estimate (include/mcl.h).  Scalar arithmetic in the
    order of (sdd - np.outer(sd, sd) / v1) / fact: the same IEEE operations without 17 us of small-array overhead
    per step."""
    v1, v2 = o[0], o[1]
    s0, s1, s2 = o[5], o[6], o[7]
    fact = v1 - v2 / v1                       # np.cov: w_sum - ddof * sum(w * aweights) / w_sum
    if fact == 0.0 or v1 == 0.0:              # one effective particle: np.cov divides by zero (nan / inf entries)
        sd = np.array([s0, s1, s2])
        sdd = np.array([[o[8], o[9], o[10]], [o[9], o[11], o[12]], [o[10], o[12], o[13]]])
        with np.errstate(divide="ignore", invalid="ignore"):
            return o[2], o[3], o[4], (sdd - np.outer(sd, sd) / v1) / fact
    c00 = (o[8] - s0 * s0 / v1) / fact
    c01 = (o[9] - s0 * s1 / v1) / fact
    c02 = (o[10] - s0 * s2 / v1) / fact
    c11 = (o[11] - s1 * s1 / v1) / fact
    c12 = (o[12] - s1 * s2 / v1) / fact
    c22 = (o[13] - s2 * s2 / v1) / fact
    cov = np.array(((c00, c01, c02), (c01, c11, c12), (c02, c12, c22)))
    return o[2], o[3], o[4], cov
