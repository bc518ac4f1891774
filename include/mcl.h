/*
 * mcl.h -- C ABI of libmcl.so, the B200-native Monte-Carlo-localization core.
 *
 * The reference (gustavorvillela/mcmh_localization) has no FFI: its hot path is the set of
 * numba functions in app/scripts/parallel_utils.py ("pu") that the ROS node
 * app/scripts/amcmh_localizer.py ("node") imports by name at node:13.  Each entry point below
 * replaces one of those functions (or one block of glue arithmetic in the node) and says which.
 * A Python maintainer binds them with ctypes (INTEGRATION.md shows the stub);
 * mcmh_localization_b200/_lib.py is that binding.
 *
 * Conventions
 *  - plain pointers and sizes only; no torch / CUDA types in the signatures
 *    (a CUDA stream is passed as void*).
 *  - pointers named d_* are DEVICE pointers owned by the caller (PyTorch tensors in the Python
 *    host code); pointers named h_* are HOST pointers.  The library owns only its handle, the
 *    device copies of map / likelihood table / scan, and reduction scratch.
 *  - particle state is SoA fp64: x[n], y[n], theta[n]  (the reference keeps an (N,3) fp64 AoS
 *    array; SoA gives coalesced 8-byte loads).
 *  - every call returns 0 on success or a negative mcl_status; mcl_last_error() gives the text.
 *    There is no CPU fallback: without a CUDA device mcl_create fails.
 *  - calls are asynchronous on the handle's stream unless they return host values
 *    (mcl_estimate, mcl_softmax_stats, mcl_sync).  One handle = one stream, not thread-safe
 *    (the node's callbacks run on two threads without a lock, SURVEY 3.3: the Python Localizer
 *    serialises them).
 */
#ifndef MCL_H_
#define MCL_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct mcl_handle mcl_handle;

typedef enum {
    MCL_OK = 0,
    MCL_ERR_ARG = -1,      /* bad argument (null pointer, n < 0, ...) */
    MCL_ERR_STATE = -2,    /* map / sensor / scan not set yet */
    MCL_ERR_CUDA = -3,     /* CUDA runtime error (text in mcl_last_error) */
    MCL_ERR_CAPACITY = -4, /* problem too large for a library limit */
    MCL_ERR_NOMEM = -5
} mcl_status;

/* Resampling arithmetic (mcl_resample_indices). */
enum {
    MCL_RESAMPLE_REFERENCE_F32 = 0, /* pu:416-446 bit-exact: sequential f32 normalising sum and f32
                                       running cumulative sum, f64 U = r + m/N */
    MCL_RESAMPLE_FIXED_POINT = 1,   /* same walk on 64-bit fixed-point weights (exact, associative => identical
                                       for any block / rank split): the sharded runs' arithmetic */
    MCL_RESAMPLE_AMCL_F32 = 2       /* pu:486-502 low_variance_resample_amcl: weights as given (no normalisation),
                                       sequential f32 running sum, U = r + m / n_out, walk bounded by n_in - 1 */
};

/* Philox stream ids (ctr[3] low byte). */
enum { MCL_STREAM_MOTION = 1, MCL_STREAM_MH = 2, MCL_STREAM_RESAMPLE = 3, MCL_STREAM_INIT = 4,
       MCL_STREAM_KLD = 5, MCL_STREAM_MOTION_RADIUS = 6, MCL_STREAM_MOTION_RADIUS2 = 7 };

/* ---- lifetime ------------------------------------------------------------------------- */
int mcl_create(mcl_handle **out, int device);
int mcl_destroy(mcl_handle *h);
const char *mcl_last_error(const mcl_handle *h);     /* h may be NULL: last create() error */
const char *mcl_version(void);
int mcl_set_stream(mcl_handle *h, void *cuda_stream); /* NULL = legacy default stream */
int mcl_sync(mcl_handle *h);
int mcl_device_info(mcl_handle *h, int *sm_count, int *smem_per_block_optin, int *cc_major,
                    int *cc_minor);

/* ---- configuration ---------------------------------------------------------------------- */
/* node:124-177 load_map: the OccupancyGrid payload (int8, 0 free / 100 occupied / -1 unknown,
 * row = y, index my*W+mx) and the EDT distance map (f32 metres) computed by the caller
 * (scipy.ndimage.distance_transform_edt, node:156).  HOST pointers; copied.  Either pointer may be
 * NULL when only the other half is needed (likelihood needs dist, predict/init need occ). */
int mcl_set_map(mcl_handle *h, const int8_t *h_occ, const float *h_dist, int W, int H,
                double resolution, double origin_x, double origin_y);
/* Same, with the distance map computed on the device: an exact Euclidean distance transform of the free
 * cells in integers, bit-identical to scipy.ndimage.distance_transform_edt(map == 0) * resolution as f32
 * (node:153-157).  h_dist_out (nullable, W*H floats) receives it. */
int mcl_set_map_edt(mcl_handle *h, const int8_t *h_occ, int W, int H, double resolution, double origin_x,
                    double origin_y, float *h_dist_out);
/* node:52-56 sensor model parameters (amhmcl.yaml: sigma_hit, z_hit, z_rand, max_range, step). */
int mcl_set_sensor(mcl_handle *h, double sigma_hit, double z_hit, double z_rand, double max_range,
                   int step);
/* node:28-33 odometry noise, float32[4] exactly as the node stores it. */
int mcl_set_motion(mcl_handle *h, const float alpha[4]);
/* node:341-348 update_scans: ranges f32[M] and per-beam angles f32[M] (np.linspace(angle_min,
 * angle_max, M, dtype=float32)).  HOST pointers; the valid-beam table is built and uploaded. */
int mcl_set_scan(mcl_handle *h, const float *h_ranges, const float *h_angles, int M);
/* Pre-stage K scans (h_ranges is K x M, one shared angle vector) on the device and switch between
 * them without any copy: replaying a bag, or a benchmark whose inputs are resident in HBM. */
int mcl_set_scan_batch(mcl_handle *h, const float *h_ranges, const float *h_angles, int M, int K);
int mcl_use_scan(mcl_handle *h, int k);
/* number of beams counted in valid_count (pu:123-124) for the current scan */
int mcl_scan_valid_count(mcl_handle *h, int *valid_count);
/* Likelihood-table staging: 0 = auto (shared-memory window when it fits, int32 or byte-coded; else, for
 * large particle counts, particles binned by map tile with one tile's neighbourhood staged at a time;
 * else global/L2 gather), 1 = force global/L2 gather, 2 = force shared-memory window (error if it does
 * not fit). For measurements. */
int mcl_set_likelihood_path(mcl_handle *h, int path);

/* ---- the hot path ------------------------------------------------------------------------ */
/* pu:85-149 compute_likelihoods -> d_score[n] (f32). */
int mcl_likelihood(mcl_handle *h, const double *d_x, const double *d_y, const double *d_theta,
                   int64_t n, float *d_score);

/* pu:151-201 compute_likelihoods_raycast (+ pu:4-29 raycast, pu:36-58 p_hit / p_rand): the reference's
 * alternative beam model by ray marching (hard-coded sigma 0.05, z_hit 0.8, z_rand 0.1, max_range 10).  Not
 * reached from the node's callbacks; provided for completeness of the parallel_utils surface.
 * h_blocked: (H, W) bytes, non-zero where the reference's grid_map > 0.5; x_min / y_min = limits[0] / limits[2]. */
int mcl_set_raycast_grid(mcl_handle *h, const uint8_t *h_blocked, int W, int H, double resolution, double x_min,
                         double y_min);
int mcl_likelihood_raycast(mcl_handle *h, const float *h_ranges, const float *h_angles, int M, const double *d_x,
                           const double *d_y, const double *d_theta, int64_t n, float *d_score);

/* node:351-358 convert_scores (softmax).  d_stats (nullable, room for 4 doubles) receives
 * {max, sum exp(s-max), the same sum as a 2^-40 fixed-point uint64 bit pattern, unused} on the device; d_weights (nullable) receives exp(s-max)/sum as f32.
 * ext_stats (nullable, HOST): if given, use these {max, sum} instead of the local ones -- the
 * multi-GPU path passes the all-reduced values. */
int mcl_softmax(mcl_handle *h, const float *d_score, int64_t n, float *d_weights, double *d_stats,
                const double *ext_stats);
/* Staged softmax for sharded particles: max -> [all-reduce MAX on d_stats[0]] -> sumexp (uses
 * d_stats[0], writes d_stats[1] and the exact 2^-40 fixed-point integer d_stats[2]) -> [all-reduce SUM
 * on the integer, d_stats[1] = integer * 2^-40] -> weights.  d_stats: device, 4 doubles. */
int mcl_softmax_max(mcl_handle *h, const float *d_score, int64_t n, double *d_stats);
int mcl_softmax_sumexp(mcl_handle *h, const float *d_score, int64_t n, double *d_stats);
int mcl_softmax_weights(mcl_handle *h, const float *d_score, int64_t n, const double *d_stats,
                        float *d_weights);
/* node:276-278 update_acml_weights: d_w /= sum(d_w) (f32 divide by the f32-rounded exact sum);
 * h_out (nullable) = {sum before, mean of the normalised weights (node:284)}.  Blocking if h_out. */
int mcl_weights_normalize(mcl_handle *h, float *d_weights, int64_t n, double h_out[2]);
/* blocking read of the two doubles written by mcl_softmax */
int mcl_softmax_stats(mcl_handle *h, const float *d_score, int64_t n, double h_stats[2]);

/* pu:332-363 apply_motion_model_parallel (+ pu:388-396 is_valid_position).
 * delta = (rot1, trans, rot2) from node:410-421 compute_motion (see mcl_compute_motion).
 * d_normals == NULL: Philox4x32-10 draws keyed (seed, step, first_index + i, attempt).
 * d_normals != NULL: injected standard normals, layout (n, A, 3) f64, attempt t uses row t % A.
 * d_attempts (nullable): 1-based index of the accepted attempt, 0 = kept old pose (pu:360-361).
 * Output may alias input. */
int mcl_predict(mcl_handle *h, const double *d_x, const double *d_y, const double *d_theta,
                int64_t n, const double delta[3], uint64_t seed, uint64_t step,
                uint64_t first_index, const double *d_normals, int A, int max_attempts,
                double *d_xo, double *d_yo, double *d_thetao, int32_t *d_attempts);

/* node:410-421 compute_motion (host scalar arithmetic, glibc atan2/hypot: within 1 ulp of the
 * node's NumPy calls; the Python host code uses NumPy itself to stay bit-exact). */
int mcl_compute_motion(const double odom1[3], const double odom2[3], double delta[3]);

/* pu:208-236 mh_resampling: current = (d_x..), proposal = (d_px..), likelihoods = weights of the
 * proposal, old_weights = weights of the current.  d_uniforms == NULL: Philox u53 keyed
 * (seed, step, first_index + i).  Outputs: new poses, new weights, accept flags (nullable). */
int mcl_mh_accept(mcl_handle *h, const double *d_x, const double *d_y, const double *d_theta,
                  const double *d_px, const double *d_py, const double *d_ptheta,
                  const float *d_likelihoods, const float *d_old_weights, int64_t n,
                  const double *d_uniforms, uint64_t seed, uint64_t step, uint64_t first_index,
                  double *d_xo, double *d_yo, double *d_thetao, float *d_weights_out,
                  uint8_t *d_accept);

/* pu:282-330 motion_model_odometry_parallel: transition density p(cur | prev, delta) per particle
 * (product of three Gaussians, pu:31-33), d_probs f64[n]; d_sum (nullable, device) receives the population
 * sum; normalise != 0 divides by it when positive (pu:326-328).  Sharded callers pass normalise = 0,
 * all-reduce the sum and call mcl_scale_by_sum. */
int mcl_motion_density(mcl_handle *h, const double *d_px, const double *d_py, const double *d_ptheta,
                       const double *d_cx, const double *d_cy, const double *d_ctheta, int64_t n,
                       const double delta[3], double *d_probs, double *d_sum, int normalise);
int mcl_scale_by_sum(mcl_handle *h, double *d_probs, int64_t n, const double *d_sum);
/* pu:238-276 assym_mh_resampling (reference quirk kept: alpha = 1 unless log_den > 0). */
int mcl_assym_mh_accept(mcl_handle *h, const double *d_x, const double *d_y, const double *d_theta,
                        const double *d_px, const double *d_py, const double *d_ptheta,
                        const float *d_likelihoods, const float *d_old_weights,
                        const double *d_trans_forward, const double *d_trans_backward, int64_t n,
                        const double *d_uniforms, uint64_t seed, uint64_t step, uint64_t first_index,
                        double *d_xo, double *d_yo, double *d_thetao, float *d_weights_out,
                        uint8_t *d_accept);

/* mcl_assym_mh_accept with corrected != 0: the Metropolis-Hastings ratio applied unconditionally (see
 * mcl_filter_set_assym). */
int mcl_assym_mh_accept_ex(mcl_handle *h, const double *d_x, const double *d_y, const double *d_theta,
                           const double *d_px, const double *d_py, const double *d_ptheta,
                           const float *d_likelihoods, const float *d_old_weights,
                           const double *d_trans_forward, const double *d_trans_backward, int64_t n,
                           const double *d_uniforms, uint64_t seed, uint64_t step, uint64_t first_index,
                           double *d_xo, double *d_yo, double *d_thetao, float *d_weights_out,
                           uint8_t *d_accept, int corrected);

/* pu:416-446 low_variance_resample_numba -> source index per output (d_idx[n_out], int32).
 * r is the single uniform draw in [0, 1/n_out) (see mcl_resample_offset). */
int mcl_resample_indices(mcl_handle *h, const float *d_weights, int64_t n_in, int64_t n_out,
                         double r, int mode, int32_t *d_idx);
/* Global systematic resampling over sharded particles (FIXED_POINT arithmetic):
 *   mcl_weights_max   -> local max (device f32)            [caller: all-reduce MAX]
 *   mcl_resample_scan -> local fixed-point cumulative sums with the global scale, local total
 *                        (device u64)                       [caller: all-gather totals -> offsets]
 *   mcl_resample_search -> local source index of every global output m in [m0, m0 + n_out_local)
 *                        that falls into this rank's interval (offset, offset + total].
 * The cumulative sums stay in the handle's scratch between scan and search. */
int mcl_weights_max(mcl_handle *h, const float *d_weights, int64_t n, float *d_wmax);
int mcl_resample_scan(mcl_handle *h, const float *d_weights, int64_t n_in, const float *d_wmax_global,
                      int64_t n_global, uint64_t *d_total);
int mcl_resample_search(mcl_handle *h, int64_t n_in, uint64_t offset, uint64_t grand_total, int64_t m0,
                        int64_t n_out_local, double r, int64_t n_out_global, int32_t *d_idx);
/* pu:529-591 kld_sampling_amcl: KLD-adaptive systematic resampling with Gaussian jitter.
 * weights are used as given (node:276-278 normalises them first).  r in [0, 1/max_samples).
 * d_normals != NULL: injected standard normals (max_samples, 3); NULL: Philox(seed, step, sample index).
 * mode: MCL_RESAMPLE_REFERENCE_F32 (sequential f32 running sum, bit-exact) or MCL_RESAMPLE_FIXED_POINT.
 * Outputs hold max_samples poses (f32-rounded like the reference's buffer); *h_count = how many are valid.
 * Blocking (the caller needs the count). */
int mcl_kld_resample(mcl_handle *h, const double *d_x, const double *d_y, const double *d_theta,
                     const float *d_weights, int64_t n_in, int64_t max_samples, int64_t min_particles,
                     double bin_size_xy, double bin_size_theta, double epsilon, double z, double r,
                     const double *d_normals, uint64_t seed, uint64_t step, int mode, double *d_xo,
                     double *d_yo, double *d_thetao, int64_t *h_count);
/* Peer-push variant of the search + exchange: after mcl_resample_scan and the all-gather of the totals
 * (d_totals_all, world u64 on the device) every rank stores its offspring directly into the destination
 * ranks' pose buffers through peer (NVLink) pointers: d_peer_ptrs = device array [3][world] of the x, y, theta
 * destination base pointers of every rank.  No host round trip; follow with a cross-rank barrier. */
int mcl_resample_push(mcl_handle *h, int64_t n_in, const uint64_t *d_totals_all, int rank, int world, double r,
                      int64_t n_global, int64_t n_per_rank, const double *d_x, const double *d_y,
                      const double *d_theta, const uint64_t *d_peer_ptrs);
/* r = 0 + (1/n_out - 0) * u53(Philox(seed, step, 0, RESAMPLE))  (np.random.uniform(0, 1/N)) */
double mcl_resample_offset(uint64_t seed, uint64_t step, int64_t n_out);
/* new_particles[m] = particles[idx[m]] (pu:445), SoA gather; outputs must not alias inputs. */
int mcl_gather(mcl_handle *h, const double *d_x, const double *d_y, const double *d_theta,
               const int32_t *d_idx, int64_t n_out, double *d_xo, double *d_yo, double *d_thetao);

/* node:586-597 publish_estimate arithmetic.  h_out[16]:
 *  [0] V1 = sum w  [1] V2 = sum w^2  [2] mean_x  [3] mean_y  [4] mean_theta
 *  [5..7]  sum w*d  (d = (x-mean_x, y-mean_y, f32(wrap(theta-mean_theta))))
 *  [8..13] sum w*d_i*d_j  in order xx, xy, xt, yy, yt, tt        [14],[15] reserved
 * mcl_estimate_moments returns the raw first-pass sums (for the multi-GPU all-reduce):
 *  h_m[6] = {sum w, sum w^2, sum w x, sum w y, sum w cos, sum w sin}; mcl_estimate_central then
 *  takes the global means and returns h_c[9] = {sum w d (3), sum w d d (6)}. Blocking calls. */
int mcl_estimate(mcl_handle *h, const double *d_x, const double *d_y, const double *d_theta,
                 const float *d_weights, int64_t n, double h_out[16]);
/* non-blocking: d_out18 (device) = {6 raw sums, mean_x, mean_y, mean_theta, 9 central sums} */
int mcl_estimate_async(mcl_handle *h, const double *d_x, const double *d_y, const double *d_theta,
                       const float *d_weights, int64_t n, double *d_out18);
/* staged, non-blocking (sharded particles): d_m9[0..5] raw sums -> [all-reduce SUM] ->
 * mcl_estimate_means_async fills d_m9[6..8] -> central sums d_c9 -> [all-reduce SUM]. */
int mcl_estimate_moments_async(mcl_handle *h, const double *d_x, const double *d_y,
                               const double *d_theta, const float *d_weights, int64_t n, double *d_m9);
int mcl_estimate_means_async(mcl_handle *h, double *d_m9);
int mcl_estimate_central_async(mcl_handle *h, const double *d_x, const double *d_y,
                               const double *d_theta, const float *d_weights, int64_t n,
                               const double *d_mean3, double *d_c9);
int mcl_estimate_moments(mcl_handle *h, const double *d_x, const double *d_y, const double *d_theta,
                         const float *d_weights, int64_t n, double h_m[6]);
int mcl_estimate_central(mcl_handle *h, const double *d_x, const double *d_y, const double *d_theta,
                         const float *d_weights, int64_t n, const double mean[3], double h_c[9]);

/* pu:69-83 normalize_angle_array: out[i] = (float) normalize_angle(angles[i] - mean_angle). */
int mcl_normalize_angle_array(mcl_handle *h, const double *d_angles, double mean_angle, int64_t n,
                              float *d_out);

/* pu:450-465 generate_valid_particles (+ pu:398-413 compute_valid_mask).
 * d_u != NULL: injected uniforms, layout (3, max_trials) f64 = ux | uy | utheta; the first n valid
 *   trials are kept, in trial order (bit-exact restatement); *h_count returns how many (<= n).
 * d_u == NULL: per-particle rejection sampling with Philox (seed, first_index + i, attempt);
 *   the 50M-particle global-localisation config cannot afford 50 N trial vectors. */
int mcl_init_uniform(mcl_handle *h, int64_t n, const double *d_u, int64_t max_trials, uint64_t seed,
                     uint64_t first_index, double *d_x, double *d_y, double *d_theta,
                     int64_t *h_count);

/* AoS (n,3) f64 <-> SoA conversion on the device (the reference's layout at the shim boundary). */
int mcl_aos_to_soa(mcl_handle *h, const double *d_aos, int64_t n, double *d_x, double *d_y,
                   double *d_theta);
int mcl_soa_to_aos(mcl_handle *h, const double *d_x, const double *d_y, const double *d_theta,
                   int64_t n, double *d_aos);

/* ---- the whole step in one call ------------------------------------------------------------- */
/* Bind the caller's device buffers once; afterwards each call below enqueues a whole stage of the
 * node's callbacks on the handle's stream, and the roles of the three pose buffer sets
 * (particles / particles_prev / spare) rotate inside the library like the node's copies at
 * node:404-405, node:370 and node:490.  x/y/th: three SoA sets of n fp64; w_a/w_b: the weights
 * double buffer.  Stochastic stages draw Philox(seed, tick) with tick += 1 per predict, MH accept and
 * resample -- the same streams as the individual entry points, so results are identical. */
int mcl_filter_bind(mcl_handle *h, int64_t n, double *const x[3], double *const y[3],
                    double *const th[3], float *score_pre, float *score_post, float *w_pre,
                    float *w_post, float *w_a, float *w_b, int32_t *idx, int use_mh, int resample_mode,
                    uint64_t seed, uint64_t first_index, int max_attempts);
int mcl_filter_configure(mcl_handle *h, int use_mh, int resample_mode, uint64_t seed,
                         uint64_t first_index, int64_t tick /* < 0: keep */);
/* asymmetric MH (localization_mode containing "AMH", node:21): update() then runs node:424-439
 * transition_probability + pu:238-276.  set_transition overrides the increments stored by predict
 * (delta_b NULL = derive it with node:429-434's formula).
 * assym = 1: the reference, quirks included (pu:269 only applies the ratio when
 * log_den > 0, i.e. never, so every proposal is accepted; node:429-434 treats (rot1, trans, rot2) as (dx, dy, dtheta)).
 * assym = 2: corrected variant -- alpha = min(1, exp(log_alpha)) unconditionally and the backward increment is the
 * odometry increment that undoes the forward one, (pi - rot2, trans, -rot1 - pi). */
int mcl_filter_set_assym(mcl_handle *h, int assym);
int mcl_filter_set_transition(mcl_handle *h, const double delta[3], const double delta_b[3]);
/* roles = {particles, particles_prev, spare (indices into x/y/th), weights slot (0 = w_a)} */
int mcl_filter_set_n(mcl_handle *h, int64_t n);   /* KLD-adaptive modes: n changes, buffers keep capacity */
int mcl_filter_roles(mcl_handle *h, int roles[4], uint64_t *tick);
/* for hosts that sequence the stages themselves (the sharded path interleaves collectives) */
int mcl_filter_set_roles(mcl_handle *h, const int roles[4], uint64_t tick);
int mcl_filter_predict(mcl_handle *h, const double delta[3], const double *d_normals, int A);   /* node:384-408 */
int mcl_filter_update(mcl_handle *h, const double *d_uniforms);                                /* node:296-322 */
/* update() with `iters` Metropolis-Hastings iterations per scan, one chain per particle (BASELINE
 * config 4): chain_0 = particles_prev, every iteration proposes a fresh motion sample of particles_prev
 * (pu:332-363), softmaxes proposal and chain scores separately (node:254-270) and accepts with pu:208-236;
 * iters = 1 is exactly mcl_filter_update in MHMCL mode.  Uses the increment stored by the last predict. */
int mcl_filter_update_chain(mcl_handle *h, int iters);
int mcl_filter_estimate(mcl_handle *h, double *d_out18, double h_out16[16]);                   /* node:586-597 */
int mcl_filter_resample(mcl_handle *h, double r /* < 0: Philox draw */);                       /* node:488-492 */
/* estimate (node:586-597; to d_out18 and/or blocking into h_out16, both nullable) followed by resample_lvr
 * (node:488-492, Philox offset) of the particles and weights as they are -- what follows mcl_filter_update or
 * mcl_filter_update_chain in lidar_callback.  One launch of the persistent tail kernel on one GPU; same results as
 * mcl_filter_estimate + mcl_filter_resample(-1). */
int mcl_filter_finish(mcl_handle *h, double *d_out18, double h_out16[16]);
/* odom + scan -> predict (delta != NULL), update on pre-staged scan `scan_slot` (or the current scan
 * if < 0), estimate (to d_out18 and/or blocking into h_out16; both nullable), resample.
 * With symmetric MH or plain MCL the step is four launches: motion, motion retries, likelihood of both particle sets, and the
 * persistent tail kernel (csrc/tail.cu: softmax x2, MH accept, estimate sums, resampling in either arithmetic;
 * on sharded handles also the cross-rank exchanges and the peer push) -- same results bit for bit as the calls
 * above issued one by one.  MCL_NO_TAIL=1 in the environment selects round 1's four fused kernels (fixed point)
 * or the stand-alone sequence, MCL_NO_FUSE=1 the stand-alone sequence (A/B measurements). */
int mcl_filter_step(mcl_handle *h, const double delta[3], int scan_slot, double *d_out18,
                    double h_out16[16]);

/* ---- functions the node imports (node:13) but never reaches from its callbacks ---------
 * pu:369-386 compute_valid_indices: indices (ascending) of the particles whose cell (int() truncation) lies in the
 * map with map_data <= 10.  Needs the occupancy grid (mcl_set_map).  Blocking; *h_count = how many. */
int mcl_compute_valid_indices(mcl_handle *h, const double *d_x, const double *d_y, int64_t n, int32_t *d_idx,
                              int64_t *h_count);
/* pu:600-614 validate_samples (second half of initialize_gaussian_parallel, node:183): in place, a sample whose
 * cell is outside the map or has distance_map >= 1.0 becomes (0, 0, 0).  Needs the distance map. */
int mcl_validate_samples(mcl_handle *h, double *d_x, double *d_y, double *d_theta, int64_t n);
/* pu:467-477 parallel_resample_simple: cum = np.cumsum(weights) as a sequential f32 sum, one uniform per output
 * (d_u injected, or Philox(seed, step, output index)), idx = searchsorted(cum, u).  Where the reference reads out
 * of bounds (u > cum[-1]) the last particle is taken. */
int mcl_resample_multinomial(mcl_handle *h, const float *d_weights, int64_t n_in, int64_t n_out, const double *d_u,
                             uint64_t seed, uint64_t step, int32_t *d_idx);
/* pu:504-526 reinitialize_particles_numba: per new particle a uniformly chosen free cell (pose = its lower-left
 * corner) and a uniform heading.  d_choice (index into the row-major list of free cells) / d_theta: injected draws
 * or NULL for Philox(seed, step, particle).  Needs the occupancy grid.  Blocking; *h_n_free = free cells. */
int mcl_reinitialize_particles(mcl_handle *h, int64_t n, const int64_t *d_choice, const double *d_theta, uint64_t seed,
                               uint64_t step, double *d_x, double *d_y, double *d_theta_out, int64_t *h_n_free);

/* Sharded operation, one process per GPU (call after mcl_filter_bind on every rank): the per-step scalar
 * exchanges (softmax max / sum, estimate sums, resampling scale and totals) and the resampling exchange then run
 * over NVLink peer memory inside the library's own kernels -- every mcl_filter_* call above becomes collective.
 *   d_mailbox        this rank's mailbox, >= 8 KiB of zero-initialised symmetric (peer-mapped) memory
 *   h_peer_mailbox   host array[world]: the address of every rank's mailbox as mapped in THIS process
 *   h_peer_pose      host array[3][3][world]: peer-mapped base address of pose set s, component c (x, y, theta)
 * Particle indices for the random streams become global (first_index = rank * n).  mcl_comm_status reports
 * whether an exchange timed out (a peer died). */
int mcl_comm_init(mcl_handle *h, int rank, int world, void *d_mailbox, const uint64_t *h_peer_mailbox,
                  const uint64_t *h_peer_pose);
int mcl_comm_status(mcl_handle *h, int *err);

/* mcl_filter_step runs the whole tail of the step (node:351-358 softmax x2, pu:208-236 MH accept,
 * node:586-597 estimate sums, pu:416-446 resampling in either arithmetic) as ONE persistent cooperative kernel
 * with grid-wide barriers (csrc/tail.cu).  Its waits are bounded; *err != 0 after a time-out (blocking call;
 * the step's results are then invalid).  mcl_filter_step with a host estimate checks it itself.
 * MCL_NO_TAIL=1 in the environment selects the multi-kernel sequence instead (A/B measurements). */
int mcl_tail_status(mcl_handle *h, int *err);
/* Test hook: systematic resampling (pu:416-446) of the given device weights through the resampling stages of
 * that kernel alone; d_c (nullable) receives the running sums (n f32 in reference mode, n u64 in fixed point). */
/* Debug (MCL_TAIL_PROF=1 in the environment): globaltimer stamps [grid][32] of the stage boundaries of the last
 * tail launch (out must hold 1024 * 32 values); *grid = CTAs of that launch. */
int mcl_tail_prof(mcl_handle *h, unsigned long long *out, int *grid);
int mcl_debug_tail_resample(mcl_handle *h, float *d_w, int64_t n, double r, int mode, int32_t *d_idx, void *d_c);
/* Debug (MCL_MOTION_STATS=1 in the environment): counters of the motion kernel's rejection loop since the last
 * call: [0] attempt 0 failed, [1] provably stuck, [2] particles retried, [3] screening rounds, [4] evaluation rounds,
 * [5] retries that found a pose, [6] retried with threshold <= 2^28, [7] attempts evaluated, [8] warps with a retry. */
int mcl_debug_motion_stats(mcl_handle *h, unsigned long long out[16]);
/* Test hook for the motion kernel's rejection loop: screening thresholds below min_thr are raised to it (a looser
 * screen is still exact; 2^28 makes every zero top nibble a candidate, 2^32 disables the screen) and small_queue = 1
 * selects a 64-entry candidate queue, so that the paths a production run reaches with probability ~0 are tested. */
int mcl_debug_motion(mcl_handle *h, unsigned long long min_thr, int small_queue);

/* ---- measurement helpers (bench.py roofline denominators; not on the product path) ------- */
/* Random 4-byte gather rate, lookups/s: table_bytes resident in shared memory (where = 0) or in
 * global memory / L2 (where = 1); n_lookups per launch, iters launches timed with CUDA events. */
int mcl_bench_gather(mcl_handle *h, int where, int64_t table_bytes, int64_t n_lookups, int iters,
                     double *lookups_per_s);
/* test hook: the sequential-f32 running sums c_i = fl32(c_{i-1} + w_i) that MCL_RESAMPLE_REFERENCE_F32
 * searches (normalise: of w_i / seq_sum(w), pu:430; else of w_i as given, pu:555-563), produced by the
 * exact parallel scan (serial = 0) or by the one-warp in-order replay (serial = 1). */
int mcl_debug_seq_cumsum(mcl_handle *h, const float *d_weights, int64_t n, int normalise, int serial, float *d_c);
/* launches of library kernels since create (the bench's gpu_launches claim) */
int64_t mcl_launch_count(const mcl_handle *h);
/* CUDA-event timing of the library's own launches: accumulate the device time of every
 * mcl_likelihood launch between start and stop (events on the handle's stream). */
int mcl_timing_start(mcl_handle *h);
int mcl_timing_stop(mcl_handle *h, double *likelihood_ms, int64_t *likelihood_launches);
/* particle sets evaluated by those launches (the fused step scores particles and particles_prev in one launch) */
int64_t mcl_timing_sets(const mcl_handle *h);

#ifdef __cplusplus
}
#endif
#endif /* MCL_H_ */
