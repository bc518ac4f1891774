// common.cuh -- handle, error plumbing, warp/block reductions, Philox4x32-10.
// Internal to libmcl.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <string>
#include <vector>

#include "../../include/mcl.h"

#define MCL_PI 3.141592653589793      // == np.pi
#define MCL_TWO_PI 6.283185307179586  // == 2*np.pi

struct BeamTable {       // per valid beam, endpoint offset in CELL units (fp64):
    double bx, by;       //   bx = r cos(a) / res,  by = r sin(a) / res
};

struct mcl_handle {
    int device = 0;
    cudaStream_t stream = nullptr;
    std::string err;
    int sm_count = 0, smem_optin = 0, cc_major = 0, cc_minor = 0;
    int64_t launches = 0;

    // map (node:124-177)
    int W = 0, H = 0;
    double res = 0, ox = 0, oy = 0;
    int8_t *d_occ = nullptr;
    float *d_dist = nullptr;
    // sensor (node:52-56)
    bool sensor_set = false;
    double sigma_hit = 0, z_hit = 0, z_rand = 0, max_range = 0;
    int step = 1;
    // motion
    bool motion_set = false;
    float alpha[4] = {0, 0, 0, 0};

    // likelihood table: logtab[c] = rint(2^25 * log(max(z_hit*p_hit(dist[c]) + z_rand/max_range, 1e-6)))
    // 2^-25 fixed point (|log p| <= 13.82 fits int32 four times over): sums of table values are exact
    // integers, so a score does not depend on the order in which beams are accumulated (lanes per
    // particle, shared vs global path, number of ranks) and carries ~1e-9 error instead of fp32's ~1e-6.
    bool tab_dirty = true;
    int32_t *d_logtab = nullptr; // W*H
    int32_t c0 = 0;              // table value on dist == 0 cells (everything outside the window)
    // The shared-memory copies of the table (d_win, d_lut) hold v - voff >= 0 (voff <= every table value), so
    // the kernels can add up to MCL_ACC_TERMS of them in one unsigned 32-bit register before widening.
    int32_t voff = 0;
    bool acc_terms_ok = false;   // MCL_ACC_TERMS * (c0 - voff) < 2^32 (c0 is the table maximum: dist == 0)
    // free-space window [wx0, wx0+ww) x [wy0, wy0+wh): outside it every in-map cell holds c0.
    // d_win is the window with a one-cell c0 border, (wh+2) x (ww+2) cells, stored with a pitch of 256
    // cells along its "minor" axis (x, or y when win_tpose) so that a cell index is one byte permute of
    // the two clamped coordinates: index = minor | major << 8 (likelihood.cu).  win_ok = false when
    // neither axis fits in 256 columns (the table is then gathered from global memory / L2).
    int wx0 = 0, wy0 = 0, ww = 0, wh = 0;
    // Geometry of the staged window (mcl_prepare_table): window index i <-> map cell i + win_of per axis, largest
    // index win_c.  A side of the free-space box that lies within beam reach of the map edge is an EDGE side: the
    // window then runs to the map edge and its border holds "outside the map" (adds 0) instead of c0, so particles
    // near that edge need no per-beam in-map test; on the low sides one extra column / row repeats cell 0, which is
    // where int() sends coordinates in (-1, 0).  Bits of win_edge: 1 low x, 2 high x, 4 low y, 8 high y.
    int win_ofx = 0, win_ofy = 0, win_cx = 0, win_cy = 0, win_edge = 0;
    bool win_ok = false, win_tpose = false;
    int win_rows = 0;            // (major extent + 2) rows of 256 cells
    int32_t *d_win = nullptr;
    size_t win_bytes = 0;
    int32_t *d_win_skew = nullptr;   // the same window with rows of 264 words (bank-skewed)
    size_t win_skew_bytes = 0;
    int lik_path = 0;            // 0 auto, 1 global, 2 smem window
    // coded window for maps whose int32 window exceeds shared memory: one byte per cell indexing a
    // table of the (<= 256) distinct values (the value depends only on dist, and an EDT on a grid takes
    // few distinct values near walls: 85 on map_world, 242 on map_house); same layout as d_win
    bool coded = false;
    uint8_t *d_win8 = nullptr;
    int32_t *d_lut = nullptr;
    size_t win8_bytes = 0;
    // Large maps (no window fits in shared memory): the whole table as one byte per cell (codes into d_lut, code
    // 255 = "outside the map", contributing 0) so that the tiled likelihood kernel can stage the neighbourhood of
    // one map tile at a time; particles are binned by tile first (likelihood.cu, k_likelihood_tiled).
    bool tiled_ok = false;
    uint8_t *d_code8 = nullptr;      // W * H
    uint8_t *d_code8p = nullptr;     // (H + 1) x (W + 16): the same with an apron (mcl_prepare_table), for k_likelihood_tiled2
    int tile_w = 0, tile_h = 0, tile_margin = 0, tiles_x = 0, tiles_y = 0;
    void *d_tiled = nullptr;         // per-call binning buffers
    size_t tiled_bytes = 0;
    // cell-index arithmetic (likelihood.cu): endpoints are evaluated relative to the window origin in the
    // "magic" form cell_M + t, cell_M = 1.5 * 2^(20 - cell_S): the high word of the double then holds
    // floor(t * 2^cell_S) + cell_K.  |t| must stay below cell_lim (particles further out see no map cell).
    int cell_S = 8, cell_K = 0;
    double cell_M = 0, cell_lim = 0;

    // scan (node:341-348): valid beams with r >= 0 first, then valid beams with r < 0
    bool scan_set = false;
    BeamTable *d_beams = nullptr;
    float *d_neg_r = nullptr;    // unused placeholder (negative ranges handled via beam table)
    int beams_cap = 0;
    int n_pos = 0, n_neg = 0;    // valid_count = n_pos + n_neg
    double rmax_cells = 0;       // max |r|/res over valid beams
    BeamTable *h_beams = nullptr; // pinned staging
    // pre-staged scans (mcl_set_scan_batch / mcl_use_scan): K tables of stride batch_stride
    struct ScanMeta { int n_pos, n_neg; double rmax_cells; };
    BeamTable *d_batch = nullptr;
    std::vector<ScanMeta> batch_meta;
    int batch_stride = 0;
    const BeamTable *d_beams_active = nullptr;
    uint64_t scan_gen = 0;       // process-wide unique id of the active scan table's contents (mcl_next_scan_uid)

    // scratch for reductions / scans
    void *d_scratch = nullptr;
    size_t scratch_bytes = 0;
    double *h_pinned = nullptr;  // 64 doubles, pinned, for blocking scalar reads
    void *d_seq = nullptr;       // work buffers of the exact sequential-f32 scan (resample.cu)
    size_t seq_bytes = 0;
    void *d_kld = nullptr;       // KLD-sampling work buffers (kld.cu)
    size_t kld_bytes = 0;
    void *d_fused = nullptr;     // work area of the fused step tail (fused.cu)
    size_t fused_bytes = 0;
    void *d_tail = nullptr;      // work area of the persistent step tail (tail.cu)
    size_t tail_bytes = 0;
    unsigned long long tail_bar = 0;   // grid-barrier arrivals consumed so far (base of the next launch)
    void *d_tail_prof = nullptr; // MCL_TAIL_PROF=1: stage time stamps of the tail kernel
    int tail_prof_grid = 0;
    int64_t tail_drift_n = -1;   // particle count the tail kernel's drift hints belong to
    void *d_motion_stats = nullptr; // MCL_MOTION_STATS=1: counters of the motion kernel's rejection loop
    unsigned long long motion_min_thr = 0;   // mcl_debug_motion (tests)
    bool motion_small_queue = false;
    int32_t *d_retry_idx = nullptr;           // motion kernel's retry list (particle, threshold), counters {count, done}
    unsigned long long *d_retry_thr = nullptr;
    unsigned *d_retry_ctr = nullptr;
    int64_t retry_cap = 0;
    int coop_launch = -1;        // cudaDevAttrCooperativeLaunch (-1: not queried yet)
    double *d_est18 = nullptr;   // device staging of the estimate sums (mcl_filter_step)
    cudaEvent_t ev_est = nullptr;
    cudaEvent_t ev_beams = nullptr;      // after the last copy out of the pinned beam staging buffer (mcl_set_scan)

    // timing of likelihood launches
    bool timing = false;
    std::vector<std::pair<cudaEvent_t, cudaEvent_t>> lik_events;
    int64_t lik_sets_timed = 0;  // particle sets evaluated by the timed launches (a pair launch evaluates two)
};

extern thread_local std::string g_create_err;

int mcl_fail(mcl_handle *h, int code, const std::string &msg);
uint64_t mcl_next_scan_uid();
void mcl_raycast_forget(const mcl_handle *h);
int mcl_ensure_scratch(mcl_handle *h, size_t bytes);
int mcl_prepare_table(mcl_handle *h);   // rebuild logtab/window if dirty
void mcl_filter_forget(const mcl_handle *h);
int mcl_cumsum_f32_seq(mcl_handle *h, const float *d_w, int64_t n, float *d_c);
int mcl_predict_cached(mcl_handle *h, const double *d_x, const double *d_y, const double *d_theta, int64_t n,
                       const double delta[3], uint64_t seed, uint64_t step, uint64_t first_index,
                       const double *d_normals, int A, int max_attempts, double *d_xo, double *d_yo, double *d_thetao,
                       int32_t *d_attempts, unsigned long long *d_thr_cache);
int mcl_softmax_pair(mcl_handle *h, const float *s0, float *w0, const float *s1, float *w1, int64_t n);
// fused.cu: the step tail in four kernels (softmax sums -> weights + MH + raw estimate sums -> central sums +
// cumulative weights -> search + gather); a sharded run exchanges the quantities of FusedPtrs between stages
struct FusedStep {
    int64_t n, n_global;
    int use_mh;
    const float *s_post, *s_pre;
    float *w_post, *w_pre, *w_out;
    const double *px, *py, *pt, *ox, *oy, *ot;
    double *nx, *ny, *nth;
    uint64_t seed, step, first_index;
    double *est18;
};
struct FusedPtrs { unsigned long long *keymax, *sumq, *total; double *msum, *csum; };
struct TailComm;
int mcl_tail_finish(mcl_handle *h, int64_t n, int64_t n_global, float *d_w, double *d_x, double *d_y, double *d_th,
                    double *d_est18, int resample_mode, double r, int32_t *idx, double *gx, double *gy, double *gt,
                    const TailComm *comm);
int mcl_fused_prepare(mcl_handle *h, int64_t n);
unsigned long long *mcl_fused_keymax(mcl_handle *h);
void mcl_fused_exchange_ptrs(mcl_handle *h, FusedPtrs *out);
const unsigned long long *mcl_fused_cumsum(mcl_handle *h, int64_t n);
int mcl_fused_sumexp(mcl_handle *h, const FusedStep &u);
int mcl_fused_weights(mcl_handle *h, const FusedStep &u);
int mcl_fused_scan(mcl_handle *h, const FusedStep &u);
int mcl_fused_chain_accept(mcl_handle *h, const FusedStep &u, float *score_chain);
int mcl_fused_resample(mcl_handle *h, int64_t n, double r, const double *nx, const double *ny, const double *nth,
                       int32_t *idx, double *gx, double *gy, double *gt);
const unsigned long long *mcl_fused_tile_prefix(mcl_handle *h, int64_t n, int *nt, int *tile);
int mcl_resample_push_from(mcl_handle *h, const unsigned long long *d_C, const unsigned long long *d_tile_prefix, int nt,
                           int tile, int64_t n_in, const uint64_t *d_totals_all, int rank, int world, double r,
                           int64_t n_global, int64_t n_per_rank, const double *d_x, const double *d_y,
                           const double *d_theta, const uint64_t *d_peer_ptrs);
// tail.cu: the same tail as ONE persistent cooperative kernel, in either resampling arithmetic
bool mcl_tail_available(mcl_handle *h, int64_t n);
// comm != NULL: sharded step -- the kernel exchanges its five global quantities over NVLink peer memory itself
// (TAIL_EXCHANGES epochs of filter.cu's mailbox) and pushes every offspring into the destination rank's set
struct TailComm {
    int rank, world;
    int64_t n_global;
    unsigned long long *mailbox;
    unsigned long long peers[16];
    unsigned long long epoch0;
    int *d_err;
    const unsigned long long *d_peer_pose_dst;   // device [3][world]
};
#define TAIL_EXCHANGES 5          /* fixed-point arithmetic */
#define TAIL_EXCHANGES_REF 7      /* the reference's arithmetic: + the two rank-to-rank hand-offs of the exact sums */
int mcl_tail_step(mcl_handle *h, const FusedStep &u, unsigned long long *d_keymax, int resample_mode, double r,
                  int32_t *idx, double *gx, double *gy, double *gt, const TailComm *comm);
const int *mcl_tail_err_ptr(mcl_handle *h);
// MH chain iteration as one cooperative launch (k_chain_tail); TAIL_CHAIN_EXCHANGES epochs when sharded
#define TAIL_CHAIN_EXCHANGES 2
unsigned long long *mcl_tail_chain_key(mcl_handle *h, int64_t n, int slot);
int mcl_tail_chain_iteration(mcl_handle *h, const FusedStep &u, unsigned long long *d_keymax, float *score_chain, int it,
                             const TailComm *comm);
int mcl_tail_chain_finish(mcl_handle *h);
int mcl_likelihood_pair(mcl_handle *h, const double *d_x, const double *d_y, const double *d_theta, float *d_score,
                        const double *d_x2, const double *d_y2, const double *d_theta2, float *d_score2, int64_t n,
                        unsigned long long *d_keymax, bool *g1_used);

#define MCL_CUDA(h, expr)                                                                   \
    do {                                                                                    \
        cudaError_t e__ = (expr);                                                           \
        if (e__ != cudaSuccess)                                                             \
            return mcl_fail((h), MCL_ERR_CUDA,                                              \
                            std::string(#expr) + ": " + cudaGetErrorString(e__));           \
    } while (0)

#define MCL_LAUNCH_CHECK(h)                                                                 \
    do {                                                                                    \
        (h)->launches++;                                                                    \
        cudaError_t e__ = cudaGetLastError();                                               \
        if (e__ != cudaSuccess)                                                             \
            return mcl_fail((h), MCL_ERR_CUDA, std::string("kernel launch: ") +             \
                                                   cudaGetErrorString(e__));                \
    } while (0)

struct DeviceGuard {
    int prev = -1;
    explicit DeviceGuard(int dev) { cudaGetDevice(&prev); if (prev != dev) cudaSetDevice(dev); else prev = -1; }
    ~DeviceGuard() { if (prev >= 0) cudaSetDevice(prev); }
};

// ------------------------------------------------------------------------------------------
// device helpers
// ------------------------------------------------------------------------------------------
#ifdef __CUDACC__

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// pu:62-67 normalize_angle with Python modulo semantics; fmod is exact.
// fmod(a, 2 pi), bit for bit, without libdevice's division loop in the common range: for |a| < 2 pi the result
// is a itself; for 2 pi <= |a| < 4 pi it is |a| - 2 pi with a's sign, which is exact (Sterbenz: y <= x <= 2y)
__device__ __forceinline__ double fmod_two_pi(double a) {
    const double aa = fabs(a);
    if (aa < MCL_TWO_PI) return a;
    if (aa < 2.0 * MCL_TWO_PI) return copysign(__dadd_rn(aa, -MCL_TWO_PI), a);
    return fmod(a, MCL_TWO_PI);
}
__device__ __forceinline__ double normalize_angle_dev(double theta) {
    double r = fmod_two_pi(__dadd_rn(theta, MCL_PI));
    if (r != 0.0 && r < 0.0) r = __dadd_rn(r, MCL_TWO_PI);
    return __dadd_rn(r, -MCL_PI);
}

// Philox4x32-10 (Salmon et al. SC'11).  Same counter layout as oracle/c/mcl_oracle.c draw4().
__device__ __forceinline__ uint4 philox4x32_10(uint4 c, uint2 k) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint32_t hi0 = __umulhi(0xD2511F53u, c.x), lo0 = 0xD2511F53u * c.x;
        const uint32_t hi1 = __umulhi(0xCD9E8D57u, c.z), lo1 = 0xCD9E8D57u * c.z;
        c = make_uint4(hi1 ^ c.y ^ k.x, lo1, hi0 ^ c.w ^ k.y, lo0);
        k.x += 0x9E3779B9u;
        k.y += 0xBB67AE85u;
    }
    return c;
}
__device__ __forceinline__ uint4 philox_draw4(uint64_t seed, uint64_t step, uint64_t item,
                                              uint32_t sub, uint32_t stream) {
    uint4 c;
    c.x = (uint32_t)item;
    c.y = (uint32_t)step;
    c.z = sub;
    c.w = (stream & 0xffu) | ((uint32_t)(item >> 32) << 8) | ((uint32_t)((step >> 32) & 0xffu) << 24);
    return philox4x32_10(c, make_uint2((uint32_t)seed, (uint32_t)(seed >> 32)));
}
// 53-bit uniform in [0,1): same bit recipe as numpy's legacy random_sample.
__device__ __forceinline__ double u53_from(uint32_t a, uint32_t b) {
    return ((double)(a >> 5) * 67108864.0 + (double)(b >> 6)) / 9007199254740992.0;
}
// Three standard normals by Box-Muller on 32-bit uniforms (u1 in (0,1], u2 in [0,1)); same construction as
// oracle/c/mcl_oracle.c normals3().  The attempt's own MOTION block (particle, step, t) gives the angle of the first
// pair (word 0), the second pair (words 1, 2) and, in word 3, the part of the RADIUS WORD that is not dealt out by
// the two screening streams below.
//
// Radius word of attempt t (uniform on [0, 2^32), independent over t -- every bit is used once):
//   t = 0 : word 3 of the own block.
//   t >= 1: top nibble = nibble (t & 31) of the MOTION_R block (particle, step, t >> 5): one Philox call deals the
//           top nibbles of 32 attempts, so a warp screens 1024 attempts of the rejection loop with one call per lane.
//           nibble != 0: the other 28 bits are the top 28 bits of word 3 of the own block.
//           nibble == 0 (one attempt in 16; the only ones that can pass a screening threshold <= 2^28): the next
//           16 bits are half-word (c & 7) of the MOTION_R2 block (particle, step, c >> 3), c = number of attempts
//           1 <= s < t with a zero nibble (dealt by RANK, so the survivors of the first level cost one call per eight);
//           the low 12 bits are the low 12 bits of word 3 of the own block.
#define MCL_STREAM_MOTION_R 6
#define MCL_STREAM_MOTION_R2 7
__device__ __forceinline__ uint32_t pick_word(const uint4 &a, uint32_t k) {
    return k == 0 ? a.x : (k == 1 ? a.y : (k == 2 ? a.z : a.w));
}
__device__ __forceinline__ void normals3_from_words(uint32_t w_radius, const uint4 &o, double &z0, double &z1,
                                                    double &z2) {
    const double k = 2.3283064365386963e-10;  // 2^-32
    const double u1 = __dmul_rn(__dadd_rn((double)w_radius, 1.0), k), u2 = __dmul_rn((double)o.x, k);
    const double u3 = __dmul_rn(__dadd_rn((double)o.y, 1.0), k), u4 = __dmul_rn((double)o.z, k);
    const double r1 = sqrt(__dmul_rn(-2.0, log(u1))), r2 = sqrt(__dmul_rn(-2.0, log(u3)));
    double s1, c1, s2, c2;
    sincos(__dmul_rn(MCL_TWO_PI, u2), &s1, &c1);
    sincos(__dmul_rn(MCL_TWO_PI, u4), &s2, &c2);
    z0 = __dmul_rn(r1, c1);
    z1 = __dmul_rn(r1, s1);
    z2 = __dmul_rn(r2, c2);
}
__device__ __forceinline__ uint32_t half_word(const uint4 &a, uint32_t k) {     // k = 0..7
    const uint32_t w = pick_word(a, k >> 1);
    return (k & 1u) ? (w >> 16) : (w & 0xffffu);
}
// bit 4k set <=> nibble k of x is zero
__device__ __forceinline__ uint32_t zero_nibble_flags(uint32_t x) {
    return ((x | (x >> 1) | (x >> 2) | (x >> 3)) & 0x11111111u) ^ 0x11111111u;
}
// radius word of an attempt t >= 1 from its top nibble, its MOTION_R2 half-word (used when the nibble is zero) and
// its own block
__device__ __forceinline__ uint32_t radius_word(uint32_t nibble, uint32_t mid16, const uint4 &o) {
    return nibble ? ((nibble << 28) | (o.w >> 4)) : ((mid16 << 12) | (o.w & 0xfffu));
}
// attempt 0 (every particle, every step): one Philox call
__device__ __forceinline__ void philox_normals3_first(uint64_t seed, uint64_t step, uint64_t item, double &z0,
                                                      double &z1, double &z2) {
    const uint4 o = philox_draw4(seed, step, item, 0u, MCL_STREAM_MOTION);
    normals3_from_words(o.w, o, z0, z1, z2);
}

// pu:388-396 is_valid_position: trunc-toward-zero cell index, cell == 0 only.
__device__ __forceinline__ bool is_valid_position_dev(double x, double y, const int8_t *__restrict__ occ,
                                                      int W, int H, double res, double ox, double oy) {
    const long long mx = __double2ll_rz(__ddiv_rn(__dadd_rn(x, -ox), res));
    const long long my = __double2ll_rz(__ddiv_rn(__dadd_rn(y, -oy), res));
    if (mx >= 0 && mx < W && my >= 0 && my < H) return occ[my * (long long)W + mx] == 0;
    return false;
}
// pu:135-142 evaluated once per map cell (the per-beam value depends only on dist[cell] when
// 0 <= r <= max_range): logtab = (float) log(max(z_hit * p_hit + z_rand / max_range, 1e-6)).
// dist ** 2 is an f32 multiply (numba types float32 ** 2 as float32; pinned by the golden vectors).
__device__ __forceinline__ double cell_p_hit(float d, double sigma_hit, double max_range) {
    const double dsq = (double)__fmul_rn(d, d);
    const double s2 = __dmul_rn(sigma_hit, sigma_hit);
    if ((double)d <= max_range)
        return __ddiv_rn(exp(__ddiv_rn(__dmul_rn(-0.5, dsq), s2)), sqrt(__dmul_rn(MCL_TWO_PI, s2)));
    return 0.0;
}
__device__ __forceinline__ double cell_logp(float d, double sigma_hit, double z_hit, double z_rand, double max_range,
                            bool with_rand) {
    const double p_hit = cell_p_hit(d, sigma_hit, max_range);
    const double p_rand = with_rand ? __ddiv_rn(1.0, max_range) : 0.0;
    double p = __dadd_rn(__dmul_rn(z_hit, p_hit), __dmul_rn(z_rand, p_rand));
    p = p > 1e-6 ? p : 1e-6;
    return log(p);
}

// order-preserving map float -> unsigned (atomicMax on the key == max on the float; key 0 is below every float)
__device__ __forceinline__ unsigned mcl_key_of_float(float v) {
    const unsigned b = __float_as_uint(v);
    return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
__device__ __forceinline__ float mcl_float_of_key(unsigned k) {
    return __uint_as_float((k & 0x80000000u) ? (k & 0x7fffffffu) : ~k);
}
#define MCL_LOGP_SCALE 33554432.0   // 2^25
#define MCL_ACC_TERMS 8
__device__ __forceinline__ int32_t quantise_logp(double lp) { return __double2int_rn(lp * MCL_LOGP_SCALE); }
#endif  // __CUDACC__
