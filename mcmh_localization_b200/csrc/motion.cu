// motion.cu -- kernel (1): odometry motion-model sample with map rejection,
// pu:332-363 apply_motion_model_parallel + pu:388-396 is_valid_position.
//
// One thread per particle draws attempt 0.  The reference retries up to 1000 times, and the
// ~2 % of particles facing a wall burn all of them -- serialised inside a warp that
// would stall 31 finished lanes for 999 iterations.  Because the generator is counter-based
// (Philox keyed by (seed, step, particle, attempt)) any lane can evaluate any attempt, so the warp
// retries cooperatively: the 32 lanes evaluate attempts t..t+31 of ONE pending particle at once
// and the lowest successful attempt wins -- identical to the sequential "first valid attempt"
// semantics, 32x fewer iterations.  The same code path serves injected draws (parity tests).
#include <algorithm>

#include "common.cuh"

struct MotionParams {
    const double *x, *y, *th;
    int64_t n;
    double rot1, trans, rot2;
    double s1, s2, s3;         // pu:345-347 sigmas, computed on the host in the reference's order
    const int8_t *occ;
    int W, H;
    double res, ox, oy;
    uint64_t seed, step, first_index;
    const double *normals;     // nullable, (n, A, 3)
    int A;
    int max_attempts;
    double *xo, *yo, *tho;
    int32_t *attempts;
    // screening levels (warp-uniform, computed on the host): sector bounds for |z| <= rho_k
    int n_levels;
    double lv_ulo[10], lv_uhi[10], lv_hv[10];      // along-heading interval and half width of the sector box
    unsigned long long lv_thr[10];                  // an attempt can only succeed if radius word + 1 <= thr
    double inv_res;
    // optional per-particle cache of the screening threshold (~0 = not computed yet): the MH chain proposes from
    // the SAME poses with the SAME increment 31 more times, and the threshold depends on nothing else
    unsigned long long *thr_cache;
    // MCL_MOTION_STATS=1: counters of the rejection loop (mcl_debug_motion_stats), else null
    unsigned long long *stats;
};
#define MOTION_STAT(p, k, v) do { if ((p).stats) atomicAdd((p).stats + (k), (unsigned long long)(v)); } while (0)

struct Pose { double x, y, th; };

__device__ __forceinline__ bool motion_candidate(const MotionParams &p, double x, double y, double th, double z0,
                                                 double z1, double z2, Pose &cand);

// one attempt t for a particle at (x,y,th); returns validity and the candidate pose
__device__ __forceinline__ bool motion_attempt(const MotionParams &p, int64_t i, int t, double x, double y,
                                               double th, Pose &cand) {
    double z0, z1, z2;
    if (p.normals) {
        const double *z = p.normals + ((size_t)i * p.A + (size_t)(t % p.A)) * 3;
        z0 = z[0]; z1 = z[1]; z2 = z[2];
    } else {
        philox_normals3(p.seed, p.step, p.first_index + (uint64_t)i, (uint32_t)t, z0, z1, z2);
    }
    return motion_candidate(p, x, y, th, z0, z1, z2, cand);
}

// ---------------------------------------------------------------------------------------------
// Exact screening of the rejection loop.  A particle facing a wall burns all max_attempts in the
// reference (~2 % of a uniform cloud, growing to >10 % of a running filter because stuck
// particles never move).  Every candidate of an attempt whose normals satisfy |z0|, |z1| <= rho lies in
// the annular sector  t in trans +- rho s2,  angle in theta + rot1 +- rho s1.  region_blocked(rho)
// proves (separating-axis test against every FREE cell near the sector, everything inflated by 1e-9 m)
// that this sector touches no free cell.  Since |z0|, |z1| <= R1 = sqrt(-2 ln u1), an attempt whose
// radius word gives R1 < rho cannot succeed -- an integer compare on the raw Philox word, no log / sqrt /
// sincos.  The high halves of the radius words of eight consecutive attempts come from one Philox call, so a lane
// screens eight attempts per call and a warp 256 attempts per iteration; only attempts that pass are evaluated in
// full.  Results are identical to evaluating every attempt in order (tests: vs the oracle's plain loop).
// ---------------------------------------------------------------------------------------------
__device__ bool region_blocked(const MotionParams &p, double x, double y, double sn, double cs, int k) {
    // oriented rectangle containing the sector of level k: along the nominal heading u in [u_lo, u_hi],
    // across it |v| <= hv (all inflated by 1e-9 m on the host); centre (ccx, ccy), half extents (hu, hv)
    const double u_lo = p.lv_ulo[k], u_hi = p.lv_uhi[k], hv = p.lv_hv[k];
    const double hu = 0.5 * (u_hi - u_lo), um = 0.5 * (u_hi + u_lo);
    const double ccx = x + um * cs, ccy = y + um * sn;
    const double ax = fabs(cs), ay = fabs(sn);
    const double bx = hu * ax + hv * ay + 1e-9, by = hu * ay + hv * ax + 1e-9;     // its axis-aligned half box
    // cells touched by the box: a conservative superset is enough, so multiply by 1/res (relative error
    // 1e-16, far inside the 1e-9 m inflation) and widen by one ulp-ish margin instead of dividing
    const double fx0 = (ccx - bx - p.ox) * p.inv_res - 1e-7, fx1 = (ccx + bx - p.ox) * p.inv_res + 1e-7;
    const double fy0 = (ccy - by - p.oy) * p.inv_res - 1e-7, fy1 = (ccy + by - p.oy) * p.inv_res + 1e-7;
    if (!(fx0 >= 1.0) || !(fy0 >= 1.0) || !(fx1 < 2.0e9) || !(fy1 < 2.0e9)) return false;   // int() quirk at the map edge
    const int mx0 = (int)fx0, mx1 = (int)fx1, my0 = (int)fy0, my1 = (int)fy1;
    if ((mx1 - mx0 + 1) * (my1 - my0 + 1) > 36) return false;         // large region: no claim
    const double hc = 0.5 * p.res + 1e-9;                             // half cell, inflated
    for (int my = my0; my <= my1; ++my)
        for (int mx = mx0; mx <= mx1; ++mx) {
            if (mx >= p.W || my >= p.H || p.occ[(size_t)my * p.W + mx] != 0) continue;
            const double qx = p.ox + ((double)mx + 0.5) * p.res - ccx, qy = p.oy + ((double)my + 0.5) * p.res - ccy;
            const bool separated = fabs(qx) > bx + hc || fabs(qy) > by + hc ||
                                   fabs(qx * cs + qy * sn) > hu + hc * (ax + ay) ||
                                   fabs(-qx * sn + qy * cs) > hv + hc * (ax + ay);
            if (!separated) return false;
        }
    return true;
}

// Largest screening level whose sector is blocked -> threshold on (radius word + 1): an attempt can only
// succeed if word + 1 <= T.  T = 2^32 (no screening) ... 0 (provably stuck: even R1 = 6.6604 stays blocked).
// blocked(rho) is monotone (a larger sector contains the smaller ones), so test the top level first -- the
// common outcome for a particle facing a wall -- and bisect otherwise.
__device__ unsigned long long screening_threshold(const MotionParams &p, double x, double y, double th) {
    if (p.n_levels == 0) return 1ull << 32;
    double sn, cs;
    sincos(th + p.rot1, &sn, &cs);
    if (region_blocked(p, x, y, sn, cs, p.n_levels - 1)) return p.lv_thr[p.n_levels - 1];
    int lo = -1, hi = p.n_levels - 1;          // blocked(lo) (or lo = -1), not blocked(hi)
    while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (region_blocked(p, x, y, sn, cs, mid)) lo = mid; else hi = mid;
    }
    return lo < 0 ? (1ull << 32) : p.lv_thr[lo];
}

__device__ __forceinline__ double shfl_d(double v, int src) { return __shfl_sync(0xffffffffu, v, src); }

// pose update of pu:350-353 from three normals; returns validity
__device__ __forceinline__ bool motion_candidate(const MotionParams &p, double x, double y, double th, double z0,
                                                 double z1, double z2, Pose &cand) {
    // np.random.normal(0, s) = 0.0 + s*z ; no FMA contraction anywhere (numba does not fuse)
    const double r1_hat = __dadd_rn(p.rot1, __dadd_rn(0.0, __dmul_rn(p.s1, z0)));
    const double t_hat = __dadd_rn(p.trans, __dadd_rn(0.0, __dmul_rn(p.s2, z1)));
    const double r2_hat = __dadd_rn(p.rot2, __dadd_rn(0.0, __dmul_rn(p.s3, z2)));
    double sn, cs;
    sincos(__dadd_rn(th, r1_hat), &sn, &cs);
    cand.x = __dadd_rn(x, __dmul_rn(t_hat, cs));                                  // pu:351
    cand.y = __dadd_rn(y, __dmul_rn(t_hat, sn));                                  // pu:352
    cand.th = normalize_angle_dev(__dadd_rn(__dadd_rn(th, r1_hat), r2_hat));      // pu:353
    return is_valid_position_dev(cand.x, cand.y, p.occ, p.W, p.H, p.res, p.ox, p.oy);
}

#define MOTION_Q 320     // >= 31 leftover + 256 new candidates per screening round
// Small CTAs: a CTA stays resident until its slowest warp is done, and the warps that hold stuck particles run
// 10-100x longer than the others -- with 8 warps per CTA the SM averaged 13 resident warps (ncu r2a: 20 %).
template <int MOTION_BLOCK>
__global__ void __launch_bounds__(MOTION_BLOCK, 1024 / MOTION_BLOCK) k_motion(const MotionParams p) {
    __shared__ unsigned short q_att[MOTION_BLOCK / 32][MOTION_Q];
    __shared__ unsigned short q_hi[MOTION_BLOCK / 32][MOTION_Q];
    const int lane = threadIdx.x & 31;
    const int64_t warp_base = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) & ~(int64_t)31;
    if (warp_base >= p.n) return;                       // warp-uniform
    const int64_t i = warp_base + lane;
    const bool live = i < p.n;
    double x = 0, y = 0, th = 0;
    if (live) { x = p.x[i]; y = p.y[i]; th = p.th[i]; }
    Pose out = {x, y, th};
    int32_t att = 0;
    unsigned long long thr = 1ull << 32;
    bool done = !live || p.max_attempts <= 0;
    if (!done) {
        Pose c;
        if (motion_attempt(p, i, 0, x, y, th, c)) { out = c; att = 1; done = true; }
        else if (!p.normals) {
            if (p.thr_cache) {
                thr = p.thr_cache[i];
                if (thr == ~0ull) { thr = screening_threshold(p, x, y, th); p.thr_cache[i] = thr; }
            } else {
                thr = screening_threshold(p, x, y, th);
            }
            if (thr == 0ull) done = true;                // provably stuck: att = 0, pose kept (pu:360-361)
        }
    }
    unsigned pending = __ballot_sync(0xffffffffu, !done);
    if (p.stats) {
        const unsigned f0 = __ballot_sync(0xffffffffu, live && att == 0 && p.max_attempts > 0);
        const unsigned ps = __ballot_sync(0xffffffffu, live && att == 0 && thr == 0ull);
        const unsigned hard = __ballot_sync(0xffffffffu, !done && thr <= (1ull << 28));
        if (lane == 0) {
            MOTION_STAT(p, 0, __popc(f0)); MOTION_STAT(p, 1, __popc(ps)); MOTION_STAT(p, 2, __popc(pending));
            MOTION_STAT(p, 6, __popc(hard)); MOTION_STAT(p, 8, pending ? 1 : 0);
        }
    }
    while (pending) {
        const int src = __ffs(pending) - 1;
        const double sx = shfl_d(x, src), sy = shfl_d(y, src), sth = shfl_d(th, src);
        const int64_t si = warp_base + src;
        Pose win = {sx, sy, sth};
        int watt = 0;
        if (p.normals) {
            // injected draws: the lanes evaluate attempts t0..t0+31 of this particle, lowest valid wins
            for (int t0 = 1; t0 < p.max_attempts; t0 += 32) {
                const int t = t0 + lane;
                Pose c = {0, 0, 0};
                const bool ok = (t < p.max_attempts) && motion_attempt(p, si, t, sx, sy, sth, c);
                const unsigned okm = __ballot_sync(0xffffffffu, ok);
                if (okm) {
                    const int w = __ffs(okm) - 1;
                    win.x = shfl_d(c.x, w); win.y = shfl_d(c.y, w); win.th = shfl_d(c.th, w);
                    watt = t0 + w + 1;
                    break;
                }
            }
        } else {
            // Philox draws.  Screening: lane L reads the high halves of the radius words of attempts 8g..8g+7
            // (g = g0 + L) from one Philox block and queues, in attempt order, the attempts whose radius can reach
            // the threshold whatever the low half.  Full evaluation (log, sqrt, sincos, map lookup: ~1000
            // instructions) then takes 32 queued attempts at a time, lowest index first -- evaluating a passing
            // attempt inside the screening loop would run it with one or two active lanes.
            const unsigned long long T = __shfl_sync(0xffffffffu, thr, src);
            const uint64_t item = p.first_index + (uint64_t)si;
            const int groups = (p.max_attempts + 7) >> 3;
            unsigned short *qt = q_att[threadIdx.x >> 5];
            unsigned short *qw = q_hi[threadIdx.x >> 5];
            int qlen = 0, qhead = 0, g0 = 0;
            bool found = false;
            while (!found && (g0 < groups || qhead < qlen)) {
                // fewer than 32 candidates left: move them to the front of the queue before screening more
                if (qhead > 0 && g0 < groups && qlen - qhead < 32) {
                    const int left = qlen - qhead;
                    unsigned short ta = 0, tw = 0;
                    if (lane < left) { ta = qt[qhead + lane]; tw = qw[qhead + lane]; }
                    __syncwarp();
                    if (lane < left) { qt[lane] = ta; qw[lane] = tw; }
                    __syncwarp();
                    qlen = left; qhead = 0;
                }
                // screen until 32 candidates are queued (or the attempts are exhausted)
                while (g0 < groups && qlen - qhead < 32) {
                    const int g = g0 + lane;
                    uint4 a = make_uint4(0u, 0u, 0u, 0u);
                    int cnt = 0;
                    unsigned passm = 0;
                    if (lane == 0) MOTION_STAT(p, 3, 1);
                    if (g < groups) {
                        a = philox_draw4(p.seed, p.step, item, (uint32_t)g, MCL_STREAM_MOTION_R);
#pragma unroll
                        for (int k = 0; k < 8; ++k) {
                            const int t = 8 * g + k;
                            // smallest radius word with this high half: (hi << 16); it must satisfy word + 1 <= T
                            const bool ps = t >= 1 && t < p.max_attempts &&
                                            ((unsigned long long)radius_hi16(a, (uint32_t)k) << 16) + 1ull <= T;
                            passm |= (ps ? 1u : 0u) << k;
                            cnt += ps;
                        }
                    }
                    if (__ballot_sync(0xffffffffu, cnt > 0) == 0u) {  // nothing passed (the usual outcome for a particle
                        g0 += 32;                                    // facing a wall): skip the queue bookkeeping
                        continue;
                    }
                    int pre = cnt;                                   // exclusive prefix over lanes -> queue order = attempt order
#pragma unroll
                    for (int o = 1; o < 32; o <<= 1) { const int v = __shfl_up_sync(0xffffffffu, pre, o); if (lane >= o) pre += v; }
                    const int total = __shfl_sync(0xffffffffu, pre, 31);
                    int pos = qlen + pre - cnt;
#pragma unroll
                    for (int k = 0; k < 8; ++k)
                        if ((passm >> k) & 1u) { qt[pos] = (unsigned short)(8 * g + k); qw[pos] = (unsigned short)radius_hi16(a, (uint32_t)k); ++pos; }
                    qlen += total;
                    g0 += 32;
                    __syncwarp();
                }
                // evaluate up to 32 queued attempts, lowest attempt index first
                const int e = qhead + lane;
                Pose c = {0, 0, 0};
                bool ok = false;
                int t = 0;
                if (lane == 0 && qlen > qhead) { MOTION_STAT(p, 4, 1); MOTION_STAT(p, 7, min(32, qlen - qhead)); }
                if (e < qlen) {
                    t = qt[e];
                    const uint4 o = philox_draw4(p.seed, p.step, item, (uint32_t)t, MCL_STREAM_MOTION);
                    double z0, z1, z2;
                    normals3_from_words(radius_word((uint32_t)qw[e], o), o, z0, z1, z2);
                    ok = motion_candidate(p, sx, sy, sth, z0, z1, z2, c);
                }
                const unsigned okm = __ballot_sync(0xffffffffu, ok);
                if (okm) {
                    const int w = __ffs(okm) - 1;                    // queue order = attempt order: lowest valid attempt
                    win.x = shfl_d(c.x, w); win.y = shfl_d(c.y, w); win.th = shfl_d(c.th, w);
                    watt = __shfl_sync(0xffffffffu, t, w) + 1;
                    found = true;
                }
                qhead += 32;
                if (qhead >= qlen) { qhead = 0; qlen = 0; }          // queue drained: reuse it from the start
                __syncwarp();
            }
        }
        if (lane == 0 && watt) MOTION_STAT(p, 5, 1);
        if (lane == src) { out = win; att = watt; }       // watt == 0: keep the old pose (pu:360-361)
        pending &= pending - 1;
    }
    if (live) {
        p.xo[i] = out.x; p.yo[i] = out.y; p.tho[i] = out.th;
        if (p.attempts) p.attempts[i] = att;
    }
}

extern "C" int mcl_predict(mcl_handle *h, const double *d_x, const double *d_y, const double *d_theta,
                           int64_t n, const double delta[3], uint64_t seed, uint64_t step,
                           uint64_t first_index, const double *d_normals, int A, int max_attempts,
                           double *d_xo, double *d_yo, double *d_thetao, int32_t *d_attempts) {
    return mcl_predict_cached(h, d_x, d_y, d_theta, n, delta, seed, step, first_index, d_normals, A, max_attempts, d_xo,
                              d_yo, d_thetao, d_attempts, nullptr);
}

// internal: d_thr_cache (n values, ~0 = unknown) is valid for ONE set of source poses and ONE increment
int mcl_predict_cached(mcl_handle *h, const double *d_x, const double *d_y, const double *d_theta, int64_t n,
                       const double delta[3], uint64_t seed, uint64_t step, uint64_t first_index,
                       const double *d_normals, int A, int max_attempts, double *d_xo, double *d_yo, double *d_thetao,
                       int32_t *d_attempts, unsigned long long *d_thr_cache) {
    if (!h) return MCL_ERR_ARG;
    if (n < 0 || !delta || (n > 0 && (!d_x || !d_y || !d_theta || !d_xo || !d_yo || !d_thetao)))
        return mcl_fail(h, MCL_ERR_ARG, "mcl_predict: bad argument");
    // the cooperative retry queue keeps attempt indices in 16 bits (pu:339 uses 1000)
    if (max_attempts > 65535) return mcl_fail(h, MCL_ERR_ARG, "mcl_predict: max_attempts > 65535");
    if (d_normals && A <= 0) return mcl_fail(h, MCL_ERR_ARG, "mcl_predict: injected normals need A > 0");
    if (!h->d_occ) return mcl_fail(h, MCL_ERR_STATE, "mcl_predict: map not set");
    if (!h->motion_set) return mcl_fail(h, MCL_ERR_STATE, "mcl_predict: motion noise not set (mcl_set_motion)");
    if (n == 0) return MCL_OK;
    DeviceGuard guard(h->device);
    MotionParams p;
    p.x = d_x; p.y = d_y; p.th = d_theta; p.n = n;
    p.rot1 = delta[0]; p.trans = delta[1]; p.rot2 = delta[2];
    // alpha is float32[4] promoted to f64 inside the reference's compiled code
    const double a1 = (double)h->alpha[0], a2 = (double)h->alpha[1], a3 = (double)h->alpha[2],
                 a4 = (double)h->alpha[3];
    volatile double t1, t2;   // keep the host compiler from contracting a*b+c
    t1 = a1 * fabs(p.rot1); t2 = a2 * fabs(p.trans); p.s1 = t1 + t2;
    t1 = a3 * fabs(p.trans); t2 = a4 * (fabs(p.rot1) + fabs(p.rot2)); p.s2 = t1 + t2;
    t1 = a1 * fabs(p.rot2); t2 = a2 * fabs(p.trans); p.s3 = t1 + t2;
    p.occ = h->d_occ; p.W = h->W; p.H = h->H; p.res = h->res; p.ox = h->ox; p.oy = h->oy;
    p.seed = seed; p.step = step; p.first_index = first_index;
    p.normals = d_normals; p.A = A; p.max_attempts = max_attempts;
    p.xo = d_xo; p.yo = d_yo; p.tho = d_thetao; p.attempts = d_attempts;
    p.thr_cache = d_normals ? nullptr : d_thr_cache;
    static const bool want_stats = getenv("MCL_MOTION_STATS") != nullptr;
    if (want_stats && !h->d_motion_stats) {
        MCL_CUDA(h, cudaMalloc((void **)&h->d_motion_stats, 16 * 8));
        MCL_CUDA(h, cudaMemsetAsync(h->d_motion_stats, 0, 16 * 8, h->stream));
    }
    p.stats = (unsigned long long *)h->d_motion_stats;
    // screening levels: every candidate with |z0|, |z1| <= rho has t in trans +- rho s2 and heading within
    // +- rho s1 of theta + rot1.  Valid only while t stays positive and the spread small; levels are nested.
    const double levels[10] = {1.0, 2.0, 3.0, 3.5, 4.0, 4.5, 5.0, 5.5, 6.0, 6.6605};
    p.n_levels = 0;
    p.inv_res = 1.0 / h->res;
    for (int k = 0; k < 10 && !d_normals; ++k) {
        const double rho = levels[k], dt = rho * p.s2, da = rho * p.s1;
        const double t_lo = p.trans - dt, t_hi = p.trans + dt;
        if (!(t_lo > 0.0) || !(da < 1.0)) break;
        p.lv_ulo[k] = t_lo * cos(da) - 1e-9;
        p.lv_uhi[k] = t_hi + 1e-9;
        p.lv_hv[k] = t_hi * sin(da) + 1e-9;
        // R1 >= rho  <=>  u1 <= exp(-rho^2/2); relative margin far above the fp64 error of log/sqrt.
        // rho = 6.6605 exceeds the largest possible radius sqrt(-2 ln 2^-32) = 6.6604: nothing passes.
        p.lv_thr[k] = rho > 6.66 ? 0ull
                                 : (unsigned long long)(4294967296.0 * exp(-0.5 * rho * rho) * (1.0 + 1e-9)) + 2ull;
        p.n_levels = k + 1;
    }
    static const int block = [] {                        // MCL_MOTION_BLOCK=32|64|128|256 for A/B measurements
        const char *e = getenv("MCL_MOTION_BLOCK");
        const int v = e ? atoi(e) : 64;
        return (v == 32 || v == 64 || v == 128 || v == 256) ? v : 64;
    }();
    const int blocks = (int)((n + block - 1) / block);
    switch (block) {
    case 32:  k_motion<32><<<blocks, 32, 0, h->stream>>>(p); break;
    case 128: k_motion<128><<<blocks, 128, 0, h->stream>>>(p); break;
    case 256: k_motion<256><<<blocks, 256, 0, h->stream>>>(p); break;
    default:  k_motion<64><<<blocks, 64, 0, h->stream>>>(p); break;
    }
    MCL_LAUNCH_CHECK(h);
    return MCL_OK;
}

// Debug (MCL_MOTION_STATS=1): counters of the rejection loop since the last call, then reset.
// [0] attempt 0 failed, [1] provably stuck, [2] particles retried, [3] screening rounds (one Philox call per lane),
// [4] evaluation rounds, [5] retries that found a pose, [6] retried with threshold <= 2^28, [7] attempts evaluated,
// [8] warps with a retry.
int mcl_debug_motion_stats(mcl_handle *h, unsigned long long out[16]) {
    if (!h || !out) return MCL_ERR_ARG;
    for (int k = 0; k < 16; ++k) out[k] = 0;
    if (!h->d_motion_stats) return MCL_OK;
    DeviceGuard guard(h->device);
    MCL_CUDA(h, cudaStreamSynchronize(h->stream));
    MCL_CUDA(h, cudaMemcpy(out, h->d_motion_stats, 16 * 8, cudaMemcpyDeviceToHost));
    MCL_CUDA(h, cudaMemset(h->d_motion_stats, 0, 16 * 8));
    return MCL_OK;
}
