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
    unsigned long long min_thr;   // test hook (mcl_debug_motion): screening thresholds below it are raised to it
    // retry list (Philox draws): particles whose attempt 0 failed and that are not provably stuck
    int32_t *retry_idx;
    unsigned long long *retry_thr;
    unsigned *retry_count, *retry_done;      // retry_count[0]: thresholds <= 2^28 (front), [1]: the others (back)
    unsigned *retry_ticket;                  // [0] dense, [1] sparse: next entry to hand out
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
        // Philox mode: only attempt 0 comes through here (the retries are screened in k_motion)
        philox_normals3_first(p.seed, p.step, p.first_index + (uint64_t)i, z0, z1, z2);
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

// Per-warp scratch of the rejection loop: the queue of attempts that passed the screen (attempt index + the top 20
// bits of its radius word) and the MOTION_R2 half-words of one screening round (<= 1024 zero nibbles + alignment).
template <int Q>
struct MotionWarpScratch {
    uint4 r2[130];                       // 1040 half-words
    uint32_t q_top[Q];
    unsigned short q_att[Q];
};
#define MOTION_BLOCK 128
#define MOTION_RETRY_BLOCK 256
#define MOTION_RETRY_BLOCKS_PER_SM 4

// pu:350-352 only: x, y of the candidate and its validity -- what decides whether an attempt is the accepted one.
// Same operations in the same order as motion_candidate (the heading and the third normal do not enter).
__device__ __forceinline__ bool motion_candidate_xy(const MotionParams &p, double x, double y, double th, double z0,
                                                    double z1) {
    const double r1_hat = __dadd_rn(p.rot1, __dadd_rn(0.0, __dmul_rn(p.s1, z0)));
    const double t_hat = __dadd_rn(p.trans, __dadd_rn(0.0, __dmul_rn(p.s2, z1)));
    double sn, cs;
    sincos(__dadd_rn(th, r1_hat), &sn, &cs);
    return is_valid_position_dev(__dadd_rn(x, __dmul_rn(t_hat, cs)), __dadd_rn(y, __dmul_rn(t_hat, sn)), p.occ, p.W, p.H,
                                 p.res, p.ox, p.oy);
}
// first pair of normals3_from_words (identical operations)
__device__ __forceinline__ void normals2_from_words(uint32_t w_radius, const uint4 &o, double &z0, double &z1) {
    const double k = 2.3283064365386963e-10;  // 2^-32
    const double u1 = __dmul_rn(__dadd_rn((double)w_radius, 1.0), k), u2 = __dmul_rn((double)o.x, k);
    const double r1 = sqrt(__dmul_rn(-2.0, log(u1)));
    double s1, c1;
    sincos(__dmul_rn(MCL_TWO_PI, u2), &s1, &c1);
    z0 = __dmul_rn(r1, c1);
    z1 = __dmul_rn(r1, s1);
}

// The retries 1 .. max_attempts - 1 of ONE particle with a screening threshold T <= 2^28 by a warp (Philox draws):
// returns the 1-based index of the first valid attempt (0: none) and its pose.  Identical to evaluating the attempts
// one by one in order.
template <int Q>
__device__ __forceinline__ int retry_particle(const MotionParams &p, MotionWarpScratch<Q> &ws, const int lane,
                                              const uint64_t item, const double sx, const double sy, const double sth,
                                              const unsigned long long T, Pose &win) {
    int watt = 0;
    uint32_t wwr = 0u;                                    // radius word of the accepted attempt
    // Philox draws, screened on the radius word (common.cuh): an attempt can only succeed if word + 1 <= T.
    // Level 1: lane L takes the MOTION_R block of attempts 32 g .. 32 g + 31 (g = g0 + L): one call per lane
    // deals the top nibbles of up to 1024 attempts.  With T <= 2^28 (a particle facing a wall) only zero
    // nibbles survive, one attempt in 16.  Level 2: the survivors' next 16 bits come from the MOTION_R2
    // stream in RANK order, eight per call.  What passes both is queued in attempt order and evaluated for validity
    // (x, y only: log, sqrt, two sincos, map lookup) 32 queued attempts at a time, lowest first; the accepted attempt
    // is then evaluated in full.
    const int nblk = (p.max_attempts + 31) >> 5;                // T <= 2^28 here: only zero nibbles can pass
    const uint32_t hmax = (uint32_t)((T - 1ull) >> 12);         // ... and of those the ones with (half << 12) + 1 <= T
    const unsigned short *r2h = reinterpret_cast<const unsigned short *>(ws.r2);
    int qlen = 0, qhead = 0, g0 = 0;
    unsigned rank_base = 0;                                    // zero nibbles among attempts 1 .. 32 g0 - 1
    bool found = false;
    while (!found && (g0 < nblk || qhead < qlen)) {
        // fewer than 32 candidates left: move them to the front of the queue before screening more
        if (qhead > 0 && g0 < nblk && qlen - qhead < 32) {
            const int left = qlen - qhead;
            unsigned short ta = 0; uint32_t tw = 0;
            if (lane < left) { ta = ws.q_att[qhead + lane]; tw = ws.q_top[qhead + lane]; }
            __syncwarp();
            if (lane < left) { ws.q_att[lane] = ta; ws.q_top[lane] = tw; }
            __syncwarp();
            qlen = left; qhead = 0;
        }
        // screen until 32 candidates are queued (or the attempts are exhausted)
        while (g0 < nblk && qlen - qhead < 32) {
            const int active = min(32, nblk - g0);
            const int g = g0 + lane;
            uint32_t zf[4] = {0u, 0u, 0u, 0u};
            int zc = 0;
            if (lane == 0) MOTION_STAT(p, 3, 1);
            if (lane < active) {
                const uint4 a = philox_draw4(p.seed, p.step, item, (uint32_t)g, MCL_STREAM_MOTION_R);
                if (g > 0 && 32 * g + 32 <= p.max_attempts) {           // all 32 attempts of the block are retries
#pragma unroll
                    for (int w = 0; w < 4; ++w) { zf[w] = zero_nibble_flags(pick_word(a, (uint32_t)w)); zc += __popc(zf[w]); }
                } else {
#pragma unroll
                    for (int w = 0; w < 4; ++w) {
                        const int lo = 32 * g + 8 * w;                  // attempts lo .. lo + 7
                        const int nv = min(8, max(0, p.max_attempts - lo));
                        uint32_t valid = nv == 8 ? 0x11111111u : (((1u << (4 * nv)) - 1u) & 0x11111111u);
                        if (lo == 0) valid &= ~1u;                      // attempt 0 is not part of the retry
                        zf[w] = zero_nibble_flags(pick_word(a, (uint32_t)w)) & valid;
                        zc += __popc(zf[w]);
                    }
                }
            }
            int zpre = zc;                                   // inclusive prefix of the zero-nibble counts
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const int v = __shfl_up_sync(0xffffffffu, zpre, o); if (lane >= o) zpre += v; }
            const int ztot = __shfl_sync(0xffffffffu, zpre, 31);
            // level 2: MOTION_R2 blocks of ranks rank_base .. rank_base + ztot - 1.  Usually <= 32 blocks (a round has
            // ~64 zero nibbles): lane j tests the eight half-words of block b_first + j in registers, and as a rule
            // none can reach the threshold -- the particle is done with this round without touching shared memory.
            const unsigned b_first = rank_base >> 3;
            const int nb2 = ztot ? (int)(((rank_base + (unsigned)ztot - 1u) >> 3) - b_first) + 1 : 0;
            if (nb2 <= 32) {
                uint4 b = make_uint4(0u, 0u, 0u, 0u);
                bool any = false;
                if (lane < nb2) {
                    b = philox_draw4(p.seed, p.step, item, b_first + (uint32_t)lane, MCL_STREAM_MOTION_R2);
                    const unsigned r0 = 8u * (b_first + (unsigned)lane);        // rank of half-word 0 of this block
#pragma unroll
                    for (int k = 0; k < 8; ++k) {
                        const unsigned r = r0 + (unsigned)k;
                        any |= r >= rank_base && r < rank_base + (unsigned)ztot && half_word(b, (uint32_t)k) <= hmax;
                    }
                }
                if (__ballot_sync(0xffffffffu, any) == 0u) {  // nothing passed (the usual outcome for a particle facing
                    g0 += active;                            // a wall)
                    rank_base += (unsigned)ztot;
                    continue;
                }
                if (lane < nb2) ws.r2[lane] = b;
            } else {
                for (int j0 = 0; j0 < nb2; j0 += 32)
                    if (j0 + lane < nb2) ws.r2[j0 + lane] = philox_draw4(p.seed, p.step, item, b_first + (uint32_t)(j0 + lane), MCL_STREAM_MOTION_R2);
            }
            __syncwarp();
            // pass flags: zero nibbles whose next 16 bits can still reach the threshold
            uint32_t ff[4];
            int cnt = 0;
            {
                unsigned r = rank_base + (unsigned)(zpre - zc) - 8u * b_first;      // index into r2h of this lane's first zero nibble
#pragma unroll
                for (int w = 0; w < 4; ++w) {
                    uint32_t f = 0u, z = zf[w];
                    while (z) {
                        const uint32_t b = z & (0u - z);
                        if ((uint32_t)r2h[r] <= hmax) f |= b;
                        ++r;
                        z ^= b;
                    }
                    ff[w] = f;
                    cnt += __popc(f);
                }
            }
            if (__ballot_sync(0xffffffffu, cnt > 0) == 0u) {  // nothing passed (the usual outcome for a particle
                g0 += active;                                // facing a wall): skip the queue bookkeeping
                rank_base += (unsigned)ztot;
                __syncwarp();
                continue;
            }
            int pre = cnt;                                   // inclusive prefix over lanes -> queue order = attempt order
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const int v = __shfl_up_sync(0xffffffffu, pre, o); if (lane >= o) pre += v; }
            // a round may not fit the queue (every nibble zero and passing: probability ~ 0, but the loop
            // must be exact): take the leading blocks that fit -- the first always does -- and redo the others
            const int space = Q - qlen;
            const int fit = min(active, __popc(__ballot_sync(0xffffffffu, pre <= space)));
            if (lane < fit) {
                int pos = qlen + pre - cnt;
                unsigned r = rank_base + (unsigned)(zpre - zc) - 8u * b_first;
#pragma unroll
                for (int w = 0; w < 4; ++w) {
                    uint32_t z = zf[w];
                    const uint32_t f = ff[w];
                    while (z) {                               // ascending over the zero nibbles of this word
                        const uint32_t b = z & (0u - z);
                        if (f & b) {
                            ws.q_att[pos] = (unsigned short)(32 * g + 8 * w + ((__ffs(b) - 1) >> 2));
                            ws.q_top[pos] = (uint32_t)r2h[r];  // top nibble 0 | the 16 bits dealt by rank
                            ++pos;
                        }
                        ++r;
                        z ^= b;
                    }
                }
            }
            qlen += __shfl_sync(0xffffffffu, pre, fit - 1);
            rank_base += (unsigned)__shfl_sync(0xffffffffu, zpre, fit - 1);
            g0 += fit;
            __syncwarp();
        }
        // evaluate up to 32 queued attempts, lowest attempt index first
        const int e = qhead + lane;
        bool ok = false;
        int t = 0;
        if (lane == 0 && qlen > qhead) { MOTION_STAT(p, 4, 1); MOTION_STAT(p, 7, min(32, qlen - qhead)); }
        uint4 o = make_uint4(0u, 0u, 0u, 0u);
        uint32_t wr = 0u;
        bool pass = false;
        if (e < qlen) {
            t = ws.q_att[e];
            const uint32_t top = ws.q_top[e];
            o = philox_draw4(p.seed, p.step, item, (uint32_t)t, MCL_STREAM_MOTION);
            wr = radius_word(top >> 16, top & 0xffffu, o);
            pass = (unsigned long long)wr + 1ull <= T;         // the exact word: drop what the screen let through
        }
        if (__ballot_sync(0xffffffffu, pass)) {
            if (pass) {
                double z0, z1;
                normals2_from_words(wr, o, z0, z1);
                ok = motion_candidate_xy(p, sx, sy, sth, z0, z1);
            }
        }
        const unsigned okm = __ballot_sync(0xffffffffu, ok);
        if (okm) {
            const int w = __ffs(okm) - 1;                    // queue order = attempt order: lowest valid attempt
            wwr = __shfl_sync(0xffffffffu, wr, w);
            watt = __shfl_sync(0xffffffffu, t, w) + 1;
            found = true;
        }
        qhead += 32;
        if (qhead >= qlen) { qhead = 0; qlen = 0; }          // queue drained: reuse it from the start
        __syncwarp();
    }
    if (watt) {                                           // the accepted attempt in full (every lane: same values)
        const uint4 o = philox_draw4(p.seed, p.step, item, (uint32_t)(watt - 1), MCL_STREAM_MOTION);
        double z0, z1, z2;
        normals3_from_words(wwr, o, z0, z1, z2);
        motion_candidate(p, sx, sy, sth, z0, z1, z2, win);
    }
    return watt;
}

// Kernel 1: attempt 0 of every particle.  A particle whose attempt 0 fails is proved stuck by the geometric screen
// (no draws: pose kept), or appended with its threshold to the retry list, which kernel 2 works off one particle
// per warp: the cloud is ordered by ancestor after resampling, so the particles facing a wall sit in runs, and a warp
// that retried its own 32 particles one after the other set the duration of the whole kernel (measured: 0.15-0.3 ms
// for the slowest warp against 0.03 ms for attempt 0).  Injected draws (parity tests) retry in place.
__global__ void __launch_bounds__(MOTION_BLOCK, 1024 / MOTION_BLOCK) k_motion(const MotionParams p) {
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
        if (!done && !p.normals && !p.retry_count) done = true;      // max_attempts == 1: nothing to retry
        if (!done && !p.normals) {
            if (p.thr_cache) {
                thr = p.thr_cache[i];
                if (thr == ~0ull) { thr = screening_threshold(p, x, y, th); p.thr_cache[i] = thr; }
            } else {
                thr = screening_threshold(p, x, y, th);
            }
            if (thr == 0ull) done = true;                // provably stuck: att = 0, pose kept (pu:360-361)
            else if (thr < p.min_thr) thr = p.min_thr;   // test hook: a looser screen is still exact
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
    if (!p.normals) {
        if (pending) {
            // append to the retry lists (order is irrelevant): thresholds <= 2^28 from the front of the arrays (a warp
            // each in kernel 2), the others -- one attempt in 16 or more passes the screen -- from the back (a CTA each)
            const unsigned dense = __ballot_sync(0xffffffffu, !done && thr > (1ull << 28));
            const unsigned sparse = pending & ~dense;
            unsigned bs = 0, bd = 0;
            if (lane == 0) {
                if (sparse) bs = atomicAdd(p.retry_count, (unsigned)__popc(sparse));
                if (dense) bd = atomicAdd(p.retry_count + 1, (unsigned)__popc(dense));
            }
            bs = __shfl_sync(0xffffffffu, bs, 0); bd = __shfl_sync(0xffffffffu, bd, 0);
            if (!done) {
                const unsigned below = (1u << lane) - 1u;
                const bool isd = (dense >> lane) & 1u;
                const int64_t slot = isd ? p.n - 1 - (int64_t)(bd + (unsigned)__popc(dense & below))
                                         : (int64_t)(bs + (unsigned)__popc(sparse & below));
                if (slot >= 0 && slot < p.n) {           // (always, unless an earlier call's second kernel never ran)
                    p.retry_idx[slot] = (int32_t)i;
                    p.retry_thr[slot] = thr;
                }
            }
        }
    } else {
        while (pending) {
            // injected draws: the lanes evaluate attempts t0..t0+31 of this particle, lowest valid wins
            const int src = __ffs(pending) - 1;
            const double sx = shfl_d(x, src), sy = shfl_d(y, src), sth = shfl_d(th, src);
            const int64_t si = warp_base + src;
            Pose win = {sx, sy, sth};
            int watt = 0;
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
            if (lane == src) { out = win; att = watt; }       // watt == 0: keep the old pose (pu:360-361)
            pending &= pending - 1;
        }
    }
    if (live) {
        p.xo[i] = out.x; p.yo[i] = out.y; p.tho[i] = out.th;     // retried particles: the old pose until kernel 2 finds one
        if (p.attempts) p.attempts[i] = att;
    }
}

// One block of 32 attempts (32 g .. 32 g + 31, lane k = attempt 32 g + k) of a particle whose threshold is above
// 2^28: every attempt whose top nibble is <= nmax is evaluated for validity.  a = its MOTION_R block, zbefore = zero
// nibbles among attempts 1 .. 32 g - 1.  Returns (attempt << 32 | radius word) of this lane's attempt if valid.
__device__ __forceinline__ unsigned long long dense_block(const MotionParams &p, const uint4 &a, unsigned zbefore, int g,
                                                          int lane, uint64_t item, double sx, double sy, double sth,
                                                          unsigned long long T, int nmax, unsigned &zeros) {
    const int t = 32 * g + lane;
    const uint32_t nib = (pick_word(a, (uint32_t)lane >> 3) >> (4 * (lane & 7))) & 15u;
    const bool valid = t >= 1 && t < p.max_attempts;
    const unsigned zmask = __ballot_sync(0xffffffffu, valid && nib == 0u);
    zeros = (unsigned)__popc(zmask);
    uint32_t mid = 0u;
    if (valid && nib == 0u) {
        const unsigned c = zbefore + (unsigned)__popc(zmask & ((1u << lane) - 1u));
        mid = half_word(philox_draw4(p.seed, p.step, item, c >> 3, MCL_STREAM_MOTION_R2), c & 7u);
    }
    unsigned long long r = ~0ull;
    if (valid && (int)nib <= nmax) {
        const uint4 o = philox_draw4(p.seed, p.step, item, (uint32_t)t, MCL_STREAM_MOTION);
        const uint32_t wr = radius_word(nib, mid, o);
        if ((unsigned long long)wr + 1ull <= T) {
            double z0, z1;
            normals2_from_words(wr, o, z0, z1);
            if (motion_candidate_xy(p, sx, sy, sth, z0, z1)) r = ((unsigned long long)(unsigned)t << 32) | wr;
        }
    }
    if (lane == 0) { MOTION_STAT(p, 3, 1); MOTION_STAT(p, 4, 1); }
    return r;
}
// the accepted attempt in full (called by a whole warp; lane 0 writes)
__device__ __forceinline__ void retry_commit(const MotionParams &p, int64_t i, uint64_t item, double sx, double sy,
                                             double sth, unsigned long long best, int lane) {
    const int watt = (int)(best >> 32) + 1;
    const uint4 o = philox_draw4(p.seed, p.step, item, (uint32_t)(watt - 1), MCL_STREAM_MOTION);
    double z0, z1, z2;
    Pose win;
    normals3_from_words((uint32_t)best, o, z0, z1, z2);
    motion_candidate(p, sx, sy, sth, z0, z1, z2, win);
    if (lane == 0) {
        MOTION_STAT(p, 5, 1);
        p.xo[i] = win.x; p.yo[i] = win.y; p.tho[i] = win.th;
        if (p.attempts) p.attempts[i] = watt;
    }
}

// Kernel 2: the retry lists, entries handed out by ticket.  Q = capacity of the candidate queue: 320 in production;
// the 64-entry instantiation exists so that tests reach the partial-round path (mcl_debug_motion).
// Dense entries (T > 2^28) first: a warp evaluates the first 64 attempts, which settles most of them.  The others
// may need every attempt evaluated -- up to 1000 evaluations, 0.08 ms on one warp, which set the duration of the
// kernel -- so the whole CTA takes them over: warp w evaluates attempt blocks 2 + w, 2 + w + 8, ... and the lowest
// valid attempt over the CTA wins.  Then the sparse entries (T <= 2^28), a warp each.
template <int Q>
__global__ void __launch_bounds__(MOTION_RETRY_BLOCK, MOTION_RETRY_BLOCKS_PER_SM) k_motion_retry(const MotionParams p) {
    constexpr int NW = MOTION_RETRY_BLOCK / 32;
    constexpr int WARP_STAGE = 2;                         // attempt blocks a single warp tries before the CTA takes over
    __shared__ MotionWarpScratch<Q> scratch[NW];
    __shared__ uint4 sm_nib[MOTION_RETRY_BLOCK];          // MOTION_R blocks of one super-chunk (256 x 32 attempts)
    __shared__ unsigned sm_zb[MOTION_RETRY_BLOCK];        // zero nibbles before each of them
    __shared__ unsigned sm_wtot[NW];
    __shared__ long long sm_heavy[NW];                    // list slots the warps pass on to the CTA (-1: none)
    __shared__ unsigned long long sm_best;                // (attempt << 32 | radius word) of the lowest valid attempt
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    unsigned n_sparse = ((volatile unsigned *)p.retry_count)[0], n_dense = ((volatile unsigned *)p.retry_count)[1];
    if ((int64_t)n_sparse + (int64_t)n_dense > p.n) { n_sparse = 0; n_dense = 0; }     // stale counters: nothing trustworthy
    const int nblk = (p.max_attempts + 31) >> 5;
    __shared__ unsigned sm_ticket;
    while (n_dense) {
        __syncthreads();
        if (threadIdx.x == 0) sm_ticket = atomicAdd(p.retry_ticket, (unsigned)NW);     // one entry per warp
        __syncthreads();
        const unsigned e = sm_ticket + (unsigned)warp;
        if (sm_ticket >= n_dense) break;                               // uniform over the CTA
        const bool have = e < n_dense;
        long long heavy = -1;
        if (have) {
            const int64_t slot = p.n - 1 - (int64_t)e;
            const int64_t i = p.retry_idx[slot];
            const unsigned long long T = p.retry_thr[slot];
            const int nmax = (int)((T - 1ull) >> 28);                  // nibbles n with (n << 28) + 1 <= T
            const uint64_t item = p.first_index + (uint64_t)i;
            const double sx = p.xo[i], sy = p.yo[i], sth = p.tho[i];
            unsigned zbefore = 0;
            unsigned long long best = ~0ull;
            for (int g = 0; g < min(nblk, WARP_STAGE) && best == ~0ull; ++g) {
                const uint4 a = philox_draw4(p.seed, p.step, item, (uint32_t)g, MCL_STREAM_MOTION_R);
                unsigned zeros;
                unsigned long long r = dense_block(p, a, zbefore, g, lane, item, sx, sy, sth, T, nmax, zeros);
                zbefore += zeros;
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) { const unsigned long long v = __shfl_xor_sync(0xffffffffu, r, o); r = v < r ? v : r; }
                best = r;
            }
            if (best != ~0ull) retry_commit(p, i, item, sx, sy, sth, best, lane);
            else if (nblk > WARP_STAGE) heavy = slot;
        }
        if (lane == 0) sm_heavy[warp] = heavy;
        __syncthreads();
        for (int hw = 0; hw < NW; ++hw) {
            const long long slot = sm_heavy[hw];
            if (slot < 0) continue;                                    // uniform over the CTA
            const int64_t i = p.retry_idx[slot];
            const unsigned long long T = p.retry_thr[slot];
            const int nmax = (int)((T - 1ull) >> 28);
            const uint64_t item = p.first_index + (uint64_t)i;
            const double sx = p.xo[i], sy = p.yo[i], sth = p.tho[i];
            unsigned carry = 0;                                        // zero nibbles among attempts 1 .. 32 G0 - 1
            if (threadIdx.x == 0) sm_best = ~0ull;
            bool found = false;
            for (int G0 = 0; G0 < nblk && !found; G0 += MOTION_RETRY_BLOCK) {
                // top nibbles of this super-chunk: thread t draws block G0 + t; ranks of the zero nibbles by a CTA scan
                const int gmine = G0 + (int)threadIdx.x;
                uint4 a = make_uint4(0u, 0u, 0u, 0u);
                int zc = 0;
                if (gmine < nblk) {
                    a = philox_draw4(p.seed, p.step, item, (uint32_t)gmine, MCL_STREAM_MOTION_R);
#pragma unroll
                    for (int w = 0; w < 4; ++w) {
                        const int lo = 32 * gmine + 8 * w;
                        const int nv = min(8, max(0, p.max_attempts - lo));
                        uint32_t valid = nv == 8 ? 0x11111111u : (((1u << (4 * nv)) - 1u) & 0x11111111u);
                        if (lo == 0) valid &= ~1u;
                        zc += __popc(zero_nibble_flags(pick_word(a, (uint32_t)w)) & valid);
                    }
                }
                int zpre = zc;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) { const int v = __shfl_up_sync(0xffffffffu, zpre, o); if (lane >= o) zpre += v; }
                __syncthreads();                                       // previous users of the shared arrays are done
                if (lane == 31) sm_wtot[warp] = (unsigned)zpre;
                sm_nib[threadIdx.x] = a;
                __syncthreads();
                unsigned before = carry, total = 0;
#pragma unroll
                for (int w = 0; w < NW; ++w) { if (w < warp) before += sm_wtot[w]; total += sm_wtot[w]; }
                sm_zb[threadIdx.x] = before + (unsigned)(zpre - zc);
                carry += total;
                __syncthreads();
                const int first = G0 == 0 ? WARP_STAGE : 0;            // the warp stage has tried blocks 0 .. WARP_STAGE - 1
                const int chunk_blocks = min(MOTION_RETRY_BLOCK, nblk - G0);
                for (int r0 = first; r0 < chunk_blocks && !found; r0 += NW) {
                    const int gl = r0 + warp;
                    if (gl < chunk_blocks) {
                        unsigned zeros;
                        const unsigned long long r = dense_block(p, sm_nib[gl], sm_zb[gl], G0 + gl, lane, item, sx, sy, sth, T,
                                                                 nmax, zeros);
                        if (r != ~0ull) atomicMin(&sm_best, r);
                    }
                    __syncthreads();                                   // the blocks of this round are complete:
                    found = sm_best != ~0ull;                          // a valid attempt now is the lowest one
                }
            }
            if (found && warp == 0) retry_commit(p, i, item, sx, sy, sth, sm_best, lane);
            __syncthreads();
        }
    }
    MotionWarpScratch<Q> &ws = scratch[warp];
    while (n_sparse) {
        unsigned e = 0;
        if (lane == 0) e = atomicAdd(p.retry_ticket + 1, 1u);
        e = __shfl_sync(0xffffffffu, e, 0);
        if (e >= n_sparse) break;
        const int64_t i = p.retry_idx[e];
        const unsigned long long T = p.retry_thr[e];
        // the source pose: kernel 1 left it in the output arrays as well, so in-place calls (xo == x) see it too
        const double sx = p.xo[i], sy = p.yo[i], sth = p.tho[i];
        Pose win = {sx, sy, sth};
        const int watt = retry_particle<Q>(p, ws, lane, p.first_index + (uint64_t)i, sx, sy, sth, T, win);
        if (lane == 0 && watt) {
            MOTION_STAT(p, 5, 1);
            p.xo[i] = win.x; p.yo[i] = win.y; p.tho[i] = win.th;
            if (p.attempts) p.attempts[i] = watt;
        }
        __syncwarp();
    }
    // the last CTA to finish re-arms the lists for the next call
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        if (atomicAdd(p.retry_done, 1u) == gridDim.x - 1) {
            p.retry_count[0] = 0u; p.retry_count[1] = 0u; *p.retry_done = 0u; p.retry_ticket[0] = 0u; p.retry_ticket[1] = 0u;
        }
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
    p.min_thr = h->motion_min_thr;
    p.retry_idx = nullptr; p.retry_thr = nullptr; p.retry_count = nullptr; p.retry_done = nullptr; p.retry_ticket = nullptr;
    const bool retry = !d_normals && max_attempts > 1;
    if (retry) {
        if (n > 0x7fffffffll) return mcl_fail(h, MCL_ERR_ARG, "mcl_predict: more than 2^31 - 1 particles per call");
        if (h->retry_cap < n) {                          // retry list: worst case every particle
            MCL_CUDA(h, cudaStreamSynchronize(h->stream));
            cudaFree(h->d_retry_idx); cudaFree(h->d_retry_thr);
            h->d_retry_idx = nullptr; h->d_retry_thr = nullptr; h->retry_cap = 0;
            MCL_CUDA(h, cudaMalloc((void **)&h->d_retry_idx, (size_t)n * sizeof(int32_t)));
            MCL_CUDA(h, cudaMalloc((void **)&h->d_retry_thr, (size_t)n * sizeof(unsigned long long)));
            h->retry_cap = n;
        }
        if (!h->d_retry_ctr) {
            MCL_CUDA(h, cudaMalloc((void **)&h->d_retry_ctr, 8 * sizeof(unsigned)));
            MCL_CUDA(h, cudaMemsetAsync(h->d_retry_ctr, 0, 8 * sizeof(unsigned), h->stream));
        }
        p.retry_idx = h->d_retry_idx; p.retry_thr = h->d_retry_thr;
        p.retry_count = h->d_retry_ctr; p.retry_done = h->d_retry_ctr + 2; p.retry_ticket = h->d_retry_ctr + 4;
    } else if (!d_normals) {
        p.max_attempts = max_attempts > 0 ? 1 : 0;       // nothing to retry: kernel 1 alone (the list is never touched)
        p.n_levels = 0;
    }
    const int blocks = (int)((n + MOTION_BLOCK - 1) / MOTION_BLOCK);
    k_motion<<<blocks, MOTION_BLOCK, 0, h->stream>>>(p);
    MCL_LAUNCH_CHECK(h);
    if (retry) {
        // one warp per entry at most: small clouds (the reference's own 1 k particles) launch a few CTAs only
        const int rblocks = (int)std::min<int64_t>((int64_t)h->sm_count * MOTION_RETRY_BLOCKS_PER_SM,
                                                   (n + MOTION_RETRY_BLOCK / 32 - 1) / (MOTION_RETRY_BLOCK / 32));
        if (h->motion_small_queue) k_motion_retry<64><<<rblocks, MOTION_RETRY_BLOCK, 0, h->stream>>>(p);
        else k_motion_retry<320><<<rblocks, MOTION_RETRY_BLOCK, 0, h->stream>>>(p);
        MCL_LAUNCH_CHECK(h);
    }
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

// Test hook: min_thr raises every screening threshold of the rejection loop below it (a looser screen is still exact:
// 2^28 makes every zero nibble a candidate, 2^32 disables the screen); small_queue = 1 selects the 64-entry queue so
// that rounds which do not fit the queue occur.  (0, 0) restores production behaviour.
int mcl_debug_motion(mcl_handle *h, unsigned long long min_thr, int small_queue) {
    if (!h) return MCL_ERR_ARG;
    if (min_thr > (1ull << 32)) return mcl_fail(h, MCL_ERR_ARG, "mcl_debug_motion: min_thr > 2^32");
    h->motion_min_thr = min_thr;
    h->motion_small_queue = small_queue != 0;
    return MCL_OK;
}
