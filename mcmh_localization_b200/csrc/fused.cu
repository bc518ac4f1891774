// fused.cu -- the tail of a filter step (softmax x2 -> MH accept -> estimate -> systematic resampling) in four
// launches instead of twelve kernels and six memsets.  Same arithmetic, element for element, as the stand-alone
// kernels in mh_softmax.cu (node:351-358 convert_scores, pu:208-236 mh_resampling), estimate.cu (node:586-597)
// and resample.cu (pu:416-446, fixed-point mode); the stand-alone forms stay for the sharded path, the other
// localization modes and the function shim.  At 1 M particles every one of those kernels is latency-bound
// (4-24 MB each, 8-30 us against 1-4 us of HBM time), so the step pays for launches, tails and grid-wide
// "last block" finalisations, not for bytes; fusing removes eight of them.
//
//   likelihood (likelihood.cu)  : also leaves max(score) of both sets as order-preserving keys
//   k_fz_sumexp                 : exact 2^-40 fixed-point sums of exp(s - max), both sets          (8 MB read)
//   k_fz_weights_mh_moments     : weights of both sets, MH accept, new pose + weight, the six raw
//                                 estimate sums and the weight maximum (-> resampling scale)       (60 MB)
//   k_fz_central_scan           : the nine central estimate sums + single-pass (decoupled look-back)
//                                 inclusive scan of the quantised weights                          (36 MB)
//   k_fz_search_gather          : per output slot, binary search in the cumulative sums + pose gather
//
// Order independence: softmax sums, resampling sums are exact integers; the estimate sums are fp64 per-tile
// partials combined in tile order (deterministic, equal to the stand-alone kernels to rounding).
#include <float.h>
#include <math.h>
#include <string.h>

#include <algorithm>

#include "common.cuh"

#define FZ_THREADS 256
#define FZ_ITEMS 4
#define FZ_TILE (FZ_THREADS * FZ_ITEMS)
#define SOFTMAX_FIX 1099511627776.0   // 2^40 (mh_softmax.cu)

struct FzHeader {               // device, zero-initialised once; every counter is reset by its last user
    unsigned long long keymax[2];   // max score per set as order-preserving keys: written by the likelihood kernel
                                    // (sharded: then MAX-exchanged in place), cleared by k_fz_sumexp
    unsigned long long sumq[2];     // sums of the softmax numerators, raw 2^-40 integers (sharded: SUM-exchanged)
    double smax[2];                 // the maxima as floats (copied from the keys before they are cleared)
    double msum[8];                 // six raw estimate sums, weight maximum (sharded: SUM / MAX-exchanged)
    double csum[10];                // nine central estimate sums of this rank; [9]: `total` again (one exchange)
    unsigned long long total;       // total of this rank's quantised weights
    unsigned cnt_sumexp[2];
    unsigned cnt_moments, cnt_central, ticket, pad;
};

struct FzArgs {
    FzHeader *hd;
    int64_t n, n_global;                       // particles on this rank / in the whole population
    int nt;                                    // tiles of FZ_TILE particles
    const float *s_post, *s_pre;
    float *w_post, *w_pre, *w_out;
    const double *px, *py, *pt;                // proposal = particles          (node:363 "particles")
    const double *ox, *oy, *ot;                // particles_prev
    double *nx, *ny, *nth;                     // result of the MH step (spare set)
    int use_mh;
    uint64_t seed, step, first_index;
    unsigned long long *part_q;                // [2][nb2] sumexp partials
    double *part_m;                            // [nt][8]  six raw sums, weight max
    double *part_c;                            // [nt][9]
    unsigned long long *status;                // [nt] look-back descriptors: flag << 62 | value
    unsigned long long *C;                     // [n] inclusive cumulative sums
    double *est18;
    // search + gather
    double r, rstep;
    int32_t *idx;
    double *gx, *gy, *gt;
};

__device__ __forceinline__ float fz_softmax_num(float s, float m) { return (float)exp((double)__fsub_rn(s, m)); }
__device__ __forceinline__ unsigned long long fz_quantise(float w, double scale) {
    const double v = __dmul_rn((double)w, scale);
    return v > 0.0 ? __double2ull_rz(v) : 0ull;     // negative / NaN weights count as 0
}
__device__ __forceinline__ unsigned long long fz_warp_sum_u64(unsigned long long v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// ---------------------------------------------------------------------------------------------
// sums of the softmax numerators (k_sumexp2 of mh_softmax.cu with the maxima taken from the keys)
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(FZ_THREADS) k_fz_sumexp(const FzArgs a) {
    __shared__ unsigned long long shq[FZ_THREADS / 32];
    __shared__ bool last;
    const int y = blockIdx.y;
    const float *__restrict__ s = y ? a.s_pre : a.s_post;
    const unsigned key = (unsigned)((volatile unsigned long long *)a.hd->keymax)[y];
    const float m = key ? mcl_float_of_key(key) : -FLT_MAX;
    unsigned long long acc = 0;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < a.n; i += (int64_t)gridDim.x * blockDim.x)
        acc += __double2ull_rz(__dmul_rn((double)fz_softmax_num(s[i], m), SOFTMAX_FIX));
    acc = fz_warp_sum_u64(acc);
    if ((threadIdx.x & 31) == 0) shq[threadIdx.x >> 5] = acc;
    __syncthreads();
    unsigned long long *part = a.part_q + (size_t)y * gridDim.x;
    if (threadIdx.x == 0) {
        unsigned long long t = 0;
        for (int k = 0; k < FZ_THREADS / 32; ++k) t += shq[k];
        part[blockIdx.x] = t;
        __threadfence();
        last = atomicAdd(&a.hd->cnt_sumexp[y], 1u) == gridDim.x - 1;
    }
    __syncthreads();
    if (last) {
        __threadfence();
        unsigned long long t = 0;
        for (unsigned b = threadIdx.x; b < gridDim.x; b += blockDim.x) t += ((volatile unsigned long long *)part)[b];
        t = fz_warp_sum_u64(t);
        __syncthreads();
        if ((threadIdx.x & 31) == 0) shq[threadIdx.x >> 5] = t;
        __syncthreads();
        if (threadIdx.x == 0) {
            unsigned long long tot = 0;
            for (int k = 0; k < FZ_THREADS / 32; ++k) tot += shq[k];
            a.hd->smax[y] = (double)m;
            a.hd->sumq[y] = tot;
            a.hd->cnt_sumexp[y] = 0;
            a.hd->keymax[y] = 0ull;     // every block of this set has read it
        }
    }
}

template <int K>
__device__ __forceinline__ void fz_block_sum(double (&v)[K], double *sh /* K * 8 */) {
#pragma unroll
    for (int k = 0; k < K; ++k) v[k] = warp_sum(v[k]);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (lane == 0)
#pragma unroll
        for (int k = 0; k < K; ++k) sh[k * 8 + warp] = v[k];
    __syncthreads();
    if (threadIdx.x == 0)
#pragma unroll
        for (int k = 0; k < K; ++k) {
            double t = 0.0;
            for (int w = 0; w < FZ_THREADS / 32; ++w) t += sh[k * 8 + w];
            v[k] = t;
        }
    __syncthreads();
}

// ---------------------------------------------------------------------------------------------
// weights (node:356-357), MH accept (pu:229-233), raw estimate sums (node:586-589), weight maximum
// ---------------------------------------------------------------------------------------------
template <bool MH>
__global__ void __launch_bounds__(FZ_THREADS, 4) k_fz_weights_mh_moments(const FzArgs a) {
    __shared__ double sh[6 * 8];
    __shared__ float shm[FZ_THREADS / 32];
    __shared__ bool last;
    // sum as f32 of the exact integer (node:356-357: f32 divide by the f32-rounded sum)
    const float m_post = (float)a.hd->smax[0], sum_post = (float)((double)a.hd->sumq[0] / SOFTMAX_FIX);
    const float m_pre = MH ? (float)a.hd->smax[1] : 0.f, sum_pre = MH ? (float)((double)a.hd->sumq[1] / SOFTMAX_FIX) : 1.f;
    double v[6] = {0, 0, 0, 0, 0, 0};
    float wmax = 0.0f;
    const int64_t base = (int64_t)blockIdx.x * FZ_TILE + threadIdx.x;
#pragma unroll 2
    for (int k = 0; k < FZ_ITEMS; ++k) {
        const int64_t i = base + (int64_t)k * FZ_THREADS;      // rows of 256 consecutive particles: coalesced
        if (i >= a.n) break;
        const float p_new = __fdiv_rn(fz_softmax_num(a.s_post[i], m_post), sum_post);
        double x = a.px[i], y = a.py[i], t = a.pt[i];
        float w = p_new;
        if (MH) {
            const float p_old = __fdiv_rn(fz_softmax_num(a.s_pre[i], m_pre), sum_pre);
            a.w_post[i] = p_new;
            a.w_pre[i] = p_old;
            double alpha = 1.0;
            if (p_old > 0.f) {
                const double q = (double)__fdiv_rn(p_new, p_old);          // f32 divide, promoted
                alpha = (q < 1.0) ? q : 1.0;
            }
            const uint4 o = philox_draw4(a.seed, a.step, a.first_index + (uint64_t)i, 0u, MCL_STREAM_MH);
            const bool acc = u53_from(o.x, o.y) < alpha;
            if (!acc) { x = a.ox[i]; y = a.oy[i]; t = a.ot[i]; w = p_old; }
            a.nx[i] = x; a.ny[i] = y; a.nth[i] = t;
        }
        a.w_out[i] = w;
        const double wi = (double)w;
        double sn, cs;
        sincos(t, &sn, &cs);
        v[0] += wi; v[1] += wi * wi; v[2] += wi * x; v[3] += wi * y; v[4] += wi * cs; v[5] += wi * sn;
        wmax = fmaxf(wmax, w);
    }
    fz_block_sum<6>(v, sh);
    wmax = warp_max(wmax);
    if ((threadIdx.x & 31) == 0) shm[threadIdx.x >> 5] = wmax;
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int k = 1; k < FZ_THREADS / 32; ++k) wmax = fmaxf(wmax, shm[k]);
        double *pm = a.part_m + (size_t)blockIdx.x * 8;
#pragma unroll
        for (int k = 0; k < 6; ++k) pm[k] = v[k];
        pm[6] = (double)wmax;
        __threadfence();
        last = atomicAdd(&a.hd->cnt_moments, 1u) == gridDim.x - 1;
    }
    __syncthreads();
    if (last) {
        __threadfence();
        // fixed summation tree (thread t takes tiles t, t + 256, ...; then the block sum): deterministic
        double t[6] = {0, 0, 0, 0, 0, 0};
        float wm2 = 0.0f;
        for (unsigned b = threadIdx.x; b < gridDim.x; b += blockDim.x) {
            const double *pm = a.part_m + (size_t)b * 8;
#pragma unroll
            for (int k = 0; k < 6; ++k) t[k] += __ldcg(pm + k);
            wm2 = fmaxf(wm2, (float)__ldcg(pm + 6));
        }
        fz_block_sum<6>(t, sh);
        wm2 = warp_max(wm2);
        if ((threadIdx.x & 31) == 0) shm[threadIdx.x >> 5] = wm2;
        __syncthreads();
        if (threadIdx.x == 0) {
            for (int k = 1; k < FZ_THREADS / 32; ++k) wm2 = fmaxf(wm2, shm[k]);
#pragma unroll
            for (int k = 0; k < 6; ++k) sh[k] = t[k];
            sh[6] = (double)wm2;
        }
        for (int b = threadIdx.x; b < a.nt; b += blockDim.x) a.status[b] = 0ull;     // look-back descriptors of the scan
        __syncthreads();
        if (threadIdx.x == 0) {
#pragma unroll
            for (int k = 0; k < 7; ++k) a.hd->msum[k] = sh[k];      // this rank's sums; means and scale: k_fz_central_scan
            a.hd->cnt_moments = 0;
            a.hd->ticket = 0;
        }
    }
}

// ---------------------------------------------------------------------------------------------
// One iteration of the MH refinement chain (filter.cu, BASELINE config 4): weights of proposal and chain from the
// two softmax statistics, accept (pu:229-233), carried score; leaves max(score_chain) as the key of set 1 for the
// next iteration's k_fz_sumexp (the proposal's key comes from the next likelihood launch).
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(FZ_THREADS) k_fz_chain_accept(const FzArgs a, float *score_chain) {
    __shared__ float shm[FZ_THREADS / 32];
    const float m_prop = (float)a.hd->smax[0], sum_prop = (float)((double)a.hd->sumq[0] / SOFTMAX_FIX);
    const float m_chain = (float)a.hd->smax[1], sum_chain = (float)((double)a.hd->sumq[1] / SOFTMAX_FIX);
    float smax = -FLT_MAX;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < a.n; i += (int64_t)gridDim.x * blockDim.x) {
        const float s_prop = a.s_post[i], s_chain = score_chain[i];
        const float p_new = __fdiv_rn(fz_softmax_num(s_prop, m_prop), sum_prop);
        const float p_old = __fdiv_rn(fz_softmax_num(s_chain, m_chain), sum_chain);
        a.w_post[i] = p_new;
        a.w_pre[i] = p_old;
        double alpha = 1.0;
        if (p_old > 0.f) {
            const double q = (double)__fdiv_rn(p_new, p_old);
            alpha = (q < 1.0) ? q : 1.0;
        }
        const uint4 o = philox_draw4(a.seed, a.step, a.first_index + (uint64_t)i, 0u, MCL_STREAM_MH);
        const bool acc = u53_from(o.x, o.y) < alpha;
        a.nx[i] = acc ? a.px[i] : a.ox[i];
        a.ny[i] = acc ? a.py[i] : a.oy[i];
        a.nth[i] = acc ? a.pt[i] : a.ot[i];
        a.w_out[i] = acc ? p_new : p_old;
        const float s_new = acc ? s_prop : s_chain;
        if (acc) score_chain[i] = s_prop;
        smax = fmaxf(smax, s_new);
    }
    smax = warp_max(smax);
    if ((threadIdx.x & 31) == 0) shm[threadIdx.x >> 5] = smax;
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int k = 1; k < FZ_THREADS / 32; ++k) smax = fmaxf(smax, shm[k]);
        if (smax > -FLT_MAX) atomicMax(&a.hd->keymax[1], (unsigned long long)mcl_key_of_float(smax));
    }
}

// ---------------------------------------------------------------------------------------------
// central sums (node:590-597, pu:69-83) + inclusive scan of the quantised weights (pu:436-443 as integers).
// Single pass: tiles take tickets in launch order, publish their aggregate, and look back over their
// predecessors' descriptors (aggregate or inclusive prefix) -- Merrill & Garland's decoupled look-back.
// ---------------------------------------------------------------------------------------------
#define FZ_FLAG_AGG (1ull << 62)
#define FZ_FLAG_INC (2ull << 62)
#define FZ_VALUE_MASK ((1ull << 62) - 1)

__global__ void __launch_bounds__(FZ_THREADS, 4) k_fz_central_scan(const FzArgs a) {
    __shared__ double sh[9 * 8];
    __shared__ unsigned long long shw[FZ_ITEMS][FZ_THREADS / 32];
    __shared__ unsigned long long sh_excl;
    __shared__ int sh_tile;
    __shared__ bool last;
    __shared__ double sh_par[4];
    if (threadIdx.x == 0) {
        const int tk = (int)atomicAdd(&a.hd->ticket, 1u);
        sh_tile = tk;
        // means (node:586-589) and resampling scale from the population's raw sums and weight maximum
        const double *ms = a.hd->msum;
        const double pmx = ms[2] / ms[0], pmy = ms[3] / ms[0], pmt = atan2(ms[5], ms[4]);   // np.average; arctan2(sin, cos)
        const float wm = (float)ms[6];
        int e = 0;
        if (wm > 0.0f) frexp((double)wm, &e);
        int lg = 0;
        while (((int64_t)1 << lg) < a.n_global) ++lg;
        // scale = 2^(62 - ceil(log2 n) - e), 2^e > wmax   (resample.cu k_wmax, oracle orc_resample_scale)
        sh_par[0] = pmx; sh_par[1] = pmy; sh_par[2] = pmt; sh_par[3] = ldexp(1.0, 62 - lg - e);
        if (tk == 0) {
#pragma unroll
            for (int k = 0; k < 6; ++k) a.est18[k] = ms[k];
            a.est18[6] = pmx; a.est18[7] = pmy; a.est18[8] = pmt;
        }
    }
    __syncthreads();
    const int tile = sh_tile;
    const double scale = sh_par[3];
    const double mx = sh_par[0], my = sh_par[1], mt = sh_par[2];
    const int64_t base = (int64_t)tile * FZ_TILE + threadIdx.x;     // rows of 256 consecutive particles
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    unsigned long long inc[FZ_ITEMS];           // inclusive scan of the row inside this warp
    double v[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
#pragma unroll
    for (int k = 0; k < FZ_ITEMS; ++k) {
        const int64_t i = base + (int64_t)k * FZ_THREADS;
        unsigned long long q = 0;
        if (i < a.n) {
            const float w = a.w_out[i];
            q = fz_quantise(w, scale);
            const double wi = (double)w;
            const double dx = a.nx[i] - mx, dy = a.ny[i] - my;
            const double dt = (double)(float)normalize_angle_dev(__dadd_rn(a.nth[i], -mt));   // pu:80-82
            v[0] += wi * dx; v[1] += wi * dy; v[2] += wi * dt;
            v[3] += wi * dx * dx; v[4] += wi * dx * dy; v[5] += wi * dx * dt;
            v[6] += wi * dy * dy; v[7] += wi * dy * dt; v[8] += wi * dt * dt;
        }
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const unsigned long long t = __shfl_up_sync(0xffffffffu, q, o);
            if (lane >= o) q += t;
        }
        inc[k] = q;
        if (lane == 31) shw[k][warp] = q;
    }
    __syncthreads();
    // offset of (row k, this warp) inside the tile, and the tile aggregate
    unsigned long long off[FZ_ITEMS], agg = 0;
#pragma unroll
    for (int k = 0; k < FZ_ITEMS; ++k) {
        off[k] = agg;
#pragma unroll
        for (int w = 0; w < FZ_THREADS / 32; ++w) {
            const unsigned long long t = shw[k][w];
            if (w < warp) off[k] += t;
            agg += t;
        }
    }
    // publish + look back (warp 0; a block-wide look-back, 256 descriptors per round, measured slower)
    if (warp == 0) {
        volatile unsigned long long *st = a.status;
        if (lane == 0) st[tile] = (tile == 0 ? FZ_FLAG_INC : FZ_FLAG_AGG) | agg;
        unsigned long long excl = 0;
        int t0 = tile - 1;                      // lanes inspect tiles t0, t0-1, ..., t0-31
        while (t0 >= 0) {
            const int t = t0 - lane;
            unsigned long long d = FZ_FLAG_INC;  // tiles before 0: empty inclusive prefix
            if (t >= 0) { do { d = st[t]; } while ((d >> 62) == 0); }
            const unsigned incl = __ballot_sync(0xffffffffu, (d >> 62) == 2);
            const int stop = incl ? __ffs(incl) - 1 : 32;       // nearest tile with an inclusive prefix
            unsigned long long part = lane <= stop ? (d & FZ_VALUE_MASK) : 0ull;
            part = fz_warp_sum_u64(part);
            excl += part;
            if (incl) break;
            t0 -= 32;
        }
        if (lane == 0) {
            if (tile > 0) st[tile] = FZ_FLAG_INC | (excl + agg);
            sh_excl = excl;
            if (tile == a.nt - 1) { a.hd->total = excl + agg; ((unsigned long long *)a.hd->csum)[9] = excl + agg; }
        }
    }
    fz_block_sum<9>(v, sh);                     // (contains the barriers that publish sh_excl)
    const unsigned long long excl = sh_excl;
#pragma unroll
    for (int k = 0; k < FZ_ITEMS; ++k) {
        const int64_t i = base + (int64_t)k * FZ_THREADS;
        if (i < a.n) a.C[i] = excl + off[k] + inc[k];
    }
    if (threadIdx.x == 0) {
        double *pc = a.part_c + (size_t)tile * 9;
#pragma unroll
        for (int k = 0; k < 9; ++k) pc[k] = v[k];
        __threadfence();
        last = atomicAdd(&a.hd->cnt_central, 1u) == gridDim.x - 1;
    }
    __syncthreads();
    if (last) {
        __threadfence();
        double t[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
        for (int b = threadIdx.x; b < a.nt; b += blockDim.x) {
            const double *pc = a.part_c + (size_t)b * 9;
#pragma unroll
            for (int k = 0; k < 9; ++k) t[k] += __ldcg(pc + k);
        }
        fz_block_sum<9>(t, sh);
        if (threadIdx.x == 0)
#pragma unroll
            for (int k = 0; k < 9; ++k) { a.est18[9 + k] = t[k]; a.hd->csum[k] = t[k]; }   // sharded: csum is SUM-exchanged into est18
        if (threadIdx.x == 0) a.hd->cnt_central = 0;
    }
}

// ---------------------------------------------------------------------------------------------
// idx[m] = min(first i with C_i >= T_m, n - 1), T_m = ceil((r + m step) total)  (k_search_fixed) + pu:445 gather
// ---------------------------------------------------------------------------------------------
#define FZ_SEARCH_SMEM_TILES 6144      // 48 KB of tile prefixes
__global__ void __launch_bounds__(256) k_fz_search_gather(const FzArgs a) {
    // two levels: the tiles' inclusive prefixes (left in the look-back descriptors) are searched in shared
    // memory, then 11 steps inside one tile of C -- half the dependent L2 round trips of a flat search
    extern __shared__ unsigned long long tp[];
    const bool coarse = a.nt <= FZ_SEARCH_SMEM_TILES;
    if (coarse) {
        for (int t = threadIdx.x; t < a.nt; t += blockDim.x) tp[t] = __ldcg(a.status + t) & FZ_VALUE_MASK;
        __syncthreads();
    }
    const double totd = (double)a.hd->total;
    const unsigned long long *__restrict__ C = a.C;
    for (int64_t j = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; j < a.n; j += (int64_t)gridDim.x * blockDim.x) {
        const double U = __dadd_rn(a.r, __dmul_rn((double)j, a.rstep));
        const double t = ceil(__dmul_rn(U, totd));
        const unsigned long long T = t >= 18446744073709551616.0 ? 0xffffffffffffffffull : (t > 0.0 ? __double2ull_rz(t) : 0ull);
        int64_t lo = 0, hi = a.n - 1;
        if (coarse) {
            int tl = 0, th = a.nt - 1;                 // first tile whose inclusive prefix reaches T (or the last tile)
            while (tl < th) {
                const int mid = (tl + th) >> 1;
                if (T > tp[mid]) tl = mid + 1; else th = mid;
            }
            lo = (int64_t)tl * FZ_TILE;
            hi = min(lo + FZ_TILE - 1, a.n - 1);
        }
        while (lo < hi) {
            const int64_t mid = (lo + hi) >> 1;
            if (T > C[mid]) lo = mid + 1; else hi = mid;
        }
        a.idx[j] = (int32_t)lo;
        a.gx[j] = a.nx[lo]; a.gy[j] = a.ny[lo]; a.gt[j] = a.nth[lo];
    }
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
static size_t align256(size_t v) { return (v + 255) & ~(size_t)255; }

// FzPlan: where everything lives inside h->d_fused
struct FzPlan {
    int nt, nb2;
    size_t o_q, o_m, o_c, o_st, o_C, bytes;
};
static FzPlan fz_plan(const mcl_handle *h, int64_t n) {
    FzPlan p;
    p.nt = (int)((n + FZ_TILE - 1) / FZ_TILE);
    p.nb2 = (int)std::max<int64_t>(1, std::min<int64_t>((n + FZ_THREADS * 4 - 1) / (FZ_THREADS * 4), (int64_t)h->sm_count * 8));
    size_t off = align256(sizeof(FzHeader));
    p.o_q = off; off += align256((size_t)2 * p.nb2 * 8);
    p.o_m = off; off += align256((size_t)p.nt * 8 * 8);
    p.o_c = off; off += align256((size_t)p.nt * 9 * 8);
    p.o_st = off; off += align256((size_t)p.nt * 8);
    p.o_C = off; off += align256((size_t)n * 8);
    p.bytes = off;
    return p;
}

int mcl_fused_prepare(mcl_handle *h, int64_t n) {
    if (n > 0x7fffffffLL) return mcl_fail(h, MCL_ERR_CAPACITY, "fused step: n exceeds int32 indices");
    const FzPlan p = fz_plan(h, n);
    if (p.bytes <= h->fused_bytes) return MCL_OK;
    MCL_CUDA(h, cudaStreamSynchronize(h->stream));
    cudaFree(h->d_fused);
    h->d_fused = nullptr; h->fused_bytes = 0;
    MCL_CUDA(h, cudaMalloc(&h->d_fused, p.bytes));
    MCL_CUDA(h, cudaMemset(h->d_fused, 0, align256(sizeof(FzHeader))));
    h->fused_bytes = p.bytes;
    return MCL_OK;
}

unsigned long long *mcl_fused_keymax(mcl_handle *h) { return reinterpret_cast<FzHeader *>(h->d_fused)->keymax; }

// device addresses of the quantities a sharded run exchanges between the stages (payload == result, in place)
void mcl_fused_exchange_ptrs(mcl_handle *h, FusedPtrs *out) {
    FzHeader *hd = reinterpret_cast<FzHeader *>(h->d_fused);
    out->keymax = hd->keymax; out->sumq = hd->sumq; out->msum = hd->msum; out->csum = hd->csum; out->total = &hd->total;
}
const unsigned long long *mcl_fused_cumsum(mcl_handle *h, int64_t n) {
    return reinterpret_cast<const unsigned long long *>((char *)h->d_fused + fz_plan(h, n).o_C);
}

// the look-back descriptors hold every tile's inclusive prefix after mcl_fused_scan (flag bits in the top two)
const unsigned long long *mcl_fused_tile_prefix(mcl_handle *h, int64_t n, int *nt, int *tile) {
    const FzPlan p = fz_plan(h, n);
    *nt = p.nt; *tile = FZ_TILE;
    return reinterpret_cast<const unsigned long long *>((char *)h->d_fused + p.o_st);
}

static void fz_fill(mcl_handle *h, const FusedStep &u, FzArgs &a) {
    const FzPlan p = fz_plan(h, u.n);
    char *b = (char *)h->d_fused;
    memset(&a, 0, sizeof(a));
    a.hd = (FzHeader *)b; a.n = u.n; a.n_global = u.n_global; a.nt = p.nt;
    a.s_post = u.s_post; a.s_pre = u.s_pre; a.w_post = u.w_post; a.w_pre = u.w_pre; a.w_out = u.w_out;
    a.px = u.px; a.py = u.py; a.pt = u.pt; a.ox = u.ox; a.oy = u.oy; a.ot = u.ot; a.nx = u.nx; a.ny = u.ny; a.nth = u.nth;
    a.use_mh = u.use_mh; a.seed = u.seed; a.step = u.step; a.first_index = u.first_index;
    a.part_q = (unsigned long long *)(b + p.o_q); a.part_m = (double *)(b + p.o_m); a.part_c = (double *)(b + p.o_c);
    a.status = (unsigned long long *)(b + p.o_st); a.C = (unsigned long long *)(b + p.o_C);
    a.est18 = u.est18;
}

// stage 1: sums of the softmax numerators (needs the population's score maxima in keymax)
int mcl_fused_sumexp(mcl_handle *h, const FusedStep &u) {
    FzArgs a;
    fz_fill(h, u, a);
    k_fz_sumexp<<<dim3(fz_plan(h, u.n).nb2, u.use_mh ? 2 : 1), FZ_THREADS, 0, h->stream>>>(a);
    MCL_LAUNCH_CHECK(h);
    return MCL_OK;
}
// stage 2: weights, MH accept into (nx, ny, nth), raw estimate sums + weight maximum of this rank (needs the
// population's sums in sumq).  Without MH the post weights go to w_out and (nx, ny, nth) are the particles.
int mcl_fused_weights(mcl_handle *h, const FusedStep &u) {
    FzArgs a;
    fz_fill(h, u, a);
    if (u.use_mh) k_fz_weights_mh_moments<true><<<a.nt, FZ_THREADS, 0, h->stream>>>(a);
    else k_fz_weights_mh_moments<false><<<a.nt, FZ_THREADS, 0, h->stream>>>(a);
    MCL_LAUNCH_CHECK(h);
    return MCL_OK;
}
// stage 3: means + scale from the population's msum, central sums and cumulative quantised weights of this rank
int mcl_fused_scan(mcl_handle *h, const FusedStep &u) {
    FzArgs a;
    fz_fill(h, u, a);
    k_fz_central_scan<<<a.nt, FZ_THREADS, 0, h->stream>>>(a);
    MCL_LAUNCH_CHECK(h);
    return MCL_OK;
}

// systematic resampling from the cumulative sums left by mcl_fused_scan: (nx, ny, nth) -> (gx, gy, gt)
int mcl_fused_resample(mcl_handle *h, int64_t n, double r, const double *nx, const double *ny, const double *nth,
                       int32_t *idx, double *gx, double *gy, double *gt) {
    const FzPlan p = fz_plan(h, n);
    char *b = (char *)h->d_fused;
    FzArgs a;
    memset(&a, 0, sizeof(a));
    a.hd = (FzHeader *)b; a.n = n; a.nt = p.nt;
    a.C = (unsigned long long *)(b + p.o_C); a.status = (unsigned long long *)(b + p.o_st);
    a.nx = const_cast<double *>(nx); a.ny = const_cast<double *>(ny); a.nth = const_cast<double *>(nth);
    a.r = r; a.rstep = 1.0 / (double)n;                       // pu:434
    a.idx = idx; a.gx = gx; a.gy = gy; a.gt = gt;
    const int blocks = (int)std::min<int64_t>((n + 255) / 256, (int64_t)h->sm_count * 16);
    k_fz_search_gather<<<blocks, 256, p.nt <= FZ_SEARCH_SMEM_TILES ? (size_t)p.nt * 8 : 0, h->stream>>>(a);
    MCL_LAUNCH_CHECK(h);
    return MCL_OK;
}

// MH chain iteration: (ox, oy, ot) chain, (px, py, pt) proposal -> (nx, ny, nth); may alias the chain (in place)
int mcl_fused_chain_accept(mcl_handle *h, const FusedStep &u, float *score_chain) {
    FzArgs a;
    fz_fill(h, u, a);
    const int blocks = (int)std::min<int64_t>((u.n + FZ_THREADS - 1) / FZ_THREADS, (int64_t)h->sm_count * 16);
    k_fz_chain_accept<<<blocks, FZ_THREADS, 0, h->stream>>>(a, score_chain);
    MCL_LAUNCH_CHECK(h);
    return MCL_OK;
}
