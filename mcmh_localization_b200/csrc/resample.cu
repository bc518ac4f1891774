// resample.cu -- kernel (4): systematic / low-variance resampling, pu:416-446
// low_variance_resample_numba, as cumulative sum + vectorised search + SoA gather.
//
// The reference walks  "while U_m > c and i < N-1: i++; c += w[i]"  with U_m = r + m/N (f64) and c
// the running sum of the normalised f32 weights.  Because c_i is non-decreasing the walk is
//     idx[m] = min( first i with c_i >= U_m , N-1 )
// i.e. an independent search per output once the cumulative sums exist.
//   REFERENCE_F32 : c_i reproduces the reference's rounding sequence exactly -- sequential f32
//                   normalising sum (numba np.sum), f32 divide, sequential f32 running sum.  The
//                   sequence is produced by ONE warp (loads and divides 32-wide, the additions
//                   replayed in order through shuffles); the search and gather stay parallel.
//   FIXED_POINT   : weights quantised to 64-bit fixed point, q_i = trunc(w_i * 2^k); the cumulative
//                   sum is exact integer arithmetic (associative => any block or rank decomposition
//                   gives identical indices) computed by a 3-kernel block scan.
#include <stdlib.h>

#include <algorithm>
#include <float.h>

#include "common.cuh"

// ----------------------------------------------------------------------------- reference mode
// sequential f32 sum of w[0..n): s = ((0 + w0) + w1) + ...   (all lanes redundantly, in order)
__device__ float seq_sum_f32_warp(const float *__restrict__ w, int64_t n) {
    const int lane = threadIdx.x & 31;
    float s = 0.0f;
    for (int64_t base = 0; base < n; base += 32) {
        const float v = (base + lane < n) ? w[base + lane] : 0.0f;   // s + 0.0f == s
#pragma unroll
        for (int k = 0; k < 32; ++k) s = __fadd_rn(s, __shfl_sync(0xffffffffu, v, k));
    }
    return s;
}

__global__ void __launch_bounds__(32) k_cumsum_ref_f32(const float *__restrict__ w, int64_t n, float *__restrict__ c_out,
                                                       int normalise) {
    const int lane = threadIdx.x & 31;
    const float sum = normalise ? seq_sum_f32_warp(w, n) : 1.0f;   // pu:430  np.sum(weights); x / 1.0f == x
    float c = 0.0f;                                           // 0 + wn[0] == wn[0]  (pu:436)
    for (int64_t base = 0; base < n; base += 32) {
        const bool in = base + lane < n;
        const float wn = in ? __fdiv_rn(w[base + lane], sum) : 0.0f;   // pu:430 weights / sum (f32)
        float mine = 0.0f;
#pragma unroll
        for (int k = 0; k < 32; ++k) {
            c = __fadd_rn(c, __shfl_sync(0xffffffffu, wn, k));         // pu:443 c += weights[i]
            if (lane == k) mine = c;
        }
        if (in) c_out[base + lane] = mine;
    }
}

// ---------------------------------------------------------------------------------------------
// Exact PARALLEL emulation of the sequential f32 accumulation  c_i = fl32(c_{i-1} + w_i)  (w_i >= 0).
//
// While c stays inside one binade [2^e, 2^(e+1)) it is an integer multiple K of u = 2^(e-23), and adding w
// is integer arithmetic: with w / u = a0 + f (a0 integer, 0 <= f < 1), round-to-nearest-even gives
//     K' = K + a0            if f < 1/2
//     K' = K + a0 + 1        if f > 1/2
//     K' = K + a0 + ((K + a0) & 1)   if f == 1/2            (ties to even)
// i.e. every element is a map  K -> K + A(K & 1)  given by the pair (A(0), A(1)); such maps compose
// associatively (the parity after one map selects the branch of the next), so a block scan over the pairs
// reproduces the sequential rounding sequence exactly.  The scan is valid up to the first element that pushes
// K to 2^24 (the sum leaves the binade); that one addition is done with a real f32 add and the scan restarts
// from there with the new unit.  One CTA streams over the array tile by tile (the sum is monotone, so there
// are at most ~25-150 restarts in total), replacing 1M dependent additions by ~130 tile scans.
// ---------------------------------------------------------------------------------------------
#define SEQ_THREADS 1024
#define SEQ_ITEMS 8
#define SEQ_TILE (SEQ_THREADS * SEQ_ITEMS)

#include "seqsum.cuh"

// w_eff[i] = divide ? w[i] / *div : w[i];  c_out (nullable) receives every partial sum, *total_out the last one
__global__ void __launch_bounds__(SEQ_THREADS) k_seq_accumulate_exact(const float *__restrict__ w, int64_t n,
                                                                      const float *div, float *__restrict__ c_out,
                                                                      float *total_out) {
    __shared__ Pair64 warp_tot[32];
    __shared__ float s_c;
    __shared__ long long s_cross;
    __shared__ int64_t s_seg0;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const float dv = div ? div[0] : 1.0f;
    if (threadIdx.x == 0) { s_c = 0.0f; s_seg0 = 0; }
    __syncthreads();
    for (int64_t base = 0; base < n; base += SEQ_TILE) {
        float wv[SEQ_ITEMS];
        const int64_t first = base + (int64_t)threadIdx.x * SEQ_ITEMS;
#pragma unroll
        for (int k = 0; k < SEQ_ITEMS; ++k) {
            const int64_t i = first + k;
            float v = i < n ? w[i] : 0.0f;
            if (div && i < n) v = __fdiv_rn(v, dv);
            wv[k] = v;
        }
        while (true) {
            const float c0 = s_c;
            const int64_t seg0 = s_seg0;            // first index of this tile not yet finalised
            const int e = seq_exponent(c0);
            const long long K0 = seq_K(c0);
            // thread-local inclusive prefixes of the element maps (identity for finished / out-of-range slots)
            Pair64 loc[SEQ_ITEMS];
            Pair64 run; run.a0 = 0; run.a1 = 0;
#pragma unroll
            for (int k = 0; k < SEQ_ITEMS; ++k) {
                const int64_t i = first + k;
                Pair64 m; m.a0 = 0; m.a1 = 0;
                if (i >= seg0 && i < n) m = seq_decode(wv[k], e);
                run = pair_compose(run, m);
                loc[k] = run;
            }
            // exclusive scan of the per-thread totals across the block
            Pair64 inc = run;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const Pair64 t = pair_shfl_up(inc, o);
                if (lane >= o) inc = pair_compose(t, inc);
            }
            if (lane == 31) warp_tot[warp] = inc;
            __syncthreads();
            if (warp == 0) {
                Pair64 t = warp_tot[lane];
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const Pair64 u2 = pair_shfl_up(t, o);
                    if (lane >= o) t = pair_compose(u2, t);
                }
                warp_tot[lane] = t;               // inclusive over warps
            }
            if (threadIdx.x == 0) s_cross = 0x7fffffffffffffffll;
            __syncthreads();
            Pair64 excl = pair_shfl_up(inc, 1);
            if (lane == 0) { excl.a0 = 0; excl.a1 = 0; }
            if (warp > 0) excl = pair_compose(warp_tot[warp - 1], excl);
            const int p0 = (int)(K0 & 1);
            long long Kprev = K0 + (p0 ? excl.a1 : excl.a0);      // K just before this thread's first element
            long long Ks[SEQ_ITEMS];
            long long my_cross = 0x7fffffffffffffffll;
#pragma unroll
            for (int k = 0; k < SEQ_ITEMS; ++k) {
                const Pair64 q = pair_compose(excl, loc[k]);
                Ks[k] = K0 + (p0 ? q.a1 : q.a0);
                const int64_t i = first + k;
                if (i >= seg0 && i < n && Ks[k] >= (1ll << 24) && my_cross == 0x7fffffffffffffffll) my_cross = i;
            }
            if (my_cross != 0x7fffffffffffffffll) atomicMin((unsigned long long *)&s_cross, (unsigned long long)my_cross);
            __syncthreads();
            const long long cross = s_cross;
            const int64_t end = cross == 0x7fffffffffffffffll ? (base + SEQ_TILE < n ? base + SEQ_TILE : n) : (int64_t)cross;
            // finalise [seg0, end)
            float last_val = c0;
            bool have_last = false;
#pragma unroll
            for (int k = 0; k < SEQ_ITEMS; ++k) {
                const int64_t i = first + k;
                if (i >= seg0 && i < end) {
                    const float v = seq_value(Ks[k], e);
                    if (c_out) c_out[i] = v;
                    if (i == end - 1) { last_val = v; have_last = true; }
                }
            }
            __syncthreads();
            if (cross == 0x7fffffffffffffffll) {
                if (have_last) s_c = last_val;            // exactly one thread owns index end-1 (if the segment is non-empty)
                if (threadIdx.x == 0) s_seg0 = end;
                __syncthreads();
                break;                                    // next tile
            }
            // the addition at index `cross` leaves the binade: do it for real and restart behind it
#pragma unroll
            for (int k = 0; k < SEQ_ITEMS; ++k) {
                const int64_t i = first + k;
                if (i == cross) {
                    const long long Kb = k == 0 ? Kprev : Ks[k - 1];
                    const float cb = (cross == seg0) ? c0 : seq_value(Kb, e);
                    const float cn = __fadd_rn(cb, wv[k]);
                    if (c_out) c_out[i] = cn;
                    s_c = cn;
                    s_seg0 = cross + 1;
                }
            }
            __syncthreads();
            if (s_seg0 >= (base + SEQ_TILE < n ? base + SEQ_TILE : n)) break;
        }
    }
    if (threadIdx.x == 0 && total_out) total_out[0] = s_c;
}

// ---------------------------------------------------------------------------------------------
// Multi-CTA version of the exact emulation (large n).  The element maps depend on the binade of the running
// sum, which a tile only knows once every earlier tile is done -- but almost every tile lies entirely inside
// the binade an APPROXIMATE (fp64) prefix predicts.  So:
//   A  per-tile fp64 sums                      (parallel)        B  their prefix -> speculated exponent per tile
//   C  per-tile aggregate map under that exponent (parallel; a saturated map = "the sum leaves the binade here")
//   D  one CTA walks the tiles in order: where the exact incoming sum has the speculated exponent and the
//      aggregate does not leave the binade, the outgoing sum is one integer add (thread 0, no barrier);
//      otherwise the CTA scans that tile exactly (with binade restarts).  Records the exact incoming sum of
//      every tile.  ~25-40 slow tiles per pass.
//   E  every tile, now knowing its exact incoming sum, writes its running sums (parallel, with restarts).
// ---------------------------------------------------------------------------------------------
#define MT_THREADS 256
#define MT_ITEMS 8
#define MT_TILE (MT_THREADS * MT_ITEMS)

struct MtShared {
    Pair64 warp_tot[MT_THREADS / 32];
    float c;
    long long cross;
    int64_t seg0;
};

__device__ __forceinline__ void mt_load(const float *__restrict__ w, int64_t n, const float *div, int64_t base,
                                        float (&wv)[MT_ITEMS]) {
    const float dv = div ? div[0] : 1.0f;
    const int64_t first = base + (int64_t)threadIdx.x * MT_ITEMS;
#pragma unroll
    for (int k = 0; k < MT_ITEMS; ++k) {
        const int64_t i = first + k;
        float v = i < n ? w[i] : 0.0f;
        if (div && i < n) v = __fdiv_rn(v, dv);
        wv[k] = v;
    }
}

// block-wide exclusive scan of per-thread maps; also returns the block aggregate
__device__ __forceinline__ Pair64 mt_block_exclusive(const Pair64 &mine, MtShared &sh, Pair64 &aggregate) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    Pair64 inc = mine;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const Pair64 t = pair_shfl_up(inc, o);
        if (lane >= o) inc = pair_compose(t, inc);
    }
    if (lane == 31) sh.warp_tot[warp] = inc;
    __syncthreads();
    Pair64 excl = pair_shfl_up(inc, 1);
    if (lane == 0) { excl.a0 = 0; excl.a1 = 0; }
    Pair64 pre; pre.a0 = 0; pre.a1 = 0;
    Pair64 tot; tot.a0 = 0; tot.a1 = 0;
#pragma unroll
    for (int q = 0; q < MT_THREADS / 32; ++q) {
        if (q == warp) pre = tot;
        tot = pair_compose(tot, sh.warp_tot[q]);
    }
    aggregate = tot;
    __syncthreads();
    return pair_compose(pre, excl);
}

// exact running sums of one tile starting from c_start (binade restarts inside); returns the outgoing sum
template <bool WRITE>
__device__ float mt_tile_exact(const float (&wv)[MT_ITEMS], int64_t base, int64_t n, float c_start, float *c_out,
                               MtShared &sh) {
    const int64_t first = base + (int64_t)threadIdx.x * MT_ITEMS;
    const int64_t tile_end = base + MT_TILE < n ? base + MT_TILE : n;
    if (threadIdx.x == 0) { sh.c = c_start; sh.seg0 = base; }
    __syncthreads();
    while (true) {
        const float c0 = sh.c;
        const int64_t seg0 = sh.seg0;
        const int e = seq_exponent(c0);
        const long long K0 = seq_K(c0);
        Pair64 loc[MT_ITEMS];
        Pair64 run; run.a0 = 0; run.a1 = 0;
#pragma unroll
        for (int k = 0; k < MT_ITEMS; ++k) {
            const int64_t i = first + k;
            Pair64 m; m.a0 = 0; m.a1 = 0;
            if (i >= seg0 && i < n) m = seq_decode(wv[k], e);
            run = pair_compose(run, m);
            loc[k] = run;
        }
        if (threadIdx.x == 0) sh.cross = 0x7fffffffffffffffll;
        Pair64 agg;
        const Pair64 excl = mt_block_exclusive(run, sh, agg);
        const int p0 = (int)(K0 & 1);
        const long long Kprev = K0 + (p0 ? excl.a1 : excl.a0);
        long long Ks[MT_ITEMS];
        long long my_cross = 0x7fffffffffffffffll;
#pragma unroll
        for (int k = 0; k < MT_ITEMS; ++k) {
            const Pair64 q = pair_compose(excl, loc[k]);
            Ks[k] = K0 + (p0 ? q.a1 : q.a0);
            const int64_t i = first + k;
            if (i >= seg0 && i < n && Ks[k] >= (1ll << 24) && my_cross == 0x7fffffffffffffffll) my_cross = i;
        }
        if (my_cross != 0x7fffffffffffffffll) atomicMin((unsigned long long *)&sh.cross, (unsigned long long)my_cross);
        __syncthreads();
        const long long cross = sh.cross;
        const int64_t end = cross == 0x7fffffffffffffffll ? tile_end : (int64_t)cross;
        float last_val = c0;
        bool have_last = false;
#pragma unroll
        for (int k = 0; k < MT_ITEMS; ++k) {
            const int64_t i = first + k;
            if (i >= seg0 && i < end) {
                const float v = seq_value(Ks[k], e);
                if (WRITE) c_out[i] = v;
                if (i == end - 1) { last_val = v; have_last = true; }
            }
        }
        __syncthreads();
        if (cross == 0x7fffffffffffffffll) {
            if (have_last) sh.c = last_val;
            __syncthreads();
            break;
        }
#pragma unroll
        for (int k = 0; k < MT_ITEMS; ++k) {
            const int64_t i = first + k;
            if (i == cross) {
                const long long Kb = k == 0 ? Kprev : Ks[k - 1];
                const float cb = (cross == seg0) ? c0 : seq_value(Kb, e);
                const float cn = __fadd_rn(cb, wv[k]);
                if (WRITE) c_out[i] = cn;
                sh.c = cn;
                sh.seg0 = cross + 1;
            }
        }
        __syncthreads();
        if (sh.seg0 >= tile_end) break;
    }
    const float r = sh.c;
    __syncthreads();
    return r;
}

__global__ void __launch_bounds__(MT_THREADS) k_mt_tilesum(const float *__restrict__ w, int64_t n, const float *div,
                                                           double *tsum) {
    __shared__ double shd[MT_THREADS / 32];
    float wv[MT_ITEMS];
    mt_load(w, n, div, (int64_t)blockIdx.x * MT_TILE, wv);
    double s = 0;
#pragma unroll
    for (int k = 0; k < MT_ITEMS; ++k) s += (double)wv[k];
    s = warp_sum(s);
    if ((threadIdx.x & 31) == 0) shd[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0;
        for (int q = 0; q < MT_THREADS / 32; ++q) t += shd[q];
        tsum[blockIdx.x] = t;
    }
}

// exclusive fp64 prefix of the tile sums -> speculated unit exponent of the sum entering each tile
__global__ void __launch_bounds__(1024) k_mt_speculate(const double *tsum, int64_t nt, int *e_spec) {
    __shared__ double sh[32];
    __shared__ double carry;
    if (threadIdx.x == 0) carry = 0.0;
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int64_t base = 0; base < nt; base += blockDim.x) {
        const int64_t i = base + threadIdx.x;
        const double v = i < nt ? tsum[i] : 0.0;
        double inc = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const double t = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += t; }
        if (lane == 31) sh[warp] = inc;
        __syncthreads();
        double off = 0;
        for (int q = 0; q < warp; ++q) off += sh[q];
        const double excl = carry + off + inc - v;
        if (i < nt) e_spec[i] = seq_exponent((float)excl);
        __syncthreads();
        if (threadIdx.x == blockDim.x - 1) carry = carry + off + inc;
        __syncthreads();
    }
}

__global__ void __launch_bounds__(MT_THREADS) k_mt_aggregate(const float *__restrict__ w, int64_t n, const float *div,
                                                             const int *e_spec, uint2 *agg) {
    __shared__ MtShared sh;
    float wv[MT_ITEMS];
    const int64_t base = (int64_t)blockIdx.x * MT_TILE;
    mt_load(w, n, div, base, wv);
    const int e = e_spec[blockIdx.x];
    Pair64 run; run.a0 = 0; run.a1 = 0;
    const int64_t first = base + (int64_t)threadIdx.x * MT_ITEMS;
#pragma unroll
    for (int k = 0; k < MT_ITEMS; ++k) {
        Pair64 m; m.a0 = 0; m.a1 = 0;
        if (first + k < n) m = seq_decode(wv[k], e);
        run = pair_compose(run, m);
    }
    Pair64 a;
    mt_block_exclusive(run, sh, a);
    if (threadIdx.x == 0) agg[blockIdx.x] = make_uint2(a.a0, a.a1);
}

__global__ void __launch_bounds__(MT_THREADS) k_mt_walk(const float *__restrict__ w, int64_t n, const float *div,
                                                        const int *__restrict__ e_spec, const uint2 *__restrict__ agg,
                                                        int64_t nt, float *c_in, float *total_out) {
    __shared__ MtShared sh;
    __shared__ float s_cur;
    __shared__ int64_t s_tile;
    if (threadIdx.x == 0) { s_cur = 0.0f; s_tile = 0; }
    __syncthreads();
    while (true) {
        if (threadIdx.x == 0) {
            // fast walk: one integer add per tile while the speculation holds
            float c = s_cur;
            int64_t j = s_tile;
            while (j < nt) {
                const int e = seq_exponent(c);
                const uint2 a = agg[j];
                const long long K = seq_K(c) + ((seq_K(c) & 1) ? a.y : a.x);
                if (e != e_spec[j] || K >= (1ll << 24)) break;        // mis-speculated or leaves the binade: slow tile
                c_in[j] = c;
                c = seq_value(K, e);
                ++j;
            }
            s_cur = c; s_tile = j;
            if (j < nt) c_in[j] = c;
        }
        __syncthreads();
        const int64_t j = s_tile;
        if (j >= nt) break;
        float wv[MT_ITEMS];
        mt_load(w, n, div, j * MT_TILE, wv);
        const float c = mt_tile_exact<false>(wv, j * MT_TILE, n, s_cur, nullptr, sh);
        if (threadIdx.x == 0) { s_cur = c; s_tile = j + 1; }
        __syncthreads();
    }
    if (threadIdx.x == 0 && total_out) total_out[0] = s_cur;
}

__global__ void __launch_bounds__(MT_THREADS) k_mt_write(const float *__restrict__ w, int64_t n, const float *div,
                                                         const float *__restrict__ c_in, float *__restrict__ c_out) {
    __shared__ MtShared sh;
    float wv[MT_ITEMS];
    const int64_t base = (int64_t)blockIdx.x * MT_TILE;
    mt_load(w, n, div, base, wv);
    mt_tile_exact<true>(wv, base, n, c_in[blockIdx.x], c_out, sh);
}

// exact sequential-f32 accumulation of w (or of w / *d_div): running sums to d_c (nullable), total to d_total
// (nullable).  Small inputs: one streaming CTA; large inputs: the multi-CTA pipeline above.
static int seq_accumulate(mcl_handle *h, const float *d_w, int64_t n, const float *d_div, float *d_c, float *d_total) {
    static int force_single = -1;
    if (force_single < 0) { const char *e = getenv("MCL_SEQ_SINGLE_CTA"); force_single = (e && atoi(e)) ? 1 : 0; }
    if (n <= 4 * MT_TILE || force_single) {
        k_seq_accumulate_exact<<<1, SEQ_THREADS, 0, h->stream>>>(d_w, n, d_div, d_c, d_total);
        MCL_LAUNCH_CHECK(h);
        return MCL_OK;
    }
    const int64_t nt = (n + MT_TILE - 1) / MT_TILE;
    // work buffers in the KLD arena (separate from the reduction scratch the callers use)
    const size_t need = (size_t)nt * (8 + 4 + 8 + 4) + 256;
    if (need > h->seq_bytes) {
        MCL_CUDA(h, cudaStreamSynchronize(h->stream));
        cudaFree(h->d_seq); h->d_seq = nullptr; h->seq_bytes = 0;
        MCL_CUDA(h, cudaMalloc(&h->d_seq, need * 2));
        h->seq_bytes = need * 2;
    }
    char *s = (char *)h->d_seq;
    const int64_t cap = (int64_t)((h->seq_bytes - 256) / 24);
    double *tsum = (double *)s;
    uint2 *agg = (uint2 *)(s + (size_t)cap * 8);
    int *e_spec = (int *)(s + (size_t)cap * 16);
    float *c_in = (float *)(s + (size_t)cap * 20);
    k_mt_tilesum<<<(int)nt, MT_THREADS, 0, h->stream>>>(d_w, n, d_div, tsum);
    MCL_LAUNCH_CHECK(h);
    k_mt_speculate<<<1, 1024, 0, h->stream>>>(tsum, nt, e_spec);
    MCL_LAUNCH_CHECK(h);
    k_mt_aggregate<<<(int)nt, MT_THREADS, 0, h->stream>>>(d_w, n, d_div, e_spec, agg);
    MCL_LAUNCH_CHECK(h);
    k_mt_walk<<<1, MT_THREADS, 0, h->stream>>>(d_w, n, d_div, e_spec, agg, nt, c_in, d_total);
    MCL_LAUNCH_CHECK(h);
    if (d_c) {
        k_mt_write<<<(int)nt, MT_THREADS, 0, h->stream>>>(d_w, n, d_div, c_in, d_c);
        MCL_LAUNCH_CHECK(h);
    }
    return MCL_OK;
}

// idx[m] = min(first i in [0, limit] with c_i >= U_m, limit); U_m = r + m*step (two f64 roundings)
// divide: U_m = r + m / n_out (pu:496 low_variance_resample_amcl) instead of r + m * (1 / n_out)
__global__ void k_search_ref_f32(const float *__restrict__ c, int64_t limit, int64_t n_out, double r, double step,
                                 int32_t *__restrict__ idx, bool divide = false) {
    for (int64_t m = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; m < n_out; m += (int64_t)gridDim.x * blockDim.x) {
        const double U = divide ? __dadd_rn(r, __ddiv_rn((double)m, (double)n_out))
                                : __dadd_rn(r, __dmul_rn((double)m, step));     // pu:440
        int64_t lo = 0, hi = limit;            // invariant: answer in [lo, hi]
        while (lo < hi) {
            const int64_t mid = (lo + hi) >> 1;
            if (U > (double)c[mid]) lo = mid + 1; else hi = mid;       // pu:441 "U > c" keeps walking
        }
        idx[m] = (int32_t)lo;
    }
}

// ----------------------------------------------------------------------------- fixed-point mode
#define SCAN_THREADS 256
#define SCAN_ITEMS 8
#define SCAN_TILE (SCAN_THREADS * SCAN_ITEMS)

__device__ __forceinline__ uint64_t quantise(float w, double scale) {
    const double v = __dmul_rn((double)w, scale);
    return v > 0.0 ? (uint64_t)__double2ull_rz(v) : 0ull;     // negative / NaN weights count as 0
}

__global__ void __launch_bounds__(256) k_wmax(const float *__restrict__ w, int64_t n, unsigned *counter,
                                              float *partials, double *out_scale, int64_t n_global) {
    __shared__ float sh[32];
    __shared__ bool last;
    float m = 0.0f;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        m = fmaxf(m, w[i]);
    m = warp_max(m);
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = m;
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int k = 1; k < (int)(blockDim.x >> 5); ++k) m = fmaxf(m, sh[k]);
        partials[blockIdx.x] = m;
        __threadfence();
        last = atomicAdd(counter, 1u) == gridDim.x - 1;
    }
    __syncthreads();
    if (last) {
        __threadfence();
        float t = 0.0f;
        for (unsigned b = threadIdx.x; b < gridDim.x; b += blockDim.x) t = fmaxf(t, ((volatile float *)partials)[b]);
        t = warp_max(t);
        __syncthreads();
        if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = t;
        __syncthreads();
        if (threadIdx.x == 0) {
            for (int k = 1; k < (int)(blockDim.x >> 5); ++k) t = fmaxf(t, sh[k]);
            // scale = 2^(62 - ceil(log2 n_global) - e), 2^e > wmax   (oracle: orc_resample_scale)
            int e = 0;
            if (t > 0.0f) frexp((double)t, &e);
            int lg = 0;
            while (((int64_t)1 << lg) < n_global) ++lg;
            out_scale[0] = ldexp(1.0, 62 - lg - e);
            out_scale[1] = (double)t;
            *counter = 0;
        }
    }
}

__device__ __forceinline__ uint64_t warp_incl_scan_u64(uint64_t v) {
    const int lane = threadIdx.x & 31;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint64_t t = __shfl_up_sync(0xffffffffu, v, o);
        if (lane >= o) v += t;
    }
    return v;
}

// tile sums
__global__ void __launch_bounds__(SCAN_THREADS) k_tile_sums(const float *__restrict__ w, int64_t n,
                                                           const double *scale_ptr, uint64_t *tile_sums) {
    __shared__ uint64_t sh[SCAN_THREADS / 32];
    const double scale = scale_ptr[0];
    const int64_t base = (int64_t)blockIdx.x * SCAN_TILE;
    uint64_t s = 0;
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; ++k) {
        const int64_t i = base + (int64_t)k * SCAN_THREADS + threadIdx.x;
        if (i < n) s += quantise(w[i], scale);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        uint64_t t = 0;
        for (int k = 0; k < SCAN_THREADS / 32; ++k) t += sh[k];
        tile_sums[blockIdx.x] = t;
    }
}

// exclusive scan of the tile sums (single block, sequential over chunks); writes total to tile_off[nt]
__global__ void __launch_bounds__(1024) k_scan_tile_sums(const uint64_t *tile_sums, int64_t nt, uint64_t *tile_off,
                                                         uint64_t carry_in) {
    __shared__ uint64_t sh[32];
    __shared__ uint64_t carry;
    if (threadIdx.x == 0) carry = carry_in;
    __syncthreads();
    for (int64_t base = 0; base < nt; base += blockDim.x) {
        const int64_t i = base + threadIdx.x;
        const uint64_t v = i < nt ? tile_sums[i] : 0;
        uint64_t inc = warp_incl_scan_u64(v);
        if ((threadIdx.x & 31) == 31) sh[threadIdx.x >> 5] = inc;
        __syncthreads();
        if (threadIdx.x < 32) {
            uint64_t t = threadIdx.x < (blockDim.x >> 5) ? sh[threadIdx.x] : 0;
            t = warp_incl_scan_u64(t);
            sh[threadIdx.x] = t;
        }
        __syncthreads();
        const uint64_t warp_off = (threadIdx.x >> 5) ? sh[(threadIdx.x >> 5) - 1] : 0;
        const uint64_t c0 = carry;
        if (i < nt) tile_off[i] = c0 + warp_off + inc - v;
        __syncthreads();
        if (threadIdx.x == blockDim.x - 1) carry = c0 + warp_off + inc;
        __syncthreads();
    }
    if (threadIdx.x == 0) tile_off[nt] = carry;
}

// inclusive cumulative sums C[i] (uint64), element order: i = base + k*SCAN_THREADS + t is NOT
// contiguous per thread, so scan item-row by item-row (each row of 256 consecutive elements).
__global__ void __launch_bounds__(SCAN_THREADS) k_tile_scan(const float *__restrict__ w, int64_t n,
                                                           const double *scale_ptr, const uint64_t *tile_off,
                                                           uint64_t *__restrict__ C) {
    __shared__ uint64_t sh[SCAN_THREADS / 32];
    __shared__ uint64_t carry;
    const double scale = scale_ptr[0];
    const int64_t base = (int64_t)blockIdx.x * SCAN_TILE;
    if (threadIdx.x == 0) carry = tile_off[blockIdx.x];
    __syncthreads();
    for (int k = 0; k < SCAN_ITEMS; ++k) {
        const int64_t i = base + (int64_t)k * SCAN_THREADS + threadIdx.x;
        const uint64_t v = i < n ? quantise(w[i], scale) : 0;
        uint64_t inc = warp_incl_scan_u64(v);
        if ((threadIdx.x & 31) == 31) sh[threadIdx.x >> 5] = inc;
        __syncthreads();
        uint64_t warp_off = 0;
        for (int q = 0; q < (int)(threadIdx.x >> 5); ++q) warp_off += sh[q];
        const uint64_t c0 = carry;
        const uint64_t mine = c0 + warp_off + inc;
        if (i < n) C[i] = mine;
        __syncthreads();
        if (threadIdx.x == SCAN_THREADS - 1) carry = mine;
        __syncthreads();
    }
}

// idx[m] = min(first i in [0, limit] with C_i >= T_m, limit), T_m = ceil((r + m*step) * total)
// m is a GLOBAL output index (m0 + local); C holds GLOBAL cumulative sums of this rank's slice.
__global__ void k_search_fixed(const uint64_t *__restrict__ C, int64_t limit, int64_t m0, int64_t n_out, double r,
                               double step, const uint64_t *total_ptr, uint64_t total_val, uint64_t offset,
                               int32_t *__restrict__ idx) {
    const double totd = (double)(total_ptr ? total_ptr[0] : total_val);
    for (int64_t j = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; j < n_out; j += (int64_t)gridDim.x * blockDim.x) {
        const double U = __dadd_rn(r, __dmul_rn((double)(m0 + j), step));
        const double t = ceil(__dmul_rn(U, totd));
        const uint64_t T = t >= 18446744073709551616.0 ? 0xffffffffffffffffull : (t > 0.0 ? __double2ull_rz(t) : 0ull);
        int64_t lo = 0, hi = limit;
        while (lo < hi) {
            const int64_t mid = (lo + hi) >> 1;
            if (T > C[mid] + offset) lo = mid + 1; else hi = mid;
        }
        idx[j] = (int32_t)lo;
    }
}

extern "C" int mcl_resample_indices(mcl_handle *h, const float *d_w, int64_t n_in, int64_t n_out, double r,
                                    int mode, int32_t *d_idx) {
    if (!h) return MCL_ERR_ARG;
    if (n_in <= 0 || n_out < 0 || !d_w || (n_out > 0 && !d_idx))
        return mcl_fail(h, MCL_ERR_ARG, "mcl_resample_indices: bad argument");
    if (n_in > 0x7fffffffLL) return mcl_fail(h, MCL_ERR_CAPACITY, "mcl_resample_indices: n_in exceeds int32 indices");
    if (n_out == 0) return MCL_OK;
    DeviceGuard guard(h->device);
    const double step = 1.0 / (double)n_out;                       // pu:434
    const int64_t limit = std::min(n_in, n_out) - 1;               // pu:441 "i < N - 1"
    const int sblocks = (int)std::min<int64_t>((n_out + 255) / 256, (int64_t)h->sm_count * 16);
    if (mode == MCL_RESAMPLE_REFERENCE_F32) {
        int rc = mcl_ensure_scratch(h, 64 + sizeof(float) * (size_t)n_in);
        if (rc) return rc;
        float *c = (float *)((char *)h->d_scratch + 64);
        float *sum = (float *)h->d_scratch;
        static int serial = -1;
        if (serial < 0) { const char *e = getenv("MCL_REF_SERIAL"); serial = (e && atoi(e)) ? 1 : 0; }
        if (serial) {        // the one-warp replay of the additions (kept as the cross-check of the parallel scan)
            k_cumsum_ref_f32<<<1, 32, 0, h->stream>>>(d_w, n_in, c, 1);
            MCL_LAUNCH_CHECK(h);
        } else {
            rc = seq_accumulate(h, d_w, n_in, nullptr, nullptr, sum);     // pu:430 np.sum
            if (rc) return rc;
            rc = seq_accumulate(h, d_w, n_in, sum, c, nullptr);           // pu:430 divide, pu:436-443 running sum
            if (rc) return rc;
        }
        k_search_ref_f32<<<sblocks, 256, 0, h->stream>>>(c, limit, n_out, r, step, d_idx);
        MCL_LAUNCH_CHECK(h);
        return MCL_OK;
    }
    if (mode == MCL_RESAMPLE_AMCL_F32) {
        // pu:486-502 low_variance_resample_amcl: the weights AS GIVEN (no normalisation), c a sequential f32 sum,
        // U = r + m / target_size, walk bounded by len(particles) - 1
        int rc = mcl_ensure_scratch(h, 64 + sizeof(float) * (size_t)n_in);
        if (rc) return rc;
        float *c = (float *)((char *)h->d_scratch + 64);
        rc = seq_accumulate(h, d_w, n_in, nullptr, c, nullptr);
        if (rc) return rc;
        k_search_ref_f32<<<sblocks, 256, 0, h->stream>>>(c, n_in - 1, n_out, r, step, d_idx, true);
        MCL_LAUNCH_CHECK(h);
        return MCL_OK;
    }
    if (mode != MCL_RESAMPLE_FIXED_POINT) return mcl_fail(h, MCL_ERR_ARG, "mcl_resample_indices: unknown mode");
    const int64_t nt = (n_in + SCAN_TILE - 1) / SCAN_TILE;
    const int wblocks = (int)std::min<int64_t>((n_in + 1023) / 1024, (int64_t)h->sm_count * 8);
    // scratch: [0,64) counter | [64,128) scale(2 doubles) | wmax partials (wblocks floats) | tile_sums nt |
    //          tile_off nt+1 | C n_in
    size_t off = 128;
    const size_t o_part = off; off += ((size_t)wblocks * sizeof(float) + 63) & ~(size_t)63;
    const size_t o_ts = off; off += ((size_t)nt * 8 + 63) & ~(size_t)63;
    const size_t o_to = off; off += ((size_t)(nt + 1) * 8 + 63) & ~(size_t)63;
    const size_t o_c = off; off += (size_t)n_in * 8;
    int rc = mcl_ensure_scratch(h, off);
    if (rc) return rc;
    char *s = (char *)h->d_scratch;
    unsigned *counter = (unsigned *)s;
    double *scale = (double *)(s + 64);
    uint64_t *tile_sums = (uint64_t *)(s + o_ts), *tile_off = (uint64_t *)(s + o_to), *C = (uint64_t *)(s + o_c);
    MCL_CUDA(h, cudaMemsetAsync(counter, 0, sizeof(unsigned), h->stream));
    k_wmax<<<wblocks, 256, 0, h->stream>>>(d_w, n_in, counter, (float *)(s + o_part), scale, n_in);
    MCL_LAUNCH_CHECK(h);
    k_tile_sums<<<(int)nt, SCAN_THREADS, 0, h->stream>>>(d_w, n_in, scale, tile_sums);
    MCL_LAUNCH_CHECK(h);
    k_scan_tile_sums<<<1, 1024, 0, h->stream>>>(tile_sums, nt, tile_off, 0ull);
    MCL_LAUNCH_CHECK(h);
    k_tile_scan<<<(int)nt, SCAN_THREADS, 0, h->stream>>>(d_w, n_in, scale, tile_off, C);
    MCL_LAUNCH_CHECK(h);
    k_search_fixed<<<sblocks, 256, 0, h->stream>>>(C, limit, 0, n_out, r, step, tile_off + nt, 0ull, 0ull, d_idx);
    MCL_LAUNCH_CHECK(h);
    return MCL_OK;
}

// ---- staged forms for the sharded (multi-GPU) path --------------------------------------------
__global__ void k_scale_from_wmax(const float *wmax, int64_t n_global, double *out_scale) {
    const float t = wmax[0];
    int e = 0;
    if (t > 0.0f) frexp((double)t, &e);
    int lg = 0;
    while (((int64_t)1 << lg) < n_global) ++lg;
    out_scale[0] = ldexp(1.0, 62 - lg - e);
    out_scale[1] = (double)t;
}
__global__ void k_store_wmax(const double *scale, float *out) { out[0] = (float)scale[1]; }

struct FixedLayout { size_t o_part, o_ts, o_to, o_c, total; int64_t nt; int wblocks; };
static FixedLayout fixed_layout(const mcl_handle *h, int64_t n_in) {
    FixedLayout L;
    L.nt = (n_in + SCAN_TILE - 1) / SCAN_TILE;
    L.wblocks = (int)std::min<int64_t>((n_in + 1023) / 1024, (int64_t)h->sm_count * 8);
    size_t off = 128;
    L.o_part = off; off += ((size_t)L.wblocks * sizeof(float) + 63) & ~(size_t)63;
    L.o_ts = off; off += ((size_t)L.nt * 8 + 63) & ~(size_t)63;
    L.o_to = off; off += ((size_t)(L.nt + 1) * 8 + 63) & ~(size_t)63;
    L.o_c = off; off += (size_t)n_in * 8;
    L.total = off;
    return L;
}

// local max of the weights -> d_wmax[0] (f32, device); the caller all-reduces it (MAX)
extern "C" int mcl_weights_max(mcl_handle *h, const float *d_w, int64_t n, float *d_wmax) {
    if (!h) return MCL_ERR_ARG;
    if (n <= 0 || !d_w || !d_wmax) return mcl_fail(h, MCL_ERR_ARG, "mcl_weights_max: bad argument");
    DeviceGuard guard(h->device);
    const FixedLayout L = fixed_layout(h, n);
    int rc = mcl_ensure_scratch(h, L.total);
    if (rc) return rc;
    char *s = (char *)h->d_scratch;
    MCL_CUDA(h, cudaMemsetAsync(s, 0, sizeof(unsigned), h->stream));
    k_wmax<<<L.wblocks, 256, 0, h->stream>>>(d_w, n, (unsigned *)s, (float *)(s + L.o_part), (double *)(s + 64), n);
    MCL_LAUNCH_CHECK(h);
    k_store_wmax<<<1, 1, 0, h->stream>>>((double *)(s + 64), d_wmax);
    MCL_LAUNCH_CHECK(h);
    return MCL_OK;
}

// fixed-point cumulative sums of this rank's weights (kept in the handle's scratch until the next
// library call that needs scratch) with the GLOBAL scale; d_total[0] = this rank's total (uint64).
extern "C" int mcl_resample_scan(mcl_handle *h, const float *d_w, int64_t n_in, const float *d_wmax_global,
                                 int64_t n_global, uint64_t *d_total) {
    if (!h) return MCL_ERR_ARG;
    if (n_in <= 0 || !d_w || !d_wmax_global || !d_total || n_global < n_in)
        return mcl_fail(h, MCL_ERR_ARG, "mcl_resample_scan: bad argument");
    DeviceGuard guard(h->device);
    const FixedLayout L = fixed_layout(h, n_in);
    int rc = mcl_ensure_scratch(h, L.total);
    if (rc) return rc;
    char *s = (char *)h->d_scratch;
    double *scale = (double *)(s + 64);
    uint64_t *tile_sums = (uint64_t *)(s + L.o_ts), *tile_off = (uint64_t *)(s + L.o_to), *C = (uint64_t *)(s + L.o_c);
    k_scale_from_wmax<<<1, 1, 0, h->stream>>>(d_wmax_global, n_global, scale);
    MCL_LAUNCH_CHECK(h);
    k_tile_sums<<<(int)L.nt, SCAN_THREADS, 0, h->stream>>>(d_w, n_in, scale, tile_sums);
    MCL_LAUNCH_CHECK(h);
    k_scan_tile_sums<<<1, 1024, 0, h->stream>>>(tile_sums, L.nt, tile_off, 0ull);
    MCL_LAUNCH_CHECK(h);
    k_tile_scan<<<(int)L.nt, SCAN_THREADS, 0, h->stream>>>(d_w, n_in, scale, tile_off, C);
    MCL_LAUNCH_CHECK(h);
    MCL_CUDA(h, cudaMemcpyAsync(d_total, tile_off + L.nt, sizeof(uint64_t), cudaMemcpyDeviceToDevice, h->stream));
    return MCL_OK;
}

// source index (local) for the global outputs m0 .. m0+n_out_local-1, which the caller has determined
// to fall into this rank's cumulative-weight interval (offset, offset + local total].
extern "C" int mcl_resample_search(mcl_handle *h, int64_t n_in, uint64_t offset, uint64_t grand_total, int64_t m0,
                                   int64_t n_out_local, double r, int64_t n_out_global, int32_t *d_idx) {
    if (!h) return MCL_ERR_ARG;
    if (n_in <= 0 || n_out_local < 0 || n_out_global <= 0 || (n_out_local > 0 && !d_idx))
        return mcl_fail(h, MCL_ERR_ARG, "mcl_resample_search: bad argument");
    if (n_out_local == 0) return MCL_OK;
    DeviceGuard guard(h->device);
    const FixedLayout L = fixed_layout(h, n_in);
    if (L.total > h->scratch_bytes) return mcl_fail(h, MCL_ERR_STATE, "mcl_resample_search: call mcl_resample_scan first");
    const uint64_t *C = (const uint64_t *)((char *)h->d_scratch + L.o_c);
    const int sblocks = (int)std::min<int64_t>((n_out_local + 255) / 256, (int64_t)h->sm_count * 16);
    k_search_fixed<<<sblocks, 256, 0, h->stream>>>(C, n_in - 1, m0, n_out_local, r, 1.0 / (double)n_out_global, nullptr,
                                                   grand_total, offset, d_idx);
    MCL_LAUNCH_CHECK(h);
    return MCL_OK;
}

// sequential f32 running sum of the weights as given (pu:555-563 kld_sampling_amcl does not renormalise)
int mcl_cumsum_f32_seq(mcl_handle *h, const float *d_w, int64_t n, float *d_c) {
    return seq_accumulate(h, d_w, n, nullptr, d_c, nullptr);
}

// ---- peer-push global resampling: gather fused with the all-to-all over NVLink peer memory --------------
// Every rank emits the offspring whose thresholds fall into its own cumulative-weight interval and stores each
// one DIRECTLY into the destination rank's particle buffer (P2P stores through NVLink / NVSwitch), so there are no
// send buffers, no split sizes and no host round trip: the interval, the output range and the destinations are
// all computed on the device from the all-gathered totals.  The caller follows with a barrier.
struct PushPlan { uint64_t offset, grand; long long m_lo, m_hi; };

__device__ __forceinline__ uint64_t push_threshold(long long m, double r, double step, double totd) {
    const double U = __dadd_rn(r, __dmul_rn((double)m, step));
    const double t = ceil(__dmul_rn(U, totd));
    return t >= 18446744073709551616.0 ? 0xffffffffffffffffull : (t > 0.0 ? __double2ull_rz(t) : 0ull);
}
__device__ long long count_thresholds_le_dev(uint64_t x, double r, double step, double totd, long long n_out) {
    long long lo = 0, hi = n_out;
    while (lo < hi) {
        const long long mid = (lo + hi) >> 1;
        if (push_threshold(mid, r, step, totd) <= x) lo = mid + 1; else hi = mid;
    }
    return lo;
}
// same partition as sharded.plan_resample (host): ranks without weight emit nothing, rank 0 starts at 0,
// the last rank ends at n_out
__global__ void __launch_bounds__(32) k_push_plan(const uint64_t *totals, int rank, int world, double r, long long n_out,
                                                  PushPlan *plan) {
    // one lane per rank boundary (the 2 x world binary searches of ~25 fp64 steps each took ~40 us on one thread)
    __shared__ uint64_t off[17];
    __shared__ long long cnt[17];
    const int t = threadIdx.x;
    if (t == 0) {
        uint64_t acc = 0;
        for (int k = 0; k < world; ++k) { off[k] = acc; acc += totals[k]; }
        off[world] = acc;
    }
    __syncwarp();
    const double step = 1.0 / (double)n_out, totd = (double)off[world];
    if (t <= world) cnt[t] = count_thresholds_le_dev(off[t], r, step, totd, n_out);
    __syncwarp();
    if (t == 0) {
        long long prev_hi = 0, my_lo = 0, my_hi = 0;
        for (int k = 0; k < world; ++k) {
            long long lo = k == 0 ? 0 : cnt[k];
            long long hi = k == world - 1 ? n_out : cnt[k + 1];
            if (hi < lo) hi = lo;
            if (k > 0) { if (lo < prev_hi) lo = prev_hi; if (hi < lo) hi = lo; }
            if (k == world - 1) hi = n_out;
            if (k == rank) { my_lo = lo; my_hi = hi; }
            prev_hi = hi;
        }
        plan->offset = off[rank]; plan->grand = off[world]; plan->m_lo = my_lo; plan->m_hi = my_hi;
    }
}

// tile_prefix (optional): inclusive cumulative sum at the end of every `tile` inputs (low 62 bits), searched in
// shared memory before the 10-11 steps inside one tile of C
__global__ void k_push(const uint64_t *__restrict__ C, int64_t limit, const PushPlan *plan, double r, long long n_out,
                       long long n_per_rank, int world, const double *__restrict__ x, const double *__restrict__ y,
                       const double *__restrict__ th, const unsigned long long *peers /* [3][world] */,
                       const unsigned long long *__restrict__ tile_prefix, int nt, int tile) {
    extern __shared__ unsigned long long tp[];
    if (tile_prefix) {
        for (int t = threadIdx.x; t < nt; t += blockDim.x) tp[t] = __ldcg(tile_prefix + t) & ((1ull << 62) - 1);
        __syncthreads();
    }
    const PushPlan pl = *plan;
    const double step = 1.0 / (double)n_out, totd = (double)pl.grand;
    for (long long m = pl.m_lo + blockIdx.x * (long long)blockDim.x + threadIdx.x; m < pl.m_hi;
         m += (long long)gridDim.x * blockDim.x) {
        const uint64_t T = push_threshold(m, r, step, totd);
        int64_t lo = 0, hi = limit;
        if (tile_prefix) {
            int tl = 0, th2 = nt - 1;
            while (tl < th2) {
                const int mid = (tl + th2) >> 1;
                if (T > tp[mid] + pl.offset) tl = mid + 1; else th2 = mid;
            }
            lo = (int64_t)tl * tile;
            hi = min(lo + tile - 1, limit);
        }
        while (lo < hi) {
            const int64_t mid = (lo + hi) >> 1;
            if (T > C[mid] + pl.offset) lo = mid + 1; else hi = mid;
        }
        long long d = m / n_per_rank;
        if (d > world - 1) d = world - 1;
        const long long j = m - d * n_per_rank;
        reinterpret_cast<double *>(peers[d])[j] = x[lo];
        reinterpret_cast<double *>(peers[world + d])[j] = y[lo];
        reinterpret_cast<double *>(peers[2 * world + d])[j] = th[lo];
    }
    __threadfence_system();
}

extern "C" int mcl_resample_push(mcl_handle *h, int64_t n_in, const uint64_t *d_totals_all, int rank, int world,
                                 double r, int64_t n_global, int64_t n_per_rank, const double *d_x, const double *d_y,
                                 const double *d_theta, const uint64_t *d_peer_ptrs) {
    if (!h) return MCL_ERR_ARG;
    if (n_in <= 0 || !d_totals_all || world < 1 || world > 16 || rank < 0 || rank >= world || n_global <= 0 ||
        n_per_rank <= 0 || !d_x || !d_y || !d_theta || !d_peer_ptrs)
        return mcl_fail(h, MCL_ERR_ARG, "mcl_resample_push: bad argument");
    DeviceGuard guard(h->device);
    const FixedLayout L = fixed_layout(h, n_in);
    if (L.total > h->scratch_bytes) return mcl_fail(h, MCL_ERR_STATE, "mcl_resample_push: call mcl_resample_scan first");
    const uint64_t *C = (const uint64_t *)((char *)h->d_scratch + L.o_c);
    PushPlan *plan = (PushPlan *)((char *)h->d_scratch + 64 + 32);      // after the scale slot
    k_push_plan<<<1, 32, 0, h->stream>>>(d_totals_all, rank, world, r, (long long)n_global, plan);
    MCL_LAUNCH_CHECK(h);
    k_push<<<h->sm_count * 8, 256, 0, h->stream>>>(C, n_in - 1, plan, r, (long long)n_global, (long long)n_per_rank, world,
                                                   d_x, d_y, d_theta, (const unsigned long long *)d_peer_ptrs, nullptr, 0, 0);
    MCL_LAUNCH_CHECK(h);
    return MCL_OK;
}

// the same push reading the cumulative sums from d_C (fused.cu's scan) instead of the scratch of mcl_resample_scan
int mcl_resample_push_from(mcl_handle *h, const unsigned long long *d_C, const unsigned long long *d_tile_prefix, int nt,
                           int tile, int64_t n_in, const uint64_t *d_totals_all, int rank, int world, double r,
                           int64_t n_global, int64_t n_per_rank, const double *d_x, const double *d_y,
                           const double *d_theta, const uint64_t *d_peer_ptrs) {
    int rc = mcl_ensure_scratch(h, 256);
    if (rc) return rc;
    PushPlan *plan = (PushPlan *)((char *)h->d_scratch + 64 + 32);
    k_push_plan<<<1, 32, 0, h->stream>>>(d_totals_all, rank, world, r, (long long)n_global, plan);
    MCL_LAUNCH_CHECK(h);
    const bool coarse = d_tile_prefix && nt > 0 && (size_t)nt * 8 <= 48 * 1024;
    k_push<<<h->sm_count * 8, 256, coarse ? (size_t)nt * 8 : 0, h->stream>>>(
        (const uint64_t *)d_C, n_in - 1, plan, r, (long long)n_global, (long long)n_per_rank, world, d_x, d_y, d_theta,
        (const unsigned long long *)d_peer_ptrs, coarse ? d_tile_prefix : nullptr, nt, tile);
    MCL_LAUNCH_CHECK(h);
    return MCL_OK;
}

// test hook: the running f32 sums themselves (normalise != 0: of w / seq_sum(w), else of w as given);
// serial != 0 uses the one-warp replay instead of the parallel scan
extern "C" int mcl_debug_seq_cumsum(mcl_handle *h, const float *d_w, int64_t n, int normalise, int serial, float *d_c) {
    if (!h) return MCL_ERR_ARG;
    if (n <= 0 || !d_w || !d_c) return mcl_fail(h, MCL_ERR_ARG, "mcl_debug_seq_cumsum: bad argument");
    DeviceGuard guard(h->device);
    int rc = mcl_ensure_scratch(h, 64);
    if (rc) return rc;
    float *sum = (float *)h->d_scratch;
    if (serial) {
        k_cumsum_ref_f32<<<1, 32, 0, h->stream>>>(d_w, n, d_c, normalise);
        MCL_LAUNCH_CHECK(h);
        return MCL_OK;
    }
    if (normalise) {
        rc = seq_accumulate(h, d_w, n, nullptr, nullptr, sum);
        if (rc) return rc;
    }
    return seq_accumulate(h, d_w, n, normalise ? sum : nullptr, d_c, nullptr);
}

// host mirror of oracle u53(Philox(seed, step, item 0, sub 0, RESAMPLE))
static void philox_host(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) {
    uint32_t c0 = ctr[0], c1 = ctr[1], c2 = ctr[2], c3 = ctr[3], k0 = key[0], k1 = key[1];
    for (int r = 0; r < 10; ++r) {
        const uint64_t p0 = (uint64_t)0xD2511F53u * c0, p1 = (uint64_t)0xCD9E8D57u * c2;
        const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0, n1 = (uint32_t)p1;
        const uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1, n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

extern "C" double mcl_resample_offset(uint64_t seed, uint64_t step, int64_t n_out) {
    uint32_t ctr[4] = {0u, (uint32_t)step, 0u,
                       (uint32_t)MCL_STREAM_RESAMPLE | ((uint32_t)((step >> 32) & 0xffu) << 24)};
    uint32_t key[2] = {(uint32_t)seed, (uint32_t)(seed >> 32)}, o[4];
    philox_host(ctr, key, o);
    const double u = ((double)(o[0] >> 5) * 67108864.0 + (double)(o[1] >> 6)) / 9007199254740992.0;
    const double hi = 1.0 / (double)n_out;
    return 0.0 + (hi - 0.0) * u;          // np.random.uniform(0.0, step)
}

// new_particles[m] = particles[idx[m]] (pu:445)
__global__ void k_gather(const double *__restrict__ x, const double *__restrict__ y, const double *__restrict__ th,
                         const int32_t *__restrict__ idx, int64_t n, double *__restrict__ xo, double *__restrict__ yo,
                         double *__restrict__ tho) {
    for (int64_t m = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; m < n; m += (int64_t)gridDim.x * blockDim.x) {
        const int32_t i = idx[m];
        xo[m] = x[i]; yo[m] = y[i]; tho[m] = th[i];
    }
}

extern "C" int mcl_gather(mcl_handle *h, const double *d_x, const double *d_y, const double *d_theta,
                          const int32_t *d_idx, int64_t n_out, double *d_xo, double *d_yo, double *d_thetao) {
    if (!h) return MCL_ERR_ARG;
    if (n_out < 0 || (n_out > 0 && (!d_x || !d_y || !d_theta || !d_idx || !d_xo || !d_yo || !d_thetao)))
        return mcl_fail(h, MCL_ERR_ARG, "mcl_gather: bad argument");
    if (d_xo == d_x || d_yo == d_y || d_thetao == d_theta)
        return mcl_fail(h, MCL_ERR_ARG, "mcl_gather: outputs must not alias inputs");
    if (n_out == 0) return MCL_OK;
    DeviceGuard guard(h->device);
    const int blocks = (int)std::min<int64_t>((n_out + 255) / 256, (int64_t)h->sm_count * 16);
    k_gather<<<blocks, 256, 0, h->stream>>>(d_x, d_y, d_theta, d_idx, n_out, d_xo, d_yo, d_thetao);
    MCL_LAUNCH_CHECK(h);
    return MCL_OK;
}
