// tail.cu -- the whole tail of a filter step in ONE persistent cooperative kernel:
//   softmax x2 (node:351-358) -> MH accept (pu:208-236) -> estimate sums (node:586-597) -> systematic
//   resampling (pu:416-446) in EITHER arithmetic: the reference's own (sequential f32 normalising sum, f32
//   divide, sequential f32 running sum -- reproduced bit for bit) or the 64-bit fixed-point one.
// One CTA per SM (1024 threads), grid-wide barriers between the stages instead of kernel boundaries:
//   S1  exact 2^-40 sums of exp(s - max)                                  | barrier 1
//   S2  weights, MH accept, new pose + weight, raw estimate sums per tile  | barrier 2
//   S3  means; central sums; REFERENCE: exact sequential-f32 sum S of the weights (pass 1)
//                            FIXED    : quantised tile totals              | barrier 3
//   S4  REFERENCE: exact sequential-f32 running sums of w / S (pass 2);  FIXED: integer scan   | barrier 4
//   S5  per output slot: search of the running sums (staged in shared memory) + pose gather
//
// The exact sequential-f32 passes.  c_i = fl32(c_{i-1} + w_i) is emulated with the integer maps of seqsum.cuh,
// which need the binade of the running sum.  An fp64 prefix predicts it for every thread (8 consecutive
// weights); threads whose prefix lies within 2^-13 (relative) of a power of two are not trusted: their
// additions are replayed with real f32 adds.  Per tile a segmented scan composes the maps of every run of
// trusted threads, so a tile is a short list of items (run map | replayed thread); one lane walks that list
// from the tile's exact incoming sum, CHECKING every run (binade as predicted, no overflow) -- if a check fails
// the tile is redone by the general restart loop, so the result never depends on the prediction.  Tiles are
// chained by decoupled look-back (aggregate map of a clean tile | exact outgoing sum), so the serial part of a
// pass is one short walk per binade crossing (~8 tiles at 1 M particles).
// Every spin loop is bounded (~2 s): a lost peer sets hd->err instead of hanging the GPU.
#include <float.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>

#include "common.cuh"
#include "seqsum.cuh"

#define TL_THREADS 1024
#define TL_WARPS 32
#define TL_MAX_IPT 8
#define TL_MAX_ROUNDS 64
#define TL_MAX_ITEMS 160
#define TL_MAX_SERIAL 256
#define TL_NBAR 4
#define TL_SPIN_LIMIT 4000000000ll
#define TL_SOFTMAX_FIX 1099511627776.0   // 2^40 (mh_softmax.cu)
#define TL_DELTA 1.220703125e-4           // 2^-13: per-thread distrust margin around a binade boundary (MULTI tiles)
#define TL_DELTA_TILE 1.953125e-3         // 2^-9: margin by which a tile's predicted sums are widened
#define TL_STAGE_BYTES 65536
#define TL_COARSE_MAX 2048

#define TL_F_AGG 1ull
#define TL_F_INC 2ull
#define TL_F_NOAGG 3ull
#define TL_SERIAL_MARK (-100000)
#define TL_FULL 0xffffffffu

struct TailHeader {              // device, zero-initialised once
    unsigned long long bar;      // grid barrier counter (monotone; the host tracks the base)
    unsigned long long ckey[2];  // MH chain: max(carried score) as a key, slot = iteration parity (read | written)
    int err, pad;
    float S, pad2;               // reference mode: exact sequential f32 sum of the weights (pu:430)
    unsigned long long total;    // fixed-point mode: total of the quantised weights
};

struct TailArgs {
    TailHeader *hd;
    unsigned long long *keymax;  // [2] order-preserving keys of max(score), left by the likelihood kernel
    unsigned long long bar_base;
    int64_t n;
    int nt, ipt, tile, use_mh;
    unsigned long long *prof;    // debug: [grid][32] globaltimer stamps of the stage boundaries (nullable)
    int stop;                    // debug: return after stage `stop` (0 = run everything)
    int raw;                     // 1 (test hook): w_out holds the weights already, no poses (S1, estimate, gather skipped)
                                 // 2: w_out holds the weights and (nx, ny, nth) the particles (MH chain: estimate + resampling)
    const float *s_post, *s_pre;
    float *w_out;
    const double *px, *py, *pt, *ox, *oy, *ot;
    double *nx, *ny, *nth;
    uint64_t seed, step, first_index;
    unsigned long long *part_q;  // [2][grid]
    double *part_m;              // [nt][8]: six raw sums, weight maximum
    double *part_c;              // [nt][9]
    unsigned long long *st1, *st2;   // [nt] look-back records of the two exact passes
    void *C;                     // [n]  f32 running sums (reference) / u64 cumulative sums (fixed)
    void *tend;                  // [nt] value of C at the end of every tile
    unsigned long long *ttot;    // [nt] fixed: tile totals
    double *est18;
    double *drift;               // [nt] relative drift (exact sequential sum / fp64 sum - 1) at the end of every tile, last step
    double r, rstep;
    int32_t *idx;
    double *gx, *gy, *gt;
    // sharded operation (SH kernels): exchanges over NVLink peer memory, offspring pushed to their destination rank
    int rank, world;
    int64_t n_global;                        // particles of the whole population (= world * n)
    unsigned long long *mailbox;             // this rank's mailbox (filter.cu: flags[2][16] | slots[2][16][16])
    unsigned long long peers[16];            // every rank's mailbox as mapped here
    unsigned long long epoch0;               // exchanges consumed before this launch
    int *comm_err;                           // sticky: a peer did not answer
    const unsigned long long *peer_pose;     // device [3][world]: x / y / theta of the destination set on every rank
    // MH chain iteration (k_chain_tail)
    float *score_chain;                      // carried scores (read and updated in place)
    unsigned long long *ckey_in, *ckey_out;  // key of max(carried score): of this iteration / for the next one
};

struct TlItems {
    int e[TL_MAX_ITEMS];             // unit exponent of a run, or TL_SERIAL_MARK
    Pair64 map[TL_MAX_ITEMS];        // run: composed map; replayed thread: .a0 = slot in sw
    float c[TL_MAX_ITEMS];           // exact sum entering the item (written by the walker)
    float sw[TL_MAX_SERIAL * TL_MAX_IPT];
};

struct TlShared {
    double d[9][TL_WARPS];
    double par[12];
    unsigned long long q[2][TL_WARPS];
    unsigned long long sumq[2];
    Pair64 pw[TL_WARPS];
    int pwf[TL_WARPS];
    int cnt[TL_WARPS];
    int code[TL_THREADS];
    Pair64 inc[TL_THREADS];
    TlItems it;
    double ownP[TL_MAX_ROUNDS], ownT[TL_MAX_ROUNDS];
    unsigned long long ownE[TL_MAX_ROUNDS];
    float c_in, c_out, cseg, cprime;
    int fail, nitems, nserial, tstar;
    unsigned long long xpay[16];             // exchange payload of this rank (CTA 0 publishes it)
    unsigned long long xch[16 * 16];         // [rank][value] what every rank published
    unsigned long long poff[17];             // cumulative weight before every rank (and the grand total)
    long long pcnt[17], m_lo, m_hi;          // output slots whose threshold lies below every rank / owned by this rank
    long long cross;
    int64_t seg0;
    int64_t irange[2];
};

// ------------------------------------------------------------------------------------------------ primitives
// Warp / lane index through inline PTX: nvcc 12.9 folds `&arr[tid >> 5]` to `base + (tid >> 2)` under a dominating
// `lane == 0` test and then reuses that address for a `lane == 31` store of another helper (misaligned address at
// run time); the opaque shifts keep the address computations apart.
__device__ __forceinline__ int tl_warp() {
    int w;
    asm("shr.u32 %0, %1, 5;" : "=r"(w) : "r"(threadIdx.x));
    return w;
}
__device__ __forceinline__ int tl_lane() {
    int l;
    asm("and.b32 %0, %1, 31;" : "=r"(l) : "r"(threadIdx.x));
    return l;
}
__device__ __forceinline__ unsigned long long tl_ld_acquire(const unsigned long long *p) {
    unsigned long long v;
    asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void tl_st_release(unsigned long long *p, unsigned long long v) {
    asm volatile("st.release.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

__device__ __forceinline__ void tl_stamp(const TailArgs &a, int k) {
    if (a.prof && threadIdx.x == 0) {
        unsigned long long t;
        asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
        a.prof[(size_t)blockIdx.x * 32 + k] = t;
    }
}

__device__ __forceinline__ void tl_grid_barrier(const TailArgs &a, int k) {
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned long long target = a.bar_base + (unsigned long long)k * gridDim.x;
        __threadfence();
        atomicAdd(&a.hd->bar, 1ull);
        const long long t0 = clock64();
        while (tl_ld_acquire(&a.hd->bar) < target) {
            if (clock64() - t0 > TL_SPIN_LIMIT) { a.hd->err = 1; break; }
        }
    }
    __syncthreads();
}

__device__ __forceinline__ void tl_st_release_sys(unsigned long long *p, unsigned long long v) {
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long tl_ld_acquire_sys(const unsigned long long *p) {
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}

// Cross-rank exchange k of this launch (all threads of all CTAs call it): CTA 0 stores this rank's payload sh.xpay[0 ..
// nvals) into every peer's mailbox over NVLink and raises its flag; every CTA (or only CTA 0) waits for the flags of
// all ranks on its OWN mailbox and copies what they published into sh.xch[rank][value].  Same mailbox, parities and
// epochs as filter.cu's k_exchange, so both forms can alternate on one stream.
// warp 0 of any CTA: wait until every rank has published exchange k, copy the payloads into sh.xch[rank][value]
__device__ void tl_wait_all(const TailArgs &a, TlShared &sh, int k, int nvals) {
    const int lane = tl_lane();
    const unsigned long long epoch = a.epoch0 + (unsigned long long)k;
    const int par = (int)(epoch & 1ull);
    bool dead = *(volatile int *)a.comm_err != 0;      // a peer was lost earlier: do not wait again
    if (!dead && lane < a.world) {
        const unsigned long long *flag = a.mailbox + par * 16 + lane;
        const long long t0 = clock64();
        while (tl_ld_acquire_sys(flag) < epoch) {
            if (clock64() - t0 > TL_SPIN_LIMIT) { dead = true; break; }
        }
    }
    if (__any_sync(TL_FULL, dead) && lane == 0) *a.comm_err = 1;
    const volatile unsigned long long *slots = a.mailbox + 32 + (size_t)par * 16 * 16;
    for (int q = lane; q < a.world * nvals; q += 32) {
        const int r = q / nvals, v = q - r * nvals;
        sh.xch[r * 16 + v] = slots[r * 16 + v];
    }
}
// one thread: publish ONE value as this rank's payload of exchange k (the hand-off of an exact running sum to the
// next rank; everybody reads it later through tl_wait_all)
__device__ void tl_publish_one(const TailArgs &a, int k, unsigned long long value) {
    const unsigned long long epoch = a.epoch0 + (unsigned long long)k;
    const int par = (int)(epoch & 1ull);
    for (int d = 0; d < a.world; ++d) {
        unsigned long long *peer = reinterpret_cast<unsigned long long *>(a.peers[d]);
        peer[32 + ((size_t)par * 16 + a.rank) * 16] = value;
    }
    __threadfence_system();
    for (int d = 0; d < a.world; ++d)
        tl_st_release_sys(reinterpret_cast<unsigned long long *>(a.peers[d]) + par * 16 + a.rank, epoch);
}
// any thread(s): the value rank `src` published for exchange k (spins until it is there)
__device__ unsigned long long tl_wait_one(const TailArgs &a, int k, int src) {
    const unsigned long long epoch = a.epoch0 + (unsigned long long)k;
    const int par = (int)(epoch & 1ull);
    const unsigned long long *flag = a.mailbox + par * 16 + src;
    if (*(volatile int *)a.comm_err == 0) {
        const long long t0 = clock64();
        while (tl_ld_acquire_sys(flag) < epoch) {
            if (clock64() - t0 > TL_SPIN_LIMIT) { *a.comm_err = 1; break; }
        }
    }
    return ((const volatile unsigned long long *)a.mailbox)[32 + ((size_t)par * 16 + src) * 16];
}

__device__ void tl_exchange(const TailArgs &a, TlShared &sh, int k, int nvals, bool all_wait) {
    const int lane = tl_lane(), warp = tl_warp();
    const unsigned long long epoch = a.epoch0 + (unsigned long long)k;
    const int par = (int)(epoch & 1ull);
    __syncthreads();
    if (warp == 0) {
        if (blockIdx.x == 0 && lane < a.world) {
            unsigned long long *peer = reinterpret_cast<unsigned long long *>(a.peers[lane]);
            unsigned long long *slot = peer + 32 + ((size_t)par * 16 + a.rank) * 16;
            for (int v = 0; v < nvals; ++v) slot[v] = sh.xpay[v];
            __threadfence_system();
            tl_st_release_sys(peer + par * 16 + a.rank, epoch);
        }
        if (all_wait || blockIdx.x == 0) tl_wait_all(a, sh, k, nvals);
    }
    __syncthreads();
}

// thresholds of the global systematic resampling (resample.cu push_threshold / count_thresholds_le_dev)
__device__ __forceinline__ unsigned long long tl_threshold(long long m, double r, double step, double totd) {
    const double U = __dadd_rn(r, __dmul_rn((double)m, step));
    const double t = ceil(__dmul_rn(U, totd));
    return t >= 18446744073709551616.0 ? 0xffffffffffffffffull : (t > 0.0 ? __double2ull_rz(t) : 0ull);
}
__device__ long long tl_count_thresholds_le(unsigned long long x, double r, double step, double totd, long long n_out) {
    long long lo = 0, hi = n_out;
    while (lo < hi) {
        const long long mid = (lo + hi) >> 1;
        if (tl_threshold(mid, r, step, totd) <= x) lo = mid + 1; else hi = mid;
    }
    return lo;
}

__device__ __forceinline__ float tl_softmax_num(float s, float m) { return (float)exp((double)__fsub_rn(s, m)); }
__device__ __forceinline__ unsigned long long tl_quantise(float w, double scale) {
    const double v = __dmul_rn((double)w, scale);
    return v > 0.0 ? __double2ull_rz(v) : 0ull;     // negative / NaN weights count as 0
}
__device__ __forceinline__ unsigned long long tl_warp_sum_u64(unsigned long long v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(TL_FULL, v, o);
    return v;
}

// sums of K doubles over the block; the totals are returned to every thread (fixed tree: deterministic)
template <int K>
__device__ __forceinline__ void tl_block_sum(double (&v)[K], TlShared &sh) {
    const int warp = tl_warp(), lane = tl_lane();
#pragma unroll
    for (int k = 0; k < K; ++k) v[k] = warp_sum(v[k]);
    if (lane == 0)
#pragma unroll
        for (int k = 0; k < K; ++k) sh.d[k][warp] = v[k];
    __syncthreads();
    if (warp == 0) {
#pragma unroll
        for (int k = 0; k < K; ++k) {
            const double t = warp_sum(sh.d[k][lane]);
            if (lane == 0) sh.par[k] = t;
        }
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < K; ++k) v[k] = sh.par[k];
    __syncthreads();
}
__device__ __forceinline__ float tl_block_max(float v, TlShared &sh) {
    const int warp = tl_warp(), lane = tl_lane();
    v = warp_max(v);
    if (lane == 0) sh.d[0][warp] = (double)v;
    __syncthreads();
    if (warp == 0) {
        const float t = warp_max((float)sh.d[0][lane]);
        if (lane == 0) sh.par[0] = (double)t;
    }
    __syncthreads();
    const float r = (float)sh.par[0];
    __syncthreads();
    return r;
}
__device__ __forceinline__ void tl_block_sum_u64x2(unsigned long long &a0, unsigned long long &a1, TlShared &sh) {
    const int warp = tl_warp(), lane = tl_lane();
    a0 = tl_warp_sum_u64(a0); a1 = tl_warp_sum_u64(a1);
    if (lane == 0) { sh.q[0][warp] = a0; sh.q[1][warp] = a1; }
    __syncthreads();
    if (warp == 0) {
        const unsigned long long t0 = tl_warp_sum_u64(sh.q[0][lane]), t1 = tl_warp_sum_u64(sh.q[1][lane]);
        if (lane == 0) { sh.sumq[0] = t0; sh.sumq[1] = t1; }
    }
    __syncthreads();
    a0 = sh.sumq[0]; a1 = sh.sumq[1];
    __syncthreads();
}
// exclusive scans over the block's threads (thread order); `total` is returned to every thread
__device__ __forceinline__ double tl_block_excl_scan_d(double v, double &total, TlShared &sh) {
    const int warp = tl_warp(), lane = tl_lane();
    double inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const double t = __shfl_up_sync(TL_FULL, inc, o); if (lane >= o) inc += t; }
    if (lane == 31) sh.d[0][warp] = inc;
    __syncthreads();
    if (warp == 0) {
        double t = sh.d[0][lane];
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const double u = __shfl_up_sync(TL_FULL, t, o); if (lane >= o) t += u; }
        sh.d[1][lane] = t;
    }
    __syncthreads();
    const double off = warp ? sh.d[1][warp - 1] : 0.0;
    total = sh.d[1][TL_WARPS - 1];
    const double r = off + inc - v;
    __syncthreads();
    return r;
}
__device__ __forceinline__ unsigned long long tl_block_excl_scan_u64(unsigned long long v, unsigned long long &total,
                                                                      TlShared &sh) {
    const int warp = tl_warp(), lane = tl_lane();
    unsigned long long inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const unsigned long long t = __shfl_up_sync(TL_FULL, inc, o); if (lane >= o) inc += t; }
    if (lane == 31) sh.q[0][warp] = inc;
    __syncthreads();
    if (warp == 0) {
        unsigned long long t = sh.q[0][lane];
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const unsigned long long u = __shfl_up_sync(TL_FULL, t, o); if (lane >= o) t += u; }
        sh.q[1][lane] = t;
    }
    __syncthreads();
    const unsigned long long off = warp ? sh.q[1][warp - 1] : 0ull;
    total = sh.q[1][TL_WARPS - 1];
    const unsigned long long r = off + inc - v;
    __syncthreads();
    return r;
}
// exclusive scan of element maps (first the earlier thread, then the later one); aggregate to every thread
__device__ __forceinline__ Pair64 tl_block_excl_scan_pair(const Pair64 &mine, Pair64 &aggregate, TlShared &sh) {
    const int warp = tl_warp(), lane = tl_lane();
    Pair64 inc = mine;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const Pair64 t = pair_shfl_up(inc, o);
        if (lane >= o) inc = pair_compose(t, inc);
    }
    if (lane == 31) sh.pw[warp] = inc;
    __syncthreads();
    if (warp == 0) {
        Pair64 t = sh.pw[lane];
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const Pair64 u = pair_shfl_up(t, o);
            if (lane >= o) t = pair_compose(u, t);
        }
        sh.pw[lane] = t;                       // inclusive over warps
    }
    __syncthreads();
    Pair64 excl = pair_shfl_up(inc, 1);
    if (lane == 0) { excl.a0 = 0; excl.a1 = 0; }
    if (warp > 0) excl = pair_compose(sh.pw[warp - 1], excl);
    aggregate = sh.pw[TL_WARPS - 1];
    __syncthreads();
    return excl;
}

__device__ __forceinline__ long long tl_apply(long long K, const Pair64 &m) { return K + ((K & 1) ? m.a1 : m.a0); }

// ------------------------------------------------------------------------------------------------ look-back
__device__ __forceinline__ unsigned long long tl_rec_agg(int e, const Pair64 &m) {
    return (TL_F_AGG << 62) | ((unsigned long long)((e + 126) & 0xff) << 52) | ((unsigned long long)m.a0 << 26) |
           (unsigned long long)m.a1;
}
__device__ __forceinline__ unsigned long long tl_rec_inc(float c) { return (TL_F_INC << 62) | (unsigned long long)__float_as_uint(c); }

// serial form: from the exact sum c leaving tile `start`, apply the records of tiles start+1 .. v-1 one by one; a
// record that does not apply (binade not as predicted, overflow) is waited for until its tile has published its sum
__device__ __noinline__ float tl_lookback_serial(unsigned long long *st, int v, int start, float c, long long t0, int *err) {
    const int lane = tl_lane();
    for (int base = start + 1; base < v; base += 32) {
        const int q = base + lane;
        unsigned long long rec = 0;
        if (q < v) rec = tl_ld_acquire(st + q);
        const int cnt = min(32, v - base);
        for (int l = 0; l < cnt; ++l) {
            unsigned long long r = __shfl_sync(TL_FULL, rec, l);
            if ((r >> 62) == TL_F_AGG) {
                const int e = (int)((r >> 52) & 0xff) - 126;
                Pair64 m;
                m.a0 = (unsigned)((r >> 26) & 0x3ffffffull); m.a1 = (unsigned)(r & 0x3ffffffull);
                const long long Kn = tl_apply(seq_K(c), m);
                if (seq_exponent(c) == e && Kn < (1ll << 24)) { c = seq_value(Kn, e); continue; }
            }
            bool dead = false;
            while ((r >> 62) != TL_F_INC) {
                r = tl_ld_acquire(st + base + l);
                if (clock64() - t0 > TL_SPIN_LIMIT) { dead = true; break; }
            }
            if (dead) { if (lane == 0) *err = 3; return 0.0f; }
            c = __uint_as_float((unsigned)r);
        }
    }
    return c;
}

// exact sum entering tile v (warp 0, all lanes; every lane returns the same value).  Backwards in windows of 32
// tiles: the aggregate maps of clean tiles are composed by a shuffle reduction while the window's nearest dirty
// tile is still being waited for, so that once an exact sum arrives it takes ONE checked application.
// xk != 0 (sharded, reference arithmetic): the sum does not start at 0 but at the exact sum leaving the previous
// rank, which that rank hands over as exchange xk
__device__ float tl_lookback(unsigned long long *st, int v, int *err, const TailArgs &a, int xk) {
    const int lane = tl_lane();
    const bool handoff = xk != 0 && a.rank > 0;
    if (v == 0) return handoff ? __uint_as_float((unsigned)tl_wait_one(a, xk, a.rank - 1)) : 0.0f;
    const long long t0 = clock64();
    int top = v - 1;
    Pair64 suffix; suffix.a0 = 0; suffix.a1 = 0;         // composed map of the clean tiles (top, v-1]
    int e_suf = 1000;                                    // their common unit exponent (1000: none yet)
    bool composable = true;
    while (true) {
        const int q = top - lane;
        unsigned long long rec = TL_F_INC << 62;         // "tile -1": the sum starts at 0 ...
        bool dead = false;
        if (q >= 0) {
            while (true) {
                rec = tl_ld_acquire(st + q);
                const unsigned f = (unsigned)(rec >> 62);
                if (f == TL_F_AGG || f == TL_F_INC) break;
                if (clock64() - t0 > TL_SPIN_LIMIT) { dead = true; break; }
            }
        } else if (handoff) {                            // ... or at what the previous rank hands over
            rec = (TL_F_INC << 62) | (tl_wait_one(a, xk, a.rank - 1) & 0xffffffffull);
        }
        if (__any_sync(TL_FULL, dead)) { if (lane == 0) *err = 2; return 0.0f; }
        // a clean tile predicted in ANOTHER binade than the tiles composed so far ends the composable stretch
        // like a dirty one: its own exact sum is waited for (it publishes it right after its own look-back)
        const int e = (int)((rec >> 52) & 0xff) - 126;
        const bool isinc = (rec >> 62) == TL_F_INC;
        int e_ref = e_suf;
        if (e_ref == 1000) {
            const int e0 = __shfl_sync(TL_FULL, e, 0);
            const bool inc0 = __shfl_sync(TL_FULL, isinc ? 1 : 0, 0) != 0;
            if (!inc0) e_ref = e0;
        }
        const unsigned stopm = __ballot_sync(TL_FULL, isinc || e != e_ref);
        const int j = stopm ? __ffs(stopm) - 1 : 32;     // lanes 0 .. j-1: clean tiles of one binade, nearer than any stop
        if (j < 32) {
            unsigned long long rj = __shfl_sync(TL_FULL, rec, j);
            bool dead2 = false;
            while ((rj >> 62) != TL_F_INC) {             // (uniform) the stop tile's exact sum
                rj = tl_ld_acquire(st + (top - j));
                if (clock64() - t0 > TL_SPIN_LIMIT) { dead2 = true; break; }
            }
            if (dead2) { if (lane == 0) *err = 4; return 0.0f; }
            if (lane == j) rec = rj;
        }
        const unsigned incm = j < 32 ? (1u << j) : 0u;
        const bool isagg = lane < j;
        Pair64 m; m.a0 = 0; m.a1 = 0;
        if (isagg) { m.a0 = (unsigned)((rec >> 26) & 0x3ffffffull); m.a1 = (unsigned)(rec & 0x3ffffffull); }
        if (j > 0 && e_suf == 1000) e_suf = e_ref;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {               // lane 0 <- first the farthest tile, ..., last the nearest
            Pair64 other;
            other.a0 = __shfl_down_sync(TL_FULL, m.a0, o);
            other.a1 = __shfl_down_sync(TL_FULL, m.a1, o);
            if (lane + o < 32) m = pair_compose(other, m);
        }
        Pair64 win;
        win.a0 = __shfl_sync(TL_FULL, m.a0, 0); win.a1 = __shfl_sync(TL_FULL, m.a1, 0);
        suffix = pair_compose(win, suffix);
        if (incm) {
            const float c = __uint_as_float((unsigned)__shfl_sync(TL_FULL, rec, j));
            if (e_suf == 1000) return c;                 // no clean tile in between
            if (composable) {
                const long long Kn = tl_apply(seq_K(c), suffix);
                if (seq_exponent(c) == e_suf && Kn < (1ll << 24)) return seq_value(Kn, e_suf);
            }
            return tl_lookback_serial(st, v, top - j, c, t0, err);
        }
        top -= 32;
    }
}

// ------------------------------------------------------------------------------------------------ exact tile
// general path: exact running sums of one tile from c_start, one block scan per binade crossing.  The thread's
// weights are read from the shared-memory copy wt[0 .. ipt) (rolled loops: this path is cold, keep it small).
__device__ __noinline__ float tl_tile_general(bool write, const float *wt, int ipt, int64_t first, int64_t tile_lo,
                                              int64_t end, float c_start, float *C, TlShared &sh) {
    const long long NONE = 0x7fffffffffffffffll;
    if (threadIdx.x == 0) { sh.cseg = c_start; sh.seg0 = tile_lo; }
    __syncthreads();
    while (true) {
        const float c0 = sh.cseg;
        const int64_t seg0 = sh.seg0;
        const int e = seq_exponent(c0);
        const long long K0 = seq_K(c0);
        Pair64 run; run.a0 = 0; run.a1 = 0;
#pragma unroll 1
        for (int k = 0; k < ipt; ++k) {
            const int64_t i = first + k;
            if (i >= seg0 && i < end) run = pair_compose(run, seq_decode(wt[k], e));
        }
        if (threadIdx.x == 0) sh.cross = NONE;
        Pair64 agg;
        const Pair64 excl = tl_block_excl_scan_pair(run, agg, sh);
        const long long Kstart = tl_apply(K0, excl);
        long long my_cross = NONE, Kb = 0, Kw = Kstart;
        float w_cross = 0.0f;
#pragma unroll 1
        for (int k = 0; k < ipt; ++k) {
            const int64_t i = first + k;
            if (i >= seg0 && i < end) {
                const float w = wt[k];
                const long long Kn = tl_apply(Kw, seq_decode(w, e));
                if (Kn >= (1ll << 24) && my_cross == NONE) { my_cross = i; Kb = Kw; w_cross = w; }
                Kw = Kn;
            }
        }
        if (my_cross != NONE) atomicMin((unsigned long long *)&sh.cross, (unsigned long long)my_cross);
        __syncthreads();
        const long long cross = sh.cross;
        const int64_t seg_end = cross == NONE ? end : (int64_t)cross;
        if (write) {
            Kw = Kstart;
#pragma unroll 1
            for (int k = 0; k < ipt; ++k) {
                const int64_t i = first + k;
                if (i >= seg0 && i < end) {
                    Kw = tl_apply(Kw, seq_decode(wt[k], e));
                    if (i < seg_end) C[i] = seq_value(Kw, e);
                }
            }
        }
        if (cross == NONE) {
            const float r = seq_value(tl_apply(K0, agg), e);
            __syncthreads();
            return r;
        }
        if (my_cross == cross) {                      // this thread owns the addition that leaves the binade
            const float cn = __fadd_rn(seq_value(Kb, e), w_cross);
            if (write) C[cross] = cn;
            sh.cseg = cn;
            sh.seg0 = cross + 1;
        }
        __syncthreads();
        if (sh.seg0 >= end) {
            const float r = sh.cseg;
            __syncthreads();
            return r;
        }
    }
}

// central sums of this CTA's tiles around the population means (node:590-597, pu:69-83) -> part_c
__device__ __forceinline__ void tl_central_sums(const TailArgs &a, TlShared &sh, double mx, double my, double mt) {
    const int t = threadIdx.x;
    for (int v = blockIdx.x; v < a.nt; v += gridDim.x) {
        double s9[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
        const int64_t base = (int64_t)v * a.tile + t;
        for (int k = 0; k < a.ipt; ++k) {
            const int64_t i = base + (int64_t)k * TL_THREADS;
            if (i >= a.n) break;
            const double wi = (double)a.w_out[i];
            const double dx = a.nx[i] - mx, dy = a.ny[i] - my;
            const double dt = (double)(float)normalize_angle_dev(__dadd_rn(a.nth[i], -mt));   // pu:80-82
            s9[0] += wi * dx; s9[1] += wi * dy; s9[2] += wi * dt;
            s9[3] += wi * dx * dx; s9[4] += wi * dx * dy; s9[5] += wi * dx * dt;
            s9[6] += wi * dy * dy; s9[7] += wi * dy * dt; s9[8] += wi * dt * dt;
        }
        tl_block_sum<9>(s9, sh);
        if (t == 0) {
            double *pc = a.part_c + (size_t)v * 9;
#pragma unroll
            for (int k = 0; k < 9; ++k) pc[k] = s9[k];
        }
    }
}

// exclusive SUFFIX scan of element maps: the composition of the maps of threads t+1 .. 1023 (t+1 applied first);
// `aggregate` (all threads, in order) is returned to every thread
__device__ __forceinline__ Pair64 tl_block_suffix_scan_pair(const Pair64 &mine, Pair64 &aggregate, TlShared &sh) {
    const int warp = tl_warp(), lane = tl_lane();
    Pair64 inc = mine;                                   // lanes lane .. 31 of this warp
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        Pair64 other;
        other.a0 = __shfl_down_sync(TL_FULL, inc.a0, o);
        other.a1 = __shfl_down_sync(TL_FULL, inc.a1, o);
        if (lane + o < 32) inc = pair_compose(inc, other);
    }
    if (lane == 0) sh.pw[warp] = inc;
    __syncthreads();
    if (warp == 0) {
        Pair64 p = sh.pw[lane];
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            Pair64 other;
            other.a0 = __shfl_down_sync(TL_FULL, p.a0, o);
            other.a1 = __shfl_down_sync(TL_FULL, p.a1, o);
            if (lane + o < 32) p = pair_compose(p, other);
        }
        sh.pw[lane] = p;                                 // warps lane .. 31
    }
    __syncthreads();
    Pair64 excl;
    excl.a0 = __shfl_down_sync(TL_FULL, inc.a0, 1);
    excl.a1 = __shfl_down_sync(TL_FULL, inc.a1, 1);
    if (lane == 31) { excl.a0 = 0; excl.a1 = 0; }
    if (warp < TL_WARPS - 1) excl = pair_compose(excl, sh.pw[warp + 1]);
    aggregate = sh.pw[0];
    __syncthreads();
    return excl;
}

#define TL_WSTRIDE 9      // shared-memory words per thread (odd: thread-consecutive reads are bank-conflict free)

#define TL_CENTRAL_EARLY 8      // CTAs below this do their central sums after the first pass (they head the chain)
#define TL_KIND_CLEAN 0
#define TL_KIND_TWO 1
#define TL_KIND_MULTI 2
#define TL_KIND_SEQ 3                      // one small tile: the plain sequential loop on one thread
#define TL_SEQ_MAX 2048                    // ... up to this many weights (2 ns per addition beats the walk of a MULTI tile)
#define TL_WSM_BYTES (TL_THREADS * TL_WSTRIDE * 4)

// One exact pass over the tiles this CTA owns (v = blockIdx.x, + gridDim.x, ...).  P2: the weights are divided
// by S first (pu:430) and every running sum is written to C (pu:436-443); otherwise only the total is wanted.
// A tile is prepared BEFORE its exact incoming sum is known, according to where the fp64 prediction of its sums
// (widened by 2^-9: real filter weights repeat, so the f32 sum drifts systematically) lies:
//   CLEAN  inside one binade e: plain scan of the maps under e; the aggregate map is published for the look-back
//   TWO    may cross from e to e+1: prefix scan under e and suffix scan under e+1; with the exact incoming sum one
//          lane finds the crossing thread by binary search, replays its <= 8 additions in real f32 and applies the
//          suffix map -- whatever the drift, as long as the crossing falls inside the tile (or not at all)
//   MULTI  several binades (the first tiles, where the sum is still tiny and the prediction sharp): per-thread
//          prediction, runs of trusted threads composed by a segmented scan, the others replayed one by one
// Every shortcut is CHECKED against the exact sum (binade as assumed, no overflow); a tile that fails is redone by
// the general restart loop, so no result depends on a prediction.
// composition of the maps of this thread's weights wt[0 .. ipt) under unit exponent e (rolled: code size)
__device__ __forceinline__ Pair64 tl_thread_map(const float *wt, int ipt, int e) {
    Pair64 F; F.a0 = 0; F.a1 = 0;
#pragma unroll
    for (int k = 0; k < TL_MAX_IPT; ++k)
        if (k < ipt) F = pair_compose(F, seq_decode(wt[k], e));
    return F;
}

template <bool P2>
__device__ __forceinline__ void tl_exact_pass(const TailArgs &a, TlShared &sh, float S, unsigned char *dyn, bool central,
                                              double mx, double my, double mt, int xk) {
    unsigned long long *st = P2 ? a.st2 : a.st1;
    float *C = (float *)a.C;
    float *wsm = (float *)dyn;
    Pair64 *pfs = (Pair64 *)(dyn + TL_WSM_BYTES), *sxs = pfs + TL_THREADS;
    const int t = threadIdx.x, lane = tl_lane(), warp = tl_warp();
    const double invS = P2 ? 1.0 / (double)S : 1.0;
    int rd = 0;
    for (int v = blockIdx.x; v < a.nt; v += gridDim.x, ++rd) {
        const int64_t tile_lo = (int64_t)v * a.tile;
        const int64_t end = min(a.n, tile_lo + a.tile);
        const int64_t first = tile_lo + (int64_t)t * a.ipt;
        // coalesced loads -> shared memory -> ipt consecutive weights per thread (padding: + 0 changes no sum)
        for (int k = 0; k < a.ipt; ++k) {
            const int j = k * TL_THREADS + t;
            const int64_t i = tile_lo + j;
            float w = 0.0f;
            if (i < end) { w = __ldcg(a.w_out + i); if (P2) w = __fdiv_rn(w, S); }
            const int tj = j / a.ipt;
            wsm[tj * TL_WSTRIDE + (j - tj * a.ipt)] = w;
        }
        __syncthreads();
        const float *wt = wsm + t * TL_WSTRIDE;          // this thread's weights: wt[0 .. ipt)
        double ls = 0.0;
#pragma unroll
        for (int k = 0; k < TL_MAX_IPT; ++k)
            if (k < a.ipt) ls += (double)wt[k];
        if (rd == 0) tl_stamp(a, P2 ? 26 : 29);
        double ttot;
        const double pe = tl_block_excl_scan_d(ls, ttot, sh);     // predicted (fp64) sum entering this thread's weights
        if (rd == 0) tl_stamp(a, P2 ? 27 : 30);
        // Where the tile's sums will lie.  The fp64 sums of S2 know nothing of the DRIFT of the sequential f32 sum,
        // which in a converged filter (long runs of equal weights, every addition rounding the same way) reaches
        // ~1 % -- more than the 2^-9 margin, so the last tiles (sum near 1 = a binade boundary) were misclassified
        // and fell back to the restart loop, 12 us per pass.  Pass 2 therefore predicts from the EXACT sums of
        // pass 1 (its records, divided by S: the same weights up to the scale), and pass 1 corrects the fp64 sums
        // by the relative drift the previous step saw at the same tile.  Predictions only: every result is checked.
        const double p_lo = sh.ownP[rd] * invS, p_hi = p_lo + ttot;
        double lo_v = p_lo, hi_v = p_hi;
        if (P2) {
            if (v > 0) lo_v = (double)__uint_as_float((unsigned)__ldcg(a.st1 + v - 1)) * invS;
            hi_v = (double)__uint_as_float((unsigned)__ldcg(a.st1 + v)) * invS;
        } else if (a.drift) {
            if (v > 0) lo_v = p_lo * (1.0 + __ldcg(a.drift + v - 1));
            hi_v = p_hi * (1.0 + __ldcg(a.drift + v));
        }
        if (!(hi_v >= lo_v)) { lo_v = p_lo; hi_v = p_hi; }
        int kind = TL_KIND_MULTI, e1 = 0;
        if (lo_v > 0.0 && hi_v < 1e38) {
            const int ea = ilogb(lo_v * (1.0 - TL_DELTA_TILE)), eb = ilogb(hi_v * (1.0 + TL_DELTA_TILE));
            if (ea >= -126 && eb < 127) { e1 = ea; kind = eb == ea ? TL_KIND_CLEAN : (eb == ea + 1 ? TL_KIND_TWO : TL_KIND_MULTI); }
        }
        // The reference's own scale (a few thousand particles): the whole population is one tile whose sum runs through
        // ~20 binades.  The definition itself -- one thread adding in order, pu:430 / pu:436-443 -- takes 2 ns per
        // weight there, less than preparing and walking a MULTI tile.
        if (a.nt == 1 && xk == 0 && a.n <= TL_SEQ_MAX) kind = TL_KIND_SEQ;
        // per-kind state that the write phase needs
        Pair64 EX; EX.a0 = 0; EX.a1 = 0;                 // CLEAN / TWO: map of the threads before me (under e1); MULTI: inside my run
        Pair64 G; G.a0 = 0; G.a1 = 0;                    // TWO: my map under e1 + 1
        int et = e1, my_item = 0, nitems = 1;
        bool serial = false;
        if (kind == TL_KIND_CLEAN) {
            const Pair64 F = tl_thread_map(wt, a.ipt, e1);
            Pair64 agg;
            EX = tl_block_excl_scan_pair(F, agg, sh);
            if (rd == 0) tl_stamp(a, P2 ? 20 : 16);
            if (t == 0) {
                const bool pub = agg.a0 < (1u << 24) && agg.a1 < (1u << 24);
                tl_st_release(st + v, pub ? tl_rec_agg(e1, agg) : (TL_F_NOAGG << 62));
            }
            if (!P2 && central && rd == 0) tl_central_sums(a, sh, mx, my, mt);      // hidden behind the chain
            if (warp == 0) {
                const float c_in = tl_lookback(st, v, &a.hd->err, a, xk);
                if (rd == 0) tl_stamp(a, P2 ? 21 : 17);
                if (lane == 0) {
                    const long long Kn = tl_apply(seq_K(c_in), agg);
                    const int fail = !(seq_exponent(c_in) == e1 && Kn < (1ll << 24));
                    const float c = fail ? c_in : seq_value(Kn, e1);
                    if (!fail) tl_st_release(st + v, tl_rec_inc(c));
                    sh.c_in = c_in; sh.c_out = c; sh.fail = fail;
                }
                if (rd == 0) tl_stamp(a, P2 ? 22 : 18);
            }
        } else if (kind == TL_KIND_TWO) {
            const Pair64 F = tl_thread_map(wt, a.ipt, e1);
            G = tl_thread_map(wt, a.ipt, e1 + 1);
            Pair64 aggF, aggG;
            EX = tl_block_excl_scan_pair(F, aggF, sh);
            const Pair64 SX = tl_block_suffix_scan_pair(G, aggG, sh);
            pfs[t] = pair_compose(EX, F);                // through my last weight, under e1
            sxs[t] = SX;                                 // everything behind me, under e1 + 1
            __syncthreads();
            if (rd == 0) tl_stamp(a, P2 ? 20 : 16);
            if (t == 0) tl_st_release(st + v, TL_F_NOAGG << 62);
            if (!P2 && central && rd == 0) tl_central_sums(a, sh, mx, my, mt);      // hidden behind the chain
            if (warp == 0) {
                const float c_in = tl_lookback(st, v, &a.hd->err, a, xk);
                if (rd == 0) tl_stamp(a, P2 ? 21 : 17);
                // first thread whose last sum leaves the binade (the sums are monotone): two 32-way steps by the whole
                // warp instead of ten dependent ones by one lane -- this sits on the serial chain of the pass
                const long long K0 = seq_K(c_in);
                const int ec = seq_exponent(c_in);
                int ts_w = TL_THREADS;
                if (ec == e1) {
                    const unsigned m1 = __ballot_sync(TL_FULL, tl_apply(K0, pfs[32 * lane + 31]) >= (1ll << 24));
                    if (m1) {
                        const int ch = __ffs(m1) - 1;
                        const unsigned m2 = __ballot_sync(TL_FULL, tl_apply(K0, pfs[32 * ch + lane]) >= (1ll << 24));
                        ts_w = 32 * ch + __ffs(m2) - 1;
                    }
                }
                if (lane == 0) {
                    int fail = 1, ts = TL_THREADS;
                    float c = c_in, cp = c_in;
                    if (ec == e1) {
                        ts = ts_w;
                        if (ts == TL_THREADS) { c = seq_value(tl_apply(K0, pfs[TL_THREADS - 1]), e1); fail = 0; }
                        else {
                            cp = seq_value(ts ? tl_apply(K0, pfs[ts - 1]) : K0, e1);
                            float w8[TL_MAX_IPT];
#pragma unroll
                            for (int k = 0; k < TL_MAX_IPT; ++k) w8[k] = k < a.ipt ? wsm[ts * TL_WSTRIDE + k] : 0.0f;
#pragma unroll
                            for (int k = 0; k < TL_MAX_IPT; ++k) cp = __fadd_rn(cp, w8[k]);
                            const long long Kn = tl_apply(seq_K(cp), sxs[ts]);
                            if (seq_exponent(cp) == e1 + 1 && Kn < (1ll << 24)) { c = seq_value(Kn, e1 + 1); fail = 0; }
                        }
                    } else if (ec == e1 + 1) {           // the crossing already happened before this tile
                        ts = -1;
                        const long long Kn = tl_apply(K0, aggG);
                        if (Kn < (1ll << 24)) { c = seq_value(Kn, e1 + 1); fail = 0; }
                    }
                    if (!fail) tl_st_release(st + v, tl_rec_inc(c));
                    sh.c_in = c_in; sh.c_out = c; sh.fail = fail; sh.tstar = ts; sh.cprime = cp;
                }
                if (rd == 0) tl_stamp(a, P2 ? 22 : 18);
            }
        } else if (kind == TL_KIND_SEQ) {
            float *dn = (float *)pfs;                    // dense copy of the weights (TL_SEQ_MAX floats fit the map arrays)
            const int cnt = (int)(end - tile_lo);
#pragma unroll
            for (int k = 0; k < TL_MAX_IPT; ++k) {
                const int j = t * a.ipt + k;
                if (k < a.ipt && j < TL_SEQ_MAX) dn[j] = j < cnt ? wt[k] : 0.0f;      // + 0 changes no sum
            }
            __syncthreads();
            if (t == 0) {
                float c = 0.0f;
                float4 *d4 = (float4 *)dn;
                for (int q = 0; q < (cnt + 7) / 8; ++q) {          // eight additions per trip, operands loaded ahead
                    float4 u = d4[2 * q], w = d4[2 * q + 1];
                    u.x = c = __fadd_rn(c, u.x); u.y = c = __fadd_rn(c, u.y); u.z = c = __fadd_rn(c, u.z); u.w = c = __fadd_rn(c, u.w);
                    w.x = c = __fadd_rn(c, w.x); w.y = c = __fadd_rn(c, w.y); w.z = c = __fadd_rn(c, w.z); w.w = c = __fadd_rn(c, w.w);
                    if (P2) { d4[2 * q] = u; d4[2 * q + 1] = w; }
                }
                tl_st_release(st + v, tl_rec_inc(c));
                sh.c_in = 0.0f; sh.c_out = c; sh.fail = 0;
            }
            __syncthreads();
            if (P2) {
#pragma unroll
                for (int k = 0; k < TL_MAX_IPT; ++k) {
                    const int j = t * a.ipt + k;
                    if (k < a.ipt && j < cnt) C[tile_lo + j] = dn[j];
                }
            }
            if (!P2 && central && rd == 0) tl_central_sums(a, sh, mx, my, mt);
        } else {
            // ---- MULTI: per-thread prediction; threads within 2^-13 of a power of two are replayed -------------
            const double Pt = lo_v + pe, Qt = Pt + ls;
            serial = true;
            if (Pt > 0.0) {
                et = ilogb(Pt);
                const double lo = ldexp(1.0, et);
                serial = !(et >= -126 && et < 127 && Pt >= lo * (1.0 + TL_DELTA) && Qt <= 2.0 * lo * (1.0 - TL_DELTA));
            }
            // threads behind the last weight of a ragged tile hold nothing: they never start a piece of their own
            const bool beyond = first >= end && t > 0;
            const int last_real = (int)((end - tile_lo - 1) / a.ipt);
            if (beyond) serial = false;
            const int code = serial ? (TL_SERIAL_MARK - t) : et;
            sh.code[t] = code;
            __syncthreads();
            const bool head = !beyond && (t == 0 || serial || sh.code[t - 1] != code);
            const bool last = !beyond && (t == last_real || serial || sh.code[t + 1] != code);
            Pair64 F; F.a0 = 0; F.a1 = 0;
            if (!serial && !beyond) F = tl_thread_map(wt, a.ipt, et);
            // segmented inclusive scan of the maps (segments start at heads) + counts of heads / replayed threads
            Pair64 inc = F;
            int hf = head ? 1 : 0;
            int cnt = (head ? 1 : 0) | (serial ? 0x10000 : 0);
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const Pair64 p2 = pair_shfl_up(inc, o);
                const int h2 = __shfl_up_sync(TL_FULL, hf, o);
                const int c2 = __shfl_up_sync(TL_FULL, cnt, o);
                if (lane >= o) {
                    if (!hf) inc = pair_compose(p2, inc);
                    hf |= h2;
                    cnt += c2;
                }
            }
            if (lane == 31) { sh.pw[warp] = inc; sh.pwf[warp] = hf; sh.cnt[warp] = cnt; }
            __syncthreads();
            if (warp == 0) {
                Pair64 p = sh.pw[lane];
                int f = sh.pwf[lane], c = sh.cnt[lane];
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const Pair64 p2 = pair_shfl_up(p, o);
                    const int h2 = __shfl_up_sync(TL_FULL, f, o);
                    const int c2 = __shfl_up_sync(TL_FULL, c, o);
                    if (lane >= o) {
                        if (!f) p = pair_compose(p2, p);
                        f |= h2;
                        c += c2;
                    }
                }
                sh.pw[lane] = p; sh.cnt[lane] = c;
            }
            __syncthreads();
            if (warp > 0) {
                if (!hf) inc = pair_compose(sh.pw[warp - 1], inc);
                cnt += sh.cnt[warp - 1];
            }
            const int total_cnt = sh.cnt[TL_WARPS - 1];
            nitems = total_cnt & 0xffff;
            const int nserial = total_cnt >> 16;
            const bool overflow = nitems > TL_MAX_ITEMS || nserial > TL_MAX_SERIAL;
            my_item = (cnt & 0xffff) - 1;
            sh.inc[t] = inc;
            if (!overflow && last) {
                if (serial) {
                    const int slot = (cnt >> 16) - 1;
                    sh.it.e[my_item] = TL_SERIAL_MARK;
                    sh.it.map[my_item].a0 = (unsigned)slot;
#pragma unroll 1
                    for (int k = 0; k < TL_MAX_IPT; ++k) sh.it.sw[slot * TL_MAX_IPT + k] = k < a.ipt ? wt[k] : 0.0f;
                } else {
                    sh.it.e[my_item] = et;
                    sh.it.map[my_item] = inc;            // composed map of the whole run
                }
            }
            __syncthreads();
            if (!head) EX = sh.inc[t - 1];               // map from the head of my run up to (excluding) me
            if (rd == 0) tl_stamp(a, P2 ? 20 : 16);
            if (t == 0) tl_st_release(st + v, TL_F_NOAGG << 62);
            if (!P2 && central && rd == 0) tl_central_sums(a, sh, mx, my, mt);      // hidden behind the chain
            if (warp == 0) {
                const float c_in = tl_lookback(st, v, &a.hd->err, a, xk);
                if (rd == 0) tl_stamp(a, P2 ? 21 : 17);
                if (lane == 0) {
                    float c = c_in;
                    int fail = overflow ? 1 : 0;
                    if (!fail) {
                        for (int it = 0; it < nitems; ++it) {
                            sh.it.c[it] = c;
                            const int e = sh.it.e[it];
                            if (e == TL_SERIAL_MARK) {
                                const float *w8 = sh.it.sw + sh.it.map[it].a0 * TL_MAX_IPT;
#pragma unroll
                                for (int k = 0; k < TL_MAX_IPT; ++k) c = __fadd_rn(c, w8[k]);
                            } else {
                                const long long Kn = tl_apply(seq_K(c), sh.it.map[it]);
                                if (seq_exponent(c) != e || Kn >= (1ll << 24)) { fail = 1; break; }
                                c = seq_value(Kn, e);
                            }
                        }
                    }
                    if (!fail) tl_st_release(st + v, tl_rec_inc(c));     // the chain goes on before this tile writes
                    sh.c_in = c_in; sh.c_out = c; sh.fail = fail;
                }
                if (rd == 0) tl_stamp(a, P2 ? 22 : 18);
            }
        }
        __syncthreads();
        float c_out = sh.c_out;
        const int failed = sh.fail;
        if (failed) {                                    // a check failed (or too many pieces): general path
            c_out = tl_tile_general(P2, wt, a.ipt, first, tile_lo, end, sh.c_in, C, sh);
            if (t == 0) tl_st_release(st + v, tl_rec_inc(c_out));
        } else if (P2 && kind != TL_KIND_SEQ) {
            // ---- every running sum of the tile, from the exact sum that enters it --------------------------------
            bool replay = false;
            float cs = sh.c_in;
            long long K = 0;
            int ew = e1;
            if (kind == TL_KIND_CLEAN) {
                K = tl_apply(seq_K(cs), EX);
            } else if (kind == TL_KIND_TWO) {
                const int ts = sh.tstar;                 // -1: the whole tile is in e1 + 1; 1024: all of it in e1
                Pair64 EXG; EXG.a0 = 0; EXG.a1 = 0;
                if (ts < TL_THREADS) {                   // (uniform) maps under e1 + 1 of the threads behind the crossing
                    Pair64 mine; mine.a0 = 0; mine.a1 = 0;
                    if (t > ts) mine = G;
                    Pair64 dummy;
                    EXG = tl_block_excl_scan_pair(mine, dummy, sh);
                }
                if (t < ts) K = tl_apply(seq_K(cs), EX);
                else if (t == ts) { replay = true; cs = seq_value(tl_apply(seq_K(cs), EX), e1); }
                else { ew = e1 + 1; K = tl_apply(seq_K(sh.cprime), EXG); }
            } else {
                cs = sh.it.c[my_item];
                replay = serial;
                ew = et;
                K = tl_apply(seq_K(cs), EX);
            }
            if (replay) {
                float c = cs;
#pragma unroll
                for (int k = 0; k < TL_MAX_IPT; ++k) {
                    const int64_t i = first + k;
                    if (k < a.ipt) {
                        c = __fadd_rn(c, wt[k]);
                        if (i < end) C[i] = c;
                    }
                }
            } else {
#pragma unroll
                for (int k = 0; k < TL_MAX_IPT; ++k) {
                    const int64_t i = first + k;
                    if (k < a.ipt) {
                        K = tl_apply(K, seq_decode(wt[k], ew));
                        if (i < end) C[i] = seq_value(K, ew);
                    }
                }
            }
        }
        if (rd == 0) {
            tl_stamp(a, P2 ? 23 : 19);
            if (t == 0 && a.prof) a.prof[(size_t)blockIdx.x * 32 + (P2 ? 25 : 24)] = (unsigned long long)(failed * 1000 + kind * 100000 + nitems);
        }
        if (t == 0) {
            if (!P2 && a.drift) a.drift[v] = p_hi > 0.0 ? (double)c_out / p_hi - 1.0 : 0.0;      // for the next step's pass 1
            if (P2) ((float *)a.tend)[v] = c_out;
            else if (v == a.nt - 1) a.hd->S = c_out;
            if (xk && v == a.nt - 1) tl_publish_one(a, xk, (unsigned long long)__float_as_uint(c_out));   // to the next rank
        }
        __syncthreads();
    }
}

// ------------------------------------------------------------------------------------------------ search helpers
template <bool REF> struct TlCum;
template <> struct TlCum<true> {
    typedef float T;
    typedef double Key;
    static __device__ __forceinline__ bool gt(double U, float c) { return U > (double)c; }          // pu:441 "U > c"
};
template <> struct TlCum<false> {
    typedef unsigned long long T;
    typedef unsigned long long Key;
    static __device__ __forceinline__ bool gt(unsigned long long T_, unsigned long long c) { return T_ > c; }
};
template <bool REF>
__device__ __forceinline__ typename TlCum<REF>::Key tl_key(int64_t m, double r, double rstep, double totd);
template <>
__device__ __forceinline__ double tl_key<true>(int64_t m, double r, double rstep, double) {
    return __dadd_rn(r, __dmul_rn((double)m, rstep));                                              // pu:440
}
template <>
__device__ __forceinline__ unsigned long long tl_key<false>(int64_t m, double r, double rstep, double totd) {
    const double U = __dadd_rn(r, __dmul_rn((double)m, rstep));
    const double t = ceil(__dmul_rn(U, totd));
    return t >= 18446744073709551616.0 ? 0xffffffffffffffffull : (t > 0.0 ? __double2ull_rz(t) : 0ull);
}

// first i in [lo, hi] with !(key > C[i]), else hi; one warp, 32 probes per round
template <bool REF>
__device__ int64_t tl_warp_search(const typename TlCum<REF>::T *C, int64_t lo, int64_t hi, typename TlCum<REF>::Key key) {
    const int lane = tl_lane();
    while (lo < hi) {
        const int64_t span = hi - lo;
        const int64_t p = lo + ((int64_t)lane * span) / 32;                  // p_0 = lo, p_31 < hi
        const bool g = TlCum<REF>::gt(key, __ldcg(C + p));
        const unsigned m = __ballot_sync(TL_FULL, g);
        const int nt = __popc(m);                                            // predicates are monotone: true...false
        if (nt == 0) { hi = lo; break; }
        const int64_t p_last_true = lo + ((int64_t)(nt - 1) * span) / 32;
        const int64_t p_first_false = nt < 32 ? lo + ((int64_t)nt * span) / 32 : hi;
        lo = p_last_true + 1;
        hi = p_first_false;
        if (lo > hi) lo = hi;
    }
    return lo;
}

// ------------------------------------------------------------------------------------------------ the kernel
template <bool MH, bool REF, bool SH>
__global__ void __launch_bounds__(TL_THREADS, 1) k_tail(const TailArgs a) {
    typedef typename TlCum<REF>::T CT;
    typedef typename TlCum<REF>::Key KeyT;
    __shared__ TlShared sh;
    extern __shared__ __align__(16) unsigned char dyn[];
    const int t = threadIdx.x, G = gridDim.x, b = blockIdx.x;
    const int64_t n = a.n;

    // ---- S1: sums of the softmax numerators (node:355-357), exact 2^-40 integers ------------------------------
    unsigned key0 = (unsigned)((volatile unsigned long long *)a.keymax)[0];
    unsigned key1 = MH ? (unsigned)((volatile unsigned long long *)a.keymax)[1] : 0u;
    if (SH) {                                            // exchange 1: the population's score maxima
        if (t == 0) { sh.xpay[0] = key0; sh.xpay[1] = key1; }
        tl_exchange(a, sh, 1, 2, true);
        for (int r = 0; r < a.world; ++r) {
            key0 = max(key0, (unsigned)sh.xch[r * 16]);
            key1 = max(key1, (unsigned)sh.xch[r * 16 + 1]);
        }
    }
    const float m_post = key0 ? mcl_float_of_key(key0) : -FLT_MAX;
    const float m_pre = key1 ? mcl_float_of_key(key1) : -FLT_MAX;
    for (int v = b; v < a.nt; v += G) {                  // look-back records of both passes start empty
        if (t == 0) { a.st1[v] = 0ull; a.st2[v] = 0ull; }
    }
    const bool raw = a.raw == 1, given = a.raw == 2;
    tl_stamp(a, 0);
    if (!raw && !given) {
        unsigned long long acc0 = 0, acc1 = 0;
        for (int v = b; v < a.nt; v += G) {
            const int64_t base = (int64_t)v * a.tile + t;
            for (int k = 0; k < a.ipt; ++k) {
                const int64_t i = base + (int64_t)k * TL_THREADS;
                if (i < n) {
                    acc0 += __double2ull_rz(__dmul_rn((double)tl_softmax_num(a.s_post[i], m_post), TL_SOFTMAX_FIX));
                    if (MH) acc1 += __double2ull_rz(__dmul_rn((double)tl_softmax_num(a.s_pre[i], m_pre), TL_SOFTMAX_FIX));
                }
            }
        }
        tl_block_sum_u64x2(acc0, acc1, sh);
        if (t == 0) { a.part_q[b] = acc0; a.part_q[G + b] = acc1; }
    }
    if (a.stop == 1) return;
    tl_stamp(a, 1);
    tl_grid_barrier(a, 1);
    tl_stamp(a, 2);
    float sum_post = 1.0f, sum_pre = 1.0f;
    if (!raw && !given) {
        unsigned long long q0 = 0, q1 = 0;
        for (int u = t; u < G; u += TL_THREADS) { q0 += __ldcg(a.part_q + u); q1 += __ldcg(a.part_q + G + u); }
        tl_block_sum_u64x2(q0, q1, sh);
        if (SH) {                                        // exchange 2: the exact softmax sums
            if (t == 0) { sh.xpay[0] = q0; sh.xpay[1] = q1; }
            tl_exchange(a, sh, 2, 2, true);
            q0 = 0; q1 = 0;
            for (int r = 0; r < a.world; ++r) { q0 += sh.xch[r * 16]; q1 += sh.xch[r * 16 + 1]; }
        }
        // sum as f32 of the exact integer (node:356-357: f32 divide by the f32-rounded sum)
        sum_post = (float)((double)q0 / TL_SOFTMAX_FIX);
        sum_pre = MH ? (float)((double)q1 / TL_SOFTMAX_FIX) : 1.0f;
        if (b == 0 && t == 0) { a.keymax[0] = 0ull; a.keymax[1] = 0ull; }    // every CTA has read the keys
    } else if (SH) {
        tl_exchange(a, sh, 2, 0, true);                  // given weights: the exchange is made empty (parities stay in step)
    }

    // ---- S2: weights (node:356-357), MH accept (pu:229-233), raw estimate sums (node:586-589), weight maximum ---
    for (int v = b; v < a.nt; v += G) {
        double s6[6] = {0, 0, 0, 0, 0, 0};
        float wmax = 0.0f;
        const int64_t base = (int64_t)v * a.tile + t;
#pragma unroll 2
        for (int k = 0; k < a.ipt; ++k) {
            const int64_t i = base + (int64_t)k * TL_THREADS;
            if (i >= n) break;
            if (raw) {
                const float w = a.w_out[i];
                s6[0] += (double)w;
                wmax = fmaxf(wmax, w);
                continue;
            }
            if (given) {                                 // weights and particles are there: estimate sums only
                const float w = a.w_out[i];
                const double wi = (double)w, x = a.nx[i], y = a.ny[i];
                double sn, cs;
                sincos(a.nth[i], &sn, &cs);
                s6[0] += wi; s6[1] += wi * wi; s6[2] += wi * x; s6[3] += wi * y; s6[4] += wi * cs; s6[5] += wi * sn;
                wmax = fmaxf(wmax, w);
                continue;
            }
            const float p_new = __fdiv_rn(tl_softmax_num(a.s_post[i], m_post), sum_post);
            double x = a.px[i], y = a.py[i], th = a.pt[i];
            float w = p_new;
            if (MH) {
                const float p_old = __fdiv_rn(tl_softmax_num(a.s_pre[i], m_pre), sum_pre);
                double alpha = 1.0;
                if (p_old > 0.f) {
                    const double q = (double)__fdiv_rn(p_new, p_old);          // f32 divide, promoted
                    alpha = (q < 1.0) ? q : 1.0;
                }
                const uint4 o = philox_draw4(a.seed, a.step, a.first_index + (uint64_t)i, 0u, MCL_STREAM_MH);
                const bool acc = u53_from(o.x, o.y) < alpha;
                if (!acc) { x = a.ox[i]; y = a.oy[i]; th = a.ot[i]; w = p_old; }
                a.nx[i] = x; a.ny[i] = y; a.nth[i] = th;
            }
            a.w_out[i] = w;
            const double wi = (double)w;
            double sn, cs;
            sincos(th, &sn, &cs);
            s6[0] += wi; s6[1] += wi * wi; s6[2] += wi * x; s6[3] += wi * y; s6[4] += wi * cs; s6[5] += wi * sn;
            wmax = fmaxf(wmax, w);
        }
        tl_block_sum<6>(s6, sh);
        wmax = tl_block_max(wmax, sh);
        if (t == 0) {
            double *pm = a.part_m + (size_t)v * 8;
#pragma unroll
            for (int k = 0; k < 6; ++k) pm[k] = s6[k];
            pm[6] = (double)wmax;
        }
    }
    if (a.stop == 2) return;
    tl_stamp(a, 3);
    tl_grid_barrier(a, 2);
    tl_stamp(a, 4);

    // ---- S3: population sums (same order in every CTA), means, central sums; first resampling pass ------------
    double mx, my, mt, scale = 0.0, rank_prefix = 0.0;     // rank_prefix: fp64 sum of the weights of the ranks before this one
    {
        double tv[6] = {0, 0, 0, 0, 0, 0};
        float wm = 0.0f;
        for (int u = t; u < a.nt; u += TL_THREADS) {
            const double *pm = a.part_m + (size_t)u * 8;
#pragma unroll
            for (int k = 0; k < 6; ++k) tv[k] += __ldcg(pm + k);
            wm = fmaxf(wm, (float)__ldcg(pm + 6));
        }
        tl_block_sum<6>(tv, sh);
        wm = tl_block_max(wm, sh);
        if (SH) {                                        // exchange 3: raw estimate sums (rank order) + weight maximum
            if (t == 0) {
#pragma unroll
                for (int k = 0; k < 6; ++k) sh.xpay[k] = (unsigned long long)__double_as_longlong(tv[k]);
                sh.xpay[6] = (unsigned long long)__double_as_longlong((double)wm);
            }
            tl_exchange(a, sh, 3, 7, true);
#pragma unroll
            for (int k = 0; k < 6; ++k) {
                double acc = __longlong_as_double((long long)sh.xch[k]);
                for (int r = 1; r < a.world; ++r) acc += __longlong_as_double((long long)sh.xch[r * 16 + k]);
                tv[k] = acc;
            }
            for (int r = 0; r < a.world; ++r) wm = fmaxf(wm, (float)__longlong_as_double((long long)sh.xch[r * 16 + 6]));
            for (int r = 0; r < a.rank; ++r) rank_prefix += __longlong_as_double((long long)sh.xch[r * 16]);
        }
        mx = tv[2] / tv[0]; my = tv[3] / tv[0]; mt = atan2(tv[5], tv[4]);      // np.average; arctan2(sin, cos)
        if (!REF) {
            int e = 0;
            if (wm > 0.0f) frexp((double)wm, &e);
            int lg = 0;
            while (((int64_t)1 << lg) < (SH ? a.n_global : n)) ++lg;
            scale = ldexp(1.0, 62 - lg - e);               // resample.cu k_wmax, oracle orc_resample_scale
        }
        if (b == 0 && t == 0 && !raw) {
#pragma unroll
            for (int k = 0; k < 6; ++k) a.est18[k] = tv[k];
            a.est18[6] = mx; a.est18[7] = my; a.est18[8] = mt;
        }
        if (REF) {
            // fp64 prefix of the tile sums of w: the prediction the exact passes start from
            double carry = 0.0;
            for (int base = 0; base < a.nt; base += TL_THREADS) {
                const int u = base + t;
                const double val = u < a.nt ? __ldcg(a.part_m + (size_t)u * 8) : 0.0;
                double tot;
                const double ex = tl_block_excl_scan_d(val, tot, sh);
                if (u < a.nt && u % G == b) { sh.ownP[u / G] = rank_prefix + carry + ex; sh.ownT[u / G] = val; }
                carry += tot;
            }
            __syncthreads();
        }
    }
    if (a.stop == 31) return;
    tl_stamp(a, 5);
    // central sums: in reference mode the CTAs whose first tile is far down the look-back chain compute them while
    // they wait for it (inside the first exact pass), the first few after their pass; fixed point: here
    const bool central_in_pass = REF && !raw && b >= TL_CENTRAL_EARLY;
    if (!raw && !REF) tl_central_sums(a, sh, mx, my, mt);
    tl_stamp(a, 6);
    if (REF) {
        tl_exact_pass<false>(a, sh, 1.0f, dyn, central_in_pass, mx, my, mt, SH ? 4 : 0);   // pu:430 np.sum(weights): sequential f32
        if (!raw && !central_in_pass) tl_central_sums(a, sh, mx, my, mt);
    } else {
        for (int v = b; v < a.nt; v += G) {
            const int64_t first = (int64_t)v * a.tile + (int64_t)t * a.ipt, end = min(n, (int64_t)(v + 1) * a.tile);
            unsigned long long s = 0, dummy = 0;
            for (int k = 0; k < a.ipt; ++k)
                if (first + k < end) s += tl_quantise(a.w_out[first + k], scale);
            tl_block_sum_u64x2(s, dummy, sh);
            if (t == 0) a.ttot[v] = s;
        }
    }
    if (a.stop == 3) return;
    tl_stamp(a, 7);
    tl_grid_barrier(a, 3);
    tl_stamp(a, 8);

    // ---- S4: running sums of the normalised weights (reference) / of the quantised weights (fixed) ------------
    double totd = 0.0;
    if (REF) {
        float S = __ldcg(&a.hd->S);
        if (SH) {                                        // exchange 4 (handed from rank to rank during pass 1): the
            if (tl_warp() == 0) tl_wait_all(a, sh, 4, 1);    // population's sum is what left the LAST rank
            __syncthreads();
            S = __uint_as_float((unsigned)sh.xch[(a.world - 1) * 16]);
            __syncthreads();
        }
        tl_exact_pass<true>(a, sh, S, dyn, false, 0.0, 0.0, 0.0, SH ? 5 : 0);
    } else {
        unsigned long long carry = 0;
        for (int base = 0; base < a.nt; base += TL_THREADS) {
            const int u = base + t;
            const unsigned long long val = u < a.nt ? __ldcg(a.ttot + u) : 0ull;
            unsigned long long tot;
            const unsigned long long ex = tl_block_excl_scan_u64(val, tot, sh);
            if (u < a.nt && u % G == b) sh.ownE[u / G] = carry + ex;
            carry += tot;
        }
        __syncthreads();
        if (a.stop == 41) return;
        totd = (double)carry;
        if (b == 0 && t == 0) a.hd->total = carry;
        if (SH) {                                        // exchange 4: central sums (rank order) + every rank's total
            double tc[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
            if (b == 0) {
                for (int u = t; u < a.nt; u += TL_THREADS) {
                    const double *pc = a.part_c + (size_t)u * 9;
#pragma unroll
                    for (int k = 0; k < 9; ++k) tc[k] += __ldcg(pc + k);
                }
                tl_block_sum<9>(tc, sh);
            }
            if (t == 0) {
#pragma unroll
                for (int k = 0; k < 9; ++k) sh.xpay[k] = (unsigned long long)__double_as_longlong(tc[k]);
                sh.xpay[9] = carry;
            }
            tl_exchange(a, sh, 4, 10, true);
            if (b == 0 && t < 9) {
                double acc = __longlong_as_double((long long)sh.xch[t]);
                for (int r = 1; r < a.world; ++r) acc += __longlong_as_double((long long)sh.xch[r * 16 + t]);
                a.est18[9 + t] = acc;
            }
            if (tl_warp() == 0) {                        // which output slots this rank emits (sharded.plan_resample)
                const int lane = tl_lane();
                if (lane == 0) {
                    unsigned long long acc = 0;
                    for (int r = 0; r < a.world; ++r) { sh.poff[r] = acc; acc += sh.xch[r * 16 + 9]; }
                    sh.poff[a.world] = acc;
                }
                __syncwarp();
                const double gt = (double)sh.poff[a.world];
                if (lane <= a.world) sh.pcnt[lane] = tl_count_thresholds_le(sh.poff[lane], a.r, a.rstep, gt, (long long)a.n_global);
                __syncwarp();
                if (lane == 0) {
                    long long prev_hi = 0, my_lo = 0, my_hi = 0;
                    for (int r = 0; r < a.world; ++r) {
                        long long lo = r == 0 ? 0 : sh.pcnt[r];
                        long long hi = r == a.world - 1 ? (long long)a.n_global : sh.pcnt[r + 1];
                        if (hi < lo) hi = lo;
                        if (r > 0) { if (lo < prev_hi) lo = prev_hi; if (hi < lo) hi = lo; }
                        if (r == a.world - 1) hi = (long long)a.n_global;
                        if (r == a.rank) { my_lo = lo; my_hi = hi; }
                        prev_hi = hi;
                    }
                    sh.m_lo = my_lo; sh.m_hi = my_hi;
                }
            }
            __syncthreads();
            totd = (double)sh.poff[a.world];
        }
        unsigned long long *C = (unsigned long long *)a.C;
        int rd = 0;
        for (int v = b; v < a.nt; v += G, ++rd) {
            const int64_t first = (int64_t)v * a.tile + (int64_t)t * a.ipt, end = min(n, (int64_t)(v + 1) * a.tile);
            unsigned long long q[TL_MAX_IPT], s = 0;
#pragma unroll
            for (int k = 0; k < TL_MAX_IPT; ++k) {
                q[k] = (k < a.ipt && first + k < end) ? tl_quantise(a.w_out[first + k], scale) : 0ull;
                s += q[k];
            }
            unsigned long long tot;
            unsigned long long run = sh.ownE[rd] + tl_block_excl_scan_u64(s, tot, sh);
#pragma unroll
            for (int k = 0; k < TL_MAX_IPT; ++k) {
                run += q[k];
                if (k < a.ipt && first + k < end) C[first + k] = run;
            }
            if (t == 0) ((unsigned long long *)a.tend)[v] = sh.ownE[rd] + tot;
        }
    }
    if (b == G - 1 && !raw && !SH) {                     // central sums of the population -> est18[9..17]
        double tc[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};      // (after this CTA's own tiles: off the look-back chain)
        for (int u = t; u < a.nt; u += TL_THREADS) {
            const double *pc = a.part_c + (size_t)u * 9;
#pragma unroll
            for (int k = 0; k < 9; ++k) tc[k] += __ldcg(pc + k);
        }
        tl_block_sum<9>(tc, sh);
        if (t == 0)
#pragma unroll
            for (int k = 0; k < 9; ++k) a.est18[9 + k] = tc[k];
    }
    if (a.stop == 4) return;
    tl_stamp(a, 9);
    tl_grid_barrier(a, 4);
    tl_stamp(a, 10);
    if (SH && REF) {
        // exchange 5 (handed from rank to rank during pass 2): the exact running sum at the end of every rank = the
        // stretch of the population's cumulative weight each rank owns.  Exchange 6: central sums, rank order.
        if (tl_warp() == 0) tl_wait_all(a, sh, 5, 1);
        __syncthreads();
        if (t <= a.world) sh.poff[t] = t == 0 ? 0ull : (sh.xch[(t - 1) * 16] & 0xffffffffull);     // float bits of the bounds
        double tc[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
        if (b == 0) {
            for (int u = t; u < a.nt; u += TL_THREADS) {
                const double *pc = a.part_c + (size_t)u * 9;
#pragma unroll
                for (int k = 0; k < 9; ++k) tc[k] += __ldcg(pc + k);
            }
            tl_block_sum<9>(tc, sh);
        }
        if (t == 0)
#pragma unroll
            for (int k = 0; k < 9; ++k) sh.xpay[k] = (unsigned long long)__double_as_longlong(tc[k]);
        tl_exchange(a, sh, 6, 9, true);
        if (b == 0 && t < 9) {
            double acc = __longlong_as_double((long long)sh.xch[t]);
            for (int r = 1; r < a.world; ++r) acc += __longlong_as_double((long long)sh.xch[r * 16 + t]);
            a.est18[9 + t] = acc;
        }
        if (tl_warp() == 0) {                            // output slots m with bound[rank-1] < U_m <= bound[rank] (pu:441)
            const int lane = tl_lane();
            if (lane <= a.world) {
                const double cb = (double)__uint_as_float((unsigned)sh.poff[lane]);
                long long lo = 0, hi = (long long)a.n_global;
                while (lo < hi) {                        // number of slots with U_m <= cb
                    const long long mid = (lo + hi) >> 1;
                    if (__dadd_rn(a.r, __dmul_rn((double)mid, a.rstep)) <= cb) lo = mid + 1; else hi = mid;
                }
                sh.pcnt[lane] = lo;
            }
            __syncwarp();
            if (lane == 0) {
                long long prev_hi = 0, my_lo = 0, my_hi = 0;
                for (int r = 0; r < a.world; ++r) {
                    long long lo = r == 0 ? 0 : sh.pcnt[r];
                    long long hi = r == a.world - 1 ? (long long)a.n_global : sh.pcnt[r + 1];     // pu:441 "i < N - 1": the rest
                    if (hi < lo) hi = lo;
                    if (r > 0) { if (lo < prev_hi) lo = prev_hi; if (hi < lo) hi = lo; }
                    if (r == a.rank) { my_lo = lo; my_hi = hi; }
                    prev_hi = hi;
                }
                sh.m_lo = my_lo; sh.m_hi = my_hi;
            }
        }
        __syncthreads();
    }

    // ---- S5: idx[m] = min(first i with C_i >= key_m, n - 1) (pu:439-444 as a search) + pu:445 gather ----------
    // Sharded: this rank emits the output slots [m_lo, m_hi) whose thresholds fall into its own stretch of the
    // population's cumulative weight and stores every offspring straight into the destination rank's set.
    {
        const CT *C = (const CT *)a.C;
        const int64_t limit = n - 1;
        const bool coarse = a.nt <= TL_COARSE_MAX;
        CT *tendS = (CT *)dyn;
        CT *stage = (CT *)(dyn + TL_COARSE_MAX * sizeof(unsigned long long));
        const int stage_cap = (int)(TL_STAGE_BYTES / sizeof(CT));
        if (coarse) {
            for (int u = t; u < a.nt; u += TL_THREADS) tendS[u] = __ldcg((const CT *)a.tend + u);
            __syncthreads();
        }
        const int warp = tl_warp(), lane = tl_lane();
        const int64_t out_lo = SH ? (int64_t)sh.m_lo : 0, out_hi = SH ? (int64_t)sh.m_hi : n;
        const unsigned long long koff = (SH && !REF) ? sh.poff[a.rank] : 0ull;     // fixed point: cumulative weight of the ranks before this one
        const int64_t n_per_rank = n;
        if (SH && t < 3 * a.world) sh.xch[t] = a.peer_pose[t];
        if (SH) __syncthreads();
        const int ntile_out = (int)((out_hi - out_lo + a.tile - 1) / a.tile);
        for (int ov = b; ov < ntile_out; ov += G) {
            const int64_t m0 = out_lo + (int64_t)ov * a.tile, m1 = min(out_hi, m0 + a.tile);
            if (warp < 2) {
                KeyT key = tl_key<REF>(warp == 0 ? m0 : m1 - 1, a.r, a.rstep, totd);
                if (SH && !REF) key = (KeyT)((unsigned long long)key > koff ? (unsigned long long)key - koff : 0ull);
                int64_t lo = 0, hi = limit;
                if (coarse) {
                    int tl = 0, th = a.nt - 1;               // first tile whose last running sum reaches the key
                    while (tl < th) {
                        const int mid = (tl + th) >> 1;
                        if (TlCum<REF>::gt(key, tendS[mid])) tl = mid + 1; else th = mid;
                    }
                    lo = (int64_t)tl * a.tile;
                    hi = min(limit, lo + a.tile - 1);
                }
                const int64_t i = tl_warp_search<REF>(C, lo, hi, key);
                if (lane == 0) sh.irange[warp] = i;
            }
            __syncthreads();
            const int64_t i_lo = sh.irange[0], i_hi = sh.irange[1];
            const int64_t cnt = i_hi - i_lo + 1;
            // the running sums of the source range go to shared memory; a range that does not fit (very uneven
            // weights: a freshly initialised cloud) is staged as every stride-th sum -- the last of each block of
            // `stride` sources, i.e. the block's largest -- and the search ends inside one block in global memory
            const bool staged = cnt <= stage_cap;
            const int64_t stride = staged ? 1 : (cnt + stage_cap - 1) / stage_cap;
            const int nblocks = (int)((cnt + stride - 1) / stride);
            for (int64_t j = t; j < nblocks; j += TL_THREADS) stage[j] = __ldcg(C + i_lo + min(cnt - 1, (j + 1) * stride - 1));
            __syncthreads();
            // four output slots at a time: the searches first, then all twelve gather loads in flight together, then
            // the stores (slot by slot the loads of one slot waited for the stores of the one before)
            for (int k0 = 0; k0 < a.ipt; k0 += 4) {
                int64_t src[4], mm[4];
                bool ok[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const int k = k0 + u;
                    const int64_t m = m0 + (int64_t)k * TL_THREADS + t;
                    mm[u] = m; ok[u] = k < a.ipt && m < m1; src[u] = 0;
                    if (!ok[u]) continue;
                    KeyT key = tl_key<REF>(m, a.r, a.rstep, totd);
                    if (SH && !REF) key = (KeyT)((unsigned long long)key > koff ? (unsigned long long)key - koff : 0ull);
                    if (staged) {
                        int lo = 0, hi = (int)cnt - 1;
                        while (lo < hi) {
                            const int mid = (lo + hi) >> 1;
                            if (TlCum<REF>::gt(key, stage[mid])) lo = mid + 1; else hi = mid;
                        }
                        src[u] = i_lo + lo;
                    } else {
                        int bl = 0, bh = nblocks - 1;                  // first block whose last sum reaches the key
                        while (bl < bh) {
                            const int mid = (bl + bh) >> 1;
                            if (TlCum<REF>::gt(key, stage[mid])) bl = mid + 1; else bh = mid;
                        }
                        int64_t lo = i_lo + (int64_t)bl * stride, hi = min(i_hi, lo + stride - 1);
                        while (lo < hi) {
                            const int64_t mid = (lo + hi) >> 1;
                            if (TlCum<REF>::gt(key, __ldcg(C + mid))) lo = mid + 1; else hi = mid;
                        }
                        src[u] = lo;
                    }
                }
                double vx[4], vy[4], vt[4];
                if (SH || !raw) {
#pragma unroll
                    for (int u = 0; u < 4; ++u)
                        if (ok[u]) { vx[u] = a.nx[src[u]]; vy[u] = a.ny[src[u]]; vt[u] = a.nth[src[u]]; }
                }
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    if (!ok[u]) continue;
                    const int64_t m = mm[u];
                    if (SH) {
                        int64_t d = m / n_per_rank;
                        if (d > a.world - 1) d = a.world - 1;
                        const int64_t j = m - d * n_per_rank;
                        reinterpret_cast<double *>(sh.xch[d])[j] = vx[u];
                        reinterpret_cast<double *>(sh.xch[a.world + d])[j] = vy[u];
                        reinterpret_cast<double *>(sh.xch[2 * a.world + d])[j] = vt[u];
                    } else {
                        a.idx[m] = (int32_t)src[u];
                        if (!raw) { a.gx[m] = vx[u]; a.gy[m] = vy[u]; a.gt[m] = vt[u]; }
                    }
                }
            }
            __syncthreads();
        }
    }
    if (SH) {                                            // exchange 5: every rank's offspring have landed everywhere
        __threadfence_system();
        tl_grid_barrier(a, 5);
        tl_exchange(a, sh, REF ? 7 : 5, 0, false);
    }
    tl_stamp(a, 11);
}

// ------------------------------------------------------------------------------------------------ MH chain iteration
// One iteration of the MH refinement chain (filter.cu mcl_filter_update_chain, BASELINE config 4) after its
// likelihood launch: weights of the proposal and of the chain from two separately normalised softmaxes
// (node:254-270), accept (pu:229-233), carried score and its maximum for the next iteration -- one cooperative
// launch with one grid barrier; sharded: the two exchanges (score maxima, exact softmax sums) inside.
//   a.s_post = proposal scores, a.score_chain = carried scores, (px..) proposal, (ox..) chain, (nx..) result
template <bool SH>
__global__ void __launch_bounds__(TL_THREADS, 1) k_chain_tail(const TailArgs a) {
    __shared__ TlShared sh;
    const int t = threadIdx.x, G = gridDim.x, b = blockIdx.x;
    const int64_t n = a.n;
    unsigned key0 = (unsigned)((volatile unsigned long long *)a.keymax)[0];
    unsigned key1 = (unsigned)((volatile unsigned long long *)a.ckey_in)[0];
    if (SH) {
        if (t == 0) { sh.xpay[0] = key0; sh.xpay[1] = key1; }
        tl_exchange(a, sh, 1, 2, true);
        for (int r = 0; r < a.world; ++r) {
            key0 = max(key0, (unsigned)sh.xch[r * 16]);
            key1 = max(key1, (unsigned)sh.xch[r * 16 + 1]);
        }
    }
    const float m_prop = key0 ? mcl_float_of_key(key0) : -FLT_MAX;
    const float m_chain = key1 ? mcl_float_of_key(key1) : -FLT_MAX;
    {
        unsigned long long acc0 = 0, acc1 = 0;
        for (int64_t i = (int64_t)b * TL_THREADS + t; i < n; i += (int64_t)G * TL_THREADS) {
            acc0 += __double2ull_rz(__dmul_rn((double)tl_softmax_num(a.s_post[i], m_prop), TL_SOFTMAX_FIX));
            acc1 += __double2ull_rz(__dmul_rn((double)tl_softmax_num(a.score_chain[i], m_chain), TL_SOFTMAX_FIX));
        }
        tl_block_sum_u64x2(acc0, acc1, sh);
        if (t == 0) { a.part_q[b] = acc0; a.part_q[G + b] = acc1; }
    }
    tl_grid_barrier(a, 1);
    unsigned long long q0 = 0, q1 = 0;
    for (int u = t; u < G; u += TL_THREADS) { q0 += __ldcg(a.part_q + u); q1 += __ldcg(a.part_q + G + u); }
    tl_block_sum_u64x2(q0, q1, sh);
    if (SH) {
        if (t == 0) { sh.xpay[0] = q0; sh.xpay[1] = q1; }
        tl_exchange(a, sh, 2, 2, true);
        q0 = 0; q1 = 0;
        for (int r = 0; r < a.world; ++r) { q0 += sh.xch[r * 16]; q1 += sh.xch[r * 16 + 1]; }
    }
    if (b == 0 && t == 0) { a.keymax[0] = 0ull; a.ckey_in[0] = 0ull; }     // every CTA has read both keys
    const float sum_prop = (float)((double)q0 / TL_SOFTMAX_FIX), sum_chain = (float)((double)q1 / TL_SOFTMAX_FIX);
    float smax = -FLT_MAX;
    for (int64_t i = (int64_t)b * TL_THREADS + t; i < n; i += (int64_t)G * TL_THREADS) {
        const float s_prop = a.s_post[i], s_chain = a.score_chain[i];
        const float p_new = __fdiv_rn(tl_softmax_num(s_prop, m_prop), sum_prop);
        const float p_old = __fdiv_rn(tl_softmax_num(s_chain, m_chain), sum_chain);
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
        if (acc) a.score_chain[i] = s_prop;
        smax = fmaxf(smax, acc ? s_prop : s_chain);
    }
    smax = tl_block_max(smax, sh);
    if (t == 0 && smax > -FLT_MAX) atomicMax(a.ckey_out, (unsigned long long)mcl_key_of_float(smax));
}

// ------------------------------------------------------------------------------------------------ host side
static size_t tl_align(size_t v) { return (v + 255) & ~(size_t)255; }

struct TailPlan {
    int ipt, tile, nt, grid;
    size_t o_drift, o_q, o_m, o_c, o_st1, o_st2, o_C, o_tend, o_ttot, bytes;
};
static TailPlan tail_plan(const mcl_handle *h, int64_t n) {
    TailPlan p;
    p.ipt = TL_MAX_IPT;
    for (int k = 1; k <= TL_MAX_IPT; ++k)
        if ((n + (int64_t)TL_THREADS * k - 1) / ((int64_t)TL_THREADS * k) <= h->sm_count) { p.ipt = k; break; }
    // a few thousand particles (the reference's own scale): ONE tile, whose exact sums a single thread adds in order
    if (n <= TL_SEQ_MAX) p.ipt = (int)std::max<int64_t>(1, (n + TL_THREADS - 1) / TL_THREADS);
    p.tile = TL_THREADS * p.ipt;
    p.nt = (int)((n + p.tile - 1) / p.tile);
    p.grid = std::min(p.nt, h->sm_count);
    size_t off = tl_align(sizeof(TailHeader));
    p.o_drift = off; off += tl_align((size_t)h->sm_count * TL_MAX_ROUNDS * 8);     // fixed place: survives a re-plan
    p.o_q = off; off += tl_align((size_t)2 * h->sm_count * 8);
    p.o_m = off; off += tl_align((size_t)p.nt * 8 * 8);
    p.o_c = off; off += tl_align((size_t)p.nt * 9 * 8);
    p.o_st1 = off; off += tl_align((size_t)p.nt * 8);
    p.o_st2 = off; off += tl_align((size_t)p.nt * 8);
    p.o_tend = off; off += tl_align((size_t)p.nt * 8);
    p.o_ttot = off; off += tl_align((size_t)p.nt * 8);
    p.o_C = off; off += tl_align((size_t)n * 8);
    p.bytes = off;
    return p;
}

static thread_local int g_tail_raw = 0;     // set by the test hook around its launch

static int tail_disabled() {
    static int off = -1;
    if (off < 0) { const char *e = getenv("MCL_NO_TAIL"); off = (e && atoi(e)) ? 1 : 0; }
    return off;
}

// can the persistent tail run this step?  (cooperative launch, every tile's round fits the per-CTA tables)
bool mcl_tail_available(mcl_handle *h, int64_t n) {
    if (tail_disabled() || n <= 0 || n > 0x7fffffffLL) return false;
    if (h->coop_launch < 0) {
        int v = 0;
        cudaDeviceGetAttribute(&v, cudaDevAttrCooperativeLaunch, h->device);
        h->coop_launch = v ? 1 : 0;
    }
    if (!h->coop_launch) return false;
    const TailPlan p = tail_plan(h, n);
    return (p.nt + p.grid - 1) / p.grid <= TL_MAX_ROUNDS;
}

static int tail_prepare(mcl_handle *h, int64_t n) {
    const TailPlan p = tail_plan(h, n);
    if (p.bytes <= h->tail_bytes) return MCL_OK;
    MCL_CUDA(h, cudaStreamSynchronize(h->stream));
    cudaFree(h->d_tail);
    h->d_tail = nullptr; h->tail_bytes = 0;
    MCL_CUDA(h, cudaMalloc(&h->d_tail, p.bytes + p.bytes / 4));
    MCL_CUDA(h, cudaMemset(h->d_tail, 0, p.o_q));      // header + drift hints (0 = no drift known)
    h->tail_bytes = p.bytes + p.bytes / 4;
    h->tail_bar = 0;
    return MCL_OK;
}

template <bool MH, bool REF, bool SH>
static cudaError_t tail_launch(const TailArgs &a, int grid, size_t dyn, cudaStream_t s) {
    cudaFuncSetAttribute(k_tail<MH, REF, SH>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn);
    void *params[] = {(void *)&a};
    return cudaLaunchCooperativeKernel((const void *)k_tail<MH, REF, SH>, dim3(grid), dim3(TL_THREADS), params, dyn, s);
}

// softmax -> (MH) -> estimate sums -> resampling of one step; see the head of this file.  u.(nx, ny, nth) receives
// the MH result (without MH it must be the particles themselves), (gx, gy, gt) the resampled set.
int mcl_tail_step(mcl_handle *h, const FusedStep &u, unsigned long long *d_keymax, int resample_mode, double r,
                  int32_t *idx, double *gx, double *gy, double *gt, const TailComm *comm) {
    if (comm && resample_mode != MCL_RESAMPLE_FIXED_POINT && resample_mode != MCL_RESAMPLE_REFERENCE_F32)
        return mcl_fail(h, MCL_ERR_ARG, "mcl_tail_step: unknown resampling arithmetic");
    int rc = tail_prepare(h, u.n);
    if (rc) return rc;
    const TailPlan p = tail_plan(h, u.n);
    char *b = (char *)h->d_tail;
    TailArgs a;
    memset(&a, 0, sizeof(a));
    a.hd = (TailHeader *)b; a.keymax = d_keymax; a.bar_base = h->tail_bar;
    a.n = u.n; a.nt = p.nt; a.ipt = p.ipt; a.tile = p.tile; a.use_mh = u.use_mh; a.raw = g_tail_raw;
    { const char *e = getenv("MCL_TAIL_STOP"); a.stop = e ? atoi(e) : 0; }
    {
        static int prof = -1;
        if (prof < 0) { const char *e = getenv("MCL_TAIL_PROF"); prof = (e && atoi(e)) ? 1 : 0; }
        if (prof) {
            if (!h->d_tail_prof) MCL_CUDA(h, cudaMalloc((void **)&h->d_tail_prof, (size_t)1024 * 32 * 8));
            a.prof = (unsigned long long *)h->d_tail_prof;
            h->tail_prof_grid = p.grid;
        }
    }
    a.s_post = u.s_post; a.s_pre = u.s_pre; a.w_out = u.w_out;
    a.px = u.px; a.py = u.py; a.pt = u.pt; a.ox = u.ox; a.oy = u.oy; a.ot = u.ot; a.nx = u.nx; a.ny = u.ny; a.nth = u.nth;
    a.seed = u.seed; a.step = u.step; a.first_index = u.first_index;
    a.part_q = (unsigned long long *)(b + p.o_q); a.part_m = (double *)(b + p.o_m); a.part_c = (double *)(b + p.o_c);
    a.st1 = (unsigned long long *)(b + p.o_st1); a.st2 = (unsigned long long *)(b + p.o_st2);
    a.C = b + p.o_C; a.tend = b + p.o_tend; a.ttot = (unsigned long long *)(b + p.o_ttot);
    a.est18 = u.est18;
    a.drift = (double *)(b + p.o_drift);
    if (h->tail_drift_n != u.n) {                    // another particle count: the tiles are other tiles
        MCL_CUDA(h, cudaMemsetAsync(a.drift, 0, (size_t)h->sm_count * TL_MAX_ROUNDS * 8, h->stream));
        h->tail_drift_n = u.n;
    }
    a.r = r; a.rstep = 1.0 / (double)(comm ? comm->n_global : u.n);     // pu:434
    if (comm) {
        a.rank = comm->rank; a.world = comm->world; a.n_global = comm->n_global;
        a.mailbox = comm->mailbox;
        for (int d = 0; d < 16; ++d) a.peers[d] = comm->peers[d];
        a.epoch0 = comm->epoch0; a.comm_err = comm->d_err; a.peer_pose = comm->d_peer_pose_dst;
    }
    a.idx = idx; a.gx = gx; a.gy = gy; a.gt = gt;
    const size_t dyn = (size_t)TL_COARSE_MAX * 8 + TL_STAGE_BYTES;
    const bool ref = resample_mode == MCL_RESAMPLE_REFERENCE_F32;
    cudaError_t e;
    if (comm && ref) e = u.use_mh ? tail_launch<true, true, true>(a, p.grid, dyn, h->stream) : tail_launch<false, true, true>(a, p.grid, dyn, h->stream);
    else if (comm) e = u.use_mh ? tail_launch<true, false, true>(a, p.grid, dyn, h->stream) : tail_launch<false, false, true>(a, p.grid, dyn, h->stream);
    else if (u.use_mh) e = ref ? tail_launch<true, true, false>(a, p.grid, dyn, h->stream) : tail_launch<true, false, false>(a, p.grid, dyn, h->stream);
    else e = ref ? tail_launch<false, true, false>(a, p.grid, dyn, h->stream) : tail_launch<false, false, false>(a, p.grid, dyn, h->stream);
    if (e != cudaSuccess) return mcl_fail(h, MCL_ERR_CUDA, std::string("k_tail launch: ") + cudaGetErrorString(e));
    h->launches++;
    h->tail_bar += (unsigned long long)(comm ? TL_NBAR + 1 : TL_NBAR) * p.grid;
    return MCL_OK;
}

// estimate sums + resampling of particles whose weights are there already (the MH chain's result): the tail kernel
// without its softmax / accept stages (sharded: the exchanges of those stages are still made, empty, so that the
// mailbox protocol stays in step).
int mcl_tail_finish(mcl_handle *h, int64_t n, int64_t n_global, float *d_w, double *d_x, double *d_y, double *d_th,
                    double *d_est18, int resample_mode, double r, int32_t *idx, double *gx, double *gy, double *gt,
                    const TailComm *comm) {
    FusedStep u;
    memset(&u, 0, sizeof(u));
    u.n = n; u.n_global = n_global; u.use_mh = 0; u.w_out = d_w; u.nx = d_x; u.ny = d_y; u.nth = d_th; u.est18 = d_est18;
    g_tail_raw = 2;
    const int rc = mcl_tail_step(h, u, mcl_fused_keymax(h), resample_mode, r, idx, gx, gy, gt, comm);
    g_tail_raw = 0;
    return rc;
}

// device address of the sticky error word of the tail kernel (0 = ok), or NULL before the first tail step
const int *mcl_tail_err_ptr(mcl_handle *h) {
    return h->d_tail ? &reinterpret_cast<TailHeader *>(h->d_tail)->err : nullptr;
}

// blocking: *err = 0 ok, != 0 a spin loop of the tail kernel timed out (the step's results are invalid)
extern "C" int mcl_tail_status(mcl_handle *h, int *err) {
    if (!h || !err) return MCL_ERR_ARG;
    *err = 0;
    if (!h->d_tail) return MCL_OK;
    DeviceGuard guard(h->device);
    MCL_CUDA(h, cudaStreamSynchronize(h->stream));
    TailHeader hd;
    MCL_CUDA(h, cudaMemcpy(&hd, h->d_tail, sizeof(hd), cudaMemcpyDeviceToHost));
    *err = hd.err;
    if (hd.err) {       // the barrier counter is out of step after a time-out: start over
        MCL_CUDA(h, cudaMemset(h->d_tail, 0, tl_align(sizeof(TailHeader))));
        h->tail_bar = 0;
    }
    return MCL_OK;
}

// test hook: systematic resampling of the given weights through the tail kernel's resampling stages (S3-S5)
// alone; d_c (nullable) receives the running sums (n f32, reference mode) / cumulative sums (n u64, fixed point)
extern "C" int mcl_debug_tail_resample(mcl_handle *h, float *d_w, int64_t n, double r, int mode, int32_t *d_idx, void *d_c) {
    if (!h) return MCL_ERR_ARG;
    if (n <= 0 || !d_w || !d_idx) return mcl_fail(h, MCL_ERR_ARG, "mcl_debug_tail_resample: bad argument");
    if (mode != MCL_RESAMPLE_REFERENCE_F32 && mode != MCL_RESAMPLE_FIXED_POINT)
        return mcl_fail(h, MCL_ERR_ARG, "mcl_debug_tail_resample: unknown mode");
    DeviceGuard guard(h->device);
    if (!mcl_tail_available(h, n)) return mcl_fail(h, MCL_ERR_STATE, "mcl_debug_tail_resample: persistent tail not available");
    int rc = mcl_fused_prepare(h, n);
    if (rc) return rc;
    FusedStep u;
    memset(&u, 0, sizeof(u));
    u.n = n; u.n_global = n; u.use_mh = 0; u.w_out = d_w;
    g_tail_raw = 1;
    rc = mcl_tail_step(h, u, mcl_fused_keymax(h), mode, r, d_idx, nullptr, nullptr, nullptr, nullptr);
    g_tail_raw = 0;
    if (rc) return rc;
    if (d_c) {
        const TailPlan p = tail_plan(h, n);
        MCL_CUDA(h, cudaMemcpyAsync(d_c, (char *)h->d_tail + p.o_C, (size_t)n * (mode == MCL_RESAMPLE_REFERENCE_F32 ? 4 : 8),
                                    cudaMemcpyDeviceToDevice, h->stream));
    }
    return MCL_OK;
}

// debug (MCL_TAIL_PROF=1): globaltimer stamps [grid][16] of the last tail launch; returns the grid size in *grid
extern "C" int mcl_tail_prof(mcl_handle *h, unsigned long long *out, int *grid) {
    if (!h || !out || !grid) return MCL_ERR_ARG;
    *grid = 0;
    if (!h->d_tail_prof) return MCL_OK;
    DeviceGuard guard(h->device);
    MCL_CUDA(h, cudaStreamSynchronize(h->stream));
    MCL_CUDA(h, cudaMemcpy(out, h->d_tail_prof, (size_t)h->tail_prof_grid * 32 * 8, cudaMemcpyDeviceToHost));
    *grid = h->tail_prof_grid;
    return MCL_OK;
}

// MH chain (mcl_filter_update_chain): where the likelihood of chain_0 leaves its maximum; one iteration
unsigned long long *mcl_tail_chain_key(mcl_handle *h, int64_t n, int slot) {
    if (tail_prepare(h, n)) return nullptr;
    return &reinterpret_cast<TailHeader *>(h->d_tail)->ckey[slot & 1];
}

int mcl_tail_chain_iteration(mcl_handle *h, const FusedStep &u, unsigned long long *d_keymax, float *score_chain, int it,
                             const TailComm *comm) {
    int rc = tail_prepare(h, u.n);
    if (rc) return rc;
    const TailPlan p = tail_plan(h, u.n);
    char *b = (char *)h->d_tail;
    TailArgs a;
    memset(&a, 0, sizeof(a));
    a.hd = (TailHeader *)b; a.keymax = d_keymax; a.bar_base = h->tail_bar;
    a.n = u.n; a.s_post = u.s_post; a.w_out = u.w_out; a.score_chain = score_chain;
    a.px = u.px; a.py = u.py; a.pt = u.pt; a.ox = u.ox; a.oy = u.oy; a.ot = u.ot; a.nx = u.nx; a.ny = u.ny; a.nth = u.nth;
    a.seed = u.seed; a.step = u.step; a.first_index = u.first_index;
    a.part_q = (unsigned long long *)(b + p.o_q);
    a.ckey_in = &a.hd->ckey[it & 1]; a.ckey_out = &a.hd->ckey[(it + 1) & 1];
    if (comm) {
        a.rank = comm->rank; a.world = comm->world; a.n_global = comm->n_global; a.mailbox = comm->mailbox;
        for (int d = 0; d < 16; ++d) a.peers[d] = comm->peers[d];
        a.epoch0 = comm->epoch0; a.comm_err = comm->d_err;
    }
    const int grid = (int)std::max<int64_t>(1, std::min<int64_t>((u.n + TL_THREADS * 4 - 1) / (TL_THREADS * 4), h->sm_count));
    void *params[] = {(void *)&a};
    cudaError_t e = comm ? cudaLaunchCooperativeKernel((const void *)k_chain_tail<true>, dim3(grid), dim3(TL_THREADS), params, 0, h->stream)
                         : cudaLaunchCooperativeKernel((const void *)k_chain_tail<false>, dim3(grid), dim3(TL_THREADS), params, 0, h->stream);
    if (e != cudaSuccess) return mcl_fail(h, MCL_ERR_CUDA, std::string("k_chain_tail launch: ") + cudaGetErrorString(e));
    h->launches++;
    h->tail_bar += (unsigned long long)grid;
    return MCL_OK;
}

// after the last iteration: both chain-key slots back to zero
int mcl_tail_chain_finish(mcl_handle *h) {
    if (!h->d_tail) return MCL_OK;
    MCL_CUDA(h, cudaMemsetAsync(reinterpret_cast<TailHeader *>(h->d_tail)->ckey, 0, 2 * sizeof(unsigned long long), h->stream));
    return MCL_OK;
}
