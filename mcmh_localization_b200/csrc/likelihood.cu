// likelihood.cu -- kernel (2): likelihood-field scan likelihood, pu:85-149 compute_likelihoods.
//
// Per particle i and valid beam j the reference evaluates
//     lx = x + r cos(theta + a_j),  mx = int((lx - ox) / res)   (fp64, trunc toward zero)
//     log p(dist[my*W+mx])  accumulated in fp64, mean over valid_count, stored as fp32.
// Here the per-cell value log p(dist[c]) is a precomputed fixed-point table (mcl_core.cu), the beam
// endpoints r (cos a_j, sin a_j) / res are a per-scan fp64 table, and a particle needs one fp64 sincos.
//
// Cell index.  Per (particle, beam) the endpoint is  t = p + c bx - s by  in cell units; the cell index
// must agree with the fp64 reference (SURVEY 7 hard part 1: fp32 flips 6e-5 of the cells), and only the
// integer part of t is wanted.  The particle coordinate is taken relative to the window origin and
// biased by the "magic" constant M = 1.5 * 2^(20-S) once per particle; the two FMAs of a coordinate then
// produce M + t directly, a double whose HIGH WORD is  K + floor(t * 2^S)  (S = 8 for maps up to ~1800
// cells).  No conversion instruction, no fp64 floor: the clamp onto the window is one VIADDMNMX on the
// high word and the table index is one byte permute of the two clamped words.  The bias rounds t to
// 2^-(32+S) (2^-40 cell for S = 8, ~1e-12 against the reference's own ~1e-14 endpoint rounding): every
// kernel and path below uses exactly this arithmetic, so cell indices -- and, the table being fixed
// point, scores -- are identical for any lanes-per-particle mapping, path or number of ranks.
//
// The table is staged in shared memory: the free-space window of the map (all cells outside
// it hold one constant c0) plus a one-cell c0 border, so out-of-window lookups clamp onto the
// border instead of branching.  Staging uses the bulk-copy engine (cp.async.bulk + mbarrier, "TMA"
// 1-D form); CTAs are persistent and stage once.  If the window does not fit in shared memory the
// table is gathered from global memory / L2.
// Mapping: G lanes per particle (G = 1 for large N: beam constants are then warp-uniform constant-bank
// loads; G up to 32 for small N to fill the machine), beams strided over the G lanes and reduced
// with __shfl_xor.
#include <float.h>
#include <stdlib.h>

#include <algorithm>

#include "common.cuh"

#define LIK_THREADS 512

struct LikParams {
    const double *x, *y, *th;
    int64_t n;
    float *score;
    // G = 1 kernels only: an optional second particle set of the same length evaluated by the same launch
    // (the MH update scores particles and particles_prev, node:254-268), and an optional pair of cells that
    // receive the maximum score of each set as an order-preserving unsigned key (mcl_key_of_float)
    const double *x2, *y2, *th2;
    float *score2;
    unsigned long long *keymax;
    const BeamTable *beams;
    int n_pos, n_neg;
    double ox, oy, res;
    int W, H;
    const int32_t *logtab, *win;
    const float *dist;
    const uint8_t *win8;       // coded window + table of distinct values (maps whose int32 window is too big)
    const int32_t *lut;
    uint32_t win8_bytes, win_bytes;
    int32_t voff;              // shared-memory table values are v - voff (mcl_handle::voff)
    int wofx, wofy;            // window origin incl. border: wx0 - 1, wy0 - 1
    int cx, cy;                // largest window index per axis: ww + 1, wh + 1
    int tpose;                 // window stored with y as the minor (pitch-256) axis
    double M, lim;             // cell arithmetic: bias and validity limit (mcl_handle::cell_*)
    int K, S;
    double sigma_hit, z_hit, z_rand, max_range;
    double margin;     // rmax_cells + 2: particles further than this from every map edge cannot
                       // produce an out-of-map endpoint
};

__device__ __forceinline__ uint32_t smem_u32(const void *p) {
    return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::
                     "r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    uint32_t done;
    do {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(smem_u32(bar)), "r"(parity)
            : "memory");
    } while (!done);
}
__device__ __forceinline__ void bulk_g2s_chunked(unsigned char *dst, const unsigned char *src, uint32_t bytes,
                                                 uint64_t *bar) {
    const uint32_t CH = 32768;
    for (uint32_t o = 0; o < bytes; o += CH) bulk_g2s(dst + o, src + o, min(CH, bytes - o), bar);
}

// ---------------------------------------------------------------------------------------------
// the cell arithmetic shared by every path
// ---------------------------------------------------------------------------------------------
struct Pose {          // one particle, ready for the beam loop
    double PX, PY;     // M + window-relative position (cells)
    double s, c;
    bool interior;     // no endpoint can leave the map (coordinates >= 1: trunc == floor, no bounds test)
    bool far;          // outside the range of the arithmetic: no endpoint can be inside the map
};

__device__ __forceinline__ Pose load_pose(const LikParams &p, const double *__restrict__ xs, const double *__restrict__ ys,
                                          const double *__restrict__ ts, int64_t i) {
    Pose q;
    const double x = xs[i], y = ys[i], th = ts[i];
    sincos(th, &q.s, &q.c);
    const double px = __ddiv_rn(__dadd_rn(x, -p.ox), p.res);          // pu:128: (lx - ox) / res, distributed
    const double py = __ddiv_rn(__dadd_rn(y, -p.oy), p.res);
    const double wx = __dadd_rn(px, -(double)p.wofx), wy = __dadd_rn(py, -(double)p.wofy);   // exact
    q.far = !(fabs(wx) < p.lim && fabs(wy) < p.lim);                   // also catches NaN poses
    q.PX = __dadd_rn(q.far ? 0.0 : wx, p.M);
    q.PY = __dadd_rn(q.far ? 0.0 : wy, p.M);
    q.interior = (px >= p.margin) && (px <= (double)p.W - p.margin) && (py >= p.margin) && (py <= (double)p.H - p.margin);
    return q;
}

__device__ __forceinline__ double end_x(const Pose &q, const double2 b) { return fma(q.c, b.x, fma(-q.s, b.y, q.PX)); }
__device__ __forceinline__ double end_y(const Pose &q, const double2 b) { return fma(q.s, b.x, fma(q.c, b.y, q.PY)); }

// clamped window coordinate scaled by 2^S (low S bits: fraction)
__device__ __forceinline__ int win_coord(double T, int K, int cmax) {
    return __viaddmin_s32_relu(__double2hiint(T), -K, cmax);
}
// In-map test and map cell with the reference's int() (trunc toward zero) semantics, exact on the 2^-(32+S) grid.
// Fm = floor(map coordinate * 2^S) = high word - Km, Km = K - (window origin << S).  int() sends (-1, 0) to cell
// 0 (SURVEY 7 hard part 2), so a coordinate is inside iff it lies in (-1, dim): Fm in [-2^S, dim * 2^S), minus
// the single point -1.0 (Fm == -2^S with a zero low word).  The cell is then max(Fm >> S, 0).
__device__ __forceinline__ bool coord_in_map(double T, int Fm, int S, unsigned lim /* (dim + 1) << S */) {
    bool in = (unsigned)(Fm + (1 << S)) < lim;
    if (Fm == -(1 << S)) in = __double2loint(T) != 0;
    return in;
}
__device__ __forceinline__ int map_coord(double T, int K, int S, int wof, int dim, bool &in) {
    const int Fm = (__double2hiint(T) - K) + (wof << S);
    in = coord_in_map(T, Fm, S, (unsigned)(dim + 1) << S);
    return max(Fm >> S, 0);
}
__device__ __forceinline__ int win_index(int rx, int ry, int S, int tpose) {   // generic (S != 8 or run-time layout)
    const int ix = rx >> S, iy = ry >> S;
    return tpose ? (ix << 8) | iy : (iy << 8) | ix;
}

template <int G, bool SMEM>
__global__ void __launch_bounds__(LIK_THREADS, 2) k_likelihood(const LikParams p) {
    extern __shared__ __align__(128) unsigned char smem[];
    uint64_t *bar = reinterpret_cast<uint64_t *>(smem);
    double2 *sb = reinterpret_cast<double2 *>(smem + 16);
    const int nb = p.n_pos + p.n_neg;
    const uint32_t beam_bytes = (uint32_t)nb * (uint32_t)sizeof(BeamTable);
    int32_t *swin = reinterpret_cast<int32_t *>(smem + 16 + beam_bytes);

    if (threadIdx.x == 0) mbar_init(bar, 1);
    __syncthreads();
    if (threadIdx.x == 0) {
        mbar_expect_tx(bar, beam_bytes + (SMEM ? p.win_bytes : 0u));
        bulk_g2s_chunked(reinterpret_cast<unsigned char *>(sb), reinterpret_cast<const unsigned char *>(p.beams),
                         beam_bytes, bar);
        if (SMEM)
            bulk_g2s_chunked(reinterpret_cast<unsigned char *>(swin), reinterpret_cast<const unsigned char *>(p.win),
                             p.win_bytes, bar);
    }
    mbar_wait(bar, 0);

    constexpr int GROUPS = LIK_THREADS / G;
    const int g = threadIdx.x & (G - 1);
    const int grp = threadIdx.x / G;
    const int cmx = (p.cx << p.S) | ((1 << p.S) - 1), cmy = (p.cy << p.S) | ((1 << p.S) - 1);

    // blockIdx.y selects the particle set (pair launches of the fused step)
    const bool second = blockIdx.y != 0;
    const double *__restrict__ xs = second ? p.x2 : p.x, *__restrict__ ys = second ? p.y2 : p.y,
                 *__restrict__ ts = second ? p.th2 : p.th;
    float *__restrict__ score = second ? p.score2 : p.score;
    float smax = -FLT_MAX;
    // warp-uniform trip count: every lane iterates while the FIRST group of its warp is in range
    const int64_t stride = (int64_t)gridDim.x * GROUPS;
    const int warp_first_grp = (threadIdx.x & ~31) / G;
    for (int64_t base = (int64_t)blockIdx.x * GROUPS; base + warp_first_grp < p.n; base += stride) {
        const int64_t i = base + grp;
        const Pose q = load_pose(p, xs, ys, ts, i < p.n ? i : p.n - 1);
        long long acc = 0;
        if (!q.far) {
            if (SMEM && __all_sync(0xffffffffu, q.interior)) {
#pragma unroll 4
                for (int j = g; j < p.n_pos; j += G) {
                    const double2 b = sb[j];
                    const int rx = win_coord(end_x(q, b), p.K, cmx), ry = win_coord(end_y(q, b), p.K, cmy);
                    acc += swin[win_index(rx, ry, p.S, p.tpose)] + (long long)p.voff;
                }
            } else {
#pragma unroll 2
                for (int j = g; j < p.n_pos; j += G) {
                    const double2 b = sb[j];
                    bool inx, iny;
                    const int mx = map_coord(end_x(q, b), p.K, p.S, p.wofx, p.W, inx);      // pu:128-129
                    const int my = map_coord(end_y(q, b), p.K, p.S, p.wofy, p.H, iny);
                    if (inx && iny) {                                                       // pu:131-132
                        if (SMEM) {
                            const int ix = min(max(mx - p.wofx, 0), p.cx), iy = min(max(my - p.wofy, 0), p.cy);
                            acc += swin[p.tpose ? (ix << 8) | iy : (iy << 8) | ix] + (long long)p.voff;
                        } else {
                            acc += __ldg(p.logtab + (size_t)my * p.W + mx);
                        }
                    }
                }
            }
            // valid beams with a negative range: p_rand = 0 (pu:139); evaluated from the distance map
            for (int j = p.n_pos + g; j < nb; j += G) {
                const double2 b = sb[j];
                bool inx, iny;
                const int mx = map_coord(end_x(q, b), p.K, p.S, p.wofx, p.W, inx);
                const int my = map_coord(end_y(q, b), p.K, p.S, p.wofy, p.H, iny);
                if (inx && iny)
                    acc += quantise_logp(cell_logp(__ldg(p.dist + (size_t)my * p.W + mx), p.sigma_hit, p.z_hit,
                                                   p.z_rand, p.max_range, false));
            }
        }
#pragma unroll
        for (int o = G / 2; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
        if (g == 0 && i < p.n) {
            const float sc = (float)(((double)acc / MCL_LOGP_SCALE) / (double)nb);   // pu:144-145
            score[i] = sc;
            smax = fmaxf(smax, sc);
        }
    }
    if (p.keymax) {                                        // maximum score of the set (first softmax pass, node:353)
        __shared__ float smx[LIK_THREADS / 32];
        smax = warp_max(smax);
        if ((threadIdx.x & 31) == 0) smx[threadIdx.x >> 5] = smax;
        __syncthreads();
        if (threadIdx.x < 32) {
            float t = threadIdx.x < LIK_THREADS / 32 ? smx[threadIdx.x] : -FLT_MAX;
            t = warp_max(t);
            if (threadIdx.x == 0 && (int64_t)blockIdx.x * GROUPS < p.n)
                atomicMax(p.keymax + (second ? 1 : 0), (unsigned long long)mcl_key_of_float(t));
        }
    }
}


// ---------------------------------------------------------------------------------------------
// G = 1 (one thread per particle, N large): every lane of a warp evaluates the SAME beam, so the
// beam constants are warp-uniform and come from the constant bank (no LSU / shared-memory traffic:
// ncu on the first version showed the shared-memory pipe at 78 % with 40 % of its wavefronts spent on
// the beam table).  Per evaluation: 4 DFMA, 2 VIADDMNMX, 1 PRMT, 1 LEA, 1 LDS, ~0.75 integer adds and,
// shared by the two particles of a thread, the uniform loads of the beam.
//
// Work split: every CTA owns an equal contiguous share of the particles and walks it in rows of one
// particle per thread; a warp takes its 32-particle slices two rows at a time and a last odd slice alone,
// so all warps of the grid carry the same number of slices +-1 (no tail of half-empty SMs).
// Table values in shared memory are v - voff >= 0, so MCL_ACC_TERMS of them are summed in one unsigned
// 32-bit register before the 64-bit accumulator is touched.
// ---------------------------------------------------------------------------------------------
#define MAX_CBEAMS 2048
__constant__ BeamTable c_beams_raw[MAX_CBEAMS];
#define c_beams (reinterpret_cast<const double2 *>(c_beams_raw))

struct G1Ctx {
    const int32_t *swin, *slut;
    const uint8_t *swin8;
    int lane, nb, cmx, cmy;
};

template <bool CODED>
__device__ __forceinline__ uint32_t g1_fetch(const G1Ctx &k, int cell) {
    return (uint32_t)(CODED ? k.slut[(int)k.swin8[cell] * 32 + k.lane] : k.swin[cell]);
}

// P slices (rows i0, i0 + row, ...) of one warp; lanes whose particle index is >= end idle on a copy of end - 1
template <bool SMEM, bool CODED, bool TPOSE, int P>
__device__ __forceinline__ void g1_slices(const LikParams &p, const G1Ctx &k, const double *__restrict__ xs,
                                          const double *__restrict__ ys, const double *__restrict__ ts,
                                          float *__restrict__ score, int64_t i0, int64_t row, int64_t end, float &smax) {
    int64_t idx[P];
    Pose q[P];
    bool interior = true, any_near = false;
#pragma unroll
    for (int u = 0; u < P; ++u) {
        idx[u] = i0 + u * row;
        q[u] = load_pose(p, xs, ys, ts, idx[u] < end ? idx[u] : end - 1);
        interior = interior && q[u].interior;
        any_near = any_near || !q[u].far;
    }
    long long acc[P];
#pragma unroll
    for (int u = 0; u < P; ++u) acc[u] = 0;
    if (SMEM && __reduce_and_sync(0xffffffffu, interior ? 1u : 0u)) {   // REDUX: the result is a uniform register, the branch stays uniform
        unsigned long long uacc[P];
#pragma unroll
        for (int u = 0; u < P; ++u) uacc[u] = 0;
        // plain register arrays (not the Pose structs): ptxas then keeps the beam loop on the uniform datapath
        double PX[P], PY[P], ss[P], cc[P];
#pragma unroll
        for (int u = 0; u < P; ++u) { PX[u] = q[u].PX; PY[u] = q[u].PY; ss[u] = q[u].s; cc[u] = q[u].c; }
        const int negK = -p.K;
        auto eval = [&](int u, const double2 b) -> uint32_t {
            const double TX = fma(cc[u], b.x, fma(-ss[u], b.y, PX[u])), TY = fma(ss[u], b.x, fma(cc[u], b.y, PY[u]));
            const int rx = __viaddmin_s32_relu(__double2hiint(TX), negK, k.cmx);
            const int ry = __viaddmin_s32_relu(__double2hiint(TY), negK, k.cmy);
            return g1_fetch<CODED>(k, (int)(TPOSE ? __byte_perm(ry, rx, 0x7651) : __byte_perm(rx, ry, 0x7651)));
        };
        int j = 0;
        for (; j + MCL_ACC_TERMS <= p.n_pos; j += MCL_ACC_TERMS) {
            uint32_t part[P];
#pragma unroll
            for (int u = 0; u < P; ++u) part[u] = 0;
#pragma unroll
            for (int t = 0; t < MCL_ACC_TERMS; ++t) {
                const double2 b = c_beams[j + t];
#pragma unroll
                for (int u = 0; u < P; ++u) part[u] += eval(u, b);
            }
#pragma unroll
            for (int u = 0; u < P; ++u) uacc[u] += part[u];
        }
        for (; j < p.n_pos; ++j) {
            const double2 b = c_beams[j];
#pragma unroll
            for (int u = 0; u < P; ++u) uacc[u] += eval(u, b);
        }
#pragma unroll
        for (int u = 0; u < P; ++u) acc[u] = (long long)uacc[u] + (long long)p.n_pos * p.voff;
    } else if (__any_sync(0xffffffffu, any_near)) {
        // an endpoint may leave the map: same loop with the in-map test of pu:131-132 (out-of-map beams add 0)
        if (SMEM) {
            unsigned long long uacc[P];
            double PX[P], PY[P], ss[P], cc[P];
#pragma unroll
            for (int u = 0; u < P; ++u) { uacc[u] = 0; PX[u] = q[u].PX; PY[u] = q[u].PY; ss[u] = q[u].s; cc[u] = q[u].c; }
            const int negK = -p.K, Kmx = p.K - (p.wofx << 8), Kmy = p.K - (p.wofy << 8);
            const unsigned limx = (unsigned)(p.W + 1) << 8, limy = (unsigned)(p.H + 1) << 8;
            const uint32_t zero_off = (uint32_t)(-p.voff);
            auto eval = [&](int u, const double2 b) -> uint32_t {
                const double TX = fma(cc[u], b.x, fma(-ss[u], b.y, PX[u])), TY = fma(ss[u], b.x, fma(cc[u], b.y, PY[u]));
                const int hx = __double2hiint(TX), hy = __double2hiint(TY);
                const int rx = __viaddmin_s32_relu(hx, negK, k.cmx), ry = __viaddmin_s32_relu(hy, negK, k.cmy);
                const uint32_t v = g1_fetch<CODED>(k, (int)(TPOSE ? __byte_perm(ry, rx, 0x7651) : __byte_perm(rx, ry, 0x7651)));
                const bool in = coord_in_map(TX, hx - Kmx, 8, limx) && coord_in_map(TY, hy - Kmy, 8, limy);
                return in ? v : zero_off;
            };
            int j = 0;
            for (; j + MCL_ACC_TERMS <= p.n_pos; j += MCL_ACC_TERMS) {
                uint32_t part[P];
#pragma unroll
                for (int u = 0; u < P; ++u) part[u] = 0;
#pragma unroll
                for (int t = 0; t < MCL_ACC_TERMS; ++t) {
                    const double2 b = c_beams[j + t];
#pragma unroll
                    for (int u = 0; u < P; ++u) part[u] += eval(u, b);
                }
#pragma unroll
                for (int u = 0; u < P; ++u) uacc[u] += part[u];
            }
            for (; j < p.n_pos; ++j) {
                const double2 b = c_beams[j];
#pragma unroll
                for (int u = 0; u < P; ++u) uacc[u] += eval(u, b);
            }
#pragma unroll
            for (int u = 0; u < P; ++u) acc[u] = q[u].far ? 0 : (long long)uacc[u] + (long long)p.n_pos * p.voff;
        } else {
#pragma unroll 4
            for (int j = 0; j < p.n_pos; ++j) {
                const double2 b = c_beams[j];
#pragma unroll
                for (int u = 0; u < P; ++u) {
                    bool inx, iny;
                    const int mx = map_coord(end_x(q[u], b), p.K, p.S, p.wofx, p.W, inx);      // pu:128-129
                    const int my = map_coord(end_y(q[u], b), p.K, p.S, p.wofy, p.H, iny);
                    acc[u] += (!q[u].far && inx && iny) ? __ldg(p.logtab + (size_t)my * p.W + mx) : 0;
                }
            }
        }
    }
    // valid beams with a negative range: p_rand = 0 (pu:139); evaluated from the distance map
    for (int j = p.n_pos; j < k.nb; ++j) {
        const double2 b = c_beams[j];
#pragma unroll
        for (int u = 0; u < P; ++u) {
            bool inx, iny;
            const int mx = map_coord(end_x(q[u], b), p.K, p.S, p.wofx, p.W, inx);
            const int my = map_coord(end_y(q[u], b), p.K, p.S, p.wofy, p.H, iny);
            if (!q[u].far && inx && iny)
                acc[u] += quantise_logp(cell_logp(__ldg(p.dist + (size_t)my * p.W + mx), p.sigma_hit, p.z_hit,
                                                  p.z_rand, p.max_range, false));
        }
    }
#pragma unroll
    for (int u = 0; u < P; ++u)
        if (idx[u] < end) {
            const float sc = (float)(((double)acc[u] / MCL_LOGP_SCALE) / (double)k.nb);   // pu:144-145
            score[idx[u]] = sc;
            smax = fmaxf(smax, sc);
        }
}

// Tunables (chosen by measurement, see profiles/): threads per CTA and minimum CTAs per SM.
template <bool SMEM, int G1_THREADS, int MINB, bool CODED = false, bool TPOSE = false>
__global__ void __launch_bounds__(G1_THREADS, MINB) k_likelihood_g1(const LikParams p) {
    extern __shared__ __align__(128) unsigned char smem[];
    uint64_t *bar = reinterpret_cast<uint64_t *>(smem);
    // plain: [16 B barrier][int32 window]
    // CODED: [16 B barrier][table of distinct values replicated per lane: 256 x 32 int32][uint8 window]
    if (SMEM) {
        if (threadIdx.x == 0) mbar_init(bar, 1);
        __syncthreads();
        if (threadIdx.x == 0) {
            if (CODED) {
                mbar_expect_tx(bar, p.win8_bytes);
                bulk_g2s_chunked(smem + 16 + 32768, p.win8, p.win8_bytes, bar);
            } else {
                mbar_expect_tx(bar, p.win_bytes);
                bulk_g2s_chunked(smem + 16, reinterpret_cast<const unsigned char *>(p.win), p.win_bytes, bar);
            }
        }
        if (CODED) {   // lane-private copies of the value table: slut[code * 32 + lane] is always conflict-free
            int32_t *slut = reinterpret_cast<int32_t *>(smem + 16);
            for (int e = threadIdx.x; e < 256 * 32; e += G1_THREADS) slut[e] = __ldg(p.lut + (e >> 5));
            __syncthreads();
        }
        mbar_wait(bar, 0);
    }
    G1Ctx k;
    k.swin = reinterpret_cast<const int32_t *>(smem + 16);
    k.slut = reinterpret_cast<const int32_t *>(smem + 16);
    k.swin8 = smem + 16 + 32768;
    k.lane = threadIdx.x & 31;
    k.nb = p.n_pos + p.n_neg;
    k.cmx = (p.cx << 8) | 255; k.cmy = (p.cy << 8) | 255;

    const int64_t per = p.n / gridDim.x, rem = p.n % gridDim.x;
    const int64_t first = (int64_t)blockIdx.x * per + min((int64_t)blockIdx.x, rem);
    const int64_t end = first + per + ((int64_t)blockIdx.x < rem ? 1 : 0);
    // The CTA's share is cut into pairs of adjacent 32-particle slices, dealt round-robin to the warps.  With
    // a second particle set, blockIdx.y selects the set (the second wave of CTAs follows the first without a
    // launch gap).  ONE instance of the beam loop in the code, reached through uniform control flow only:
    // otherwise ptxas moves the beam constants from the uniform datapath (LDCU + UR operands) to per-thread
    // LDC, which costs 15 % (ncu: idc pipe 46 % busy, dispatch stalls).
    const bool second = blockIdx.y != 0;
    const double *__restrict__ xs = second ? p.x2 : p.x, *__restrict__ ys = second ? p.y2 : p.y,
                 *__restrict__ ts = second ? p.th2 : p.th;
    float *__restrict__ score = second ? p.score2 : p.score;
    float smax0 = -FLT_MAX;
    int64_t i = first + 2 * (threadIdx.x & ~31);
    for (; i < end; i += 2 * G1_THREADS)
        g1_slices<SMEM, CODED, TPOSE, 2>(p, k, xs, ys, ts, score, i + k.lane, 32, end, smax0);
    // Never taken (the host rejects n < 0).  With this second, cold copy of the slice code in the kernel
    // ptxas keeps the hot copy above on the uniform datapath; without it the beam constants are fetched with
    // per-thread LDC (found by bisection on the SASS, CUDA 12.9).
    if (p.n < 0) g1_slices<SMEM, CODED, TPOSE, 1>(p, k, xs, ys, ts, score, i + k.lane, 0, end, smax0);
    if (p.keymax) {                                        // maximum score of the set (first softmax pass, node:353)
        __shared__ float smx[32];
        smax0 = warp_max(smax0);
        if (k.lane == 0) smx[threadIdx.x >> 5] = smax0;
        __syncthreads();
        if (threadIdx.x < 32 && end > first) {
            float t = k.lane < (G1_THREADS >> 5) ? smx[k.lane] : -FLT_MAX;
            t = warp_max(t);
            if (k.lane == 0) atomicMax(p.keymax + (second ? 1 : 0), (unsigned long long)mcl_key_of_float(t));
        }
    }
}

__global__ void k_fill_f32(float *out, int64_t n, float v) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        out[i] = v;
}

template <typename K>
static int launch_lik_kernel(mcl_handle *h, K kern, const LikParams &p, size_t smem_bytes, int G,
                             int threads = LIK_THREADS, bool balanced = false) {
    // attribute + occupancy queries are cached per (kernel, smem size): they cost tens of microseconds
    static thread_local const void *c_kern[32];
    static thread_local size_t c_smem[32];
    static thread_local int c_occ[32], c_dev[32], c_n = 0;
    int occ = 0;
    for (int k = 0; k < c_n; ++k)
        if (c_kern[k] == (const void *)kern && c_smem[k] == smem_bytes && c_dev[k] == h->device) occ = c_occ[k];
    if (occ == 0) {
        // opt in to the device maximum once (a later, smaller request must not lower the limit)
        cudaFuncAttributes fa;
        MCL_CUDA(h, cudaFuncGetAttributes(&fa, kern));
        MCL_CUDA(h, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         h->smem_optin - (int)fa.sharedSizeBytes));     // static + dynamic <= opt-in limit
        MCL_CUDA(h, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, threads, smem_bytes));
        if (occ < 1) return mcl_fail(h, MCL_ERR_CAPACITY, "likelihood kernel does not fit on an SM");
        const int k = c_n < 32 ? c_n++ : 31;
        c_kern[k] = (const void *)kern; c_smem[k] = smem_bytes; c_occ[k] = occ; c_dev[k] = h->device;
    }
    const int64_t groups = (int64_t)threads / G;
    const int64_t need = (p.n + groups - 1) / groups;
    int blocks = (int)std::min<int64_t>(need, (int64_t)h->sm_count * occ);
    if (balanced)   // equal contiguous shares: the same number of CTAs on every SM
        blocks = (int)std::min<int64_t>((p.n + 31) / 32, (int64_t)h->sm_count * occ);
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    if (h->timing) {
        MCL_CUDA(h, cudaEventCreate(&e0));
        MCL_CUDA(h, cudaEventCreate(&e1));
        MCL_CUDA(h, cudaEventRecord(e0, h->stream));
    }
    kern<<<dim3(blocks, p.x2 ? 2 : 1), threads, smem_bytes, h->stream>>>(p);
    MCL_LAUNCH_CHECK(h);
    if (h->timing) {
        MCL_CUDA(h, cudaEventRecord(e1, h->stream));
        h->lik_events.emplace_back(e0, e1);
        h->lik_sets_timed += p.x2 ? 2 : 1;
    }
    return MCL_OK;
}

template <int G, bool SMEM>
static int launch_lik(mcl_handle *h, const LikParams &p, size_t smem_bytes) {
    return launch_lik_kernel(h, k_likelihood<G, SMEM>, p, smem_bytes, G);
}

// G = 1: <threads, CTAs per SM> by variant (MCL_LIK_VARIANT, for tuning runs); layout by the handle
template <bool SMEM, bool CODED, bool TPOSE>
static int launch_g1(mcl_handle *h, const LikParams &p, size_t smem_bytes) {
    static int variant = -1;
    if (variant < 0) { const char *e = getenv("MCL_LIK_VARIANT"); variant = e ? atoi(e) : 0; }
    // does a second CTA fit next to the first?
    const bool two = 2 * (smem_bytes + 1024) <= (size_t)h->smem_optin;
    switch (variant) {
        case 1: if (two) return launch_lik_kernel(h, k_likelihood_g1<SMEM, 512, 2, CODED, TPOSE>, p, smem_bytes, 1, 512, true);
        case 2: return launch_lik_kernel(h, k_likelihood_g1<SMEM, 768, 1, CODED, TPOSE>, p, smem_bytes, 1, 768, true);
        case 3: return launch_lik_kernel(h, k_likelihood_g1<SMEM, 896, 1, CODED, TPOSE>, p, smem_bytes, 1, 896, true);
        case 4: return launch_lik_kernel(h, k_likelihood_g1<SMEM, 1024, 1, CODED, TPOSE>, p, smem_bytes, 1, 1024, true);
        default: return launch_lik_kernel(h, k_likelihood_g1<SMEM, 896, 1, CODED, TPOSE>, p, smem_bytes, 1, 896, true);   // 72 registers: fastest measured
    }
}

// the constant-bank beam table is module-global: re-upload when the active scan (or handle) changed
static const void *g_cbeams_src = nullptr;
static uint64_t g_cbeams_gen = 0;

static int likelihood_impl(mcl_handle *h, const double *d_x, const double *d_y, const double *d_theta, int64_t n,
                           float *d_score, const double *d_x2, const double *d_y2, const double *d_theta2,
                           float *d_score2, unsigned long long *d_keymax, bool *g1_used);

extern "C" int mcl_likelihood(mcl_handle *h, const double *d_x, const double *d_y, const double *d_theta,
                              int64_t n, float *d_score) {
    return likelihood_impl(h, d_x, d_y, d_theta, n, d_score, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr);
}

// internal (filter.cu): both score sets of an MH update in one launch + their maxima as keys.  Only the
// one-thread-per-particle kernels implement it: *g1_used = false and NOTHING is launched otherwise.
int mcl_likelihood_pair(mcl_handle *h, const double *d_x, const double *d_y, const double *d_theta, float *d_score,
                        const double *d_x2, const double *d_y2, const double *d_theta2, float *d_score2, int64_t n,
                        unsigned long long *d_keymax, bool *g1_used) {
    return likelihood_impl(h, d_x, d_y, d_theta, n, d_score, d_x2, d_y2, d_theta2, d_score2, d_keymax, g1_used);
}

static int likelihood_impl(mcl_handle *h, const double *d_x, const double *d_y, const double *d_theta, int64_t n,
                           float *d_score, const double *d_x2, const double *d_y2, const double *d_theta2,
                           float *d_score2, unsigned long long *d_keymax, bool *g1_used) {
    if (g1_used) *g1_used = false;
    if (!h) return MCL_ERR_ARG;
    if (n < 0 || (n > 0 && (!d_x || !d_y || !d_theta || !d_score)))
        return mcl_fail(h, MCL_ERR_ARG, "mcl_likelihood: bad argument");
    if (!h->scan_set) return mcl_fail(h, MCL_ERR_STATE, "mcl_likelihood: scan not set (mcl_set_scan)");
    if (n == 0) return MCL_OK;
    DeviceGuard guard(h->device);
    int rc = mcl_prepare_table(h);
    if (rc) return rc;
    const int nb = h->n_pos + h->n_neg;
    if (g1_used && nb == 0) return MCL_OK;
    if (nb == 0) {   // pu:146-147: no valid beam -> -50 for every particle
        k_fill_f32<<<(int)std::min<int64_t>((n + 255) / 256, h->sm_count * 8), 256, 0, h->stream>>>(d_score, n, -50.0f);
        MCL_LAUNCH_CHECK(h);
        return MCL_OK;
    }
    LikParams p;
    p.x = d_x; p.y = d_y; p.th = d_theta; p.n = n; p.score = d_score;
    p.x2 = d_x2; p.y2 = d_y2; p.th2 = d_theta2; p.score2 = d_score2; p.keymax = d_keymax;
    p.beams = h->d_beams_active; p.n_pos = h->n_pos; p.n_neg = h->n_neg;
    p.ox = h->ox; p.oy = h->oy; p.res = h->res; p.W = h->W; p.H = h->H;
    p.logtab = h->d_logtab; p.dist = h->d_dist; p.win = h->d_win;
    p.win8 = h->d_win8; p.lut = h->d_lut; p.win8_bytes = (uint32_t)h->win8_bytes;
    p.win_bytes = (uint32_t)h->win_bytes;
    p.voff = h->voff;
    p.wofx = h->wx0 - 1; p.wofy = h->wy0 - 1; p.cx = h->ww + 1; p.cy = h->wh + 1; p.tpose = h->win_tpose ? 1 : 0;
    p.M = h->cell_M; p.lim = h->cell_lim; p.K = h->cell_K; p.S = h->cell_S;
    p.sigma_hit = h->sigma_hit; p.z_hit = h->z_hit; p.z_rand = h->z_rand; p.max_range = h->max_range;
    p.margin = h->rmax_cells + 2.0;

    const size_t beam_bytes = (size_t)nb * sizeof(BeamTable);
    const size_t smem_glob = 16 + beam_bytes;
    const size_t smem_win = smem_glob + h->win_bytes;
    const size_t smem_limit = (size_t)h->smem_optin - 512;      // room for the kernels' static shared memory
    if (smem_glob > smem_limit) return mcl_fail(h, MCL_ERR_CAPACITY, "mcl_likelihood: too many beams for shared memory");
    bool use_smem = h->win_ok && smem_win <= smem_limit;
    if (h->lik_path == 1) use_smem = false;
    const bool use_coded = !use_smem && h->coded && h->lik_path != 1 && h->cell_S == 8;
    if (h->lik_path == 2 && !use_smem && !use_coded)
        return mcl_fail(h, MCL_ERR_CAPACITY, "mcl_likelihood: free-space window does not fit in shared memory");

    int G = 1;
    const int64_t target = (int64_t)h->sm_count * 2048;
    while (G < 32 && n * G < target) G *= 2;
    if (G == 1 && nb <= MAX_CBEAMS && h->acc_terms_ok) {
        if (g1_used) *g1_used = true;
        if (g_cbeams_src != (const void *)h->d_beams_active || g_cbeams_gen != h->scan_gen) {
            MCL_CUDA(h, cudaMemcpyToSymbolAsync(c_beams_raw, h->d_beams_active, beam_bytes, 0, cudaMemcpyDeviceToDevice,
                                                h->stream));
            g_cbeams_src = (const void *)h->d_beams_active;
            g_cbeams_gen = h->scan_gen;
        }
        if (use_coded) {
            const size_t sm = 16 + 32768 + h->win8_bytes;
            return h->win_tpose ? launch_g1<true, true, true>(h, p, sm) : launch_g1<true, true, false>(h, p, sm);
        }
        if (use_smem && h->cell_S == 8)
            return h->win_tpose ? launch_g1<true, false, true>(h, p, 16 + h->win_bytes)
                                : launch_g1<true, false, false>(h, p, 16 + h->win_bytes);
        return launch_g1<false, false, false>(h, p, 16);
    }
    if (g1_used) *g1_used = true;   // the lanes-per-particle kernels implement the pair / key form as well
    if (use_coded && !use_smem && h->lik_path == 2)
        return mcl_fail(h, MCL_ERR_CAPACITY, "mcl_likelihood: the coded window needs the one-thread-per-particle kernel (larger n)");
#define LIK_CASE(GV)                                                                   \
    case GV:                                                                           \
        return use_smem ? launch_lik<GV, true>(h, p, smem_win) : launch_lik<GV, false>(h, p, smem_glob);
    switch (G) {
        LIK_CASE(1) LIK_CASE(2) LIK_CASE(4) LIK_CASE(8) LIK_CASE(16) LIK_CASE(32)
    }
#undef LIK_CASE
    return mcl_fail(h, MCL_ERR_ARG, "mcl_likelihood: internal");
}
