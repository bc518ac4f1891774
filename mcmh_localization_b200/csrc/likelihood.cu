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
// 1-D form); CTAs are persistent and stage once.  If the window does not fit in shared memory even
// byte-coded, the particles are binned by map tile and the neighbourhood of one tile is staged at a time
// (k_likelihood_tiled); small particle counts on such maps gather from global memory / L2.
// Mapping: G lanes per particle (G = 1 for large N: beam constants are then warp-uniform constant-bank
// loads; G up to 32 for small N to fill the machine), beams strided over the G lanes and reduced
// with __shfl_xor.
#include <float.h>
#include <stdlib.h>

#include <algorithm>

#include <type_traits>

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
    int wofx, wofy;            // map cell of window index 0 per axis (mcl_handle::win_of*)
    int cx, cy;                // largest window index per axis
    int tpose;                 // window stored with y as the minor (pitch-256) axis
    double M, lim;             // cell arithmetic: bias and validity limit (mcl_handle::cell_*)
    int K, S;
    double sigma_hit, z_hit, z_rand, max_range;
    double margin;     // rmax_cells + 2: particles further than this from every map edge cannot
                       // produce an out-of-map endpoint
    double ilox, ihix, iloy, ihiy;   // "no per-beam in-map test needed" box of the particle position (cells): the
                                     // margin on ordinary sides, unbounded on EDGE sides (mcl_handle::win_edge)
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
    bool mfree;        // no endpoint can leave the 256 columns of the window's minor axis: no clamp on that coordinate
    bool inmap;        // the particle lies on the map (tiled kernel: inside the tile it was binned to)
};

__device__ __forceinline__ Pose load_pose(const LikParams &p, const double *__restrict__ xs, const double *__restrict__ ys,
                                          const double *__restrict__ ts, int64_t i, int wofx, int wofy) {
    Pose q;
    const double x = xs[i], y = ys[i], th = ts[i];
    sincos(th, &q.s, &q.c);
    const double px = __ddiv_rn(__dadd_rn(x, -p.ox), p.res);          // pu:128: (lx - ox) / res, distributed
    const double py = __ddiv_rn(__dadd_rn(y, -p.oy), p.res);
    const double wx = __dadd_rn(px, -(double)wofx), wy = __dadd_rn(py, -(double)wofy);       // exact
    q.far = !(fabs(wx) < p.lim && fabs(wy) < p.lim);                   // also catches NaN poses
    q.PX = __dadd_rn(q.far ? 0.0 : wx, p.M);
    q.PY = __dadd_rn(q.far ? 0.0 : wy, p.M);
    q.interior = !q.far && (px >= p.ilox) && (px <= p.ihix) && (py >= p.iloy) && (py <= p.ihiy);
    const double wm = p.tpose ? wy : wx;                               // (tiled kernels pass their own origin: unused there)
    q.mfree = (wm >= p.margin) && (wm + p.margin < 256.0);
    q.inmap = (px >= 0.0) && (px < (double)p.W) && (py >= 0.0) && (py < (double)p.H);
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
    const int Fm = (__double2hiint(T) - K) + wof * (1 << S);
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
        const Pose q = load_pose(p, xs, ys, ts, i < p.n ? i : p.n - 1, p.wofx, p.wofy);
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
// Work split: every CTA owns an equal contiguous share of the particles, cut into pairs of adjacent
// 32-particle slices dealt round-robin to its warps (a thread evaluates two particles per beam), so all
// warps of the grid carry the same number of pairs +-1 (no tail of half-empty SMs).
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
    int wofx, wofy;            // tiled kernel only: origin of the staged sub-window (map cells)
    uint32_t swin8_s, slut_s;  // PERM == 2: the sub-window (column 0 of row 0) and this lane's column of the value table
                               // as 32-bit shared-memory addresses (the buffer changes per item: no 64-bit pointer math)
};

// PERM == 2 fetch: rows of 272 bytes; index, pitch and buffer base in one 3-input add
__device__ __forceinline__ uint32_t g1_fetch_t2(const G1Ctx &k, unsigned cell) {
    uint32_t code, v;
    const uint32_t a = cell + ((cell >> 4) & 0xfffffff0u) + k.swin8_s;
    asm("ld.shared.u8 %0, [%1];" : "=r"(code) : "r"(a));
    asm("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(code * 128u + k.slut_s));
    return v;
}

template <bool CODED>
__device__ __forceinline__ uint32_t g1_fetch(const G1Ctx &k, int cell) {
    return (uint32_t)(CODED ? k.slut[(int)k.swin8[cell] * 32 + k.lane] : k.swin[cell]);
}

// P slices (rows i0, i0 + row, ...) of one warp; lanes whose particle index is >= end idle on a copy of end - 1
// PERM: i0 / end are positions in a tile-sorted order and perm[] maps them to particle indices (tiled kernel)
// PERM: 0 no; 1 tiled kernel with rows of 260 bytes (manual staging); 2 tiled kernel with rows of 272 bytes (bulk copies)
template <bool SMEM, bool CODED, bool TPOSE, int P, int PERM = 0, bool SKEW = false>
__device__ __forceinline__ void g1_slices(const LikParams &p, const G1Ctx &k, const double *__restrict__ xs,
                                          const double *__restrict__ ys, const double *__restrict__ ts,
                                          float *__restrict__ score, int64_t i0, int64_t row, int64_t end, float &smax,
                                          const int32_t *__restrict__ perm = nullptr) {
    const int wofx = PERM ? k.wofx : p.wofx, wofy = PERM ? k.wofy : p.wofy;     // tiled: origin of the staged tile
    int64_t idx[P];
    int64_t pidx[P];
    Pose q[P];
    bool interior = true, any_near = false, mfree = true, inmap = true;
#pragma unroll
    for (int u = 0; u < P; ++u) {
        idx[u] = i0 + u * row;
        const int64_t il = idx[u] < end ? idx[u] : end - 1;
        pidx[u] = PERM ? (int64_t)perm[il] : il;
        q[u] = load_pose(p, xs, ys, ts, pidx[u], wofx, wofy);
        interior = interior && q[u].interior;
        mfree = mfree && q[u].mfree;
        inmap = inmap && q[u].inmap;
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
        // MF: no endpoint of these particles leaves the 256 columns of the minor axis, so that coordinate needs no
        // clamp: the byte permute takes byte 1 of the raw high word (the low 16 bits of K are zero) -- one issue cycle
        // in 15.  (Windowed kernels: particles whose beam footprint lies inside the columns; tiled kernel: particles
        // inside the tile they were binned to -- the staged sub-window covers every endpoint.  Its 304 rows need nine
        // bits, which a raw high word does not deliver cleanly, so the row coordinate keeps its clamp.)
        auto beam_loop = [&](auto mf_tag) {
            constexpr bool MF = decltype(mf_tag)::value;
            auto eval = [&](int u, const double2 b) -> uint32_t {
                const double TX = fma(cc[u], b.x, fma(-ss[u], b.y, PX[u])), TY = fma(ss[u], b.x, fma(cc[u], b.y, PY[u]));
                const int rx = (MF && !TPOSE) ? __double2hiint(TX) : __viaddmin_s32_relu(__double2hiint(TX), negK, k.cmx);
                const int ry = (MF && TPOSE) ? __double2hiint(TY) : __viaddmin_s32_relu(__double2hiint(TY), negK, k.cmy);
                unsigned cell = TPOSE ? __byte_perm(ry, rx, 0x7651) : __byte_perm(rx, ry, 0x7651);
                if (PERM == 1) cell += (cell >> 8) << 2;     // tiled kernel: rows of 260 bytes (see k_likelihood_tiled)
                if (SKEW) cell += cell >> 5;                 // rows of 264 words: consecutive rows 8 banks apart (k_pack_window)
                if (PERM == 2) return g1_fetch_t2(k, cell);  // rows of 272 bytes (k_likelihood_tiled2)
                return g1_fetch<CODED>(k, (int)cell);
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
        };
        const bool mf = PERM == 0 ? mfree : inmap;
        if ((PERM == 0 || (PERM == 2 && !TPOSE)) && !SKEW && __reduce_and_sync(0xffffffffu, mf ? 1u : 0u)) beam_loop(std::true_type{});
        else beam_loop(std::false_type{});
#pragma unroll
        for (int u = 0; u < P; ++u) acc[u] = (long long)uacc[u] + (long long)p.n_pos * p.voff;
    } else if (__any_sync(0xffffffffu, any_near)) {
        // an endpoint may leave the map: same loop with the in-map test of pu:131-132 (out-of-map beams add 0)
        if (SMEM) {
            unsigned long long uacc[P];
            double PX[P], PY[P], ss[P], cc[P];
#pragma unroll
            for (int u = 0; u < P; ++u) { uacc[u] = 0; PX[u] = q[u].PX; PY[u] = q[u].PY; ss[u] = q[u].s; cc[u] = q[u].c; }
            const int negK = -p.K, Kmx = p.K - wofx * 256, Kmy = p.K - wofy * 256;
            const unsigned limx = (unsigned)(p.W + 1) << 8, limy = (unsigned)(p.H + 1) << 8;
            const uint32_t zero_off = (uint32_t)(-p.voff);
            auto eval = [&](int u, const double2 b) -> uint32_t {
                const double TX = fma(cc[u], b.x, fma(-ss[u], b.y, PX[u])), TY = fma(ss[u], b.x, fma(cc[u], b.y, PY[u]));
                const int hx = __double2hiint(TX), hy = __double2hiint(TY);
                const int fx = hx - Kmx, fy = hy - Kmy;               // floor(map coordinate * 256)
                // int() sends a coordinate in (-1, 0) to map cell 0: read that cell, not the one left of it
                const int rx = __viaddmin_s32_relu(hx - min(fx, 0), negK, k.cmx);
                const int ry = __viaddmin_s32_relu(hy - min(fy, 0), negK, k.cmy);
                unsigned cell = TPOSE ? __byte_perm(ry, rx, 0x7651) : __byte_perm(rx, ry, 0x7651);
                if (PERM == 1) cell += (cell >> 8) << 2;
                if (SKEW) cell += cell >> 5;
                const uint32_t v = PERM == 2 ? g1_fetch_t2(k, cell) : g1_fetch<CODED>(k, (int)cell);
                const bool in = coord_in_map(TX, fx, 8, limx) && coord_in_map(TY, fy, 8, limy);
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
                    const int mx = map_coord(end_x(q[u], b), p.K, p.S, wofx, p.W, inx);      // pu:128-129
                    const int my = map_coord(end_y(q[u], b), p.K, p.S, wofy, p.H, iny);
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
            const int mx = map_coord(end_x(q[u], b), p.K, p.S, wofx, p.W, inx);
            const int my = map_coord(end_y(q[u], b), p.K, p.S, wofy, p.H, iny);
            if (!q[u].far && inx && iny)
                acc[u] += quantise_logp(cell_logp(__ldg(p.dist + (size_t)my * p.W + mx), p.sigma_hit, p.z_hit,
                                                  p.z_rand, p.max_range, false));
        }
    }
#pragma unroll
    for (int u = 0; u < P; ++u)
        if (idx[u] < end) {
            const float sc = (float)(((double)acc[u] / MCL_LOGP_SCALE) / (double)k.nb);   // pu:144-145
            score[PERM ? pidx[u] : idx[u]] = sc;
            smax = fmaxf(smax, sc);
        }
}

// Tunables (chosen by measurement, see profiles/): threads per CTA and minimum CTAs per SM.
template <bool SMEM, int G1_THREADS, int MINB, bool CODED = false, bool TPOSE = false, bool SKEW = false>
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
        g1_slices<SMEM, CODED, TPOSE, 2, 0, SKEW>(p, k, xs, ys, ts, score, i + k.lane, 32, end, smax0);
    // Never taken (the host rejects n < 0).  With this second, cold copy of the slice code in the kernel
    // ptxas keeps the hot copy above on the uniform datapath; without it the beam constants are fetched with
    // per-thread LDC (found by bisection on the SASS, CUDA 12.9).
    if (p.n < 0) g1_slices<SMEM, CODED, TPOSE, 1, 0, SKEW>(p, k, xs, ys, ts, score, i + k.lane, 0, end, smax0);
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


// ---------------------------------------------------------------------------------------------
// Large maps (BASELINE config 5: 4096 x 4096, 64 MB of table): no window fits in shared memory and a gather
// through L1/L2 runs at a quarter of the shared-memory rate.  The particles are binned by map tile (counting
// sort: histogram, offsets, scatter of indices), and persistent CTAs walk the sorted order in chunks, staging
// for every tile they meet the byte-coded neighbourhood of that tile -- tile + the beam reach on every side,
// 256 columns wide so that the byte-permuted index of the coded kernel applies unchanged -- and then running
// the same beam loop.  Coordinates are taken relative to the staged sub-window with S = 8.
// ---------------------------------------------------------------------------------------------
struct TiledArgs {
    const uint8_t *code8p;                // (H + 1) x (W + 16) padded coded map (k_likelihood_tiled2)
    const uint8_t *code8;
    const int32_t *lut;
    const int32_t *perm, *offsets;        // particle index by sorted position; first sorted position of every tile
    const int32_t *items;                 // work items (tile, first position, end position): one tile, <= piece particles
    int *counters;                        // [0] number of items, [1] next item to hand out
    int tile_w, tile_h, margin, tiles_x, tiles_y, ntiles, sub_rows, piece;
};

__device__ __forceinline__ int tile_index_1d(double pc, int tile, int ntile) {
    if (!(pc >= 0.0)) return 0;                                   // negative or NaN
    const double q = floor(pc / (double)tile);
    return q < (double)ntile ? (int)q : ntile - 1;
}
// Histogram of the tiles.  SH: per-block histogram in shared memory, flushed once (a few thousand counters hit by
// millions of global atomics cost 260 us at 6 M particles); the block walks a CONTIGUOUS share so that the scatter
// kernel sees the same particles.
template <bool SH>
__global__ void __launch_bounds__(512) k_tile_count(const double *__restrict__ x, const double *__restrict__ y, int64_t n,
                                                    double ox, double oy, double res, int tile_w, int tile_h, int tiles_x,
                                                    int tiles_y, int ntiles, int32_t *__restrict__ tile_of, int *hist) {
    extern __shared__ int sh_hist[];
    if (SH) {
        for (int t = threadIdx.x; t < ntiles; t += blockDim.x) sh_hist[t] = 0;
        __syncthreads();
    }
    const int64_t per = (n + gridDim.x - 1) / gridDim.x, lo = blockIdx.x * per, hi = min(n, lo + per);
    for (int64_t i = lo + threadIdx.x; i < hi; i += blockDim.x) {
        const double px = __ddiv_rn(__dadd_rn(x[i], -ox), res), py = __ddiv_rn(__dadd_rn(y[i], -oy), res);
        const int t = tile_index_1d(py, tile_h, tiles_y) * tiles_x + tile_index_1d(px, tile_w, tiles_x);
        tile_of[i] = t;
        atomicAdd(SH ? sh_hist + t : hist + t, 1);
    }
    if (SH) {
        __syncthreads();
        for (int t = threadIdx.x; t < ntiles; t += blockDim.x) { const int c = sh_hist[t]; if (c) atomicAdd(hist + t, c); }
    }
}
// exclusive offsets of the tiles (single block); hist is zeroed to serve as the scatter cursors.  Also cuts the
// sorted order into work items for the tiled kernel: every non-empty tile in pieces of at most `piece` particles.
__global__ void __launch_bounds__(1024) k_tile_offsets(int *hist, int ntiles, int32_t *offsets, int piece, int32_t *items,
                                                       int *counters) {
    __shared__ int sh[32], sh2[32];
    __shared__ int carry, carry2;
    if (threadIdx.x == 0) { carry = 0; carry2 = 0; }
    __syncthreads();
    for (int base = 0; base < ntiles; base += 1024) {
        const int i = base + threadIdx.x;
        const int v = i < ntiles ? hist[i] : 0;
        const int np = (v + piece - 1) / piece;                 // pieces of this tile
        int inc = v, inc2 = np;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int t = __shfl_up_sync(0xffffffffu, inc, o), t2 = __shfl_up_sync(0xffffffffu, inc2, o);
            if ((threadIdx.x & 31) >= o) { inc += t; inc2 += t2; }
        }
        if ((threadIdx.x & 31) == 31) { sh[threadIdx.x >> 5] = inc; sh2[threadIdx.x >> 5] = inc2; }
        __syncthreads();
        if (threadIdx.x < 32) {
            int t = sh[threadIdx.x], t2 = sh2[threadIdx.x];
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int u = __shfl_up_sync(0xffffffffu, t, o), u2 = __shfl_up_sync(0xffffffffu, t2, o);
                if (threadIdx.x >= o) { t += u; t2 += u2; }
            }
            sh[threadIdx.x] = t; sh2[threadIdx.x] = t2;
        }
        __syncthreads();
        const int woff = (threadIdx.x >> 5) ? sh[(threadIdx.x >> 5) - 1] : 0;
        const int woff2 = (threadIdx.x >> 5) ? sh2[(threadIdx.x >> 5) - 1] : 0;
        const int c0 = carry, c2 = carry2;
        if (i < ntiles) {
            const int first = c0 + woff + inc - v;
            offsets[i] = first;
            hist[i] = 0;
        }
        __syncthreads();
        if (threadIdx.x == 1023) { carry = c0 + woff + inc; carry2 = c2 + woff2 + inc2; }
        __syncthreads();
    }
    if (threadIdx.x == 0) { offsets[ntiles] = carry; counters[0] = carry2; counters[1] = 0; }
    // Work items, LARGEST FIRST (four size classes; the order inside a class is irrelevant): the CTAs of the tiled
    // kernel take items from a counter and the kernel ends with the slowest CTA, so the small pieces must come last.
    __shared__ int ccount[4], cbase[4], ccur[4];
    if (threadIdx.x < 4) { ccount[threadIdx.x] = 0; ccur[threadIdx.x] = 0; }
    __syncthreads();                                    // also: offsets[] written above are visible to the block
    auto size_class = [piece](int size) { return size >= piece ? 0 : 1 + min(2, (3 * (piece - size)) / piece); };
    for (int i = threadIdx.x; i < ntiles; i += 1024) {
        const int v = offsets[i + 1] - offsets[i];
        for (int j = 0; j * piece < v; ++j) atomicAdd(&ccount[size_class(min(piece, v - j * piece))], 1);
    }
    __syncthreads();
    if (threadIdx.x == 0) { cbase[0] = 0; for (int c = 1; c < 4; ++c) cbase[c] = cbase[c - 1] + ccount[c - 1]; }
    __syncthreads();
    for (int i = threadIdx.x; i < ntiles; i += 1024) {
        const int first = offsets[i], v = offsets[i + 1] - first;
        for (int j = 0; j * piece < v; ++j) {
            const int size = min(piece, v - j * piece), c = size_class(size);
            const int it = cbase[c] + atomicAdd(&ccur[c], 1);
            items[3 * it] = i;
            items[3 * it + 1] = first + j * piece;
            items[3 * it + 2] = first + j * piece + size;
        }
    }
}
// perm[first position of the tile + rank inside the tile] = particle.  SH: the block counts its share per tile in
// shared memory, reserves one range per tile with a single global atomic, then ranks its particles locally.
template <bool SH>
__global__ void __launch_bounds__(512) k_tile_scatter(const int32_t *__restrict__ tile_of, int64_t n, int ntiles,
                                                      const int32_t *__restrict__ offsets, int *cursor,
                                                      int32_t *__restrict__ perm) {
    extern __shared__ int sh_cnt[];          // [ntiles] counts, then local cursors; [ntiles] base of this block in the tile
    const int64_t per = (n + gridDim.x - 1) / gridDim.x, lo = blockIdx.x * per, hi = min(n, lo + per);
    if (!SH) {
        for (int64_t i = lo + threadIdx.x; i < hi; i += blockDim.x) {
            const int t = tile_of[i];
            perm[offsets[t] + atomicAdd(cursor + t, 1)] = (int32_t)i;     // order inside a tile is irrelevant to the results
        }
        return;
    }
    int *sh_base = sh_cnt + ntiles;
    for (int t = threadIdx.x; t < ntiles; t += blockDim.x) sh_cnt[t] = 0;
    __syncthreads();
    for (int64_t i = lo + threadIdx.x; i < hi; i += blockDim.x) atomicAdd(sh_cnt + tile_of[i], 1);
    __syncthreads();
    for (int t = threadIdx.x; t < ntiles; t += blockDim.x) {
        const int c = sh_cnt[t];
        sh_base[t] = c ? offsets[t] + atomicAdd(cursor + t, c) : 0;
        sh_cnt[t] = 0;
    }
    __syncthreads();
    for (int64_t i = lo + threadIdx.x; i < hi; i += blockDim.x) {
        const int t = tile_of[i];
        perm[sh_base[t] + atomicAdd(sh_cnt + t, 1)] = (int32_t)i;
    }
}

template <int THREADS>
__global__ void __launch_bounds__(THREADS, 2) k_likelihood_tiled(const LikParams p, const TiledArgs t) {
    extern __shared__ __align__(128) unsigned char smem[];
    // [16 B][table of distinct values replicated per lane: 256 x 32 int32][uint8 sub-window: sub_rows x 260]
    // Rows of 260 bytes, index = v + 4 (v >> 8): with 256-byte rows the bank would depend on the column only, and the
    // 32 particles of a warp share a 48-column tile -- a dozen banks for 32 lanes (measured: the kernel ran at half rate)
    int32_t *slut = reinterpret_cast<int32_t *>(smem + 16);
    uint8_t *swin8 = smem + 16 + 32768;
    for (int e = threadIdx.x; e < 256 * 32; e += THREADS) slut[e] = __ldg(t.lut + (e >> 5));
    G1Ctx k;
    k.swin = nullptr; k.slut = slut; k.swin8 = swin8;
    k.lane = threadIdx.x & 31;
    k.nb = p.n_pos + p.n_neg;
    k.cmx = (255 << 8) | 255; k.cmy = ((t.sub_rows - 1) << 8) | 255;
    k.wofx = 0; k.wofy = 0;
    const int warp = threadIdx.x >> 5;
    float smax = -FLT_MAX;
    int staged = -1;
    __shared__ int s_item;
    const int nitems = t.counters[0];
    for (;;) {                                                 // work items are handed out dynamically: their sizes vary
        __syncthreads();                                       // previous item done (sub-window, s_item free)
        if (threadIdx.x == 0) s_item = atomicAdd(t.counters + 1, 1);
        __syncthreads();
        const int item = s_item;
        if (item >= nitems) break;
        const int tile = __ldg(t.items + 3 * item);
        const int64_t seg_lo = __ldg(t.items + 3 * item + 1), seg_hi = __ldg(t.items + 3 * item + 2);
        const int sub_x0 = (tile % t.tiles_x) * t.tile_w - t.margin, sub_y0 = (tile / t.tiles_x) * t.tile_h - t.margin;
        if (tile != staged) {
            uint32_t *dst = reinterpret_cast<uint32_t *>(swin8);
            const bool aligned = (p.W & 3) == 0;
            for (int e = threadIdx.x; e < t.sub_rows * 64; e += THREADS) {
                const int my = sub_y0 + (e >> 6), mx = sub_x0 + ((e & 63) << 2);
                uint32_t *d = dst + (e >> 6) * 65 + (e & 63);          // rows of 65 words
                if ((unsigned)my < (unsigned)p.H && aligned && mx >= 0 && mx < p.W) {
                    // asynchronous 4-byte copies: all of a thread's ~40 loads in flight at once
                    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(smem_u32(d)),
                                 "l"(t.code8 + (size_t)my * p.W + mx) : "memory");
                } else {
                    uint32_t v = 0u;                                   // code 0: outside the map
                    if ((unsigned)my < (unsigned)p.H && !aligned) {
                        const uint8_t *row = t.code8 + (size_t)my * p.W;
                        v = 0;
#pragma unroll
                        for (int b = 0; b < 4; ++b)
                            v |= (uint32_t)((unsigned)(mx + b) < (unsigned)p.W ? row[mx + b] : 0) << (8 * b);
                    }
                    *d = v;
                }
            }
            asm volatile("cp.async.wait_all;" ::: "memory");
            staged = tile;
            __syncthreads();
        }
        k.wofx = sub_x0; k.wofy = sub_y0;
        // one 32-particle slice per warp and round: twice the work units of the two-particle form, so that the
        // warps of a CTA stay busy on tiles with few particles (the uniform beam loads are no longer shared)
        int64_t pos = seg_lo + 32 * warp;
        for (; pos < seg_hi; pos += 32 * (THREADS / 32))
            g1_slices<true, true, false, 1, 1>(p, k, p.x, p.y, p.th, p.score, pos + k.lane, 0, seg_hi, smax, t.perm);
        // never taken: keeps the hot copy of the slice code on the uniform datapath (see k_likelihood_g1)
        if (p.n < 0) g1_slices<true, true, false, 2, 1>(p, k, p.x, p.y, p.th, p.score, pos + k.lane, 32, seg_hi, smax, t.perm);
    }
    if (p.keymax) {
        __shared__ float smx[32];
        smax = warp_max(smax);
        if (k.lane == 0) smx[warp] = smax;
        __syncthreads();
        if (threadIdx.x < 32) {
            float v = k.lane < (THREADS >> 5) ? smx[k.lane] : -FLT_MAX;
            v = warp_max(v);
            if (k.lane == 0 && v > -FLT_MAX) atomicMax(p.keymax, (unsigned long long)mcl_key_of_float(v));
        }
    }
}


// ---------------------------------------------------------------------------------------------
// Tiled kernel, second form: one persistent CTA per SM, NCONS consumer warps + one producer warp, TWO sub-window
// buffers.  The producer takes the next work item and copies the byte-coded neighbourhood of its tile row by row
// with bulk asynchronous copies (cp.async.bulk, the TMA engine: 272-byte rows so that the bank depends on the row;
// rows / columns beyond the map are zero-filled = code 0 = "outside, adds 0") while the consumers are still on
// the previous item; full / empty mbarriers hand the buffers over, there is no __syncthreads in the loop.  The
// consumers take 32-particle slices of the item from a shared counter (no static imbalance) and a warp that finds
// none left moves on to the next item on its own.
// ---------------------------------------------------------------------------------------------
#define T2_PITCH 272
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

template <int NCONS>
__global__ void __launch_bounds__((NCONS + 1) * 32, 1) k_likelihood_tiled2(const LikParams p, const TiledArgs t) {
    extern __shared__ __align__(1024) unsigned char smem[];
    // [0,32) full[2], empty[2] | [64,96) ring[2][4] | [96,104) slice counters | [1024, +32 KB) value table per lane |
    // two sub-window buffers of sub_rows x 272 bytes
    uint64_t *full = reinterpret_cast<uint64_t *>(smem), *empty = full + 2;
    int *ring = reinterpret_cast<int *>(smem + 64);
    int *sctr = reinterpret_cast<int *>(smem + 96);
    int32_t *slut = reinterpret_cast<int32_t *>(smem + 1024);
    unsigned char *bufs = smem + 1024 + 32768;
    const uint32_t buf_bytes = (uint32_t)t.sub_rows * T2_PITCH;
    // REDUX results live in uniform registers: every value that steers a loop below goes through one, so that ptxas
    // keeps the beam loop on the uniform datapath (LDCU + UR operands; see k_likelihood_g1)
    const int warp = __reduce_max_sync(0xffffffffu, (int)(threadIdx.x >> 5)), lane = threadIdx.x & 31;
    const int lpad = (16 - (t.margin & 15)) & 15;                     // bytes in front of column 0: 16-byte aligned sources
    if (threadIdx.x == 0) {
        mbar_init(full, 1); mbar_init(full + 1, 1);
        mbar_init(empty, NCONS + 1); mbar_init(empty + 1, NCONS + 1);
    }
    for (int e = threadIdx.x; e < 256 * 32; e += (NCONS + 1) * 32) slut[e] = __ldg(t.lut + (e >> 5));
    __syncthreads();
    const int nitems = t.counters[0];
    float smax = -FLT_MAX;
    // Every warp consumes; the last one also stages: before it turns to use s it stages use s + 1 into the other
    // buffer (as soon as every warp is done with use s - 1, which held that buffer).  With a warp that only staged,
    // one of the four schedulers carried six consumers instead of seven.
    auto stage_use = [&](int s) {                                      // producer warp only; s = use to be staged
        const int b = s & 1;
        if (s >= 2) mbar_wait(empty + b, ((s >> 1) - 1) & 1);          // every warp is done with use s - 2
        int item = 0;
        if (lane == 0) item = atomicAdd(t.counters + 1, 1);
        item = __shfl_sync(0xffffffffu, item, 0);
        if (item >= nitems) {
            if (lane == 0) { ring[b * 4] = -1; mbar_arrive(full + b); }
            return;
        }
        const int tile = __ldg(t.items + 3 * item);
        const int sub_x0 = (tile % t.tiles_x) * t.tile_w - t.margin, sub_y0 = (tile / t.tiles_x) * t.tile_h - t.margin;
        unsigned char *dst = bufs + (size_t)b * buf_bytes;
        // columns [xs, xe) of every row come from the padded map (x >= -16, y >= -1; 16-byte aligned), the rest is
        // zero-filled
        const int x_start = sub_x0 - lpad;
        const int xs = max(x_start, -16), xe = min(x_start + T2_PITCH, p.W);
        const int y_lo = max(sub_y0, -1), y_hi = min(sub_y0 + t.sub_rows, p.H);
        const uint32_t row_bytes = xe > xs ? (uint32_t)(xe - xs) : 0u;
        const uint32_t total = row_bytes * (uint32_t)max(y_hi - y_lo, 0);
        if (row_bytes < T2_PITCH || y_lo > sub_y0 || y_hi < sub_y0 + t.sub_rows) {
            // edge tile: zero what the copies will not write (whole rows above / below, strips left / right)
            for (int r = 0; r < t.sub_rows; ++r) {
                const int my = sub_y0 + r;
                uint32_t *row = reinterpret_cast<uint32_t *>(dst + (size_t)r * T2_PITCH);
                if (my < y_lo || my >= y_hi || row_bytes == 0) {
                    for (int w = lane; w < T2_PITCH / 4; w += 32) row[w] = 0u;
                } else {
                    const int a = (xs - x_start) >> 2, z = (xe - x_start) >> 2;
                    for (int w = lane; w < T2_PITCH / 4; w += 32)
                        if (w < a || w >= z) row[w] = 0u;
                }
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        }
        __syncwarp();
        if (lane == 0) {
            ring[b * 4] = tile; ring[b * 4 + 1] = __ldg(t.items + 3 * item + 1); ring[b * 4 + 2] = __ldg(t.items + 3 * item + 2);
            sctr[b] = 0;
            if (total) mbar_expect_tx(full + b, total); else mbar_arrive(full + b);
        }
        __syncwarp();
        if (row_bytes)
            for (int my = y_lo + lane; my < y_hi; my += 32)
                bulk_g2s(dst + (size_t)(my - sub_y0) * T2_PITCH + (xs - x_start),
                         t.code8p + (size_t)(my + 1) * (p.W + 16) + (xs + 16), row_bytes, full + b);
    };
    const bool producer = warp == NCONS;                               // (uniform: warp comes out of a REDUX)
    if (producer) stage_use(0);
    G1Ctx k;
    k.swin = nullptr; k.slut = slut;
    k.lane = lane;
    k.nb = p.n_pos + p.n_neg;
    k.cmx = (255 << 8) | 255; k.cmy = ((t.sub_rows - 1) << 8) | 255;
    k.slut_s = smem_u32(slut) + 4u * (uint32_t)lane;
    for (int s = 0;; ++s) {
        const int b = s & 1;
        if (producer) stage_use(s + 1);
        mbar_wait(full + b, (s >> 1) & 1);
        const int tile = __reduce_min_sync(0xffffffffu, ring[b * 4]);
        if (tile < 0) break;
        const int seg_lo = __reduce_min_sync(0xffffffffu, ring[b * 4 + 1]), seg_hi = __reduce_min_sync(0xffffffffu, ring[b * 4 + 2]);
        k.swin8 = bufs + (size_t)b * buf_bytes + lpad;
        k.swin8_s = smem_u32(k.swin8);
        k.wofx = (tile % t.tiles_x) * t.tile_w - t.margin; k.wofy = (tile / t.tiles_x) * t.tile_h - t.margin;
        int pos = 0;
        for (;;) {                                        // pairs of adjacent 32-particle slices from the shared counter
            const int sl = __reduce_max_sync(0xffffffffu, lane == 0 ? atomicAdd(sctr + b, 2) : 0);
            pos = seg_lo + 32 * sl;
            if (pos >= seg_hi) break;
            g1_slices<true, true, false, 2, 2>(p, k, p.x, p.y, p.th, p.score, (int64_t)pos + lane, 32, (int64_t)seg_hi, smax, t.perm);
        }
        // never taken: keeps the hot copy of the slice code on the uniform datapath (see k_likelihood_g1)
        if (p.n < 0) g1_slices<true, true, false, 1, 2>(p, k, p.x, p.y, p.th, p.score, (int64_t)pos + lane, 0, (int64_t)seg_hi, smax, t.perm);
        __syncwarp();
        if (lane == 0) mbar_arrive(empty + b);
    }
    if (p.keymax) {
        __shared__ float smx[32];
        smax = warp_max(smax);
        if (lane == 0) smx[warp] = smax;
        __syncthreads();
        if (threadIdx.x < 32) {
            float v = lane < NCONS + 1 ? smx[lane] : -FLT_MAX;
            v = warp_max(v);
            if (lane == 0 && v > -FLT_MAX) atomicMax(p.keymax, (unsigned long long)mcl_key_of_float(v));
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
        // measured at 1 M x 351 (two sets per launch): 896 threads 0.310 ms, 960 (64 registers) 0.329, 1024 0.328, 832 0.343;
        // three / four particles per thread (9.6 / 9.4 instructions per evaluation instead of 10.06, but 640 / 512 threads
        // at 96 / 128 registers): 0.278 / 0.295 ms against 0.274 -- fewer warps cost more than fewer instructions save
        case 5: return launch_lik_kernel(h, k_likelihood_g1<SMEM, 960, 1, CODED, TPOSE>, p, smem_bytes, 1, 960, true);
        case 6: return launch_lik_kernel(h, k_likelihood_g1<SMEM, 832, 1, CODED, TPOSE>, p, smem_bytes, 1, 832, true);
        default: return launch_lik_kernel(h, k_likelihood_g1<SMEM, 896, 1, CODED, TPOSE>, p, smem_bytes, 1, 896, true);   // 72 registers: fastest measured
    }
}

// tiled path: bin the particles of one set by map tile, then the tiled kernel (timed together)
static int launch_tiled(mcl_handle *h, LikParams p, unsigned long long *keymax) {
    const int ntiles = h->tiles_x * h->tiles_y;
    const size_t o_tile = 0, o_perm = o_tile + (((size_t)p.n * 4 + 255) & ~(size_t)255);
    const size_t o_hist = o_perm + (((size_t)p.n * 4 + 255) & ~(size_t)255);
    const size_t o_off = o_hist + (((size_t)(ntiles + 1) * 4 + 255) & ~(size_t)255);
    // particles per work item at most (MCL_TILED_PIECE for A/B).  An item costs a CTA ~27 us per 1000 particles and
    // the kernel ends with the slowest CTA: in tile order 7168 left the SMs idle 16 % of the time (ncu:
    // sm__cycles_active vs elapsed), and smaller pieces pay for more staging; with the items handed out largest
    // first (k_tile_offsets) 7168 and 3584 measure the same, 11 % faster.
    static const int PIECE = [] {
        const char *e = getenv("MCL_TILED_PIECE");
        const int v = e ? atoi(e) : 0;
        return v >= 64 && v <= 16 * 448 ? (v & ~63) : 16 * 448;
    }();
    const size_t max_items = (size_t)ntiles + (size_t)(p.n / PIECE) + 2;
    const size_t o_items = o_off + (((size_t)(ntiles + 1) * 4 + 255) & ~(size_t)255);
    const size_t o_cnt = o_items + ((max_items * 12 + 255) & ~(size_t)255);
    const size_t bytes = o_cnt + 256;
    if (bytes > h->tiled_bytes) {
        MCL_CUDA(h, cudaStreamSynchronize(h->stream));
        cudaFree(h->d_tiled);
        h->d_tiled = nullptr; h->tiled_bytes = 0;
        MCL_CUDA(h, cudaMalloc(&h->d_tiled, bytes));
        h->tiled_bytes = bytes;
    }
    char *b = (char *)h->d_tiled;
    int32_t *tile_of = (int32_t *)(b + o_tile), *perm = (int32_t *)(b + o_perm), *offsets = (int32_t *)(b + o_off);
    int *hist = (int *)(b + o_hist);
    int32_t *items = (int32_t *)(b + o_items);
    int *counters = (int *)(b + o_cnt);
    // 14 warps per CTA, two CTAs per SM: a tile's ~1700 particles (27 slice pairs) fill two rounds of the warps;
    // with one 28-warp CTA most tiles left a third of the warps idle
    constexpr int THREADS = 448;
    auto kern = k_likelihood_tiled<THREADS>;
    const int sub_rows = h->tile_h + 2 * h->tile_margin;
    const size_t smem_bytes = 16 + 32768 + (size_t)sub_rows * 260;
    static thread_local int attr_dev = -1;
    if (attr_dev != h->device) {
        cudaFuncAttributes fa;
        MCL_CUDA(h, cudaFuncGetAttributes(&fa, kern));
        MCL_CUDA(h, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         h->smem_optin - (int)fa.sharedSizeBytes));
        attr_dev = h->device;
    }
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    if (h->timing) {
        MCL_CUDA(h, cudaEventCreate(&e0));
        MCL_CUDA(h, cudaEventCreate(&e1));
        MCL_CUDA(h, cudaEventRecord(e0, h->stream));
    }
    MCL_CUDA(h, cudaMemsetAsync(hist, 0, (size_t)(ntiles + 1) * 4, h->stream));
    const int gb = (int)std::max<int64_t>(1, std::min<int64_t>((p.n + 4095) / 4096, (int64_t)h->sm_count * 2));
    const bool shist = (size_t)ntiles * 8 <= 96 * 1024;
    if (shist) {
        static thread_local int bin_dev = -1;
        if (bin_dev != h->device) {
            MCL_CUDA(h, cudaFuncSetAttribute(k_tile_count<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
            MCL_CUDA(h, cudaFuncSetAttribute(k_tile_scatter<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
            bin_dev = h->device;
        }
        k_tile_count<true><<<gb, 512, (size_t)ntiles * 4, h->stream>>>(p.x, p.y, p.n, p.ox, p.oy, p.res, h->tile_w, h->tile_h,
                                                                      h->tiles_x, h->tiles_y, ntiles, tile_of, hist);
    } else {
        k_tile_count<false><<<gb, 512, 0, h->stream>>>(p.x, p.y, p.n, p.ox, p.oy, p.res, h->tile_w, h->tile_h, h->tiles_x,
                                                      h->tiles_y, ntiles, tile_of, hist);
    }
    MCL_LAUNCH_CHECK(h);
    k_tile_offsets<<<1, 1024, 0, h->stream>>>(hist, ntiles, offsets, PIECE, items, counters);
    MCL_LAUNCH_CHECK(h);
    if (shist) k_tile_scatter<true><<<gb, 512, (size_t)ntiles * 8, h->stream>>>(tile_of, p.n, ntiles, offsets, hist, perm);
    else k_tile_scatter<false><<<gb, 512, 0, h->stream>>>(tile_of, p.n, ntiles, offsets, hist, perm);
    MCL_LAUNCH_CHECK(h);
    // coordinates relative to the staged sub-window: S = 8 whatever the map size
    p.M = ldexp(1.5, 12); p.K = (int)(((uint32_t)(1023 + 12) << 20) + (1u << 19)); p.S = 8;
    p.lim = 2048.0 - (double)h->tile_margin - 4.0;
    p.x2 = nullptr; p.y2 = nullptr; p.th2 = nullptr; p.score2 = nullptr; p.keymax = keymax;
    // staged sub-windows mark cells beyond the map but do not repeat cell 0 for the int() quirk: keep the margin test
    p.ilox = p.margin; p.ihix = (double)p.W - p.margin; p.iloy = p.margin; p.ihiy = (double)p.H - p.margin;
    TiledArgs t;
    t.code8 = h->d_code8; t.code8p = h->d_code8p; t.lut = h->d_lut; t.perm = perm; t.offsets = offsets;
    t.tile_w = h->tile_w; t.tile_h = h->tile_h; t.margin = h->tile_margin; t.tiles_x = h->tiles_x; t.tiles_y = h->tiles_y;
    t.ntiles = ntiles; t.sub_rows = sub_rows; t.piece = PIECE; t.items = items; t.counters = counters;
    // second form (double-buffered bulk copies, one CTA per SM) when its two buffers fit and the rows can be copied
    // 16 bytes aligned; MCL_TILED_OLD=1 keeps the first form (A/B measurements)
    static int old_form = -1;
    if (old_form < 0) { const char *e = getenv("MCL_TILED_OLD"); old_form = (e && atoi(e)) ? 1 : 0; }
    constexpr int NCONS = 27;
    const size_t smem2 = 1024 + 32768 + 2 * (size_t)sub_rows * T2_PITCH;
    if (!old_form && (h->W & 15) == 0 && (h->tile_w & 15) == 0 && smem2 + 1024 <= (size_t)h->smem_optin) {
        auto kern2 = k_likelihood_tiled2<NCONS>;
        static thread_local int attr2_dev = -1;
        if (attr2_dev != h->device) {
            cudaFuncAttributes fa;
            MCL_CUDA(h, cudaFuncGetAttributes(&fa, kern2));
            MCL_CUDA(h, cudaFuncSetAttribute(kern2, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             h->smem_optin - (int)fa.sharedSizeBytes));
            attr2_dev = h->device;
        }
        // the staged sub-windows carry the map border (zero = outside, apron for the int() quirk): no particle needs the
        // per-beam in-map test
        p.ilox = -1e300; p.ihix = 1e300; p.iloy = -1e300; p.ihiy = 1e300;
        kern2<<<h->sm_count, (NCONS + 1) * 32, smem2, h->stream>>>(p, t);
    } else {
        const int blocks = h->sm_count * 2;
        kern<<<blocks, THREADS, smem_bytes, h->stream>>>(p, t);
    }
    MCL_LAUNCH_CHECK(h);
    if (h->timing) {
        MCL_CUDA(h, cudaEventRecord(e1, h->stream));
        h->lik_events.emplace_back(e0, e1);
        h->lik_sets_timed += 1;
    }
    return MCL_OK;
}

// The constant-bank beam table is one __constant__ symbol per device, shared by every handle on that device.  What it
// holds is identified by the scan's process-wide unique id (mcl_next_scan_uid: never reused, so a new handle at a
// recycled address can never match a stale entry); a handle on ANOTHER stream waits for the previous user's kernels
// (event recorded after every launch that reads the table) before it overwrites the symbol.
#include <mutex>
struct CbeamState { uint64_t uid = 0; cudaStream_t stream = nullptr; cudaEvent_t ev = nullptr; };
static CbeamState g_cbeams[64];
static std::mutex g_cbeams_mu;

static int cbeams_upload(mcl_handle *h, size_t beam_bytes) {
    std::lock_guard<std::mutex> lk(g_cbeams_mu);
    CbeamState &c = g_cbeams[h->device & 63];
    if (c.uid == h->scan_gen && c.stream == h->stream) return MCL_OK;
    if (c.ev && c.stream != h->stream) MCL_CUDA(h, cudaStreamWaitEvent(h->stream, c.ev, 0));
    if (c.uid != h->scan_gen)
        MCL_CUDA(h, cudaMemcpyToSymbolAsync(c_beams_raw, h->d_beams_active, beam_bytes, 0, cudaMemcpyDeviceToDevice, h->stream));
    c.uid = h->scan_gen;
    c.stream = h->stream;
    return MCL_OK;
}
static int cbeams_used(mcl_handle *h) {         // after the launches that read the table
    std::lock_guard<std::mutex> lk(g_cbeams_mu);
    CbeamState &c = g_cbeams[h->device & 63];
    if (!c.ev) MCL_CUDA(h, cudaEventCreateWithFlags(&c.ev, cudaEventDisableTiming));
    MCL_CUDA(h, cudaEventRecord(c.ev, h->stream));
    return MCL_OK;
}

// the one-thread-per-particle launches (tiled / coded window / int32 window / global path); they read the beam
// table from the constant bank
static int likelihood_g1_dispatch(mcl_handle *h, const LikParams &p, unsigned long long *d_keymax, const double *d_x2,
                                  const double *d_y2, const double *d_theta2, float *d_score2, int64_t n, bool use_smem,
                                  bool use_coded) {
    int rc;
    static int notile = -1;
    if (notile < 0) { const char *e = getenv("MCL_NO_TILED"); notile = (e && atoi(e)) ? 1 : 0; }
    if (!use_smem && !use_coded && h->tiled_ok && !notile && h->lik_path != 1 &&
        n >= 64 * (int64_t)h->tiles_x * h->tiles_y && h->rmax_cells + 2.0 <= (double)h->tile_margin) {
        rc = launch_tiled(h, p, d_keymax);
        if (rc || !d_x2) return rc;
        LikParams p2 = p;
        p2.x = d_x2; p2.y = d_y2; p2.th = d_theta2; p2.score = d_score2;
        return launch_tiled(h, p2, d_keymax ? d_keymax + 1 : nullptr);
    }
    if (use_coded) {
        const size_t sm = 16 + 32768 + h->win8_bytes;
        return h->win_tpose ? launch_g1<true, true, true>(h, p, sm) : launch_g1<true, true, false>(h, p, sm);
    }
    if (use_smem && h->cell_S == 8) {
        static int skew = -1;
        if (skew < 0) { const char *e = getenv("MCL_LIK_SKEW"); skew = e ? atoi(e) : 0; }
        if (skew && h->d_win_skew) {
            LikParams ps = p;
            ps.win = h->d_win_skew; ps.win_bytes = (uint32_t)h->win_skew_bytes;
            return h->win_tpose ? launch_lik_kernel(h, k_likelihood_g1<true, 896, 1, false, true, true>, ps, 16 + h->win_skew_bytes, 1, 896, true)
                                : launch_lik_kernel(h, k_likelihood_g1<true, 896, 1, false, false, true>, ps, 16 + h->win_skew_bytes, 1, 896, true);
        }
        return h->win_tpose ? launch_g1<true, false, true>(h, p, 16 + h->win_bytes)
                            : launch_g1<true, false, false>(h, p, 16 + h->win_bytes);
    }
    return launch_g1<false, false, false>(h, p, 16);
}

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
    p.wofx = h->win_ofx; p.wofy = h->win_ofy; p.cx = h->win_cx; p.cy = h->win_cy; p.tpose = h->win_tpose ? 1 : 0;
    p.M = h->cell_M; p.lim = h->cell_lim; p.K = h->cell_K; p.S = h->cell_S;
    p.sigma_hit = h->sigma_hit; p.z_hit = h->z_hit; p.z_rand = h->z_rand; p.max_range = h->max_range;
    p.margin = h->rmax_cells + 2.0;
    {   // EDGE sides (window runs to the map edge, border = outside) need no margin; only valid with the window
        const bool win = h->win_ok;
        const double big = 1e300;
        p.ilox = (win && (h->win_edge & 1)) ? -big : p.margin;
        p.ihix = (win && (h->win_edge & 2)) ? big : (double)h->W - p.margin;
        p.iloy = (win && (h->win_edge & 4)) ? -big : p.margin;
        p.ihiy = (win && (h->win_edge & 8)) ? big : (double)h->H - p.margin;
    }

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
        rc = cbeams_upload(h, beam_bytes);
        if (rc) return rc;
        rc = likelihood_g1_dispatch(h, p, d_keymax, d_x2, d_y2, d_theta2, d_score2, n, use_smem, use_coded);
        if (rc) return rc;
        return cbeams_used(h);
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
