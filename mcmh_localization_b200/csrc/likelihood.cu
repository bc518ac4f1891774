// likelihood.cu -- kernel (2): likelihood-field scan likelihood, pu:85-149 compute_likelihoods.
//
// Per particle i and valid beam j the reference evaluates
//     lx = x + r cos(theta + a_j),  mx = int((lx - ox) / res)   (fp64, trunc toward zero)
//     log p(dist[my*W+mx])  accumulated in fp64, mean over valid_count, stored as fp32.
// Here the per-cell value log p(dist[c]) is a precomputed fp32 table (mcl_core.cu), the beam
// endpoints r (cos a_j, sin a_j) / res are a per-scan fp64 table, and a particle needs one
// fp64 sincos; per (particle, beam) that leaves 4 DFMA + 2 F2I + one 4-byte table gather:
//     tx = px + c bx - s by,   ty = py + s bx + c by      (cell units, fp64: cell-index parity
//     with the fp64 reference needs ~1e-9 cell accuracy, SURVEY 7 hard part 1)
// The table is staged in shared memory: the free-space window of the map (all cells outside
// it hold one constant c0) plus a one-cell c0 border, so out-of-window lookups clamp onto the
// border instead of branching.  Staging uses the bulk-copy engine (cp.async.bulk + mbarrier, "TMA"
// 1-D form); CTAs are persistent and stage once.  If the window does not fit in shared memory the
// table is gathered from global memory / L2.
// Mapping: G lanes per particle (G = 1 for large N: beam constants are then warp-uniform shared
// loads; G up to 32 for small N to fill the machine), beams strided over the G lanes and reduced
// with __shfl_xor.
#include <stdlib.h>

#include <algorithm>

#include "common.cuh"

#define LIK_THREADS 512

struct LikParams {
    const double *x, *y, *th;
    int64_t n;
    float *score;
    const BeamTable *beams;
    int n_pos, n_neg;
    double ox, oy, res;
    int W, H;
    const int32_t *logtab, *win;
    const float *dist;
    const uint8_t *win8;       // coded window + table of distinct values (maps whose int32 window is too big)
    const int32_t *lut;
    uint32_t win8_bytes;
    int wx0, wy0, ww, wh;
    uint32_t win_bytes;
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

template <int G, bool SMEM>
__global__ void __launch_bounds__(LIK_THREADS, 2) k_likelihood(const LikParams p) {
    extern __shared__ __align__(128) unsigned char smem[];
    uint64_t *bar = reinterpret_cast<uint64_t *>(smem);
    BeamTable *sb = reinterpret_cast<BeamTable *>(smem + 16);
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
    const int pw = p.ww + 2;
    const int cx = p.ww + 1, cy = p.wh + 1;
    const int ofx = 1 - p.wx0, ofy = 1 - p.wy0;
    const double inv_count = 1.0;  // (division done in fp64 below, like the reference)
    (void)inv_count;
    const double lo = p.margin, hix = (double)p.W - p.margin, hiy = (double)p.H - p.margin;

    // warp-uniform trip count: every lane iterates while the FIRST group of its warp is in range
    const int64_t stride = (int64_t)gridDim.x * GROUPS;
    const int warp_first_grp = (threadIdx.x & ~31) / G;
    for (int64_t base = (int64_t)blockIdx.x * GROUPS; base + warp_first_grp < p.n; base += stride) {
        const int64_t i = base + grp;
        const int64_t il = i < p.n ? i : p.n - 1;
        const double x = p.x[il], y = p.y[il], th = p.th[il];
        double s, c;
        sincos(th, &s, &c);
        const double px = __ddiv_rn(__dadd_rn(x, -p.ox), p.res);
        const double py = __ddiv_rn(__dadd_rn(y, -p.oy), p.res);
        const bool interior = (px >= lo) && (px <= hix) && (py >= lo) && (py <= hiy);
        long long acc = 0;
        if (SMEM) {
            if (__all_sync(0xffffffffu, interior)) {
                // no endpoint can leave the map: coordinates are >= 1, trunc == floor, no bounds test
#pragma unroll 4
                for (int j = g; j < p.n_pos; j += G) {
                    const BeamTable b = sb[j];
                    const double tx = fma(c, b.bx, fma(-s, b.by, px));
                    const double ty = fma(s, b.bx, fma(c, b.by, py));
                    const int ix = min(max(__double2int_rz(tx) + ofx, 0), cx);
                    const int iy = min(max(__double2int_rz(ty) + ofy, 0), cy);
                    acc += swin[iy * pw + ix];
                }
            } else {
#pragma unroll 2
                for (int j = g; j < p.n_pos; j += G) {
                    const BeamTable b = sb[j];
                    const double tx = fma(c, b.bx, fma(-s, b.by, px));
                    const double ty = fma(s, b.bx, fma(c, b.by, py));
                    const int mx = __double2int_rz(tx), my = __double2int_rz(ty);  // pu:128-129 int()
                    const bool inmap = ((unsigned)mx < (unsigned)p.W) && ((unsigned)my < (unsigned)p.H);
                    const int ix = min(max(mx + ofx, 0), cx);
                    const int iy = min(max(my + ofy, 0), cy);
                    const int v = swin[iy * pw + ix];
                    acc += inmap ? v : 0;                                           // pu:131-132
                }
            }
        } else {
#pragma unroll 4
            for (int j = g; j < p.n_pos; j += G) {
                const BeamTable b = sb[j];
                const double tx = fma(c, b.bx, fma(-s, b.by, px));
                const double ty = fma(s, b.bx, fma(c, b.by, py));
                const int mx = __double2int_rz(tx), my = __double2int_rz(ty);
                const bool inmap = ((unsigned)mx < (unsigned)p.W) && ((unsigned)my < (unsigned)p.H);
                if (inmap) acc += __ldg(p.logtab + (size_t)my * p.W + mx);
            }
        }
        // valid beams with a negative range: p_rand = 0 (pu:139); evaluated from the distance map
        for (int j = p.n_pos + g; j < nb; j += G) {
            const BeamTable b = sb[j];
            const double tx = fma(c, b.bx, fma(-s, b.by, px));
            const double ty = fma(s, b.bx, fma(c, b.by, py));
            const int mx = __double2int_rz(tx), my = __double2int_rz(ty);
            if (((unsigned)mx < (unsigned)p.W) && ((unsigned)my < (unsigned)p.H))
                acc += quantise_logp(cell_logp(__ldg(p.dist + (size_t)my * p.W + mx), p.sigma_hit, p.z_hit,
                                               p.z_rand, p.max_range, false));
        }
#pragma unroll
        for (int o = G / 2; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
        if (g == 0 && i < p.n) p.score[i] = (float)(((double)acc / MCL_LOGP_SCALE) / (double)nb);   // pu:144-145
    }
}


// ---------------------------------------------------------------------------------------------
// G = 1 (one thread per particle, N large): every lane of a warp evaluates the SAME beam, so the
// beam constants are warp-uniform and come from the constant bank (no LSU / shared-memory traffic:
// ncu on the first version showed the shared-memory pipe at 78 % with 40 % of its wavefronts spent on
// the beam table, and the XU pipe at 66 % on the two F2I.F64 per evaluation).  The cell index is
// extracted with the round-down magic-number add (FP64 pipe, exact floor for |t| < 2^31) and the
// window offset + clamp is one VIADDMNMX (__viaddmin_s32_relu).
// ---------------------------------------------------------------------------------------------
#define MAX_CBEAMS 2048
__constant__ BeamTable c_beams[MAX_CBEAMS];
#define MCL_FLOOR_MAGIC 6755399441055744.0   // 2^52 + 2^51

__device__ __forceinline__ int floor_to_int(double t) {   // exact floor(t) for |t| < 2^31
    return __double2loint(__dadd_rd(t, MCL_FLOOR_MAGIC));
}

// Tunables (chosen by measurement, see profiles/): threads per CTA, particles per thread (the uniform beam
// loads and loop overhead are shared), minimum CTAs per SM, and MIXED = take the y index with F2I (XU pipe)
// instead of the magic add (FP64 pipe) to spread the conversions over two pipes.
template <bool SMEM, int G1_THREADS, int G1_P, int MINB, bool MIXED, bool CODED = false>
__global__ void __launch_bounds__(G1_THREADS, MINB) k_likelihood_g1(const LikParams p) {
    extern __shared__ __align__(128) unsigned char smem[];
    uint64_t *bar = reinterpret_cast<uint64_t *>(smem);
    int32_t *swin = reinterpret_cast<int32_t *>(smem + 16);
    // CODED: [16 B barrier][table of distinct values replicated per lane: 256 x 32 int32][uint8 window]
    int32_t *slut = reinterpret_cast<int32_t *>(smem + 16);
    const uint8_t *swin8 = smem + 16 + 32768;
    const int nb = p.n_pos + p.n_neg;
    const int lane = threadIdx.x & 31;
    if (SMEM) {
        if (threadIdx.x == 0) mbar_init(bar, 1);
        __syncthreads();
        if (threadIdx.x == 0) {
            if (CODED) {
                mbar_expect_tx(bar, p.win8_bytes);
                bulk_g2s_chunked(smem + 16 + 32768, p.win8, p.win8_bytes, bar);
            } else {
                mbar_expect_tx(bar, p.win_bytes);
                bulk_g2s_chunked(reinterpret_cast<unsigned char *>(swin), reinterpret_cast<const unsigned char *>(p.win),
                                 p.win_bytes, bar);
            }
        }
        if (CODED) {   // lane-private copies of the value table: slut[code * 32 + lane] is always conflict-free
            for (int e = threadIdx.x; e < 256 * 32; e += G1_THREADS) slut[e] = __ldg(p.lut + (e >> 5));
            __syncthreads();
        }
        mbar_wait(bar, 0);
    }
    auto fetch = [&](int cell) -> int { return CODED ? slut[(int)swin8[cell] * 32 + lane] : swin[cell]; };
    const int pw = p.ww + 2;
    const int cx = p.ww + 1, cy = p.wh + 1;
    const int ofx = 1 - p.wx0, ofy = 1 - p.wy0;
    const double lo = p.margin, hix = (double)p.W - p.margin, hiy = (double)p.H - p.margin;
    const double2 *cb = reinterpret_cast<const double2 *>(c_beams);
    const int64_t stride = (int64_t)gridDim.x * (G1_THREADS * G1_P);
    const int warp_first = threadIdx.x & ~31;
    for (int64_t base = (int64_t)blockIdx.x * (G1_THREADS * G1_P); base + warp_first < p.n; base += stride) {
        int64_t idx[G1_P];
        double px[G1_P], py[G1_P], s[G1_P], c[G1_P];
        bool interior = true;
#pragma unroll
        for (int q = 0; q < G1_P; ++q) {
            idx[q] = base + q * G1_THREADS + threadIdx.x;
            const int64_t il = idx[q] < p.n ? idx[q] : p.n - 1;
            const double x = p.x[il], y = p.y[il], th = p.th[il];
            sincos(th, &s[q], &c[q]);
            px[q] = __ddiv_rn(__dadd_rn(x, -p.ox), p.res);
            py[q] = __ddiv_rn(__dadd_rn(y, -p.oy), p.res);
            interior = interior && (px[q] >= lo) && (px[q] <= hix) && (py[q] >= lo) && (py[q] <= hiy);
        }
        long long acc[G1_P];
#pragma unroll
        for (int q = 0; q < G1_P; ++q) acc[q] = 0;
        if (SMEM && __all_sync(0xffffffffu, interior)) {
            // no endpoint can leave the map: coordinates >= 1, floor == trunc, no bounds test
            int j = 0;
#pragma unroll 2
            for (; j + 1 < p.n_pos; j += 2) {
                const double2 b0 = cb[j], b1 = cb[j + 1];
#pragma unroll
                for (int q = 0; q < G1_P; ++q) {
                    const double tx0 = fma(c[q], b0.x, fma(-s[q], b0.y, px[q])), ty0 = fma(s[q], b0.x, fma(c[q], b0.y, py[q]));
                    const double tx1 = fma(c[q], b1.x, fma(-s[q], b1.y, px[q])), ty1 = fma(s[q], b1.x, fma(c[q], b1.y, py[q]));
                    const int ix0 = __viaddmin_s32_relu(floor_to_int(tx0), ofx, cx);
                    const int iy0 = __viaddmin_s32_relu(MIXED ? __double2int_rz(ty0) : floor_to_int(ty0), ofy, cy);
                    const int ix1 = __viaddmin_s32_relu(floor_to_int(tx1), ofx, cx);
                    const int iy1 = __viaddmin_s32_relu(MIXED ? __double2int_rz(ty1) : floor_to_int(ty1), ofy, cy);
                    acc[q] += fetch(iy0 * pw + ix0) + fetch(iy1 * pw + ix1);      // two terms fit int32
                }
            }
            if (j < p.n_pos) {
                const double2 b0 = cb[j];
#pragma unroll
                for (int q = 0; q < G1_P; ++q) {
                    const double tx0 = fma(c[q], b0.x, fma(-s[q], b0.y, px[q])), ty0 = fma(s[q], b0.x, fma(c[q], b0.y, py[q]));
                    const int ix0 = __viaddmin_s32_relu(floor_to_int(tx0), ofx, cx);
                    const int iy0 = __viaddmin_s32_relu(floor_to_int(ty0), ofy, cy);
                    acc[q] += fetch(iy0 * pw + ix0);
                }
            }
        } else {
            for (int j = 0; j < p.n_pos; ++j) {
                const double2 b = cb[j];
#pragma unroll
                for (int q = 0; q < G1_P; ++q) {
                    const double tx = fma(c[q], b.x, fma(-s[q], b.y, px[q]));
                    const double ty = fma(s[q], b.x, fma(c[q], b.y, py[q]));
                    const int mx = __double2int_rz(tx), my = __double2int_rz(ty);      // pu:128-129 int()
                    const bool inmap = ((unsigned)mx < (unsigned)p.W) && ((unsigned)my < (unsigned)p.H);
                    int v;
                    if (SMEM) {
                        const int ix = min(max(mx + ofx, 0), cx), iy = min(max(my + ofy, 0), cy);
                        v = fetch(iy * pw + ix);
                        v = inmap ? v : 0;                                             // pu:131-132
                    } else {
                        v = inmap ? __ldg(p.logtab + (size_t)my * p.W + mx) : 0;
                    }
                    acc[q] += v;
                }
            }
        }
        // valid beams with a negative range: p_rand = 0 (pu:139); evaluated from the distance map
        for (int j = p.n_pos; j < nb; ++j) {
            const double2 b = cb[j];
#pragma unroll
            for (int q = 0; q < G1_P; ++q) {
                const double tx = fma(c[q], b.x, fma(-s[q], b.y, px[q]));
                const double ty = fma(s[q], b.x, fma(c[q], b.y, py[q]));
                const int mx = __double2int_rz(tx), my = __double2int_rz(ty);
                if (((unsigned)mx < (unsigned)p.W) && ((unsigned)my < (unsigned)p.H))
                    acc[q] += quantise_logp(cell_logp(__ldg(p.dist + (size_t)my * p.W + mx), p.sigma_hit, p.z_hit,
                                                      p.z_rand, p.max_range, false));
            }
        }
#pragma unroll
        for (int q = 0; q < G1_P; ++q)
            if (idx[q] < p.n) p.score[idx[q]] = (float)(((double)acc[q] / MCL_LOGP_SCALE) / (double)nb);   // pu:144-145
    }
}

__global__ void k_fill_f32(float *out, int64_t n, float v) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        out[i] = v;
}

template <typename K>
static int launch_lik_kernel(mcl_handle *h, K kern, const LikParams &p, size_t smem_bytes, int G,
                             int threads = LIK_THREADS, int per_thread = 1) {
    // attribute + occupancy queries are cached per (kernel, smem size): they cost tens of microseconds
    static thread_local const void *c_kern[16];
    static thread_local size_t c_smem[16];
    static thread_local int c_occ[16], c_dev[16], c_n = 0;
    int occ = 0;
    for (int k = 0; k < c_n; ++k)
        if (c_kern[k] == (const void *)kern && c_smem[k] == smem_bytes && c_dev[k] == h->device) occ = c_occ[k];
    if (occ == 0) {
        // opt in to the device maximum once (a later, smaller request must not lower the limit)
        MCL_CUDA(h, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, h->smem_optin));
        MCL_CUDA(h, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, threads, smem_bytes));
        if (occ < 1) return mcl_fail(h, MCL_ERR_CAPACITY, "likelihood kernel does not fit on an SM");
        const int k = c_n < 16 ? c_n++ : 15;
        c_kern[k] = (const void *)kern; c_smem[k] = smem_bytes; c_occ[k] = occ; c_dev[k] = h->device;
    }
    const int64_t groups = (int64_t)threads * per_thread / G;
    const int64_t need = (p.n + groups - 1) / groups;
    const int blocks = (int)std::min<int64_t>(need, (int64_t)h->sm_count * occ);
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    if (h->timing) {
        MCL_CUDA(h, cudaEventCreate(&e0));
        MCL_CUDA(h, cudaEventCreate(&e1));
        MCL_CUDA(h, cudaEventRecord(e0, h->stream));
    }
    kern<<<blocks, threads, smem_bytes, h->stream>>>(p);
    MCL_LAUNCH_CHECK(h);
    if (h->timing) {
        MCL_CUDA(h, cudaEventRecord(e1, h->stream));
        h->lik_events.emplace_back(e0, e1);
    }
    return MCL_OK;
}

template <int G, bool SMEM>
static int launch_lik(mcl_handle *h, const LikParams &p, size_t smem_bytes) {
    return launch_lik_kernel(h, k_likelihood<G, SMEM>, p, smem_bytes, G);
}

// the constant-bank beam table is module-global: re-upload when the active scan (or handle) changed
static const void *g_cbeams_src = nullptr;
static uint64_t g_cbeams_gen = 0;

extern "C" int mcl_likelihood(mcl_handle *h, const double *d_x, const double *d_y, const double *d_theta,
                              int64_t n, float *d_score) {
    if (!h) return MCL_ERR_ARG;
    if (n < 0 || (n > 0 && (!d_x || !d_y || !d_theta || !d_score)))
        return mcl_fail(h, MCL_ERR_ARG, "mcl_likelihood: bad argument");
    if (!h->scan_set) return mcl_fail(h, MCL_ERR_STATE, "mcl_likelihood: scan not set (mcl_set_scan)");
    if (n == 0) return MCL_OK;
    DeviceGuard guard(h->device);
    int rc = mcl_prepare_table(h);
    if (rc) return rc;
    const int nb = h->n_pos + h->n_neg;
    if (nb == 0) {   // pu:146-147: no valid beam -> -50 for every particle
        k_fill_f32<<<(int)std::min<int64_t>((n + 255) / 256, h->sm_count * 8), 256, 0, h->stream>>>(d_score, n, -50.0f);
        MCL_LAUNCH_CHECK(h);
        return MCL_OK;
    }
    LikParams p;
    p.x = d_x; p.y = d_y; p.th = d_theta; p.n = n; p.score = d_score;
    p.beams = h->d_beams_active; p.n_pos = h->n_pos; p.n_neg = h->n_neg;
    p.ox = h->ox; p.oy = h->oy; p.res = h->res; p.W = h->W; p.H = h->H;
    p.logtab = h->d_logtab; p.dist = h->d_dist; p.win = h->d_win;
    p.win8 = h->d_win8; p.lut = h->d_lut; p.win8_bytes = (uint32_t)h->win8_bytes;
    p.wx0 = h->wx0; p.wy0 = h->wy0; p.ww = h->ww; p.wh = h->wh; p.win_bytes = (uint32_t)h->win_bytes;
    p.sigma_hit = h->sigma_hit; p.z_hit = h->z_hit; p.z_rand = h->z_rand; p.max_range = h->max_range;
    p.margin = h->rmax_cells + 2.0;

    const size_t beam_bytes = (size_t)nb * sizeof(BeamTable);
    const size_t smem_glob = 16 + beam_bytes;
    const size_t smem_win = smem_glob + h->win_bytes;
    const size_t smem_limit = (size_t)h->smem_optin;
    if (smem_glob > smem_limit) return mcl_fail(h, MCL_ERR_CAPACITY, "mcl_likelihood: too many beams for shared memory");
    bool use_smem = smem_win <= smem_limit;
    if (h->lik_path == 1) use_smem = false;
    if (h->lik_path == 2 && !use_smem && !h->coded)
        return mcl_fail(h, MCL_ERR_CAPACITY, "mcl_likelihood: free-space window does not fit in shared memory");

    int G = 1;
    const int64_t target = (int64_t)h->sm_count * 2048;
    while (G < 32 && n * G < target) G *= 2;
    if (G == 1 && nb <= MAX_CBEAMS) {
        if (g_cbeams_src != (const void *)h->d_beams_active || g_cbeams_gen != h->scan_gen) {
            MCL_CUDA(h, cudaMemcpyToSymbolAsync(c_beams, h->d_beams_active, beam_bytes, 0, cudaMemcpyDeviceToDevice,
                                                h->stream));
            g_cbeams_src = (const void *)h->d_beams_active;
            g_cbeams_gen = h->scan_gen;
        }
        if (!use_smem && h->coded && h->lik_path != 1)
            return launch_lik_kernel(h, k_likelihood_g1<true, 512, 2, 1, false, true>, p, 16 + 32768 + h->win8_bytes, 1, 512, 2);
        static int variant = -1;
        if (variant < 0) { const char *e = getenv("MCL_LIK_VARIANT"); variant = e ? atoi(e) : 0; }
#define G1_CASE(V, T, P, B, M)                                                                              \
    case V:                                                                                                 \
        if (use_smem) return launch_lik_kernel(h, k_likelihood_g1<true, T, P, B, M>, p, 16 + h->win_bytes, 1, T, P); \
        return launch_lik_kernel(h, k_likelihood_g1<false, T, P, B, M>, p, 16, 1, T, P);
        switch (variant) {
            G1_CASE(1, 256, 2, 4, false)
            G1_CASE(2, 512, 1, 2, false)
            G1_CASE(3, 256, 1, 4, false)
            G1_CASE(4, 256, 2, 3, true)
            G1_CASE(5, 128, 2, 6, false)
            G1_CASE(6, 256, 1, 4, true)
            G1_CASE(7, 256, 4, 2, false)
            default:
            G1_CASE(0, 256, 2, 4, false)
        }
#undef G1_CASE
    }
#define LIK_CASE(GV)                                                                   \
    case GV:                                                                           \
        return use_smem ? launch_lik<GV, true>(h, p, smem_win) : launch_lik<GV, false>(h, p, smem_glob);
    switch (G) {
        LIK_CASE(1) LIK_CASE(2) LIK_CASE(4) LIK_CASE(8) LIK_CASE(16) LIK_CASE(32)
    }
#undef LIK_CASE
    return mcl_fail(h, MCL_ERR_ARG, "mcl_likelihood: internal");
}
