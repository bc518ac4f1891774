// init_misc.cu -- pu:450-465 generate_valid_particles (+ pu:398-413 compute_valid_mask),
// AoS<->SoA conversion at the shim boundary, and the gather-rate microbenchmarks that give
// bench.py its shared-memory / L2 gather roofline denominators.
#include <algorithm>

#include "common.cuh"

// ---- injected-uniform restatement: x = lo + (hi - lo) * u (np.random.uniform), keep the first n
//      valid trials in trial order.  Single CTA, order-preserving compaction (tests / small N). ----
__global__ void __launch_bounds__(1024) k_init_injected(const double *__restrict__ u, int64_t max_trials, int64_t n,
                                                        const int8_t *__restrict__ occ, int W, int H, double res,
                                                        double ox, double oy, double *xo, double *yo, double *tho,
                                                        unsigned long long *count_out) {
    __shared__ unsigned warp_cnt[32];
    __shared__ unsigned long long base;
    if (threadIdx.x == 0) base = 0;
    __syncthreads();
    const double hix = __dadd_rn(ox, __dmul_rn((double)W, res)), hiy = __dadd_rn(oy, __dmul_rn((double)H, res));
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int64_t t0 = 0; t0 < max_trials; t0 += blockDim.x) {
        const int64_t t = t0 + threadIdx.x;
        bool ok = false;
        double x = 0, y = 0, th = 0;
        if (t < max_trials) {
            x = __dadd_rn(ox, __dmul_rn(__dadd_rn(hix, -ox), u[t]));
            y = __dadd_rn(oy, __dmul_rn(__dadd_rn(hiy, -oy), u[max_trials + t]));
            th = __dadd_rn(-MCL_PI, __dmul_rn(__dadd_rn(MCL_PI, MCL_PI), u[2 * max_trials + t]));
            ok = is_valid_position_dev(x, y, occ, W, H, res, ox, oy);
        }
        const unsigned m = __ballot_sync(0xffffffffu, ok);
        if (lane == 0) warp_cnt[warp] = __popc(m);
        __syncthreads();
        unsigned before = 0, total = 0;
        for (int k = 0; k < (int)(blockDim.x >> 5); ++k) {
            if (k < warp) before += warp_cnt[k];
            total += warp_cnt[k];
        }
        const unsigned long long pos = base + before + __popc(m & ((1u << lane) - 1));
        if (ok && pos < (unsigned long long)n) { xo[pos] = x; yo[pos] = y; tho[pos] = th; }
        __syncthreads();
        if (threadIdx.x == 0) base += total;
        __syncthreads();
        if (base >= (unsigned long long)n) break;
    }
    if (threadIdx.x == 0) *count_out = base < (unsigned long long)n ? base : (unsigned long long)n;
}

// ---- production: per-particle rejection sampling with Philox (attempt a uses sub-counters 2a, 2a+1) ----
__global__ void k_init_philox(int64_t n, uint64_t seed, uint64_t first_index, const int8_t *__restrict__ occ, int W,
                              int H, double res, double ox, double oy, int max_attempts, double *xo, double *yo,
                              double *tho) {
    const double hix = __dadd_rn(ox, __dmul_rn((double)W, res)), hiy = __dadd_rn(oy, __dmul_rn((double)H, res));
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        double x = ox, y = oy, th = 0;
        for (int a = 0; a < max_attempts; ++a) {
            const uint4 o0 = philox_draw4(seed, 0, first_index + (uint64_t)i, 2u * a, MCL_STREAM_INIT);
            const uint4 o1 = philox_draw4(seed, 0, first_index + (uint64_t)i, 2u * a + 1u, MCL_STREAM_INIT);
            x = __dadd_rn(ox, __dmul_rn(__dadd_rn(hix, -ox), u53_from(o0.x, o0.y)));
            y = __dadd_rn(oy, __dmul_rn(__dadd_rn(hiy, -oy), u53_from(o0.z, o0.w)));
            th = __dadd_rn(-MCL_PI, __dmul_rn(__dadd_rn(MCL_PI, MCL_PI), u53_from(o1.x, o1.y)));
            if (is_valid_position_dev(x, y, occ, W, H, res, ox, oy)) break;
        }
        xo[i] = x; yo[i] = y; tho[i] = th;
    }
}

extern "C" int mcl_init_uniform(mcl_handle *h, int64_t n, const double *d_u, int64_t max_trials, uint64_t seed,
                                uint64_t first_index, double *d_x, double *d_y, double *d_theta, int64_t *h_count) {
    if (!h) return MCL_ERR_ARG;
    if (n < 0 || (n > 0 && (!d_x || !d_y || !d_theta))) return mcl_fail(h, MCL_ERR_ARG, "mcl_init_uniform: bad argument");
    if (!h->d_occ) return mcl_fail(h, MCL_ERR_STATE, "mcl_init_uniform: map not set");
    if (h_count) *h_count = 0;
    if (n == 0) return MCL_OK;
    DeviceGuard guard(h->device);
    if (d_u) {
        if (max_trials <= 0) return mcl_fail(h, MCL_ERR_ARG, "mcl_init_uniform: max_trials <= 0");
        int rc = mcl_ensure_scratch(h, 64);
        if (rc) return rc;
        k_init_injected<<<1, 1024, 0, h->stream>>>(d_u, max_trials, n, h->d_occ, h->W, h->H, h->res, h->ox, h->oy,
                                                   d_x, d_y, d_theta, (unsigned long long *)h->d_scratch);
        MCL_LAUNCH_CHECK(h);
        MCL_CUDA(h, cudaMemcpyAsync(h->h_pinned, h->d_scratch, 8, cudaMemcpyDeviceToHost, h->stream));
        MCL_CUDA(h, cudaStreamSynchronize(h->stream));
        if (h_count) *h_count = (int64_t) * (unsigned long long *)h->h_pinned;
        return MCL_OK;
    }
    const int blocks = (int)std::min<int64_t>((n + 255) / 256, (int64_t)h->sm_count * 16);
    k_init_philox<<<blocks, 256, 0, h->stream>>>(n, seed, first_index, h->d_occ, h->W, h->H, h->res, h->ox, h->oy,
                                                 100000, d_x, d_y, d_theta);
    MCL_LAUNCH_CHECK(h);
    if (h_count) *h_count = n;
    return MCL_OK;
}

// ---- AoS (n,3) <-> SoA ----
__global__ void k_aos_to_soa(const double *__restrict__ a, int64_t n, double *__restrict__ x, double *__restrict__ y,
                             double *__restrict__ t) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        x[i] = a[3 * i]; y[i] = a[3 * i + 1]; t[i] = a[3 * i + 2];
    }
}
__global__ void k_soa_to_aos(const double *__restrict__ x, const double *__restrict__ y, const double *__restrict__ t,
                             int64_t n, double *__restrict__ a) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        a[3 * i] = x[i]; a[3 * i + 1] = y[i]; a[3 * i + 2] = t[i];
    }
}
extern "C" int mcl_aos_to_soa(mcl_handle *h, const double *d_aos, int64_t n, double *d_x, double *d_y, double *d_t) {
    if (!h) return MCL_ERR_ARG;
    if (n < 0 || (n > 0 && (!d_aos || !d_x || !d_y || !d_t))) return mcl_fail(h, MCL_ERR_ARG, "mcl_aos_to_soa: bad argument");
    if (n == 0) return MCL_OK;
    DeviceGuard guard(h->device);
    k_aos_to_soa<<<(int)std::min<int64_t>((n + 255) / 256, (int64_t)h->sm_count * 16), 256, 0, h->stream>>>(d_aos, n, d_x, d_y, d_t);
    MCL_LAUNCH_CHECK(h);
    return MCL_OK;
}
extern "C" int mcl_soa_to_aos(mcl_handle *h, const double *d_x, const double *d_y, const double *d_t, int64_t n, double *d_aos) {
    if (!h) return MCL_ERR_ARG;
    if (n < 0 || (n > 0 && (!d_aos || !d_x || !d_y || !d_t))) return mcl_fail(h, MCL_ERR_ARG, "mcl_soa_to_aos: bad argument");
    if (n == 0) return MCL_OK;
    DeviceGuard guard(h->device);
    k_soa_to_aos<<<(int)std::min<int64_t>((n + 255) / 256, (int64_t)h->sm_count * 16), 256, 0, h->stream>>>(d_x, d_y, d_t, n, d_aos);
    MCL_LAUNCH_CHECK(h);
    return MCL_OK;
}

// ---- gather-rate microbenchmarks (roofline denominators for the likelihood kernel) ----
// Every thread performs `per_thread` independent pseudo-random 4-byte lookups (LCG indices), 4-way
// unrolled.  where = 0: table in shared memory (bank-conflict statistics of a random gather are part
// of the ceiling); where = 1: table in global memory (L1/L2 path).
template <bool SMEM>
__global__ void __launch_bounds__(512, 2) k_gather_bench(const float *__restrict__ table, uint32_t entries,
                                                         int per_thread, float *out) {
    extern __shared__ float st[];
    if (SMEM) {
        for (uint32_t i = threadIdx.x; i < entries; i += blockDim.x) st[i] = table[i];
        __syncthreads();
    }
    uint32_t s = (blockIdx.x * blockDim.x + threadIdx.x) * 2654435761u + 12345u;
    float a0 = 0, a1 = 0, a2 = 0, a3 = 0;
    for (int k = 0; k < per_thread; k += 4) {
        s = s * 1664525u + 1013904223u; const uint32_t i0 = __umulhi(s, entries);
        s = s * 1664525u + 1013904223u; const uint32_t i1 = __umulhi(s, entries);
        s = s * 1664525u + 1013904223u; const uint32_t i2 = __umulhi(s, entries);
        s = s * 1664525u + 1013904223u; const uint32_t i3 = __umulhi(s, entries);
        if (SMEM) { a0 += st[i0]; a1 += st[i1]; a2 += st[i2]; a3 += st[i3]; }
        else { a0 += __ldg(table + i0); a1 += __ldg(table + i1); a2 += __ldg(table + i2); a3 += __ldg(table + i3); }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3;
}

__global__ void k_fill_table(float *t, int64_t n) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        t[i] = (float)(i & 1023) * 1e-3f;
}

extern "C" int mcl_bench_gather(mcl_handle *h, int where, int64_t table_bytes, int64_t n_lookups, int iters,
                                double *lookups_per_s) {
    if (!h || !lookups_per_s || table_bytes < 4 || n_lookups <= 0 || iters <= 0)
        return h ? mcl_fail(h, MCL_ERR_ARG, "mcl_bench_gather: bad argument") : MCL_ERR_ARG;
    DeviceGuard guard(h->device);
    const int64_t entries = table_bytes / 4;
    if (entries > 0x7fffffffLL) return mcl_fail(h, MCL_ERR_CAPACITY, "mcl_bench_gather: table too large");
    if (where == 0 && table_bytes > h->smem_optin) return mcl_fail(h, MCL_ERR_CAPACITY, "mcl_bench_gather: table exceeds shared memory");
    float *table = nullptr, *out = nullptr;
    const int threads = 512;
    int occ = 0;
    const size_t smem = where == 0 ? (size_t)table_bytes : 0;
    if (where == 0) {
        MCL_CUDA(h, cudaFuncSetAttribute(k_gather_bench<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        MCL_CUDA(h, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_gather_bench<true>, threads, smem));
    } else {
        MCL_CUDA(h, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_gather_bench<false>, threads, 0));
    }
    if (occ < 1) return mcl_fail(h, MCL_ERR_CAPACITY, "mcl_bench_gather: kernel does not fit");
    const int blocks = h->sm_count * occ;
    const int64_t nthreads = (int64_t)blocks * threads;
    int per_thread = (int)std::max<int64_t>(4, ((n_lookups + nthreads - 1) / nthreads + 3) / 4 * 4);
    MCL_CUDA(h, cudaMalloc((void **)&table, (size_t)entries * 4));
    MCL_CUDA(h, cudaMalloc((void **)&out, (size_t)nthreads * 4));
    k_fill_table<<<h->sm_count * 4, 256, 0, h->stream>>>(table, entries);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int it = 0; it < iters + 2; ++it) {
        if (it == 2) cudaEventRecord(e0, h->stream);
        if (where == 0) k_gather_bench<true><<<blocks, threads, smem, h->stream>>>(table, (uint32_t)entries, per_thread, out);
        else k_gather_bench<false><<<blocks, threads, 0, h->stream>>>(table, (uint32_t)entries, per_thread, out);
        h->launches++;
    }
    cudaEventRecord(e1, h->stream);
    cudaError_t e = cudaStreamSynchronize(h->stream);
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    cudaFree(table); cudaFree(out);
    if (e != cudaSuccess) return mcl_fail(h, MCL_ERR_CUDA, cudaGetErrorString(e));
    *lookups_per_s = (double)nthreads * per_thread * iters / (ms * 1e-3);
    return MCL_OK;
}
