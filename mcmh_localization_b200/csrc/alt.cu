// alt.cu -- the functions of app/scripts/parallel_utils.py that the node imports (node:13) but that its two
// callbacks never reach: compute_valid_indices (pu:369-386), parallel_resample_simple
// (pu:467-477), reinitialize_particles_numba (pu:504-526), validate_samples (pu:600-614).
// low_variance_resample_amcl (pu:486-502) is mode MCL_RESAMPLE_AMCL_F32 of mcl_resample_indices (resample.cu).
// They exist so that the node's import line resolves unchanged against the shim and so that the node's unused
// wrappers (resample_simple, resample_amcl_simple, resample_amcl_lvr, node:441-487) keep working.
#include <algorithm>

#include "common.cuh"

// ---- order-preserving compaction by one CTA streaming over the input (these are cold paths) ------------------
struct PredValidLe10 {      // pu:376-383: cell in the map (int() truncation) and map_data[cell] <= 10
    const double *x, *y;
    const int8_t *occ;
    int W, H;
    double res, ox, oy;
    __device__ bool operator()(int64_t i) const {
        const long long mx = __double2ll_rz(__ddiv_rn(__dadd_rn(x[i], -ox), res));
        const long long my = __double2ll_rz(__ddiv_rn(__dadd_rn(y[i], -oy), res));
        if (mx >= 0 && mx < W && my >= 0 && my < H) return occ[my * (long long)W + mx] <= 10;
        return false;
    }
};
struct PredFreeCell {       // pu:506 np.argwhere(occupancy_map == 0), row-major order
    const int8_t *occ;
    __device__ bool operator()(int64_t i) const { return occ[i] == 0; }
};

template <class Pred, class Out>
__global__ void __launch_bounds__(1024) k_compact(Pred pred, int64_t n, Out *out, unsigned long long *count_out) {
    __shared__ unsigned warp_cnt[32];
    __shared__ unsigned long long base;
    if (threadIdx.x == 0) base = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int64_t t0 = 0; t0 < n; t0 += blockDim.x) {
        const int64_t i = t0 + threadIdx.x;
        const bool ok = i < n && pred(i);
        const unsigned m = __ballot_sync(0xffffffffu, ok);
        if (lane == 0) warp_cnt[warp] = __popc(m);
        __syncthreads();
        unsigned before = 0, total = 0;
        for (int k = 0; k < (int)(blockDim.x >> 5); ++k) {
            if (k < warp) before += warp_cnt[k];
            total += warp_cnt[k];
        }
        if (ok) out[base + before + __popc(m & ((1u << lane) - 1))] = (Out)i;
        __syncthreads();
        if (threadIdx.x == 0) base += total;
        __syncthreads();
    }
    if (threadIdx.x == 0) *count_out = base;
}

static int read_count(mcl_handle *h, int64_t *h_count) {
    MCL_CUDA(h, cudaMemcpyAsync(h->h_pinned, h->d_scratch, 8, cudaMemcpyDeviceToHost, h->stream));
    MCL_CUDA(h, cudaStreamSynchronize(h->stream));
    *h_count = (int64_t) * (unsigned long long *)h->h_pinned;
    return MCL_OK;
}

extern "C" int mcl_compute_valid_indices(mcl_handle *h, const double *d_x, const double *d_y, int64_t n, int32_t *d_idx,
                                         int64_t *h_count) {
    if (!h) return MCL_ERR_ARG;
    if (n < 0 || !h_count || (n > 0 && (!d_x || !d_y || !d_idx)))
        return mcl_fail(h, MCL_ERR_ARG, "mcl_compute_valid_indices: bad argument");
    if (n > 0x7fffffffLL) return mcl_fail(h, MCL_ERR_CAPACITY, "mcl_compute_valid_indices: n exceeds int32 indices");
    if (!h->d_occ) return mcl_fail(h, MCL_ERR_STATE, "mcl_compute_valid_indices: map not set");
    *h_count = 0;
    if (n == 0) return MCL_OK;
    DeviceGuard guard(h->device);
    int rc = mcl_ensure_scratch(h, 64);
    if (rc) return rc;
    PredValidLe10 p{d_x, d_y, h->d_occ, h->W, h->H, h->res, h->ox, h->oy};
    k_compact<<<1, 1024, 0, h->stream>>>(p, n, d_idx, (unsigned long long *)h->d_scratch);
    MCL_LAUNCH_CHECK(h);
    return read_count(h, h_count);
}

// ---- pu:600-614 validate_samples: a sample whose cell lies outside the map or has distance_map >= 1.0 becomes
//      (0, 0, 0); int() truncation, so coordinates in (-1, 0) cells count as cell 0 ------------------------------
__global__ void k_validate_samples(double *x, double *y, double *t, int64_t n, const float *__restrict__ dist, int W, int H,
                                   double res, double ox, double oy) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const long long mx = __double2ll_rz(__ddiv_rn(__dadd_rn(x[i], -ox), res));
        const long long my = __double2ll_rz(__ddiv_rn(__dadd_rn(y[i], -oy), res));
        const bool ok = mx >= 0 && mx < W && my >= 0 && my < H && dist[my * (long long)W + mx] < 1.0f;
        if (!ok) { x[i] = 0.0; y[i] = 0.0; t[i] = 0.0; }
    }
}

extern "C" int mcl_validate_samples(mcl_handle *h, double *d_x, double *d_y, double *d_theta, int64_t n) {
    if (!h) return MCL_ERR_ARG;
    if (n < 0 || (n > 0 && (!d_x || !d_y || !d_theta))) return mcl_fail(h, MCL_ERR_ARG, "mcl_validate_samples: bad argument");
    if (!h->d_dist) return mcl_fail(h, MCL_ERR_STATE, "mcl_validate_samples: distance map not set");
    if (n == 0) return MCL_OK;
    DeviceGuard guard(h->device);
    const int blocks = (int)std::min<int64_t>((n + 255) / 256, (int64_t)h->sm_count * 16);
    k_validate_samples<<<blocks, 256, 0, h->stream>>>(d_x, d_y, d_theta, n, h->d_dist, h->W, h->H, h->res, h->ox, h->oy);
    MCL_LAUNCH_CHECK(h);
    return MCL_OK;
}

// ---- pu:467-477 parallel_resample_simple: cum = np.cumsum(weights) (sequential f32), u ~ U[0,1) per output,
//      idx = np.searchsorted(cum, u) = first i with cum[i] >= u.  The reference reads out of bounds when
//      u > cum[-1]; here that case takes the last particle. -----------------------------
__global__ void k_search_multinomial(const float *__restrict__ c, int64_t n_in, int64_t n_out, const double *__restrict__ u_in,
                                     uint64_t seed, uint64_t step, int32_t *__restrict__ idx) {
    for (int64_t m = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; m < n_out; m += (int64_t)gridDim.x * blockDim.x) {
        double u;
        if (u_in) u = u_in[m];
        else { const uint4 o = philox_draw4(seed, step, (uint64_t)m, 0u, MCL_STREAM_RESAMPLE); u = u53_from(o.x, o.y); }
        int64_t lo = 0, hi = n_in - 1;
        while (lo < hi) {
            const int64_t mid = (lo + hi) >> 1;
            if ((double)c[mid] < u) lo = mid + 1; else hi = mid;
        }
        idx[m] = (int32_t)lo;
    }
}

extern "C" int mcl_resample_multinomial(mcl_handle *h, const float *d_w, int64_t n_in, int64_t n_out, const double *d_u,
                                        uint64_t seed, uint64_t step, int32_t *d_idx) {
    if (!h) return MCL_ERR_ARG;
    if (n_in <= 0 || n_out < 0 || !d_w || (n_out > 0 && !d_idx))
        return mcl_fail(h, MCL_ERR_ARG, "mcl_resample_multinomial: bad argument");
    if (n_in > 0x7fffffffLL) return mcl_fail(h, MCL_ERR_CAPACITY, "mcl_resample_multinomial: n_in exceeds int32 indices");
    if (n_out == 0) return MCL_OK;
    DeviceGuard guard(h->device);
    int rc = mcl_ensure_scratch(h, 64 + sizeof(float) * (size_t)n_in);
    if (rc) return rc;
    float *c = (float *)((char *)h->d_scratch + 64);
    rc = mcl_cumsum_f32_seq(h, d_w, n_in, c);                       // pu:470 np.cumsum, exact sequential f32
    if (rc) return rc;
    const int blocks = (int)std::min<int64_t>((n_out + 255) / 256, (int64_t)h->sm_count * 16);
    k_search_multinomial<<<blocks, 256, 0, h->stream>>>(c, n_in, n_out, d_u, seed, step, d_idx);
    MCL_LAUNCH_CHECK(h);
    return MCL_OK;
}

// ---- pu:504-526 reinitialize_particles_numba: a uniformly chosen FREE cell (its lower-left corner) and a uniform
//      heading per new particle.  choice / theta: injected draws (tests); otherwise Philox. ---------------------
__global__ void k_reinit(int64_t n, const int32_t *__restrict__ free_cells, int64_t n_free, int W, double res, double ox,
                         double oy, const int64_t *__restrict__ choice, const double *__restrict__ theta, uint64_t seed,
                         uint64_t step, double *xo, double *yo, double *to) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const uint4 o = philox_draw4(seed, step, (uint64_t)i, 0u, MCL_STREAM_INIT);
        const double th = theta ? theta[i] : __dadd_rn(-MCL_PI, __dmul_rn(__dadd_rn(MCL_PI, MCL_PI), u53_from(o.z, o.w)));
        if (n_free == 0) { xo[i] = ox; yo[i] = oy; to[i] = th; continue; }          // pu:509-512
        int64_t c = choice ? choice[i] : (int64_t)(u53_from(o.x, o.y) * (double)n_free);
        c = c < 0 ? 0 : (c >= n_free ? n_free - 1 : c);
        const int cell = free_cells[c];
        const int my = cell / W, mx = cell - my * W;
        xo[i] = __dadd_rn(__dmul_rn((double)mx, res), ox);                           // pu:520-521
        yo[i] = __dadd_rn(__dmul_rn((double)my, res), oy);
        to[i] = th;
    }
}

extern "C" int mcl_reinitialize_particles(mcl_handle *h, int64_t n, const int64_t *d_choice, const double *d_theta,
                                          uint64_t seed, uint64_t step, double *d_x, double *d_y, double *d_t,
                                          int64_t *h_n_free) {
    if (!h) return MCL_ERR_ARG;
    if (n < 0 || (n > 0 && (!d_x || !d_y || !d_t))) return mcl_fail(h, MCL_ERR_ARG, "mcl_reinitialize_particles: bad argument");
    if (!h->d_occ) return mcl_fail(h, MCL_ERR_STATE, "mcl_reinitialize_particles: map not set");
    DeviceGuard guard(h->device);
    const int64_t cells = (int64_t)h->W * h->H;
    int rc = mcl_ensure_scratch(h, 64 + sizeof(int32_t) * (size_t)cells);
    if (rc) return rc;
    int32_t *free_cells = (int32_t *)((char *)h->d_scratch + 64);
    PredFreeCell p{h->d_occ};
    k_compact<<<1, 1024, 0, h->stream>>>(p, cells, free_cells, (unsigned long long *)h->d_scratch);
    MCL_LAUNCH_CHECK(h);
    int64_t n_free = 0;
    rc = read_count(h, &n_free);
    if (rc) return rc;
    if (h_n_free) *h_n_free = n_free;
    if (n == 0) return MCL_OK;
    const int blocks = (int)std::min<int64_t>((n + 255) / 256, (int64_t)h->sm_count * 16);
    k_reinit<<<blocks, 256, 0, h->stream>>>(n, free_cells, n_free, h->W, h->res, h->ox, h->oy, d_choice, d_theta, seed, step,
                                            d_x, d_y, d_t);
    MCL_LAUNCH_CHECK(h);
    return MCL_OK;
}
