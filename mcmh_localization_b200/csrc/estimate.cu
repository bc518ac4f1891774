// estimate.cu -- node:586-597 publish_estimate arithmetic (weighted mean, circular mean,
// weighted covariance of (dx, dy, wrap(dtheta))) as two fp64 block reductions.
//   pass 1: V1 = sum w, V2 = sum w^2, sum w x, sum w y, sum w cos(theta), sum w sin(theta)
//   pass 2: d = (x - mean_x, y - mean_y, f32(normalize_angle(theta - mean_theta)))   (pu:69-83: f32!)
//           sum w d (3), sum w d d^T (6)
// np.cov(diffs.T, aweights=w) is then assembled from those 9 numbers by the caller:
//   cov = (S_dd - S_d S_d^T / V1) / (V1 - V2 / V1).
// Per-block partials are combined in a fixed order by the last block: results are deterministic.
#include <algorithm>

#include "common.cuh"

#define EST_THREADS 256

template <int K>
__device__ __forceinline__ void block_sum_k(double (&v)[K], double *sh /* K*32 */) {
#pragma unroll
    for (int k = 0; k < K; ++k) v[k] = warp_sum(v[k]);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (lane == 0)
#pragma unroll
        for (int k = 0; k < K; ++k) sh[k * 32 + warp] = v[k];
    __syncthreads();
    if (warp == 0) {
#pragma unroll
        for (int k = 0; k < K; ++k) {
            double t = lane < (int)(blockDim.x >> 5) ? sh[k * 32 + lane] : 0.0;
            t = warp_sum(t);
            v[k] = t;
        }
    }
    __syncthreads();
}

template <int K>
__device__ __forceinline__ void finish_partials(double (&v)[K], double *partials, unsigned *counter, double *out,
                                                double *sh) {
    __shared__ bool last;
    if (threadIdx.x == 0) {
#pragma unroll
        for (int k = 0; k < K; ++k) partials[(size_t)blockIdx.x * K + k] = v[k];
        __threadfence();
        last = atomicAdd(counter, 1u) == gridDim.x - 1;
    }
    __syncthreads();
    if (last) {
        __threadfence();
        double t[K];
#pragma unroll
        for (int k = 0; k < K; ++k) t[k] = 0.0;
        for (unsigned b = threadIdx.x; b < gridDim.x; b += blockDim.x)
#pragma unroll
            for (int k = 0; k < K; ++k) t[k] += ((volatile double *)partials)[(size_t)b * K + k];
        block_sum_k<K>(t, sh);
        if (threadIdx.x == 0) {
#pragma unroll
            for (int k = 0; k < K; ++k) out[k] = t[k];
            *counter = 0;
        }
    }
}

// out[0..5] raw sums; out[6..8] = mean_x, mean_y, mean_theta
__global__ void __launch_bounds__(EST_THREADS) k_est_moments(const double *__restrict__ x, const double *__restrict__ y,
                                                             const double *__restrict__ th, const float *__restrict__ w,
                                                             int64_t n, double *partials, unsigned *counter, double *out) {
    __shared__ double sh[6 * 32];
    double v[6] = {0, 0, 0, 0, 0, 0};
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const double wi = (double)w[i];
        double s, c;
        sincos(th[i], &s, &c);
        v[0] += wi; v[1] += wi * wi; v[2] += wi * x[i]; v[3] += wi * y[i]; v[4] += wi * c; v[5] += wi * s;
    }
    block_sum_k<6>(v, sh);
    finish_partials<6>(v, partials, counter, out, sh);
}

__global__ void k_est_means(double *out) {
    out[6] = out[2] / out[0];              // np.average: sum(w x) / sum(w)
    out[7] = out[3] / out[0];
    out[8] = atan2(out[5], out[4]);        // node:589 arctan2(sin_mean, cos_mean)
}

__global__ void __launch_bounds__(EST_THREADS) k_est_central(const double *__restrict__ x, const double *__restrict__ y,
                                                             const double *__restrict__ th, const float *__restrict__ w,
                                                             int64_t n, const double *mean_ptr, double mx_, double my_,
                                                             double mt_, int use_args, double *partials,
                                                             unsigned *counter, double *out) {
    __shared__ double sh[9 * 32];
    // use_args = 2: mean_ptr points at the six raw sums; derive the means here (saves the k_est_means launch)
    double mx, my, mt;
    if (use_args == 2) {
        mx = mean_ptr[2] / mean_ptr[0]; my = mean_ptr[3] / mean_ptr[0]; mt = atan2(mean_ptr[5], mean_ptr[4]);
        if (blockIdx.x == 0 && threadIdx.x == 0) { out[-3] = mx; out[-2] = my; out[-1] = mt; }
    } else {
        mx = use_args ? mx_ : mean_ptr[0]; my = use_args ? my_ : mean_ptr[1]; mt = use_args ? mt_ : mean_ptr[2];
    }
    double v[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const double wi = (double)w[i];
        const double dx = x[i] - mx, dy = y[i] - my;
        const double dt = (double)(float)normalize_angle_dev(__dadd_rn(th[i], -mt));   // pu:80-82
        v[0] += wi * dx; v[1] += wi * dy; v[2] += wi * dt;
        v[3] += wi * dx * dx; v[4] += wi * dx * dy; v[5] += wi * dx * dt;
        v[6] += wi * dy * dy; v[7] += wi * dy * dt; v[8] += wi * dt * dt;
    }
    block_sum_k<9>(v, sh);
    finish_partials<9>(v, partials, counter, out, sh);
}

static int est_blocks(const mcl_handle *h, int64_t n) {
    return (int)std::max<int64_t>(1, std::min<int64_t>((n + EST_THREADS * 2 - 1) / (EST_THREADS * 2), (int64_t)h->sm_count * 4));
}

// scratch: [0,64) counter | [64, 64+32*8) results (moments 0..8, central 9..17) | partials
static int est_setup(mcl_handle *h, int64_t n, int &nb, unsigned *&counter, double *&res, double *&partials) {
    nb = est_blocks(h, n);
    int rc = mcl_ensure_scratch(h, 64 + 32 * 8 + (size_t)nb * 9 * 8);
    if (rc) return rc;
    char *s = (char *)h->d_scratch;
    counter = (unsigned *)s;
    res = (double *)(s + 64);
    partials = (double *)(s + 64 + 32 * 8);
    return MCL_OK;
}

static int check_est_args(mcl_handle *h, const double *x, const double *y, const double *t, const float *w, int64_t n) {
    if (n <= 0 || !x || !y || !t || !w) return mcl_fail(h, MCL_ERR_ARG, "mcl_estimate: bad argument");
    return MCL_OK;
}

extern "C" int mcl_estimate_moments(mcl_handle *h, const double *d_x, const double *d_y, const double *d_theta,
                                    const float *d_w, int64_t n, double h_m[6]) {
    if (!h || !h_m) return MCL_ERR_ARG;
    int rc = check_est_args(h, d_x, d_y, d_theta, d_w, n);
    if (rc) return rc;
    DeviceGuard guard(h->device);
    int nb; unsigned *counter; double *res, *partials;
    rc = est_setup(h, n, nb, counter, res, partials);
    if (rc) return rc;
    MCL_CUDA(h, cudaMemsetAsync(counter, 0, sizeof(unsigned), h->stream));
    k_est_moments<<<nb, EST_THREADS, 0, h->stream>>>(d_x, d_y, d_theta, d_w, n, partials, counter, res);
    MCL_LAUNCH_CHECK(h);
    MCL_CUDA(h, cudaMemcpyAsync(h->h_pinned, res, 6 * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    MCL_CUDA(h, cudaStreamSynchronize(h->stream));
    for (int k = 0; k < 6; ++k) h_m[k] = h->h_pinned[k];
    return MCL_OK;
}

extern "C" int mcl_estimate_central(mcl_handle *h, const double *d_x, const double *d_y, const double *d_theta,
                                    const float *d_w, int64_t n, const double mean[3], double h_c[9]) {
    if (!h || !mean || !h_c) return MCL_ERR_ARG;
    int rc = check_est_args(h, d_x, d_y, d_theta, d_w, n);
    if (rc) return rc;
    DeviceGuard guard(h->device);
    int nb; unsigned *counter; double *res, *partials;
    rc = est_setup(h, n, nb, counter, res, partials);
    if (rc) return rc;
    MCL_CUDA(h, cudaMemsetAsync(counter, 0, sizeof(unsigned), h->stream));
    k_est_central<<<nb, EST_THREADS, 0, h->stream>>>(d_x, d_y, d_theta, d_w, n, nullptr, mean[0], mean[1], mean[2], 1,
                                                     partials, counter, res + 9);
    MCL_LAUNCH_CHECK(h);
    MCL_CUDA(h, cudaMemcpyAsync(h->h_pinned, res + 9, 9 * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    MCL_CUDA(h, cudaStreamSynchronize(h->stream));
    for (int k = 0; k < 9; ++k) h_c[k] = h->h_pinned[k];
    return MCL_OK;
}

// non-blocking form: the 18 doubles {moments[6], means[3], central[9]} land in d_out18 (device)
extern "C" int mcl_estimate_async(mcl_handle *h, const double *d_x, const double *d_y, const double *d_theta,
                                  const float *d_w, int64_t n, double *d_out18) {
    if (!h || !d_out18) return MCL_ERR_ARG;
    int rc = check_est_args(h, d_x, d_y, d_theta, d_w, n);
    if (rc) return rc;
    DeviceGuard guard(h->device);
    int nb; unsigned *counter; double *res, *partials;
    rc = est_setup(h, n, nb, counter, res, partials);
    if (rc) return rc;
    MCL_CUDA(h, cudaMemsetAsync(counter, 0, sizeof(unsigned), h->stream));
    k_est_moments<<<nb, EST_THREADS, 0, h->stream>>>(d_x, d_y, d_theta, d_w, n, partials, counter, d_out18);
    MCL_LAUNCH_CHECK(h);
    // means derived inside the central pass from the raw sums (written to d_out18[6..8] by block 0)
    k_est_central<<<nb, EST_THREADS, 0, h->stream>>>(d_x, d_y, d_theta, d_w, n, d_out18, 0, 0, 0, 2, partials, counter,
                                                     d_out18 + 9);
    MCL_LAUNCH_CHECK(h);
    return MCL_OK;
}

// staged forms (sharded path): raw sums -> [all-reduce] -> means -> central sums -> [all-reduce]
extern "C" int mcl_estimate_moments_async(mcl_handle *h, const double *d_x, const double *d_y, const double *d_theta,
                                          const float *d_w, int64_t n, double *d_m9) {
    if (!h || !d_m9) return MCL_ERR_ARG;
    int rc = check_est_args(h, d_x, d_y, d_theta, d_w, n);
    if (rc) return rc;
    DeviceGuard guard(h->device);
    int nb; unsigned *counter; double *res, *partials;
    rc = est_setup(h, n, nb, counter, res, partials);
    if (rc) return rc;
    MCL_CUDA(h, cudaMemsetAsync(counter, 0, sizeof(unsigned), h->stream));
    k_est_moments<<<nb, EST_THREADS, 0, h->stream>>>(d_x, d_y, d_theta, d_w, n, partials, counter, d_m9);
    MCL_LAUNCH_CHECK(h);
    return MCL_OK;
}
extern "C" int mcl_estimate_means_async(mcl_handle *h, double *d_m9) {
    if (!h || !d_m9) return MCL_ERR_ARG;
    DeviceGuard guard(h->device);
    k_est_means<<<1, 1, 0, h->stream>>>(d_m9);
    MCL_LAUNCH_CHECK(h);
    return MCL_OK;
}
extern "C" int mcl_estimate_central_async(mcl_handle *h, const double *d_x, const double *d_y, const double *d_theta,
                                          const float *d_w, int64_t n, const double *d_mean3, double *d_c9) {
    if (!h || !d_mean3 || !d_c9) return MCL_ERR_ARG;
    int rc = check_est_args(h, d_x, d_y, d_theta, d_w, n);
    if (rc) return rc;
    DeviceGuard guard(h->device);
    int nb; unsigned *counter; double *res, *partials;
    rc = est_setup(h, n, nb, counter, res, partials);
    if (rc) return rc;
    MCL_CUDA(h, cudaMemsetAsync(counter, 0, sizeof(unsigned), h->stream));
    k_est_central<<<nb, EST_THREADS, 0, h->stream>>>(d_x, d_y, d_theta, d_w, n, d_mean3, 0, 0, 0, 0, partials, counter, d_c9);
    MCL_LAUNCH_CHECK(h);
    return MCL_OK;
}

extern "C" int mcl_estimate(mcl_handle *h, const double *d_x, const double *d_y, const double *d_theta,
                            const float *d_w, int64_t n, double h_out[16]) {
    if (!h || !h_out) return MCL_ERR_ARG;
    int rc = check_est_args(h, d_x, d_y, d_theta, d_w, n);
    if (rc) return rc;
    DeviceGuard guard(h->device);
    int nb; unsigned *counter; double *res, *partials;
    rc = est_setup(h, n, nb, counter, res, partials);
    if (rc) return rc;
    MCL_CUDA(h, cudaMemsetAsync(counter, 0, sizeof(unsigned), h->stream));
    k_est_moments<<<nb, EST_THREADS, 0, h->stream>>>(d_x, d_y, d_theta, d_w, n, partials, counter, res);
    MCL_LAUNCH_CHECK(h);
    k_est_means<<<1, 1, 0, h->stream>>>(res);
    MCL_LAUNCH_CHECK(h);
    k_est_central<<<nb, EST_THREADS, 0, h->stream>>>(d_x, d_y, d_theta, d_w, n, res + 6, 0, 0, 0, 0, partials, counter,
                                                     res + 9);
    MCL_LAUNCH_CHECK(h);
    MCL_CUDA(h, cudaMemcpyAsync(h->h_pinned, res, 18 * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    MCL_CUDA(h, cudaStreamSynchronize(h->stream));
    const double *r = h->h_pinned;
    h_out[0] = r[0]; h_out[1] = r[1]; h_out[2] = r[6]; h_out[3] = r[7]; h_out[4] = r[8];
    for (int k = 0; k < 9; ++k) h_out[5 + k] = r[9 + k];
    h_out[14] = 0; h_out[15] = 0;
    return MCL_OK;
}

// pu:69-83 normalize_angle_array (f32 result)
__global__ void k_normalize_angle_array(const double *__restrict__ a, double mean, int64_t n, float *__restrict__ out) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        out[i] = (float)normalize_angle_dev(__dadd_rn(a[i], -mean));
}
extern "C" int mcl_normalize_angle_array(mcl_handle *h, const double *d_angles, double mean_angle, int64_t n,
                                         float *d_out) {
    if (!h) return MCL_ERR_ARG;
    if (n < 0 || (n > 0 && (!d_angles || !d_out))) return mcl_fail(h, MCL_ERR_ARG, "mcl_normalize_angle_array: bad argument");
    if (n == 0) return MCL_OK;
    DeviceGuard guard(h->device);
    k_normalize_angle_array<<<(int)std::min<int64_t>((n + 255) / 256, (int64_t)h->sm_count * 16), 256, 0, h->stream>>>(
        d_angles, mean_angle, n, d_out);
    MCL_LAUNCH_CHECK(h);
    return MCL_OK;
}
