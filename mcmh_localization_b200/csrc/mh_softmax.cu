// mh_softmax.cu -- node:351-358 convert_scores (softmax) and kernel (3): pu:208-236 mh_resampling.
#include <algorithm>
#include <float.h>

#include "common.cuh"

#define RED_THREADS 256

__device__ __forceinline__ float block_max(float v, float *sh) {
    v = warp_max(v);
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
    __syncthreads();
    if (threadIdx.x < 32) {
        float t = threadIdx.x < (blockDim.x >> 5) ? sh[threadIdx.x] : -FLT_MAX;
        t = warp_max(t);
        if (threadIdx.x == 0) sh[0] = t;
    }
    __syncthreads();
    const float r = sh[0];
    __syncthreads();
    return r;
}
__device__ __forceinline__ double block_sum(double v, double *sh) {
    v = warp_sum(v);
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
    __syncthreads();
    if (threadIdx.x < 32) {
        double t = threadIdx.x < (blockDim.x >> 5) ? sh[threadIdx.x] : 0.0;
        t = warp_sum(t);
        if (threadIdx.x == 0) sh[0] = t;
    }
    __syncthreads();
    const double r = sh[0];
    __syncthreads();
    return r;
}

// scratch layout (bytes): [0,8) uint counter | [64, 64+8*nb) partials | stats live in caller memory
struct RedScratch {
    unsigned *counter;
    double *partials;
};

// pass 1: max over scores.  NaN scores are ignored by fmaxf (np.max would propagate NaN).
__global__ void __launch_bounds__(RED_THREADS) k_max(const float *__restrict__ s, int64_t n, RedScratch rs,
                                                     double *stats) {
    __shared__ float shf[32];
    __shared__ bool last;
    float m = -FLT_MAX;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        m = fmaxf(m, s[i]);
    m = block_max(m, shf);
    if (threadIdx.x == 0) {
        rs.partials[blockIdx.x] = (double)m;
        __threadfence();
        last = atomicAdd(rs.counter, 1u) == gridDim.x - 1;
    }
    __syncthreads();
    if (last) {
        __threadfence();
        float t = -FLT_MAX;
        for (int b = threadIdx.x; b < gridDim.x; b += blockDim.x) t = fmaxf(t, (float)((volatile double *)rs.partials)[b]);
        t = block_max(t, shf);
        if (threadIdx.x == 0) { stats[0] = (double)t; *rs.counter = 0; }
    }
}

// exp(s - max) exactly as the weights are formed: f32 subtract (NumPy f32 array minus f32 scalar),
// exponential evaluated in f64 and rounded to f32.
__device__ __forceinline__ float softmax_num(float s, float m) { return (float)exp((double)__fsub_rn(s, m)); }

// pass 2: sum of the f32 numerators in 2^-40 fixed point (uint64): exact integer arithmetic, so the
// sum -- and therefore every weight -- is identical for any block or rank decomposition.  (Numerators
// below 2^-17 lose the bits under 2^-40: < 1e-6 absolute on a sum that is >= 1.)
// stats[1] = sum as f64, stats[2] = the raw integer (bit pattern) for the cross-rank all-reduce.
#define SOFTMAX_FIX 1099511627776.0   // 2^40
__device__ __forceinline__ unsigned long long block_sum_u64(unsigned long long v, unsigned long long *sh) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
    __syncthreads();
    unsigned long long t = 0;
    if (threadIdx.x == 0)
        for (int k = 0; k < (int)(blockDim.x >> 5); ++k) t += sh[k];
    __syncthreads();
    return t;   // valid in thread 0
}
__global__ void __launch_bounds__(RED_THREADS) k_sumexp(const float *__restrict__ s, int64_t n, RedScratch rs,
                                                        double *stats, double ext_max, int use_ext) {
    __shared__ unsigned long long shq[32];
    __shared__ bool last;
    const float m = (float)(use_ext ? ext_max : stats[0]);
    unsigned long long acc = 0;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        acc += __double2ull_rz(__dmul_rn((double)softmax_num(s[i], m), SOFTMAX_FIX));
    acc = block_sum_u64(acc, shq);
    if (threadIdx.x == 0) {
        ((unsigned long long *)rs.partials)[blockIdx.x] = acc;
        __threadfence();
        last = atomicAdd(rs.counter, 1u) == gridDim.x - 1;
    }
    __syncthreads();
    if (last) {
        __threadfence();
        unsigned long long t = 0;
        for (int b = threadIdx.x; b < gridDim.x; b += blockDim.x) t += ((volatile unsigned long long *)rs.partials)[b];
        t = block_sum_u64(t, shq);
        if (threadIdx.x == 0) {
            stats[1] = (double)t / SOFTMAX_FIX;
            ((unsigned long long *)stats)[2] = t;
            *rs.counter = 0;
        }
    }
}

// weights = exp(s - max) / sum, f32 divide by the f32-rounded sum (node:356-357)
__global__ void k_softmax_weights(const float *__restrict__ s, int64_t n, const double *stats, double ext_max,
                                  double ext_sum, int use_ext, float *__restrict__ w) {
    const float m = (float)(use_ext ? ext_max : stats[0]);
    const float sum = (float)(use_ext ? ext_sum : stats[1]);
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        w[i] = __fdiv_rn(softmax_num(s[i], m), sum);
}

static int red_blocks(const mcl_handle *h, int64_t n) {
    return (int)std::max<int64_t>(1, std::min<int64_t>((n + RED_THREADS * 4 - 1) / (RED_THREADS * 4), (int64_t)h->sm_count * 8));
}

static int softmax_impl(mcl_handle *h, const float *d_score, int64_t n, float *d_weights, double *d_stats,
                        const double *ext) {
    const int nb = red_blocks(h, n);
    int rc = mcl_ensure_scratch(h, 128 + sizeof(double) * (size_t)nb + 64);
    if (rc) return rc;
    RedScratch rs;
    rs.counter = (unsigned *)h->d_scratch;
    rs.partials = (double *)((char *)h->d_scratch + 64);
    double *stats = d_stats ? d_stats : (double *)((char *)h->d_scratch + 64 + sizeof(double) * (size_t)nb);
    static_assert(sizeof(double) == 8, "");
    if (!ext) {
        MCL_CUDA(h, cudaMemsetAsync(rs.counter, 0, sizeof(unsigned), h->stream));
        k_max<<<nb, RED_THREADS, 0, h->stream>>>(d_score, n, rs, stats);
        MCL_LAUNCH_CHECK(h);
        k_sumexp<<<nb, RED_THREADS, 0, h->stream>>>(d_score, n, rs, stats, 0.0, 0);
        MCL_LAUNCH_CHECK(h);
    } else if (d_stats) {
        MCL_CUDA(h, cudaMemcpyAsync(d_stats, ext, 2 * sizeof(double), cudaMemcpyHostToDevice, h->stream));
    }
    if (d_weights) {
        k_softmax_weights<<<red_blocks(h, n), RED_THREADS, 0, h->stream>>>(d_score, n, stats, ext ? ext[0] : 0.0,
                                                                          ext ? ext[1] : 0.0, ext ? 1 : 0, d_weights);
        MCL_LAUNCH_CHECK(h);
    }
    return MCL_OK;
}

extern "C" int mcl_softmax(mcl_handle *h, const float *d_score, int64_t n, float *d_weights, double *d_stats,
                           const double *ext_stats) {
    if (!h) return MCL_ERR_ARG;
    if (n <= 0 || !d_score) return mcl_fail(h, MCL_ERR_ARG, "mcl_softmax: bad argument");
    DeviceGuard guard(h->device);
    return softmax_impl(h, d_score, n, d_weights, d_stats, ext_stats);
}

// staged forms for the sharded (multi-GPU) path: the caller all-reduces d_stats between stages
extern "C" int mcl_softmax_max(mcl_handle *h, const float *d_score, int64_t n, double *d_stats) {
    if (!h) return MCL_ERR_ARG;
    if (n <= 0 || !d_score || !d_stats) return mcl_fail(h, MCL_ERR_ARG, "mcl_softmax_max: bad argument");
    DeviceGuard guard(h->device);
    const int nb = red_blocks(h, n);
    int rc = mcl_ensure_scratch(h, 128 + sizeof(double) * (size_t)nb + 64);
    if (rc) return rc;
    RedScratch rs;
    rs.counter = (unsigned *)h->d_scratch;
    rs.partials = (double *)((char *)h->d_scratch + 64);
    MCL_CUDA(h, cudaMemsetAsync(rs.counter, 0, sizeof(unsigned), h->stream));
    k_max<<<nb, RED_THREADS, 0, h->stream>>>(d_score, n, rs, d_stats);
    MCL_LAUNCH_CHECK(h);
    return MCL_OK;
}
extern "C" int mcl_softmax_sumexp(mcl_handle *h, const float *d_score, int64_t n, double *d_stats) {
    if (!h) return MCL_ERR_ARG;
    if (n <= 0 || !d_score || !d_stats) return mcl_fail(h, MCL_ERR_ARG, "mcl_softmax_sumexp: bad argument");
    DeviceGuard guard(h->device);
    const int nb = red_blocks(h, n);
    int rc = mcl_ensure_scratch(h, 128 + sizeof(double) * (size_t)nb + 64);
    if (rc) return rc;
    RedScratch rs;
    rs.counter = (unsigned *)h->d_scratch;
    rs.partials = (double *)((char *)h->d_scratch + 64);
    MCL_CUDA(h, cudaMemsetAsync(rs.counter, 0, sizeof(unsigned), h->stream));
    k_sumexp<<<nb, RED_THREADS, 0, h->stream>>>(d_score, n, rs, d_stats, 0.0, 0);
    MCL_LAUNCH_CHECK(h);
    return MCL_OK;
}
extern "C" int mcl_softmax_weights(mcl_handle *h, const float *d_score, int64_t n, const double *d_stats,
                                   float *d_weights) {
    if (!h) return MCL_ERR_ARG;
    if (n <= 0 || !d_score || !d_stats || !d_weights) return mcl_fail(h, MCL_ERR_ARG, "mcl_softmax_weights: bad argument");
    DeviceGuard guard(h->device);
    k_softmax_weights<<<red_blocks(h, n), RED_THREADS, 0, h->stream>>>(d_score, n, d_stats, 0.0, 0.0, 0, d_weights);
    MCL_LAUNCH_CHECK(h);
    return MCL_OK;
}

extern "C" int mcl_softmax_stats(mcl_handle *h, const float *d_score, int64_t n, double h_stats[2]) {
    if (!h) return MCL_ERR_ARG;
    if (n <= 0 || !d_score || !h_stats) return mcl_fail(h, MCL_ERR_ARG, "mcl_softmax_stats: bad argument");
    DeviceGuard guard(h->device);
    int rc = softmax_impl(h, d_score, n, nullptr, nullptr, nullptr);
    if (rc) return rc;
    const int nb = red_blocks(h, n);
    double *stats = (double *)((char *)h->d_scratch + 64 + sizeof(double) * (size_t)nb);
    MCL_CUDA(h, cudaMemcpyAsync(h->h_pinned, stats, 2 * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    MCL_CUDA(h, cudaStreamSynchronize(h->stream));
    h_stats[0] = h->h_pinned[0];
    h_stats[1] = h->h_pinned[1];
    return MCL_OK;
}

// ------------------------------------------------------------------------------------------
// pu:208-236 mh_resampling -- one chain per particle.
//   alpha = min(1.0, p_new / p_old) if p_old > 0 else 1.0   (f32 IEEE divide, promoted to f64)
//   accept iff u < alpha (u f64 in [0,1))
// ------------------------------------------------------------------------------------------
__global__ void k_mh_accept(const double *x, const double *y,
                            const double *th, const double *px,
                            const double *py, const double *pth,
                            const float *__restrict__ lik, const float *__restrict__ oldw, int64_t n,
                            const double *__restrict__ uniforms, uint64_t seed, uint64_t step,
                            uint64_t first_index, double *xo, double *yo, double *tho, float *wo,
                            uint8_t *accept) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const float p_old = oldw[i], p_new = lik[i];
        double alpha = 1.0;
        if (p_old > 0.f) {
            const double q = (double)__fdiv_rn(p_new, p_old);
            alpha = (q < 1.0) ? q : 1.0;     // min(1.0, q); NaN -> 1.0
        }
        double u;
        if (uniforms) u = uniforms[i];
        else {
            const uint4 o = philox_draw4(seed, step, first_index + (uint64_t)i, 0u, MCL_STREAM_MH);
            u = u53_from(o.x, o.y);
        }
        const bool acc = u < alpha;
        const double nx = acc ? px[i] : x[i], ny = acc ? py[i] : y[i], nt = acc ? pth[i] : th[i];
        xo[i] = nx; yo[i] = ny; tho[i] = nt;
        wo[i] = acc ? p_new : p_old;
        if (accept) accept[i] = acc ? 1 : 0;
    }
}

extern "C" int mcl_mh_accept(mcl_handle *h, const double *d_x, const double *d_y, const double *d_theta,
                             const double *d_px, const double *d_py, const double *d_ptheta,
                             const float *d_lik, const float *d_oldw, int64_t n, const double *d_uniforms,
                             uint64_t seed, uint64_t step, uint64_t first_index, double *d_xo, double *d_yo,
                             double *d_thetao, float *d_wo, uint8_t *d_accept) {
    if (!h) return MCL_ERR_ARG;
    if (n < 0 || (n > 0 && (!d_x || !d_y || !d_theta || !d_px || !d_py || !d_ptheta || !d_lik || !d_oldw ||
                            !d_xo || !d_yo || !d_thetao || !d_wo)))
        return mcl_fail(h, MCL_ERR_ARG, "mcl_mh_accept: bad argument");
    if (n == 0) return MCL_OK;
    DeviceGuard guard(h->device);
    const int blocks = (int)std::min<int64_t>((n + 255) / 256, (int64_t)h->sm_count * 16);
    k_mh_accept<<<blocks, 256, 0, h->stream>>>(d_x, d_y, d_theta, d_px, d_py, d_ptheta, d_lik, d_oldw, n,
                                               d_uniforms, seed, step, first_index, d_xo, d_yo, d_thetao,
                                               d_wo, d_accept);
    MCL_LAUNCH_CHECK(h);
    return MCL_OK;
}

// node:276-278 update_acml_weights: weights = weights / np.sum(weights).  The sum is taken in 2^-40 fixed
// point (exact, decomposition-independent) and rounded to f32 like NumPy's f32 sum; h_out = {sum before,
// mean of the normalised weights (node:284 w_avg)}.  Blocking.
__global__ void __launch_bounds__(RED_THREADS) k_wsum_fixed(const float *__restrict__ w, int64_t n, RedScratch rs,
                                                            double *out) {
    __shared__ unsigned long long shq[32];
    __shared__ bool last;
    unsigned long long acc = 0;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const double v = __dmul_rn((double)w[i], SOFTMAX_FIX);
        acc += v > 0.0 ? __double2ull_rz(v) : 0ull;
    }
    acc = block_sum_u64(acc, shq);
    if (threadIdx.x == 0) {
        ((unsigned long long *)rs.partials)[blockIdx.x] = acc;
        __threadfence();
        last = atomicAdd(rs.counter, 1u) == gridDim.x - 1;
    }
    __syncthreads();
    if (last) {
        __threadfence();
        unsigned long long t = 0;
        for (int b = threadIdx.x; b < gridDim.x; b += blockDim.x) t += ((volatile unsigned long long *)rs.partials)[b];
        t = block_sum_u64(t, shq);
        if (threadIdx.x == 0) { out[0] = (double)t / SOFTMAX_FIX; *rs.counter = 0; }
    }
}
__global__ void k_wdiv(float *w, int64_t n, const double *sum) {
    const float s = (float)sum[0];
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        w[i] = __fdiv_rn(w[i], s);
}

extern "C" int mcl_weights_normalize(mcl_handle *h, float *d_w, int64_t n, double h_out[2]) {
    if (!h) return MCL_ERR_ARG;
    if (n <= 0 || !d_w) return mcl_fail(h, MCL_ERR_ARG, "mcl_weights_normalize: bad argument");
    DeviceGuard guard(h->device);
    const int nb = red_blocks(h, n);
    int rc = mcl_ensure_scratch(h, 128 + sizeof(double) * (size_t)nb + 64);
    if (rc) return rc;
    RedScratch rs;
    rs.counter = (unsigned *)h->d_scratch;
    rs.partials = (double *)((char *)h->d_scratch + 64);
    double *out = (double *)((char *)h->d_scratch + 64 + sizeof(double) * (size_t)nb);
    MCL_CUDA(h, cudaMemsetAsync(rs.counter, 0, sizeof(unsigned), h->stream));
    k_wsum_fixed<<<nb, RED_THREADS, 0, h->stream>>>(d_w, n, rs, out);
    MCL_LAUNCH_CHECK(h);
    k_wdiv<<<nb, RED_THREADS, 0, h->stream>>>(d_w, n, out);
    MCL_LAUNCH_CHECK(h);
    k_wsum_fixed<<<nb, RED_THREADS, 0, h->stream>>>(d_w, n, rs, out + 1);
    MCL_LAUNCH_CHECK(h);
    if (h_out) {
        MCL_CUDA(h, cudaMemcpyAsync(h->h_pinned, out, 2 * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
        MCL_CUDA(h, cudaStreamSynchronize(h->stream));
        h_out[0] = h->h_pinned[0];
        h_out[1] = h->h_pinned[1] / (double)n;
    }
    return MCL_OK;
}

// ---------------------------------------------------------------------------------------------
// Both softmaxes of an MH update (scores of particles and of particles_prev, node:254-270) in three launches
// instead of six: blockIdx.y selects the score set.  Same arithmetic as k_max / k_sumexp / k_softmax_weights.
// ---------------------------------------------------------------------------------------------
struct Softmax2 {
    const float *s[2];
    float *w[2];
    double *stats[2];
    unsigned *counter[2];
    double *partials[2];
};

__global__ void __launch_bounds__(RED_THREADS) k_max2(const Softmax2 a, int64_t n) {
    __shared__ float shf[32];
    __shared__ bool last;
    const int y = blockIdx.y;
    const float *s = a.s[y];
    float m = -FLT_MAX;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        m = fmaxf(m, s[i]);
    m = block_max(m, shf);
    if (threadIdx.x == 0) {
        a.partials[y][blockIdx.x] = (double)m;
        __threadfence();
        last = atomicAdd(a.counter[y], 1u) == gridDim.x - 1;
    }
    __syncthreads();
    if (last) {
        __threadfence();
        float t = -FLT_MAX;
        for (int b = threadIdx.x; b < gridDim.x; b += blockDim.x) t = fmaxf(t, (float)((volatile double *)a.partials[y])[b]);
        t = block_max(t, shf);
        if (threadIdx.x == 0) { a.stats[y][0] = (double)t; *a.counter[y] = 0; }
    }
}

__global__ void __launch_bounds__(RED_THREADS) k_sumexp2(const Softmax2 a, int64_t n) {
    __shared__ unsigned long long shq[32];
    __shared__ bool last;
    const int y = blockIdx.y;
    const float *s = a.s[y];
    const float m = (float)a.stats[y][0];
    unsigned long long acc = 0;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        acc += __double2ull_rz(__dmul_rn((double)softmax_num(s[i], m), SOFTMAX_FIX));
    acc = block_sum_u64(acc, shq);
    if (threadIdx.x == 0) {
        ((unsigned long long *)a.partials[y])[blockIdx.x] = acc;
        __threadfence();
        last = atomicAdd(a.counter[y], 1u) == gridDim.x - 1;
    }
    __syncthreads();
    if (last) {
        __threadfence();
        unsigned long long t = 0;
        for (int b = threadIdx.x; b < gridDim.x; b += blockDim.x) t += ((volatile unsigned long long *)a.partials[y])[b];
        t = block_sum_u64(t, shq);
        if (threadIdx.x == 0) {
            a.stats[y][1] = (double)t / SOFTMAX_FIX;
            ((unsigned long long *)a.stats[y])[2] = t;
            *a.counter[y] = 0;
        }
    }
}

__global__ void k_softmax_weights2(const Softmax2 a, int64_t n) {
    const int y = blockIdx.y;
    const float *s = a.s[y];
    float *w = a.w[y];
    const float m = (float)a.stats[y][0], sum = (float)a.stats[y][1];
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        w[i] = __fdiv_rn(softmax_num(s[i], m), sum);
}

// internal (filter.cu): softmax of two score sets of equal length
int mcl_softmax_pair(mcl_handle *h, const float *s0, float *w0, const float *s1, float *w1, int64_t n) {
    const int nb = red_blocks(h, n);
    // scratch: [0,64) counters | 2 x partials | 2 x 4 stats doubles
    int rc = mcl_ensure_scratch(h, 128 + 2 * sizeof(double) * (size_t)nb + 128);
    if (rc) return rc;
    char *sc = (char *)h->d_scratch;
    Softmax2 a;
    a.s[0] = s0; a.s[1] = s1; a.w[0] = w0; a.w[1] = w1;
    a.counter[0] = (unsigned *)sc; a.counter[1] = (unsigned *)(sc + 16);
    a.partials[0] = (double *)(sc + 64); a.partials[1] = a.partials[0] + nb;
    a.stats[0] = a.partials[1] + nb; a.stats[1] = a.stats[0] + 4;
    MCL_CUDA(h, cudaMemsetAsync(sc, 0, 32, h->stream));
    const dim3 grid(nb, 2);
    k_max2<<<grid, RED_THREADS, 0, h->stream>>>(a, n);
    MCL_LAUNCH_CHECK(h);
    k_sumexp2<<<grid, RED_THREADS, 0, h->stream>>>(a, n);
    MCL_LAUNCH_CHECK(h);
    k_softmax_weights2<<<grid, RED_THREADS, 0, h->stream>>>(a, n);
    MCL_LAUNCH_CHECK(h);
    return MCL_OK;
}
