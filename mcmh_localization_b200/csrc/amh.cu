// amh.cu -- asymmetric Metropolis-Hastings:
//   pu:282-330 motion_model_odometry_parallel  (transition density p(x_t | x_{t-1}, u), normalised by the
//                                               population sum)  + pu:31-33 gaussian_prob
//   pu:238-276 assym_mh_resampling
// The reference's quirks are reproduced: alpha = min(1, exp(log_alpha)) only if
// log_den > 0, which never holds for probabilities < 1, so every proposal is accepted; the densities are
// normalised by their sum over the population.
#include <algorithm>

#include "common.cuh"

__device__ __forceinline__ double gaussian_prob_dev(double diff, double sigma) {   // pu:31-33
    const double q = __ddiv_rn(diff, sigma);
    return __ddiv_rn(exp(__dmul_rn(-0.5, __dmul_rn(q, q))), sqrt(__dmul_rn(MCL_TWO_PI, __dmul_rn(sigma, sigma))));
}

struct DensityParams {
    const double *px, *py, *pt, *cx, *cy, *ct;
    int64_t n;
    double rot1, trans, rot2, s_rot1, s_trans, s_rot2;
    double *probs;
};

__global__ void __launch_bounds__(256) k_motion_density(const DensityParams p, double *partials, unsigned *counter,
                                                        double *sum_out) {
    __shared__ double sh[32];
    __shared__ bool last;
    double acc = 0.0;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < p.n; i += (int64_t)gridDim.x * blockDim.x) {
        const double dx = __dadd_rn(p.cx[i], -p.px[i]), dy = __dadd_rn(p.cy[i], -p.py[i]);       // pu:306-307
        const double th_prev = p.pt[i], th_curr = p.ct[i];
        const double trans_hat = sqrt(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)));         // pu:309
        const double rot1_hat = normalize_angle_dev(__dadd_rn(atan2(dy, dx), -th_prev));        // pu:310
        const double rot2_hat = normalize_angle_dev(__dadd_rn(__dadd_rn(th_curr, -th_prev), -rot1_hat));   // pu:311
        const double p1 = gaussian_prob_dev(normalize_angle_dev(__dadd_rn(p.rot1, -rot1_hat)), p.s_rot1);
        const double p2 = gaussian_prob_dev(__dadd_rn(p.trans, -trans_hat), p.s_trans);
        const double p3 = gaussian_prob_dev(normalize_angle_dev(__dadd_rn(p.rot2, -rot2_hat)), p.s_rot2);
        const double v = __dmul_rn(__dmul_rn(p1, p2), p3);                                      // pu:323
        p.probs[i] = v;
        acc += v;
    }
    acc = warp_sum(acc);
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0;
        for (int k = 0; k < (int)(blockDim.x >> 5); ++k) t += sh[k];
        partials[blockIdx.x] = t;
        __threadfence();
        last = atomicAdd(counter, 1u) == gridDim.x - 1;
    }
    __syncthreads();
    if (last) {
        __threadfence();
        double t = 0;
        for (unsigned b = threadIdx.x; b < gridDim.x; b += blockDim.x) t += ((volatile double *)partials)[b];
        t = warp_sum(t);
        __syncthreads();
        if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = t;
        __syncthreads();
        if (threadIdx.x == 0) {
            double s = 0;
            for (int k = 0; k < (int)(blockDim.x >> 5); ++k) s += sh[k];
            sum_out[0] = s;
            *counter = 0;
        }
    }
}

// pu:326-328: probs /= s if s > 0
__global__ void k_scale_by_sum(double *probs, int64_t n, const double *sum) {
    const double s = sum[0];
    if (!(s > 0)) return;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        probs[i] = __ddiv_rn(probs[i], s);
}

extern "C" int mcl_motion_density(mcl_handle *h, const double *d_px, const double *d_py, const double *d_pt,
                                  const double *d_cx, const double *d_cy, const double *d_ct, int64_t n,
                                  const double delta[3], double *d_probs, double *d_sum, int normalise) {
    if (!h) return MCL_ERR_ARG;
    if (n <= 0 || !d_px || !d_py || !d_pt || !d_cx || !d_cy || !d_ct || !delta || !d_probs)
        return mcl_fail(h, MCL_ERR_ARG, "mcl_motion_density: bad argument");
    if (!h->motion_set) return mcl_fail(h, MCL_ERR_STATE, "mcl_motion_density: motion noise not set");
    DeviceGuard guard(h->device);
    DensityParams p;
    p.px = d_px; p.py = d_py; p.pt = d_pt; p.cx = d_cx; p.cy = d_cy; p.ct = d_ct; p.n = n; p.probs = d_probs;
    p.rot1 = delta[0]; p.trans = delta[1]; p.rot2 = delta[2];
    const double a1 = (double)h->alpha[0], a2 = (double)h->alpha[1], a3 = (double)h->alpha[2], a4 = (double)h->alpha[3];
    volatile double t1, t2;
    t1 = a1 * fabs(p.rot1); t2 = a2 * fabs(p.trans); p.s_rot1 = t1 + t2;                       // pu:314
    t1 = a3 * fabs(p.trans); t2 = a4 * (fabs(p.rot1) + fabs(p.rot2)); p.s_trans = t1 + t2;    // pu:315
    t1 = a1 * fabs(p.rot2); t2 = a2 * fabs(p.trans); p.s_rot2 = t1 + t2;                       // pu:316
    const int nb = (int)std::max<int64_t>(1, std::min<int64_t>((n + 511) / 512, (int64_t)h->sm_count * 4));
    int rc = mcl_ensure_scratch(h, 128 + (size_t)nb * 8);
    if (rc) return rc;
    unsigned *counter = (unsigned *)h->d_scratch;
    double *sum = d_sum ? d_sum : (double *)((char *)h->d_scratch + 64);
    double *partials = (double *)((char *)h->d_scratch + 128);
    MCL_CUDA(h, cudaMemsetAsync(counter, 0, sizeof(unsigned), h->stream));
    k_motion_density<<<nb, 256, 0, h->stream>>>(p, partials, counter, sum);
    MCL_LAUNCH_CHECK(h);
    if (normalise) {
        k_scale_by_sum<<<(int)std::min<int64_t>((n + 255) / 256, (int64_t)h->sm_count * 16), 256, 0, h->stream>>>(d_probs, n, sum);
        MCL_LAUNCH_CHECK(h);
    }
    return MCL_OK;
}

extern "C" int mcl_scale_by_sum(mcl_handle *h, double *d_probs, int64_t n, const double *d_sum) {
    if (!h) return MCL_ERR_ARG;
    if (n <= 0 || !d_probs || !d_sum) return mcl_fail(h, MCL_ERR_ARG, "mcl_scale_by_sum: bad argument");
    DeviceGuard guard(h->device);
    k_scale_by_sum<<<(int)std::min<int64_t>((n + 255) / 256, (int64_t)h->sm_count * 16), 256, 0, h->stream>>>(d_probs, n, d_sum);
    MCL_LAUNCH_CHECK(h);
    return MCL_OK;
}

// pu:238-276 assym_mh_resampling
__global__ void k_assym_mh(const double *x, const double *y, const double *th, const double *px, const double *py,
                           const double *pth, const float *__restrict__ lik, const float *__restrict__ oldw,
                           const double *__restrict__ tf, const double *__restrict__ tb, int64_t n,
                           const double *__restrict__ uniforms, uint64_t seed, uint64_t step, uint64_t first_index,
                           double *xo, double *yo, double *tho, float *wo, uint8_t *accept, bool corrected) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const double log_pre = log(__dadd_rn((double)oldw[i], 1e-10));       // pu:259
        const double log_post = log(__dadd_rn((double)lik[i], 1e-10));       // pu:260
        const double log_tf = log(__dadd_rn(tf[i], 1e-10));                  // pu:261
        const double log_tb = log(__dadd_rn(tb[i], 1e-10));                  // pu:262
        const double log_num = __dadd_rn(log_post, log_tb);
        const double log_den = __dadd_rn(log_pre, log_tf);
        double alpha = 1.0;
        // pu:269 tests log_den > 0, which never holds for probabilities < 1: the reference accepts everything
        // (SURVEY Appendix C #1).  corrected: the Metropolis-Hastings ratio is applied unconditionally.
        if (corrected || log_den > 0) {
            const double e = exp(__dadd_rn(log_num, -log_den));
            alpha = e < 1.0 ? e : 1.0;
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
        wo[i] = acc ? lik[i] : oldw[i];
        if (accept) accept[i] = acc ? 1 : 0;
    }
}

extern "C" int mcl_assym_mh_accept(mcl_handle *h, const double *d_x, const double *d_y, const double *d_theta,
                                   const double *d_px, const double *d_py, const double *d_ptheta,
                                   const float *d_lik, const float *d_oldw, const double *d_tf, const double *d_tb,
                                   int64_t n, const double *d_uniforms, uint64_t seed, uint64_t step,
                                   uint64_t first_index, double *d_xo, double *d_yo, double *d_thetao, float *d_wo,
                                   uint8_t *d_accept) {
    return mcl_assym_mh_accept_ex(h, d_x, d_y, d_theta, d_px, d_py, d_ptheta, d_lik, d_oldw, d_tf, d_tb, n, d_uniforms, seed,
                                  step, first_index, d_xo, d_yo, d_thetao, d_wo, d_accept, 0);
}

extern "C" int mcl_assym_mh_accept_ex(mcl_handle *h, const double *d_x, const double *d_y, const double *d_theta,
                                      const double *d_px, const double *d_py, const double *d_ptheta,
                                      const float *d_lik, const float *d_oldw, const double *d_tf, const double *d_tb,
                                      int64_t n, const double *d_uniforms, uint64_t seed, uint64_t step,
                                      uint64_t first_index, double *d_xo, double *d_yo, double *d_thetao, float *d_wo,
                                      uint8_t *d_accept, int corrected) {
    if (!h) return MCL_ERR_ARG;
    if (n <= 0 || !d_x || !d_y || !d_theta || !d_px || !d_py || !d_ptheta || !d_lik || !d_oldw || !d_tf || !d_tb ||
        !d_xo || !d_yo || !d_thetao || !d_wo)
        return mcl_fail(h, MCL_ERR_ARG, "mcl_assym_mh_accept: bad argument");
    DeviceGuard guard(h->device);
    const int blocks = (int)std::min<int64_t>((n + 255) / 256, (int64_t)h->sm_count * 16);
    k_assym_mh<<<blocks, 256, 0, h->stream>>>(d_x, d_y, d_theta, d_px, d_py, d_ptheta, d_lik, d_oldw, d_tf, d_tb, n,
                                              d_uniforms, seed, step, first_index, d_xo, d_yo, d_thetao, d_wo, d_accept,
                                              corrected != 0);
    MCL_LAUNCH_CHECK(h);
    return MCL_OK;
}
