// kld.cu -- KLD-adaptive systematic resampling, pu:529-591 kld_sampling_amcl.
//
// The reference draws samples one at a time: systematic pick of a source particle, Gaussian jitter
// (sigma 1 mm, 1 mm, 0.02 rad), insertion of the jittered pose's (x, y, theta) bin into a Python set, and
// when a NEW bin appears it evaluates the Wilson-Hilferty bound and may stop.  The loop looks sequential
// but every quantity is a function of the sample index `count` alone:
//   sample(count)  = jitter(source(U_count), normals(count))
//   is_new(count)  = no earlier sample fell into the same bin      -> hash set keyed by bin, value = the
//                    smallest count that hit it (atomicMin), built by all samples in parallel
//   k(count)       = number of new bins among samples 0..count     -> prefix count
//   stop(count)    = is_new && k > 1 && count >= min_particles && count > chi2(k) / (2 eps)
// and the result is the first `count` with stop(count) (that sample is NOT stored, pu:586-590).
#include <math.h>

#include <algorithm>

#include "common.cuh"

#define KLD_EMPTY 0xffffffffffffffffull

struct KldParams {
    const double *x, *y, *th;
    const float *cf;              // sequential f32 cumulative sums (reference arithmetic) or NULL
    const uint64_t *cq;           // fixed-point cumulative sums (production) or NULL
    const uint64_t *total_q;
    int64_t n_in, max_samples;
    double r, bin_xy, bin_theta;
    const double *normals;        // nullable (max_samples, 3)
    uint64_t seed, step;
    double *xo, *yo, *tho;
    uint64_t *keys;               // bin key per sample
    unsigned long long *tab_keys; // hash table
    unsigned *tab_min;
    uint64_t tab_mask;
    int *overflow;
};

__device__ __forceinline__ uint64_t kld_hash(uint64_t k) {
    k ^= k >> 33; k *= 0xff51afd7ed558ccdull; k ^= k >> 33; k *= 0xc4ceb9fe1a85ec53ull; k ^= k >> 33;
    return k;
}

__global__ void k_kld_sample(const KldParams p) {
    const double noise_std[3] = {0.001, 0.001, 0.02};                                   // pu:552
    for (int64_t c = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; c < p.max_samples;
         c += (int64_t)gridDim.x * blockDim.x) {
        const double U = __dadd_rn(p.r, __ddiv_rn((double)c, (double)p.max_samples));   // pu:560
        int64_t lo = 0, hi = p.n_in - 1;
        if (p.cf) {
            while (lo < hi) { const int64_t mid = (lo + hi) >> 1; if (U > (double)p.cf[mid]) lo = mid + 1; else hi = mid; }
        } else {
            const double t = ceil(__dmul_rn(U, (double)p.total_q[0]));
            const uint64_t T = t >= 18446744073709551616.0 ? 0xffffffffffffffffull : (t > 0.0 ? __double2ull_rz(t) : 0ull);
            while (lo < hi) { const int64_t mid = (lo + hi) >> 1; if (T > p.cq[mid]) lo = mid + 1; else hi = mid; }
        }
        double z0, z1, z2;
        if (p.normals) { z0 = p.normals[3 * c]; z1 = p.normals[3 * c + 1]; z2 = p.normals[3 * c + 2]; }
        else {
            const uint4 o = philox_draw4(p.seed, p.step, (uint64_t)c, 0u, MCL_STREAM_KLD);
            const uint4 a = make_uint4(o.y, o.z, o.w, 0u);
            normals3_from_words(o.x, a, z0, z1, z2);     // u1 = o.x, u2 = o.y, u3 = o.z, u4 = o.w
        }
        const double nx = __dadd_rn(p.x[lo], __dadd_rn(0.0, __dmul_rn(noise_std[0], z0)));   // pu:570
        const double ny = __dadd_rn(p.y[lo], __dadd_rn(0.0, __dmul_rn(noise_std[1], z1)));
        const double nt = __dadd_rn(p.th[lo], __dadd_rn(0.0, __dmul_rn(noise_std[2], z2)));
        p.xo[c] = (double)(float)nx; p.yo[c] = (double)(float)ny; p.tho[c] = (double)(float)nt;  // pu:550: f32 buffer
        const long long xb = __double2ll_rz(__ddiv_rn(nx, p.bin_xy));                     // pu:573-575 int()
        const long long yb = __double2ll_rz(__ddiv_rn(ny, p.bin_xy));
        const long long tb = __double2ll_rz(__ddiv_rn(nt, p.bin_theta));
        if (xb < -(1ll << 20) || xb >= (1ll << 20) || yb < -(1ll << 20) || yb >= (1ll << 20) || tb < -(1ll << 21) ||
            tb >= (1ll << 21))
            *p.overflow = 1;
        const uint64_t key = ((uint64_t)(xb + (1ll << 20)) << 43) | ((uint64_t)(yb + (1ll << 20)) << 22) |
                             (uint64_t)(tb + (1ll << 21));
        p.keys[c] = key;
        uint64_t slot = kld_hash(key) & p.tab_mask;
        while (true) {
            const unsigned long long prev = atomicCAS(&p.tab_keys[slot], KLD_EMPTY, (unsigned long long)key);
            if (prev == KLD_EMPTY || prev == key) { atomicMin(&p.tab_min[slot], (unsigned)c); break; }
            slot = (slot + 1) & p.tab_mask;
        }
    }
}

// first count with stop(count); single CTA walking the samples in chunks with a running bin count
__global__ void __launch_bounds__(1024) k_kld_stop(const uint64_t *keys, const unsigned long long *tab_keys,
                                                   const unsigned *tab_min, uint64_t tab_mask, int64_t max_samples,
                                                   int64_t min_particles, double epsilon, double z, long long *count_out) {
    __shared__ int wsum[32];
    __shared__ long long carry;
    __shared__ long long found;
    if (threadIdx.x == 0) { carry = 0; found = -1; }
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int64_t base = 0; base < max_samples; base += blockDim.x) {
        const int64_t c = base + threadIdx.x;
        int is_new = 0;
        if (c < max_samples) {
            const uint64_t key = keys[c];
            uint64_t slot = kld_hash(key) & tab_mask;
            while (tab_keys[slot] != key) slot = (slot + 1) & tab_mask;
            is_new = tab_min[slot] == (unsigned)c;
        }
        int inc = is_new;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += t; }
        if (lane == 31) wsum[warp] = inc;
        __syncthreads();
        int off = 0;
        for (int q = 0; q < warp; ++q) off += wsum[q];
        const long long k = carry + off + inc;                              // len(bins) after this sample
        bool stop = false;
        if (is_new && k > 1 && c >= min_particles) {                        // pu:582-586
            const double km1 = (double)(k - 1);
            const double a = __dadd_rn(__dadd_rn(1.0, -__ddiv_rn(2.0, __dmul_rn(9.0, km1))),
                                       __dmul_rn(sqrt(__ddiv_rn(2.0, __dmul_rn(9.0, km1))), z));
            const double chi2 = __dmul_rn(km1, __dmul_rn(__dmul_rn(a, a), a));
            stop = (double)c > __ddiv_rn(chi2, __dmul_rn(2.0, epsilon));
        }
        const unsigned m = __ballot_sync(0xffffffffu, stop);
        if (m && lane == __ffs(m) - 1) atomicMin((unsigned long long *)&found, (unsigned long long)c);   // found = -1 = max
        __syncthreads();
        if (threadIdx.x == blockDim.x - 1) carry = k;
        __syncthreads();
        if (found != -1) break;
    }
    if (threadIdx.x == 0) *count_out = found != -1 ? found : (long long)max_samples;
}

__global__ void k_fill_u64(unsigned long long *p, int64_t n, unsigned long long v) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) p[i] = v;
}
__global__ void k_fill_u32(unsigned *p, int64_t n, unsigned v) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) p[i] = v;
}

extern "C" int mcl_kld_resample(mcl_handle *h, const double *d_x, const double *d_y, const double *d_theta,
                                const float *d_weights, int64_t n_in, int64_t max_samples, int64_t min_particles,
                                double bin_xy, double bin_theta, double epsilon, double z, double r,
                                const double *d_normals, uint64_t seed, uint64_t step, int mode, double *d_xo,
                                double *d_yo, double *d_thetao, int64_t *h_count) {
    if (!h) return MCL_ERR_ARG;
    if (h_count) *h_count = 0;
    if (n_in <= 0 || max_samples < 0 || !d_x || !d_y || !d_theta || !d_weights || !h_count ||
        (max_samples > 0 && (!d_xo || !d_yo || !d_thetao)) || !(bin_xy > 0) || !(bin_theta > 0) || !(epsilon > 0))
        return mcl_fail(h, MCL_ERR_ARG, "mcl_kld_resample: bad argument");
    if (max_samples == 0) return MCL_OK;
    if (max_samples > 0xfffffff0ll) return mcl_fail(h, MCL_ERR_CAPACITY, "mcl_kld_resample: too many samples");
    DeviceGuard guard(h->device);
    uint64_t cap = 64;
    while (cap < 2 * (uint64_t)max_samples) cap <<= 1;
    // work buffers: cumulative sums | per-sample keys | table keys | table min | result + overflow flag
    size_t off = 0;
    const size_t o_c = off; off += (((size_t)n_in * 8) + 255) & ~(size_t)255;
    const size_t o_keys = off; off += (((size_t)max_samples * 8) + 255) & ~(size_t)255;
    const size_t o_tk = off; off += cap * 8;
    const size_t o_tm = off; off += cap * 4;
    const size_t o_res = off; off += 256;
    if (off > h->kld_bytes) {
        MCL_CUDA(h, cudaStreamSynchronize(h->stream));
        cudaFree(h->d_kld); h->d_kld = nullptr; h->kld_bytes = 0;
        MCL_CUDA(h, cudaMalloc(&h->d_kld, off));
        h->kld_bytes = off;
    }
    char *s = (char *)h->d_kld;
    KldParams p;
    p.x = d_x; p.y = d_y; p.th = d_theta; p.n_in = n_in; p.max_samples = max_samples;
    p.r = r; p.bin_xy = bin_xy; p.bin_theta = bin_theta; p.normals = d_normals; p.seed = seed; p.step = step;
    p.xo = d_xo; p.yo = d_yo; p.tho = d_thetao;
    p.keys = (uint64_t *)(s + o_keys);
    p.tab_keys = (unsigned long long *)(s + o_tk); p.tab_min = (unsigned *)(s + o_tm); p.tab_mask = cap - 1;
    p.overflow = (int *)(s + o_res + 8);
    p.cf = nullptr; p.cq = nullptr; p.total_q = nullptr;
    int rc;
    if (mode == MCL_RESAMPLE_REFERENCE_F32) {
        rc = mcl_cumsum_f32_seq(h, d_weights, n_in, (float *)(s + o_c));
        if (rc) return rc;
        p.cf = (const float *)(s + o_c);
    } else if (mode == MCL_RESAMPLE_FIXED_POINT) {
        // reuse the fixed-point scan of resample.cu through its staged entry points
        float *wmax = (float *)(s + o_res + 16);
        rc = mcl_weights_max(h, d_weights, n_in, wmax);
        if (rc) return rc;
        rc = mcl_resample_scan(h, d_weights, n_in, wmax, n_in, (uint64_t *)(s + o_res + 24));
        if (rc) return rc;
        // the cumulative sums live in the handle's scratch (resample.cu layout): copy them out
        const int64_t nt = (n_in + 2047) / 2048;
        const int wblocks = (int)std::min<int64_t>((n_in + 1023) / 1024, (int64_t)h->sm_count * 8);
        size_t so = 128;
        so += ((size_t)wblocks * sizeof(float) + 63) & ~(size_t)63;
        so += ((size_t)nt * 8 + 63) & ~(size_t)63;
        so += ((size_t)(nt + 1) * 8 + 63) & ~(size_t)63;
        MCL_CUDA(h, cudaMemcpyAsync(s + o_c, (char *)h->d_scratch + so, (size_t)n_in * 8, cudaMemcpyDeviceToDevice, h->stream));
        p.cq = (const uint64_t *)(s + o_c);
        p.total_q = (const uint64_t *)(s + o_res + 24);
    } else {
        return mcl_fail(h, MCL_ERR_ARG, "mcl_kld_resample: unknown mode");
    }
    const int fb = (int)std::min<uint64_t>((cap + 255) / 256, (uint64_t)h->sm_count * 16);
    k_fill_u64<<<fb, 256, 0, h->stream>>>(p.tab_keys, (int64_t)cap, KLD_EMPTY);
    MCL_LAUNCH_CHECK(h);
    k_fill_u32<<<fb, 256, 0, h->stream>>>(p.tab_min, (int64_t)cap, 0xffffffffu);
    MCL_LAUNCH_CHECK(h);
    MCL_CUDA(h, cudaMemsetAsync(s + o_res, 0, 16, h->stream));
    const int blocks = (int)std::min<int64_t>((max_samples + 255) / 256, (int64_t)h->sm_count * 16);
    k_kld_sample<<<blocks, 256, 0, h->stream>>>(p);
    MCL_LAUNCH_CHECK(h);
    k_kld_stop<<<1, 1024, 0, h->stream>>>(p.keys, p.tab_keys, p.tab_min, p.tab_mask, max_samples, min_particles, epsilon,
                                          z, (long long *)(s + o_res));
    MCL_LAUNCH_CHECK(h);
    MCL_CUDA(h, cudaMemcpyAsync(h->h_pinned, s + o_res, 16, cudaMemcpyDeviceToHost, h->stream));
    MCL_CUDA(h, cudaStreamSynchronize(h->stream));
    long long cnt; int ovf;
    memcpy(&cnt, h->h_pinned, 8);
    memcpy(&ovf, (char *)h->h_pinned + 8, 4);
    if (ovf) return mcl_fail(h, MCL_ERR_CAPACITY, "mcl_kld_resample: a bin index exceeds the 21/22-bit key range");
    *h_count = cnt;
    return MCL_OK;
}
