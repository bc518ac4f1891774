// raycast.cu -- the reference's alternative beam model (not on the node's live path):
//   pu:4-29   raycast: march 0.1 m steps along the beam until an occupied cell (grid > 0.5) or the map edge
//   pu:151-201 compute_likelihoods_raycast: p = 0.8 N(r_meas; r_pred, 0.05) + 0.1 / 10, mean log p per valid beam
// One warp per particle, lanes over beams, fp64 throughout (the marching positions decide cells, like the
// likelihood-field endpoints).  The occupancy grid is a bitmap (W*H/8 bytes: 18 KB for 384 x 384) staged in
// shared memory when it fits, read through L1/L2 otherwise.
#include <math.h>

#include <algorithm>
#include <vector>

#include "common.cuh"

struct RcParams {
    const double *x, *y, *th;
    int64_t n;
    float *score;
    const float *ranges, *angles;     // device copies of the scan
    int M;
    const uint32_t *bits;
    int W, H, words;
    double res, xmin, ymin, max_range, step_size, sigma_hit, z_hit, z_rand;
    int max_steps;
};

template <bool SMEM>
__global__ void __launch_bounds__(256) k_likelihood_raycast(const RcParams p) {
    extern __shared__ uint32_t sbits[];
    if (SMEM) {
        for (int i = threadIdx.x; i < p.words; i += blockDim.x) sbits[i] = p.bits[i];
        __syncthreads();
    }
    const uint32_t *bits = SMEM ? sbits : p.bits;
    const int lane = threadIdx.x & 31;
    const int64_t warps = (int64_t)gridDim.x * (blockDim.x >> 5);
    for (int64_t i = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); i < p.n; i += warps) {
        const double x = p.x[i], y = p.y[i], th = p.th[i];
        double acc = 0.0;
        int valid = 0;
        for (int j = lane; j < p.M; j += 32) {
            const double r_meas = (double)p.ranges[j];
            if (!(isfinite(r_meas) && r_meas < p.max_range)) continue;                // pu:172
            valid += 1;
            double dy, dx;
            sincos(__dadd_rn(th, (double)p.angles[j]), &dy, &dx);                    // pu:7-8
            double r_pred = p.max_range;
            for (int s = 1; s <= p.max_steps; ++s) {
                const double t = __dmul_rn((double)s, p.step_size);                   // i * step_size
                const double cx = __dadd_rn(x, __dmul_rn(t, dx)), cy = __dadd_rn(y, __dmul_rn(t, dy));
                const long long gx = __double2ll_rz(__ddiv_rn(__dadd_rn(cx, -p.xmin), p.res));   // pu:17-18 int()
                const long long gy = __double2ll_rz(__ddiv_rn(__dadd_rn(cy, -p.ymin), p.res));
                if (!(gx >= 0 && gx < p.W && gy >= 0 && gy < p.H)) break;             // pu:21-22 -> max_range
                const long long c = gy * (long long)p.W + gx;
                if ((bits[c >> 5] >> (c & 31)) & 1u) { r_pred = t; break; }           // pu:25-26
            }
            double prob_hit = 0.0, prob_rand = 0.0;
            if (0 <= r_meas && r_meas <= p.max_range) {                               // pu:36-41, 55-58
                const double q = __ddiv_rn(__dadd_rn(r_meas, -r_pred), p.sigma_hit);
                prob_hit = __dmul_rn(__ddiv_rn(1.0, __dmul_rn(sqrt(MCL_TWO_PI), p.sigma_hit)),
                                     exp(__dmul_rn(-0.5, __dmul_rn(q, q))));
                prob_rand = __ddiv_rn(1.0, p.max_range);
            }
            double pr = __dadd_rn(__dmul_rn(p.z_hit, prob_hit), __dmul_rn(p.z_rand, prob_rand));
            pr = pr > 1e-6 ? pr : 1e-6;
            acc += log(pr);
        }
        acc = warp_sum(acc);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) valid += __shfl_xor_sync(0xffffffffu, valid, o);
        if (lane == 0) p.score[i] = valid > 0 ? (float)(acc / (double)valid) : -INFINITY;   // pu:196-199
    }
}

struct RcState {
    uint32_t *d_bits = nullptr;
    float *d_scan = nullptr;
    int scan_cap = 0;
    int W = 0, H = 0, words = 0;
    double res = 0, xmin = 0, ymin = 0;
};
#include <map>
#include <mutex>
static std::map<const mcl_handle *, RcState> g_rc;
static std::mutex g_rc_mu;
static RcState *rc_of(mcl_handle *h) {
    std::lock_guard<std::mutex> lk(g_rc_mu);
    return &g_rc[h];
}
// mcl_destroy: free the grid and drop the entry (a later handle at the same address must not inherit it)
void mcl_raycast_forget(const mcl_handle *h) {
    std::lock_guard<std::mutex> lk(g_rc_mu);
    auto it = g_rc.find(h);
    if (it == g_rc.end()) return;
    cudaFree(it->second.d_bits); cudaFree(it->second.d_scan);
    g_rc.erase(it);
}

extern "C" int mcl_set_raycast_grid(mcl_handle *h, const uint8_t *h_blocked, int W, int H, double res, double x_min,
                                    double y_min) {
    if (!h) return MCL_ERR_ARG;
    if (!h_blocked || W <= 0 || H <= 0 || !(res > 0)) return mcl_fail(h, MCL_ERR_ARG, "mcl_set_raycast_grid: bad argument");
    DeviceGuard guard(h->device);
    RcState *r = rc_of(h);
    const size_t cells = (size_t)W * H;
    const int words = (int)((cells + 31) / 32);
    std::vector<uint32_t> bits((size_t)words, 0u);
    for (size_t c = 0; c < cells; ++c)
        if (h_blocked[c]) bits[c >> 5] |= 1u << (c & 31);
    MCL_CUDA(h, cudaStreamSynchronize(h->stream));
    cudaFree(r->d_bits);
    r->d_bits = nullptr;
    MCL_CUDA(h, cudaMalloc((void **)&r->d_bits, (size_t)words * 4));
    MCL_CUDA(h, cudaMemcpy(r->d_bits, bits.data(), (size_t)words * 4, cudaMemcpyHostToDevice));
    r->W = W; r->H = H; r->words = words; r->res = res; r->xmin = x_min; r->ymin = y_min;
    return MCL_OK;
}

extern "C" int mcl_likelihood_raycast(mcl_handle *h, const float *h_ranges, const float *h_angles, int M,
                                      const double *d_x, const double *d_y, const double *d_theta, int64_t n,
                                      float *d_score) {
    if (!h) return MCL_ERR_ARG;
    if (n < 0 || M < 0 || (M > 0 && (!h_ranges || !h_angles)) || (n > 0 && (!d_x || !d_y || !d_theta || !d_score)))
        return mcl_fail(h, MCL_ERR_ARG, "mcl_likelihood_raycast: bad argument");
    RcState *r = rc_of(h);
    if (!r->d_bits) return mcl_fail(h, MCL_ERR_STATE, "mcl_likelihood_raycast: grid not set (mcl_set_raycast_grid)");
    if (n == 0) return MCL_OK;
    DeviceGuard guard(h->device);
    if (M > r->scan_cap) {
        MCL_CUDA(h, cudaStreamSynchronize(h->stream));
        cudaFree(r->d_scan);
        r->d_scan = nullptr;
        MCL_CUDA(h, cudaMalloc((void **)&r->d_scan, (size_t)std::max(M, 512) * 2 * sizeof(float)));
        r->scan_cap = std::max(M, 512);
    }
    if (M > 0) {
        MCL_CUDA(h, cudaMemcpyAsync(r->d_scan, h_ranges, (size_t)M * 4, cudaMemcpyHostToDevice, h->stream));
        MCL_CUDA(h, cudaMemcpyAsync(r->d_scan + r->scan_cap, h_angles, (size_t)M * 4, cudaMemcpyHostToDevice, h->stream));
        MCL_CUDA(h, cudaStreamSynchronize(h->stream));     // the caller's host arrays may go away
    }
    RcParams p;
    p.x = d_x; p.y = d_y; p.th = d_theta; p.n = n; p.score = d_score;
    p.ranges = r->d_scan; p.angles = r->d_scan + r->scan_cap; p.M = M;
    p.bits = r->d_bits; p.W = r->W; p.H = r->H; p.words = r->words;
    p.res = r->res; p.xmin = r->xmin; p.ymin = r->ymin;
    p.max_range = 10.0; p.step_size = 0.1; p.sigma_hit = 0.05; p.z_hit = 0.8; p.z_rand = 0.1;   // pu:159-162, pu:9
    p.max_steps = (int)(p.max_range / p.step_size);                                            // pu:10
    const size_t smem = (size_t)r->words * 4;
    const int64_t need = (n + 7) / 8;
    if (smem <= (size_t)h->smem_optin) {
        MCL_CUDA(h, cudaFuncSetAttribute(k_likelihood_raycast<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, h->smem_optin));
        int occ = 0;
        MCL_CUDA(h, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_likelihood_raycast<true>, 256, smem));
        const int blocks = (int)std::min<int64_t>(need, (int64_t)h->sm_count * std::max(occ, 1));
        k_likelihood_raycast<true><<<blocks, 256, smem, h->stream>>>(p);
    } else {
        const int blocks = (int)std::min<int64_t>(need, (int64_t)h->sm_count * 8);
        k_likelihood_raycast<false><<<blocks, 256, 0, h->stream>>>(p);
    }
    MCL_LAUNCH_CHECK(h);
    return MCL_OK;
}
