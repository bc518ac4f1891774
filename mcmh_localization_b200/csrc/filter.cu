// filter.cu -- the whole filter step sequenced inside the library: one C call enqueues
// predict -> likelihood x2 -> softmax x2 -> MH accept -> estimate -> resample -> gather on the
// handle's stream (node:384-408 move_particles + node:294-338 lidar_callback).  The Python host
// pays one ctypes call per step instead of ~15; buffer roles (particles / particles_prev / spare)
// rotate inside the handle exactly like the node's copies at node:404-405, node:370 and node:490.
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>

#include "common.cuh"

struct FilterState {
    bool bound = false;
    int64_t n = 0;
    double *x[3] = {nullptr, nullptr, nullptr}, *y[3] = {nullptr, nullptr, nullptr}, *th[3] = {nullptr, nullptr, nullptr};
    float *score_pre = nullptr, *score_post = nullptr, *w_pre = nullptr, *w_post = nullptr, *w[2] = {nullptr, nullptr};
    int32_t *idx = nullptr;
    int cur = 0, prev = 1, spare = 2, wslot = 0;
    int use_mh = 1, resample_mode = MCL_RESAMPLE_FIXED_POINT, max_attempts = 1000;
    uint64_t seed = 0, first_index = 0, tick = 0;
    // asymmetric MH (node:365-367): last odometry increment, the reference's "backward" increment
    // (node:429-434) and the two transition-density buffers
    int assym = 0;
    double delta[3] = {0, 0, 0}, delta_b[3] = {0, 0, 0};
    double *tf = nullptr, *tb = nullptr;
    int64_t t_cap = 0;
    unsigned long long *thr = nullptr;      // screening thresholds of particles_prev during an MH chain (motion.cu)
    int64_t thr_cap = 0;
    // sharded operation (one process per GPU): peer-memory mailboxes for the small exchanges and peer
    // pointers of every rank's pose buffers for the resampling push
    bool comm = false;
    int rank = 0, world = 1;
    int64_t n_global = 0;
    unsigned long long *mailbox = nullptr;           // local mailbox (symmetric memory)
    unsigned long long peer_mailbox[16] = {0};
    uint64_t *d_peer_pose = nullptr;                  // device [3 sets][3 comps][world]
    uint64_t epoch = 0;
    double *d_x8 = nullptr;                           // device staging for exchange payloads / results (64 doubles)
    int *d_comm_err = nullptr;
};

// one FilterState per handle, kept out of common.cuh: keyed by handle pointer
#include <map>
#include <mutex>
static std::map<const mcl_handle *, FilterState> g_filters;
static std::mutex g_filters_mu;

static FilterState *filter_of(mcl_handle *h, bool create) {
    std::lock_guard<std::mutex> lk(g_filters_mu);
    auto it = g_filters.find(h);
    if (it == g_filters.end()) {
        if (!create) return nullptr;
        it = g_filters.emplace(h, FilterState()).first;
    }
    return &it->second;
}

void mcl_filter_forget(const mcl_handle *h) {
    std::lock_guard<std::mutex> lk(g_filters_mu);
    auto it = g_filters.find(h);
    if (it != g_filters.end()) {
        cudaFree(it->second.tf); cudaFree(it->second.tb); cudaFree(it->second.d_peer_pose); cudaFree(it->second.thr);
        cudaFree(it->second.d_x8); cudaFree(it->second.d_comm_err);
    }
    g_filters.erase(h);
}

// node:429-434: the reference builds the backward increment as if (rot1, trans, rot2) were (dx, dy, dtheta)
static void backward_delta(const double d[3], double out[3]) {
    out[0] = -d[0] * cos(d[2]) - d[1] * sin(d[2]);
    out[1] = d[0] * sin(d[2]) - d[1] * cos(d[2]);
    out[2] = -d[2];
}
// corrected variant: the odometry increment that undoes (rot1, trans, rot2) -- turn to face
// back along the travelled segment, drive it, turn to the old heading
static double wrap_pi(double a) {
    a = fmod(a + MCL_PI, MCL_TWO_PI);
    if (a < 0) a += MCL_TWO_PI;
    return a - MCL_PI;
}
static void backward_delta_corrected(const double d[3], double out[3]) {
    out[0] = wrap_pi(MCL_PI - d[2]);
    out[1] = d[1];
    out[2] = wrap_pi(-d[0] - MCL_PI);
}

extern "C" int mcl_filter_set_assym(mcl_handle *h, int assym) {
    if (!h) return MCL_ERR_ARG;
    FilterState *f = filter_of(h, false);
    if (!f || !f->bound) return mcl_fail(h, MCL_ERR_STATE, "mcl_filter_set_assym: no filter bound");
    f->assym = assym == 2 ? 2 : (assym ? 1 : 0);     // 2: corrected variant (see include/mcl.h)
    return MCL_OK;
}

// delta_b == NULL: computed here with glibc cos/sin (the Python host passes NumPy's, to stay bit-exact)
extern "C" int mcl_filter_set_transition(mcl_handle *h, const double delta[3], const double delta_b[3]) {
    if (!h || !delta) return MCL_ERR_ARG;
    FilterState *f = filter_of(h, false);
    if (!f || !f->bound) return mcl_fail(h, MCL_ERR_STATE, "mcl_filter_set_transition: no filter bound");
    memcpy(f->delta, delta, sizeof(f->delta));
    if (delta_b) memcpy(f->delta_b, delta_b, sizeof(f->delta_b));
    else if (f->assym == 2) backward_delta_corrected(delta, f->delta_b);
    else backward_delta(delta, f->delta_b);
    return MCL_OK;
}

extern "C" int mcl_filter_bind(mcl_handle *h, int64_t n, double *const x[3], double *const y[3], double *const th[3],
                               float *score_pre, float *score_post, float *w_pre, float *w_post, float *w_a,
                               float *w_b, int32_t *idx, int use_mh, int resample_mode, uint64_t seed,
                               uint64_t first_index, int max_attempts) {
    if (!h) return MCL_ERR_ARG;
    if (n <= 0 || !x || !y || !th || !score_pre || !score_post || !w_pre || !w_post || !w_a || !w_b || !idx)
        return mcl_fail(h, MCL_ERR_ARG, "mcl_filter_bind: bad argument");
    if (max_attempts > 65535) return mcl_fail(h, MCL_ERR_ARG, "mcl_filter_bind: max_attempts > 65535");
    FilterState *f = filter_of(h, true);
    for (int k = 0; k < 3; ++k) {
        if (!x[k] || !y[k] || !th[k]) return mcl_fail(h, MCL_ERR_ARG, "mcl_filter_bind: null pose buffer");
        f->x[k] = x[k]; f->y[k] = y[k]; f->th[k] = th[k];
    }
    f->n = n;
    f->score_pre = score_pre; f->score_post = score_post; f->w_pre = w_pre; f->w_post = w_post;
    f->w[0] = w_a; f->w[1] = w_b; f->idx = idx;
    f->cur = 0; f->prev = 1; f->spare = 2; f->wslot = 0;
    f->use_mh = use_mh; f->resample_mode = resample_mode; f->seed = seed; f->first_index = first_index;
    f->max_attempts = max_attempts; f->tick = 0;
    f->bound = true;
    return MCL_OK;
}

extern "C" int mcl_filter_roles(mcl_handle *h, int roles[4], uint64_t *tick) {
    if (!h || !roles) return MCL_ERR_ARG;
    FilterState *f = filter_of(h, false);
    if (!f || !f->bound) return mcl_fail(h, MCL_ERR_STATE, "mcl_filter_roles: no filter bound");
    roles[0] = f->cur; roles[1] = f->prev; roles[2] = f->spare; roles[3] = f->wslot;
    if (tick) *tick = f->tick;
    return MCL_OK;
}

extern "C" int mcl_filter_set_roles(mcl_handle *h, const int roles[4], uint64_t tick) {
    if (!h || !roles) return MCL_ERR_ARG;
    FilterState *f = filter_of(h, false);
    if (!f || !f->bound) return mcl_fail(h, MCL_ERR_STATE, "mcl_filter_set_roles: no filter bound");
    int seen = 0;
    for (int k = 0; k < 3; ++k) {
        if (roles[k] < 0 || roles[k] > 2) return mcl_fail(h, MCL_ERR_ARG, "mcl_filter_set_roles: bad role");
        seen |= 1 << roles[k];
    }
    if (seen != 7 || roles[3] < 0 || roles[3] > 1) return mcl_fail(h, MCL_ERR_ARG, "mcl_filter_set_roles: bad roles");
    f->cur = roles[0]; f->prev = roles[1]; f->spare = roles[2]; f->wslot = roles[3]; f->tick = tick;
    return MCL_OK;
}

// KLD-adaptive modes change the particle count every scan (node:496-527); the buffers keep their capacity
extern "C" int mcl_filter_set_n(mcl_handle *h, int64_t n) {
    if (!h) return MCL_ERR_ARG;
    FilterState *f = filter_of(h, false);
    if (!f || !f->bound) return mcl_fail(h, MCL_ERR_STATE, "mcl_filter_set_n: no filter bound");
    if (n <= 0) return mcl_fail(h, MCL_ERR_ARG, "mcl_filter_set_n: n <= 0");
    f->n = n;
    return MCL_OK;
}

extern "C" int mcl_filter_configure(mcl_handle *h, int use_mh, int resample_mode, uint64_t seed, uint64_t first_index,
                                    int64_t tick /* < 0: keep */) {
    if (!h) return MCL_ERR_ARG;
    FilterState *f = filter_of(h, false);
    if (!f || !f->bound) return mcl_fail(h, MCL_ERR_STATE, "mcl_filter_configure: no filter bound");
    f->use_mh = use_mh; f->resample_mode = resample_mode; f->seed = seed; f->first_index = first_index;
    if (tick >= 0) f->tick = (uint64_t)tick;
    return MCL_OK;
}

// ---------------------------------------------------------------------------------------------
// Small all-gathers over NVLink peer memory.  The per-step exchanges are a handful of scalars (softmax
// max / sum, estimate sums, resampling scale and totals), latency-bound through NCCL (20-50 us each with 8
// ranks).  Here every rank stores its payload straight into every peer's mailbox (symmetric memory), raises
// a flag with system-scope release, and spins with acquire on its own flags until all peers have written:
// one ~5 us kernel, no host involvement, and the reduction is done locally in rank order (deterministic).
// Mailbox (u64 units): flags[2][16] | slots[2][16][16]; two parities so that a fast rank's exchange k+2
// cannot overwrite what a slow rank is still reading from exchange k (a rank cannot finish k+1 before every
// peer has entered k+1, i.e. finished k).
// ---------------------------------------------------------------------------------------------
enum { XCH_SUM_F64 = 0, XCH_MAX_F64 = 1, XCH_SUM_U64 = 2, XCH_GATHER = 3, XCH_MAX_U64 = 4,
       XCH_SUM_F64_MAXLAST = 5 /* sums, except the last value: maximum */,
       XCH_SUM_F64_GATHERLAST = 6 /* sums, except the last value: gathered (raw 64 bits) into out2[rank] */ };
#define XCH_MAXV 16

struct XchArgs {
    unsigned long long *mailbox;
    unsigned long long peers[16];
    int rank, world;
    unsigned long long epoch;
    int *err;
};

__device__ __forceinline__ void st_release_sys(unsigned long long *p, unsigned long long v) {
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long *p) {
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}

__global__ void __launch_bounds__(32) k_exchange(const XchArgs a, const unsigned long long *payload, int nvals, int op,
                                                 unsigned long long *out, unsigned long long *out2) {
    const int t = threadIdx.x;
    const int par = (int)(a.epoch & 1ull);
    if (*(volatile int *)a.err) return;        // a peer was lost earlier: sticky, the host reports it (comm_check)
    if (t < a.world) {
        unsigned long long *peer = reinterpret_cast<unsigned long long *>(a.peers[t]);
        unsigned long long *slot = peer + 32 + ((size_t)par * 16 + a.rank) * XCH_MAXV;
        for (int v = 0; v < nvals; ++v) slot[v] = payload[v];
        __threadfence_system();
        st_release_sys(peer + par * 16 + a.rank, a.epoch);
    }
    bool ok = true;
    if (t < a.world) {
        const unsigned long long *flag = a.mailbox + par * 16 + t;
        const long long t0 = clock64();
        while (ld_acquire_sys(flag) < a.epoch) {
            if (clock64() - t0 > 8000000000ll) { ok = false; break; }      // ~4 s: a peer died; do not hang the GPU
        }
    }
    if (!__all_sync(0xffffffffu, ok)) { if (t == 0) *a.err = 1; return; }
    if (t < nvals) {
        const unsigned long long *slots = a.mailbox + 32 + (size_t)par * 16 * XCH_MAXV;
        if (op == XCH_SUM_F64_GATHERLAST && t == nvals - 1) {
            for (int d = 0; d < a.world; ++d) out2[d] = ((volatile const unsigned long long *)slots)[d * XCH_MAXV + t];
        } else if (op == XCH_GATHER) {
            for (int d = 0; d < a.world; ++d) out[d * nvals + t] = ((volatile const unsigned long long *)slots)[d * XCH_MAXV + t];
        } else if (op == XCH_SUM_U64 || op == XCH_MAX_U64) {
            unsigned long long acc = 0;
            for (int d = 0; d < a.world; ++d) {
                const unsigned long long v = ((volatile const unsigned long long *)slots)[d * XCH_MAXV + t];
                acc = op == XCH_MAX_U64 ? (v > acc ? v : acc) : acc + v;
            }
            out[t] = acc;
        } else {
            double acc = __longlong_as_double((long long)((volatile const unsigned long long *)slots)[t]);
            for (int d = 1; d < a.world; ++d) {
                const double v = __longlong_as_double((long long)((volatile const unsigned long long *)slots)[d * XCH_MAXV + t]);
                const bool mx = op == XCH_MAX_F64 || (op == XCH_SUM_F64_MAXLAST && t == nvals - 1);
                acc = mx ? fmax(acc, v) : acc + v;                           // rank order: deterministic
            }
            out[t] = (unsigned long long)__double_as_longlong(acc);
        }
    }
}

static int comm_exchange(mcl_handle *h, FilterState *f, const void *d_payload, int nvals, int op, void *d_out,
                         void *d_out2 = nullptr) {
    if (nvals > XCH_MAXV) return mcl_fail(h, MCL_ERR_ARG, "comm_exchange: payload too large");
    XchArgs a;
    a.mailbox = f->mailbox;
    for (int d = 0; d < 16; ++d) a.peers[d] = f->peer_mailbox[d];
    a.rank = f->rank; a.world = f->world; a.epoch = ++f->epoch; a.err = f->d_comm_err;
    k_exchange<<<1, 32, 0, h->stream>>>(a, (const unsigned long long *)d_payload, nvals, op, (unsigned long long *)d_out,
                                        (unsigned long long *)d_out2);
    MCL_LAUNCH_CHECK(h);
    return MCL_OK;
}

// helpers on the staging buffer d_x8: [0..15] payload, [16..47] result
__global__ void k_pack2(const double *a, int ia, const double *b, int ib, double *out) { out[0] = a[ia]; out[1] = b ? b[ib] : 0.0; }
__global__ void k_unpack_max(const double *in, double *a, double *b) { a[0] = in[0]; if (b) b[0] = in[1]; }
__global__ void k_unpack_qsum(const unsigned long long *in, double *a, double *b) {
    a[1] = (double)in[0] / 1099511627776.0; ((unsigned long long *)a)[2] = in[0];
    if (b) { b[1] = (double)in[1] / 1099511627776.0; ((unsigned long long *)b)[2] = in[1]; }
}
__global__ void k_f32_to_f64(const float *in, double *out) { out[0] = (double)in[0]; }
__global__ void k_f64_to_f32(const double *in, float *out) { out[0] = (float)in[0]; }

// mcl_comm_init: see include/mcl.h
extern "C" int mcl_comm_init(mcl_handle *h, int rank, int world, void *d_mailbox, const uint64_t *h_peer_mailbox,
                             const uint64_t *h_peer_pose /* [3][3][world] */) {
    if (!h) return MCL_ERR_ARG;
    FilterState *f = filter_of(h, false);
    if (!f || !f->bound) return mcl_fail(h, MCL_ERR_STATE, "mcl_comm_init: bind the filter first");
    if (world < 1 || world > 16 || rank < 0 || rank >= world || !d_mailbox || !h_peer_mailbox || !h_peer_pose)
        return mcl_fail(h, MCL_ERR_ARG, "mcl_comm_init: bad argument");
    DeviceGuard guard(h->device);
    f->rank = rank; f->world = world; f->n_global = f->n * world;
    f->mailbox = (unsigned long long *)d_mailbox;
    for (int d = 0; d < world; ++d) f->peer_mailbox[d] = h_peer_mailbox[d];
    if (!f->d_peer_pose) MCL_CUDA(h, cudaMalloc((void **)&f->d_peer_pose, 9 * 16 * sizeof(uint64_t)));
    if (!f->d_x8) MCL_CUDA(h, cudaMalloc((void **)&f->d_x8, 64 * sizeof(double)));
    if (!f->d_comm_err) { MCL_CUDA(h, cudaMalloc((void **)&f->d_comm_err, sizeof(int))); MCL_CUDA(h, cudaMemset(f->d_comm_err, 0, sizeof(int))); }
    MCL_CUDA(h, cudaMemcpy(f->d_peer_pose, h_peer_pose, (size_t)9 * world * sizeof(uint64_t), cudaMemcpyHostToDevice));
    f->first_index = (uint64_t)rank * (uint64_t)f->n;
    f->epoch = 0;
    f->comm = true;
    return MCL_OK;
}

extern "C" int mcl_comm_status(mcl_handle *h, int *err) {
    if (!h || !err) return MCL_ERR_ARG;
    FilterState *f = filter_of(h, false);
    *err = 0;
    if (!f || !f->comm) return MCL_OK;
    DeviceGuard guard(h->device);
    MCL_CUDA(h, cudaMemcpyAsync(h->h_pinned, f->d_comm_err, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
    MCL_CUDA(h, cudaStreamSynchronize(h->stream));
    memcpy(err, h->h_pinned, sizeof(int));
    return MCL_OK;
}

// softmax of both score sets with the two global statistics exchanged over peer memory
static int sharded_softmax_arrays(mcl_handle *h, FilterState *f, const float *s_post, float *w_post, const float *s_pre,
                                  float *w_pre);
static int sharded_softmax(mcl_handle *h, FilterState *f, bool both) {
    return sharded_softmax_arrays(h, f, f->score_post, both ? f->w_post : f->w[f->wslot], both ? f->score_pre : nullptr,
                                  both ? f->w_pre : nullptr);
}
static int sharded_softmax_arrays(mcl_handle *h, FilterState *f, const float *s_post, float *w_post, const float *s_pre,
                                  float *w_pre) {
    const bool both = s_pre != nullptr;
    double *st_post = f->d_x8 + 48, *st_pre = f->d_x8 + 52;          // 4 doubles each
    int rc = mcl_softmax_max(h, s_post, f->n, st_post);
    if (rc) return rc;
    if (both) { rc = mcl_softmax_max(h, s_pre, f->n, st_pre); if (rc) return rc; }
    k_pack2<<<1, 1, 0, h->stream>>>(st_post, 0, both ? st_pre : nullptr, 0, f->d_x8);
    MCL_LAUNCH_CHECK(h);
    rc = comm_exchange(h, f, f->d_x8, 2, XCH_MAX_F64, f->d_x8 + 16);
    if (rc) return rc;
    k_unpack_max<<<1, 1, 0, h->stream>>>(f->d_x8 + 16, st_post, both ? st_pre : nullptr);
    MCL_LAUNCH_CHECK(h);
    rc = mcl_softmax_sumexp(h, s_post, f->n, st_post);
    if (rc) return rc;
    if (both) { rc = mcl_softmax_sumexp(h, s_pre, f->n, st_pre); if (rc) return rc; }
    k_pack2<<<1, 1, 0, h->stream>>>(st_post, 2, both ? st_pre : nullptr, 2, f->d_x8);     // the exact 2^-40 integers
    MCL_LAUNCH_CHECK(h);
    rc = comm_exchange(h, f, f->d_x8, 2, XCH_SUM_U64, f->d_x8 + 16);
    if (rc) return rc;
    k_unpack_qsum<<<1, 1, 0, h->stream>>>((const unsigned long long *)(f->d_x8 + 16), st_post, both ? st_pre : nullptr);
    MCL_LAUNCH_CHECK(h);
    rc = mcl_softmax_weights(h, s_post, f->n, st_post, w_post);
    if (rc) return rc;
    if (both) rc = mcl_softmax_weights(h, s_pre, f->n, st_pre, w_pre);
    return rc;
}

// after a host read-back that has synchronised the stream: a timed-out exchange invalidates everything since
static int comm_check(mcl_handle *h, FilterState *f) {
    if (!f->comm) return MCL_OK;
    int e = 0;
    MCL_CUDA(h, cudaMemcpyAsync(h->h_pinned + 22, f->d_comm_err, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
    MCL_CUDA(h, cudaStreamSynchronize(h->stream));
    memcpy(&e, h->h_pinned + 22, sizeof(int));
    if (e) return mcl_fail(h, MCL_ERR_CUDA, "sharded step: a peer-memory exchange timed out (a rank died?); the particle sets "
                                            "are no longer consistent -- rebuild the ShardedLocalizer");
    return MCL_OK;
}

#define FILTER_OR_FAIL(name)                                                                          \
    if (!h) return MCL_ERR_ARG;                                                                       \
    FilterState *f = filter_of(h, false);                                                             \
    if (!f || !f->bound) return mcl_fail(h, MCL_ERR_STATE, name ": no filter bound (mcl_filter_bind)");

// node:397-405: particles_prop = motion(particles); particles_prev = particles; particles = particles_prop
extern "C" int mcl_filter_predict(mcl_handle *h, const double delta[3], const double *d_normals, int A) {
    FILTER_OR_FAIL("mcl_filter_predict");
    memcpy(f->delta, delta, sizeof(f->delta));
    if (f->assym == 2) backward_delta_corrected(delta, f->delta_b); else backward_delta(delta, f->delta_b);
    f->tick++;
    int rc = mcl_predict(h, f->x[f->cur], f->y[f->cur], f->th[f->cur], f->n, delta, f->seed, f->tick, f->first_index,
                         d_normals, A, f->max_attempts, f->x[f->spare], f->y[f->spare], f->th[f->spare], nullptr);
    if (rc) return rc;
    const int old_prev = f->prev;
    f->prev = f->cur; f->cur = f->spare; f->spare = old_prev;
    return MCL_OK;
}

// node:252-273 update_weights + node:307-322 (MH by mode); uses the active scan
extern "C" int mcl_filter_update(mcl_handle *h, const double *d_uniforms) {
    FILTER_OR_FAIL("mcl_filter_update");
    int rc = mcl_likelihood(h, f->x[f->cur], f->y[f->cur], f->th[f->cur], f->n, f->score_post);
    if (rc) return rc;
    if (!f->use_mh) {
        // MCL: weights = weights_post (node:313); scores_pre would be computed and discarded by the reference
        if (f->comm) return sharded_softmax(h, f, false);
        return mcl_softmax(h, f->score_post, f->n, f->w[f->wslot], nullptr, nullptr);
    }
    if (f->comm) {
        rc = mcl_likelihood(h, f->x[f->prev], f->y[f->prev], f->th[f->prev], f->n, f->score_pre);
        if (rc) return rc;
        rc = sharded_softmax(h, f, true);
        if (rc) return rc;
    } else {
        rc = mcl_likelihood(h, f->x[f->prev], f->y[f->prev], f->th[f->prev], f->n, f->score_pre);
        if (rc) return rc;
        rc = mcl_softmax_pair(h, f->score_post, f->w_post, f->score_pre, f->w_pre, f->n);    // node:261, 270
        if (rc) return rc;
    }
    f->tick++;
    if (f->assym) {
        // node:366-367: transition_probability() then assym_mh_resampling(prev, cur, w_post, w_pre, fwd, bwd)
        if (f->t_cap < f->n) {
            DeviceGuard guard(h->device);
            MCL_CUDA(h, cudaStreamSynchronize(h->stream));
            cudaFree(f->tf); cudaFree(f->tb); f->tf = f->tb = nullptr; f->t_cap = 0;
            MCL_CUDA(h, cudaMalloc((void **)&f->tf, (size_t)f->n * sizeof(double)));
            MCL_CUDA(h, cudaMalloc((void **)&f->tb, (size_t)f->n * sizeof(double)));
            f->t_cap = f->n;
        }
        if (f->comm) {
            // pu:326-328 normalises by the sum over the WHOLE population: local sums, rank-ordered sum over peers
            rc = mcl_motion_density(h, f->x[f->prev], f->y[f->prev], f->th[f->prev], f->x[f->cur], f->y[f->cur],
                                    f->th[f->cur], f->n, f->delta, f->tf, f->d_x8, 0);
            if (rc) return rc;
            rc = mcl_motion_density(h, f->x[f->cur], f->y[f->cur], f->th[f->cur], f->x[f->prev], f->y[f->prev],
                                    f->th[f->prev], f->n, f->delta_b, f->tb, f->d_x8 + 1, 0);
            if (rc) return rc;
            rc = comm_exchange(h, f, f->d_x8, 2, XCH_SUM_F64, f->d_x8 + 16);
            if (rc) return rc;
            rc = mcl_scale_by_sum(h, f->tf, f->n, f->d_x8 + 16);
            if (rc) return rc;
            rc = mcl_scale_by_sum(h, f->tb, f->n, f->d_x8 + 17);
            if (rc) return rc;
        } else {
            rc = mcl_motion_density(h, f->x[f->prev], f->y[f->prev], f->th[f->prev], f->x[f->cur], f->y[f->cur],
                                    f->th[f->cur], f->n, f->delta, f->tf, nullptr, 1);
            if (rc) return rc;
            rc = mcl_motion_density(h, f->x[f->cur], f->y[f->cur], f->th[f->cur], f->x[f->prev], f->y[f->prev],
                                    f->th[f->prev], f->n, f->delta_b, f->tb, nullptr, 1);
            if (rc) return rc;
        }
        rc = mcl_assym_mh_accept_ex(h, f->x[f->prev], f->y[f->prev], f->th[f->prev], f->x[f->cur], f->y[f->cur],
                                    f->th[f->cur], f->w_post, f->w_pre, f->tf, f->tb, f->n, d_uniforms, f->seed,
                                    f->tick, f->first_index, f->x[f->spare], f->y[f->spare], f->th[f->spare],
                                    f->w[f->wslot], nullptr, f->assym == 2);
        if (rc) return rc;
        const int t = f->cur; f->cur = f->spare; f->spare = t;
        return MCL_OK;
    }
    // mh_resampling(particles_prev, particles, weights_post, weights_pre)  (node:363)
    rc = mcl_mh_accept(h, f->x[f->prev], f->y[f->prev], f->th[f->prev], f->x[f->cur], f->y[f->cur], f->th[f->cur],
                       f->w_post, f->w_pre, f->n, d_uniforms, f->seed, f->tick, f->first_index, f->x[f->spare],
                       f->y[f->spare], f->th[f->spare], f->w[f->wslot], nullptr);
    if (rc) return rc;
    const int t = f->cur; f->cur = f->spare; f->spare = t;     // self.particles = mh_particles (node:370)
    return MCL_OK;
}

// ---------------------------------------------------------------------------------------------
// MH refinement chain (BASELINE config 4: k MH iterations per scan, one chain per particle).
// Composition of reference primitives (SURVEY 8(a) "MH-chain"): chain_0 = particles_prev; iteration i
// proposes prop_i = motion(particles_prev, delta) with fresh noise (pu:332-363), weighs proposal and chain
// by two separately normalised softmaxes (node:254-270) and accepts with pu:208-236.  Iteration 1 is exactly
// the reference's single accept (the proposal is the predict() output).  The chain's scores are carried
// (score_chain = accepted ? score_prop : score_chain), which equals re-evaluating compute_likelihoods on
// the chain because the likelihood is a deterministic function of the pose.
// ---------------------------------------------------------------------------------------------
__global__ void k_chain_accept(double *cx, double *cy, double *ct, const double *px, const double *py, const double *pt,
                               const float *__restrict__ w_prop, const float *__restrict__ w_chain,
                               float *score_chain, const float *__restrict__ score_prop, int64_t n, uint64_t seed,
                               uint64_t step, uint64_t first_index, double *ox, double *oy, double *ot, float *w_out) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const float p_old = w_chain[i], p_new = w_prop[i];
        double alpha = 1.0;
        if (p_old > 0.f) {
            const double q = (double)__fdiv_rn(p_new, p_old);
            alpha = (q < 1.0) ? q : 1.0;
        }
        const uint4 o = philox_draw4(seed, step, first_index + (uint64_t)i, 0u, MCL_STREAM_MH);
        const bool acc = u53_from(o.x, o.y) < alpha;
        const double nx = acc ? px[i] : cx[i], ny = acc ? py[i] : cy[i], nt = acc ? pt[i] : ct[i];
        ox[i] = nx; oy[i] = ny; ot[i] = nt;
        w_out[i] = acc ? p_new : p_old;
        if (acc) score_chain[i] = score_prop[i];
    }
}

extern "C" int mcl_filter_update_chain(mcl_handle *h, int iters) {
    FILTER_OR_FAIL("mcl_filter_update_chain");
    if (iters < 1) return mcl_fail(h, MCL_ERR_ARG, "mcl_filter_update_chain: iters < 1");
    if (!f->use_mh || f->assym) return mcl_fail(h, MCL_ERR_STATE, "mcl_filter_update_chain: needs the symmetric MH mode");

    DeviceGuard guard(h->device);
    const int blocks = (int)std::min<int64_t>((f->n + 255) / 256, (int64_t)h->sm_count * 16);
    // roles during the chain: prev = particles_prev (fixed), prop = cur buffer, chain = spare buffer
    const int prev = f->prev, prop = f->cur, chain = f->spare;
    float *score_chain = f->score_pre, *score_prop = f->score_post;
    // single GPU: the likelihood launches leave the score maxima as keys and the iteration is two fused kernels
    // (sum-exp of both sets; weights + accept + max of the carried scores) instead of five launches and a memset
    static int nofuse = -1;
    if (nofuse < 0) { const char *e = getenv("MCL_NO_FUSE"); nofuse = (e && atoi(e)) ? 1 : 0; }
    bool fused = !f->comm && !nofuse;
    int rc;
    // every iteration after its likelihood launch = ONE cooperative kernel (tail.cu k_chain_tail): softmax sums of
    // both score sets, accept, carried score and its maximum; sharded: with the two exchanges inside
    const bool ctail = !nofuse && mcl_tail_available(h, f->n);
    if (ctail) {
        rc = mcl_fused_prepare(h, f->n);
        if (rc) return rc;
        unsigned long long *ck0 = mcl_tail_chain_key(h, f->n, 0);
        if (!ck0) return mcl_fail(h, MCL_ERR_NOMEM, "mcl_filter_update_chain: work area");
        bool ok = false;
        rc = mcl_likelihood_pair(h, f->x[prev], f->y[prev], f->th[prev], score_chain, nullptr, nullptr, nullptr, nullptr, f->n,
                                 ck0, &ok);                                                  // chain_0 = particles_prev
        if (rc) return rc;
        // (a scan without a valid beam launches nothing: !ok, the stand-alone sequence below scores -50, pu:147)
        if (ok && iters > 1) {
            if (f->thr_cap < f->n) {
                MCL_CUDA(h, cudaStreamSynchronize(h->stream));
                cudaFree(f->thr); f->thr = nullptr; f->thr_cap = 0;
                MCL_CUDA(h, cudaMalloc((void **)&f->thr, (size_t)f->n * sizeof(unsigned long long)));
                f->thr_cap = f->n;
            }
            MCL_CUDA(h, cudaMemsetAsync(f->thr, 0xff, (size_t)f->n * sizeof(unsigned long long), h->stream));
        }
        for (int it = 0; ok && it < iters; ++it) {
            if (it > 0) {   // fresh proposal from particles_prev
                f->tick++;
                rc = mcl_predict_cached(h, f->x[prev], f->y[prev], f->th[prev], f->n, f->delta, f->seed, f->tick,
                                        f->first_index, nullptr, 0, f->max_attempts, f->x[prop], f->y[prop], f->th[prop],
                                        nullptr, f->thr);
                if (rc) return rc;
            }
            const int src = it == 0 ? prev : chain;
            bool ok2 = false;
            rc = mcl_likelihood_pair(h, f->x[prop], f->y[prop], f->th[prop], score_prop, nullptr, nullptr, nullptr, nullptr,
                                     f->n, mcl_fused_keymax(h), &ok2);
            if (rc) return rc;
            f->tick++;
            FusedStep u;
            memset(&u, 0, sizeof(u));
            u.n = f->n; u.n_global = f->comm ? f->n_global : f->n; u.use_mh = 1;
            u.s_post = score_prop; u.w_out = f->w[f->wslot];
            u.px = f->x[prop]; u.py = f->y[prop]; u.pt = f->th[prop];
            u.ox = f->x[src]; u.oy = f->y[src]; u.ot = f->th[src];
            u.nx = f->x[chain]; u.ny = f->y[chain]; u.nth = f->th[chain];
            u.seed = f->seed; u.step = f->tick; u.first_index = f->first_index;
            TailComm tc;
            if (f->comm) {
                tc.rank = f->rank; tc.world = f->world; tc.n_global = f->n_global; tc.mailbox = f->mailbox;
                for (int d = 0; d < 16; ++d) tc.peers[d] = f->peer_mailbox[d];
                tc.epoch0 = f->epoch; tc.d_err = f->d_comm_err; tc.d_peer_pose_dst = nullptr;
                f->epoch += TAIL_CHAIN_EXCHANGES;
            }
            rc = mcl_tail_chain_iteration(h, u, mcl_fused_keymax(h), score_chain, it, f->comm ? &tc : nullptr);
            if (rc) return rc;
        }
        if (ok) {
            rc = mcl_tail_chain_finish(h);
            if (rc) return rc;
            f->cur = chain; f->spare = prop;              // self.particles = chain
            return MCL_OK;
        }
        fused = false;
    }
    if (fused) {
        rc = mcl_fused_prepare(h, f->n);
        if (rc) return rc;
        bool ok = false;
        rc = mcl_likelihood_pair(h, f->x[prev], f->y[prev], f->th[prev], score_chain, nullptr, nullptr, nullptr, nullptr, f->n,
                                 mcl_fused_keymax(h) + 1, &ok);                                // chain_0 = particles_prev
        if (rc) return rc;
        if (!ok) fused = false;      // no valid beam in the scan: nothing was launched
    }
    if (!fused) {
        rc = mcl_likelihood(h, f->x[prev], f->y[prev], f->th[prev], f->n, score_chain);       // chain_0 = particles_prev
        if (rc) return rc;
    }
    // every further proposal starts from particles_prev with the same increment: the geometric screening
    // threshold of a particle (motion.cu) is computed once per scan and reused by the other iterations
    if (iters > 1) {
        if (f->thr_cap < f->n) {
            MCL_CUDA(h, cudaStreamSynchronize(h->stream));
            cudaFree(f->thr); f->thr = nullptr; f->thr_cap = 0;
            MCL_CUDA(h, cudaMalloc((void **)&f->thr, (size_t)f->n * sizeof(unsigned long long)));
            f->thr_cap = f->n;
        }
        MCL_CUDA(h, cudaMemsetAsync(f->thr, 0xff, (size_t)f->n * sizeof(unsigned long long), h->stream));
    }
    for (int it = 0; it < iters; ++it) {
        if (it > 0) {   // fresh proposal from particles_prev
            f->tick++;
            rc = mcl_predict_cached(h, f->x[prev], f->y[prev], f->th[prev], f->n, f->delta, f->seed, f->tick,
                                    f->first_index, nullptr, 0, f->max_attempts, f->x[prop], f->y[prop], f->th[prop],
                                    nullptr, f->thr);
            if (rc) return rc;
        }
        const int src = it == 0 ? prev : chain;       // iteration 1 reads particles_prev and writes the chain buffer
        if (fused) {
            bool ok = false;
            rc = mcl_likelihood_pair(h, f->x[prop], f->y[prop], f->th[prop], score_prop, nullptr, nullptr, nullptr, nullptr,
                                     f->n, mcl_fused_keymax(h), &ok);
            if (rc) return rc;
            f->tick++;
            FusedStep u;
            memset(&u, 0, sizeof(u));
            u.n = f->n; u.n_global = f->n; u.use_mh = 1;
            u.s_post = score_prop; u.s_pre = score_chain; u.w_post = f->w_post; u.w_pre = f->w_pre; u.w_out = f->w[f->wslot];
            u.px = f->x[prop]; u.py = f->y[prop]; u.pt = f->th[prop];
            u.ox = f->x[src]; u.oy = f->y[src]; u.ot = f->th[src];
            u.nx = f->x[chain]; u.ny = f->y[chain]; u.nth = f->th[chain];
            u.seed = f->seed; u.step = f->tick; u.first_index = f->first_index; u.est18 = h->d_est18;
            rc = mcl_fused_sumexp(h, u);
            if (rc) return rc;
            rc = mcl_fused_chain_accept(h, u, score_chain);
            if (rc) return rc;
            continue;
        }
        rc = mcl_likelihood(h, f->x[prop], f->y[prop], f->th[prop], f->n, score_prop);
        if (rc) return rc;
        rc = f->comm ? sharded_softmax_arrays(h, f, score_prop, f->w_post, score_chain, f->w_pre)
                     : mcl_softmax_pair(h, score_prop, f->w_post, score_chain, f->w_pre, f->n);
        if (rc) return rc;
        f->tick++;
        k_chain_accept<<<blocks, 256, 0, h->stream>>>(f->x[src], f->y[src], f->th[src], f->x[prop], f->y[prop], f->th[prop],
                                                      f->w_post, f->w_pre, score_chain, score_prop, f->n, f->seed, f->tick,
                                                      f->first_index, f->x[chain], f->y[chain], f->th[chain], f->w[f->wslot]);
        MCL_LAUNCH_CHECK(h);
    }
    // the last iteration left the key of the carried scores behind: the next user of the keys starts from zero
    if (fused) MCL_CUDA(h, cudaMemsetAsync(mcl_fused_keymax(h), 0, 2 * sizeof(unsigned long long), h->stream));
    f->cur = chain; f->spare = prop;                  // self.particles = chain
    return MCL_OK;
}

extern "C" int mcl_filter_estimate(mcl_handle *h, double *d_out18, double h_out16[16]) {
    FILTER_OR_FAIL("mcl_filter_estimate");
    if (f->comm) {
        // raw sums -> rank-ordered sum over peers -> means -> central sums -> rank-ordered sum
        DeviceGuard guard(h->device);
        double *o = d_out18 ? d_out18 : h->d_est18;
        int rc = mcl_estimate_moments_async(h, f->x[f->cur], f->y[f->cur], f->th[f->cur], f->w[f->wslot], f->n, f->d_x8);
        if (rc) return rc;
        rc = comm_exchange(h, f, f->d_x8, 6, XCH_SUM_F64, o);
        if (rc) return rc;
        rc = mcl_estimate_means_async(h, o);
        if (rc) return rc;
        rc = mcl_estimate_central_async(h, f->x[f->cur], f->y[f->cur], f->th[f->cur], f->w[f->wslot], f->n, o + 6, f->d_x8);
        if (rc) return rc;
        rc = comm_exchange(h, f, f->d_x8, 9, XCH_SUM_F64, o + 9);
        if (rc) return rc;
        if (h_out16) {
            MCL_CUDA(h, cudaMemcpyAsync(h->h_pinned, o, 18 * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
            MCL_CUDA(h, cudaStreamSynchronize(h->stream));
            const double *r = h->h_pinned;
            h_out16[0] = r[0]; h_out16[1] = r[1]; h_out16[2] = r[6]; h_out16[3] = r[7]; h_out16[4] = r[8];
            for (int k = 0; k < 9; ++k) h_out16[5 + k] = r[9 + k];
            h_out16[14] = 0; h_out16[15] = 0;
            return comm_check(h, f);
        }
        return MCL_OK;
    }
    if (h_out16) return mcl_estimate(h, f->x[f->cur], f->y[f->cur], f->th[f->cur], f->w[f->wslot], f->n, h_out16);
    if (!d_out18) return mcl_fail(h, MCL_ERR_ARG, "mcl_filter_estimate: no output");
    return mcl_estimate_async(h, f->x[f->cur], f->y[f->cur], f->th[f->cur], f->w[f->wslot], f->n, d_out18);
}

// node:488-492 resample_lvr; r < 0: draw r = U(0, 1/N) from Philox(seed, tick)
extern "C" int mcl_filter_resample(mcl_handle *h, double r) {
    FILTER_OR_FAIL("mcl_filter_resample");
    f->tick++;
    if (f->comm && f->resample_mode != MCL_RESAMPLE_FIXED_POINT)
        return mcl_fail(h, MCL_ERR_STATE, "mcl_filter_resample: on sharded particles the reference's resampling arithmetic "
                                          "exists only inside the persistent step (mcl_filter_step, MCL / MHMCL modes)");
    if (f->comm) {
        // global systematic resampling (fixed-point arithmetic): scale from the global weight maximum, local
        // cumulative sums, all-gather of the totals, offspring pushed into the destination ranks' spare set
        DeviceGuard guard(h->device);
        if (r < 0) r = mcl_resample_offset(f->seed, f->tick, f->n_global);
        float *wmax = (float *)(f->d_x8 + 56);
        int rc = mcl_weights_max(h, f->w[f->wslot], f->n, wmax);
        if (rc) return rc;
        k_f32_to_f64<<<1, 1, 0, h->stream>>>(wmax, f->d_x8);
        MCL_LAUNCH_CHECK(h);
        rc = comm_exchange(h, f, f->d_x8, 1, XCH_MAX_F64, f->d_x8 + 16);
        if (rc) return rc;
        k_f64_to_f32<<<1, 1, 0, h->stream>>>(f->d_x8 + 16, wmax);
        MCL_LAUNCH_CHECK(h);
        rc = mcl_resample_scan(h, f->w[f->wslot], f->n, wmax, f->n_global, (uint64_t *)f->d_x8);
        if (rc) return rc;
        rc = comm_exchange(h, f, f->d_x8, 1, XCH_GATHER, f->d_x8 + 16);          // totals of every rank
        if (rc) return rc;
        rc = mcl_resample_push(h, f->n, (const uint64_t *)(f->d_x8 + 16), f->rank, f->world, r, f->n_global, f->n,
                               f->x[f->cur], f->y[f->cur], f->th[f->cur], f->d_peer_pose + (size_t)f->spare * 3 * f->world);
        if (rc) return rc;
        rc = comm_exchange(h, f, f->d_x8, 0, XCH_GATHER, f->d_x8 + 16);          // barrier: all pushes have landed
        if (rc) return rc;
        const int t = f->cur; f->cur = f->spare; f->spare = t;
        return MCL_OK;
    }
    if (r < 0) r = mcl_resample_offset(f->seed, f->tick, f->n);
    int rc = mcl_resample_indices(h, f->w[f->wslot], f->n, f->n, r, f->resample_mode, f->idx);
    if (rc) return rc;
    rc = mcl_gather(h, f->x[f->cur], f->y[f->cur], f->th[f->cur], f->idx, f->n, f->x[f->spare], f->y[f->spare],
                    f->th[f->spare]);
    if (rc) return rc;
    const int t = f->cur; f->cur = f->spare; f->spare = t;
    // self.weights keeps its pre-resampling values (node:490 discards the returned uniform weights)
    return MCL_OK;
}

// node:586-597 estimate, then node:488-492 resample_lvr, of the particles and weights as they are (after
// mcl_filter_update / mcl_filter_update_chain): one launch of the tail kernel without its softmax / accept stages on
// (sharded: with the exchanges and the peer push inside, in either arithmetic); else the two calls one after the
// other.  Same results as mcl_filter_estimate + mcl_filter_resample(-1).
extern "C" int mcl_filter_finish(mcl_handle *h, double *d_out18, double h_out16[16]) {
    FILTER_OR_FAIL("mcl_filter_finish");
    const bool arith_ok = f->resample_mode == MCL_RESAMPLE_FIXED_POINT || f->resample_mode == MCL_RESAMPLE_REFERENCE_F32;
    if (!arith_ok || !mcl_tail_available(h, f->n)) {
        if (d_out18 || h_out16) { const int rc = mcl_filter_estimate(h, d_out18, h_out16); if (rc) return rc; }
        return mcl_filter_resample(h, -1.0);
    }
    DeviceGuard guard(h->device);
    int rc = mcl_fused_prepare(h, f->n);             // (the key words the kernel reads at its start live there)
    if (rc) return rc;
    f->tick++;                                       // node:488-492 resample_lvr draws r
    const double r = mcl_resample_offset(f->seed, f->tick, f->comm ? f->n_global : f->n);
    double *est = h->d_est18;
    const int dst = f->spare;
    TailComm tc;
    if (f->comm) {                                   // sharded: the exchanges and the peer push run inside the kernel
        tc.rank = f->rank; tc.world = f->world; tc.n_global = f->n_global; tc.mailbox = f->mailbox;
        for (int d = 0; d < 16; ++d) tc.peers[d] = f->peer_mailbox[d];
        tc.epoch0 = f->epoch; tc.d_err = f->d_comm_err;
        tc.d_peer_pose_dst = (const unsigned long long *)(f->d_peer_pose + (size_t)dst * 3 * f->world);
        f->epoch += f->resample_mode == MCL_RESAMPLE_REFERENCE_F32 ? TAIL_EXCHANGES_REF : TAIL_EXCHANGES;
    }
    rc = mcl_tail_finish(h, f->n, f->comm ? f->n_global : f->n, f->w[f->wslot], f->x[f->cur], f->y[f->cur], f->th[f->cur], est,
                         f->resample_mode, r, f->idx, f->x[dst], f->y[dst], f->th[dst], f->comm ? &tc : nullptr);
    if (rc) return rc;
    const int t = f->cur; f->cur = f->spare; f->spare = t;
    if (d_out18) MCL_CUDA(h, cudaMemcpyAsync(d_out18, est, 18 * sizeof(double), cudaMemcpyDeviceToDevice, h->stream));
    if (h_out16) {
        MCL_CUDA(h, cudaMemcpyAsync(h->h_pinned, est, 18 * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
        MCL_CUDA(h, cudaMemcpyAsync(h->h_pinned + 20, mcl_tail_err_ptr(h), sizeof(int), cudaMemcpyDeviceToHost, h->stream));
        MCL_CUDA(h, cudaStreamSynchronize(h->stream));
        int terr = 0;
        memcpy(&terr, h->h_pinned + 20, sizeof(int));
        if (terr) return mcl_fail(h, MCL_ERR_CUDA, "step tail: a grid barrier / look-back wait timed out (mcl_tail_status)");
        if (f->comm) { rc = comm_check(h, f); if (rc) return rc; }
        const double *o = h->h_pinned;
        h_out16[0] = o[0]; h_out16[1] = o[1]; h_out16[2] = o[6]; h_out16[3] = o[7]; h_out16[4] = o[8];
        for (int k = 0; k < 9; ++k) h_out16[5 + k] = o[9 + k];
        h_out16[14] = 0; h_out16[15] = 0;
    }
    return MCL_OK;
}

// update -> estimate -> resample through the fused kernels (fused.cu) when the configuration allows it:
// symmetric MH or plain MCL, fixed-point resampling.  Sharded (f->comm): the same four kernels with the peer-memory exchanges between them (score maxima,
// softmax sums, raw estimate sums + weight maximum, central sums, totals) and the peer-push gather.
// *done = false: nothing was enqueued, the caller runs the stand-alone sequence.  MCL_NO_FUSE=1 disables it (A/B).
static int fused_tail(mcl_handle *h, FilterState *f, double *d_out18, double h_out16[16], bool *done) {
    *done = false;
    static int off = -1;
    if (off < 0) { const char *e = getenv("MCL_NO_FUSE"); off = (e && atoi(e)) ? 1 : 0; }
    if (off || f->assym) return MCL_OK;
    // single GPU: the whole tail is ONE persistent cooperative kernel (tail.cu), in either resampling arithmetic;
    // sharded: four kernels with the peer-memory exchanges between them (fixed-point arithmetic only)
    const bool tail = mcl_tail_available(h, f->n);
    if (!tail && f->resample_mode != MCL_RESAMPLE_FIXED_POINT) return MCL_OK;
    DeviceGuard guard(h->device);
    int rc = mcl_fused_prepare(h, f->n);
    if (rc) return rc;
    bool g1 = false;
    const int cur = f->cur, prev = f->prev, spare = f->spare;
    if (f->use_mh)
        rc = mcl_likelihood_pair(h, f->x[cur], f->y[cur], f->th[cur], f->score_post, f->x[prev], f->y[prev], f->th[prev],
                                 f->score_pre, f->n, mcl_fused_keymax(h), &g1);
    else
        rc = mcl_likelihood_pair(h, f->x[cur], f->y[cur], f->th[cur], f->score_post, nullptr, nullptr, nullptr, nullptr,
                                 f->n, mcl_fused_keymax(h), &g1);
    if (rc || !g1) return rc;
    *done = true;
    FusedPtrs xp;
    mcl_fused_exchange_ptrs(h, &xp);
    double *est = h->d_est18;
    FusedStep u;
    memset(&u, 0, sizeof(u));
    u.n = f->n; u.n_global = f->comm ? f->n_global : f->n; u.use_mh = f->use_mh;
    u.s_post = f->score_post; u.w_out = f->w[f->wslot];
    u.px = f->x[cur]; u.py = f->y[cur]; u.pt = f->th[cur];
    u.seed = f->seed; u.first_index = f->first_index; u.est18 = est;
    int res;                                   // the set that holds the particles after the MH step
    if (f->use_mh) {
        f->tick++;
        // mh_resampling(particles_prev, particles, weights_post, weights_pre)  (node:363) -> spare set
        u.s_pre = f->score_pre; u.w_post = f->w_post; u.w_pre = f->w_pre;
        u.ox = f->x[prev]; u.oy = f->y[prev]; u.ot = f->th[prev];
        u.nx = f->x[spare]; u.ny = f->y[spare]; u.nth = f->th[spare];
        res = spare;
    } else {
        u.nx = f->x[cur]; u.ny = f->y[cur]; u.nth = f->th[cur];
        res = cur;
    }
    u.step = f->tick;
    if (tail) {
        f->tick++;                                   // node:488-492 resample_lvr draws r
        const double r = mcl_resample_offset(f->seed, f->tick, f->comm ? f->n_global : f->n);
        const int dst = f->use_mh ? cur : spare;     // MH: the proposal set is free once the accept has read it
        TailComm tc;
        if (f->comm) {
            tc.rank = f->rank; tc.world = f->world; tc.n_global = f->n_global; tc.mailbox = f->mailbox;
            for (int d = 0; d < 16; ++d) tc.peers[d] = f->peer_mailbox[d];
            tc.epoch0 = f->epoch; tc.d_err = f->d_comm_err;
            tc.d_peer_pose_dst = (const unsigned long long *)(f->d_peer_pose + (size_t)dst * 3 * f->world);
            f->epoch += f->resample_mode == MCL_RESAMPLE_REFERENCE_F32 ? TAIL_EXCHANGES_REF : TAIL_EXCHANGES;
        }
        rc = mcl_tail_step(h, u, mcl_fused_keymax(h), f->resample_mode, r, f->idx, f->x[dst], f->y[dst], f->th[dst],
                           f->comm ? &tc : nullptr);
        if (rc) return rc;
        // roles: particles = resampled set; the MH result (or, without MH, the old particles) becomes spare
        f->cur = dst; f->spare = res;
        if (d_out18) MCL_CUDA(h, cudaMemcpyAsync(d_out18, est, 18 * sizeof(double), cudaMemcpyDeviceToDevice, h->stream));
        if (h_out16) {
            MCL_CUDA(h, cudaMemcpyAsync(h->h_pinned, est, 18 * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
            MCL_CUDA(h, cudaMemcpyAsync(h->h_pinned + 20, mcl_tail_err_ptr(h), sizeof(int), cudaMemcpyDeviceToHost, h->stream));
            MCL_CUDA(h, cudaStreamSynchronize(h->stream));
            int terr = 0;
            memcpy(&terr, h->h_pinned + 20, sizeof(int));
            if (terr) return mcl_fail(h, MCL_ERR_CUDA, "step tail: a grid barrier / look-back wait timed out (mcl_tail_status)");
            if (f->comm) { rc = comm_check(h, f); if (rc) return rc; }
            const double *o = h->h_pinned;
            h_out16[0] = o[0]; h_out16[1] = o[1]; h_out16[2] = o[6]; h_out16[3] = o[7]; h_out16[4] = o[8];
            for (int k = 0; k < 9; ++k) h_out16[5 + k] = o[9 + k];
            h_out16[14] = 0; h_out16[15] = 0;
        }
        return MCL_OK;
    }
    if (f->comm) { rc = comm_exchange(h, f, xp.keymax, 2, XCH_MAX_U64, xp.keymax); if (rc) return rc; }
    rc = mcl_fused_sumexp(h, u);
    if (rc) return rc;
    if (f->comm) { rc = comm_exchange(h, f, xp.sumq, 2, XCH_SUM_U64, xp.sumq); if (rc) return rc; }
    rc = mcl_fused_weights(h, u);
    if (rc) return rc;
    if (f->use_mh) { f->cur = spare; f->spare = cur; }      // self.particles = mh_particles (node:370)
    if (f->comm) { rc = comm_exchange(h, f, xp.msum, 7, XCH_SUM_F64_MAXLAST, xp.msum); if (rc) return rc; }
    rc = mcl_fused_scan(h, u);
    if (rc) return rc;
    // central sums -> est[9..17]; the tenth value is this rank's total of quantised weights -> totals of every rank
    if (f->comm) { rc = comm_exchange(h, f, xp.csum, 10, XCH_SUM_F64_GATHERLAST, est + 9, f->d_x8 + 16); if (rc) return rc; }
    if (h_out16) {
        MCL_CUDA(h, cudaMemcpyAsync(h->h_pinned, est, 18 * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
        MCL_CUDA(h, cudaEventRecord(h->ev_est, h->stream));
    }
    if (d_out18) MCL_CUDA(h, cudaMemcpyAsync(d_out18, est, 18 * sizeof(double), cudaMemcpyDeviceToDevice, h->stream));
    // node:488-492 resample_lvr: gather from the MH result into the free set
    f->tick++;
    const int dst = f->spare;
    if (f->comm) {
        const double r = mcl_resample_offset(f->seed, f->tick, f->n_global);
        int pnt = 0, ptile = 0;
        const unsigned long long *tprefix = mcl_fused_tile_prefix(h, f->n, &pnt, &ptile);
        rc = mcl_resample_push_from(h, mcl_fused_cumsum(h, f->n), tprefix, pnt, ptile, f->n, (const uint64_t *)(f->d_x8 + 16), f->rank, f->world,
                                    r, f->n_global, f->n, f->x[res], f->y[res], f->th[res],
                                    f->d_peer_pose + (size_t)dst * 3 * f->world);
        if (rc) return rc;
        rc = comm_exchange(h, f, f->d_x8, 0, XCH_GATHER, f->d_x8 + 16);          // barrier: all pushes have landed
        if (rc) return rc;
    } else {
        const double r = mcl_resample_offset(f->seed, f->tick, f->n);
        rc = mcl_fused_resample(h, f->n, r, f->x[res], f->y[res], f->th[res], f->idx, f->x[dst], f->y[dst], f->th[dst]);
        if (rc) return rc;
    }
    f->cur = dst; f->spare = res;
    if (h_out16) {
        MCL_CUDA(h, cudaEventSynchronize(h->ev_est));
        const double *o = h->h_pinned;
        h_out16[0] = o[0]; h_out16[1] = o[1]; h_out16[2] = o[6]; h_out16[3] = o[7]; h_out16[4] = o[8];
        for (int k = 0; k < 9; ++k) h_out16[5 + k] = o[9 + k];
        h_out16[14] = 0; h_out16[15] = 0;
        return comm_check(h, f);
    }
    return MCL_OK;
}

// one odom message followed by one scan: predict -> update -> estimate -> resample
extern "C" int mcl_filter_step(mcl_handle *h, const double delta[3], int scan_slot, double *d_out18,
                               double h_out16[16]) {
    FILTER_OR_FAIL("mcl_filter_step");
    int rc;
    if (delta) { rc = mcl_filter_predict(h, delta, nullptr, 0); if (rc) return rc; }
    if (scan_slot >= 0) { rc = mcl_use_scan(h, scan_slot); if (rc) return rc; }
    {
        bool done = false;
        rc = fused_tail(h, f, d_out18, h_out16, &done);
        if (rc || done) return rc;
    }
    rc = mcl_filter_update(h, nullptr);
    if (rc) return rc;
    if (!h_out16) {
        if (d_out18) { rc = mcl_filter_estimate(h, d_out18, nullptr); if (rc) return rc; }
        return mcl_filter_resample(h, -1.0);
    }
    // host estimate wanted: enqueue the estimate and its read-back, then the resampling kernels, and
    // wait only for the read-back -- the resampling overlaps the host's handling of the estimate.
    DeviceGuard guard(h->device);
    rc = mcl_filter_estimate(h, h->d_est18, nullptr);
    if (rc) return rc;
    MCL_CUDA(h, cudaMemcpyAsync(h->h_pinned, h->d_est18, 18 * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    MCL_CUDA(h, cudaEventRecord(h->ev_est, h->stream));
    if (d_out18) MCL_CUDA(h, cudaMemcpyAsync(d_out18, h->d_est18, 18 * sizeof(double), cudaMemcpyDeviceToDevice, h->stream));
    rc = mcl_filter_resample(h, -1.0);
    if (rc) return rc;
    MCL_CUDA(h, cudaEventSynchronize(h->ev_est));
    const double *r = h->h_pinned;
    h_out16[0] = r[0]; h_out16[1] = r[1]; h_out16[2] = r[6]; h_out16[3] = r[7]; h_out16[4] = r[8];
    for (int k = 0; k < 9; ++k) h_out16[5 + k] = r[9 + k];
    h_out16[14] = 0; h_out16[15] = 0;
    return MCL_OK;
}
