// mcl_core.cu -- handle lifetime, configuration (map / sensor / scan), likelihood-table build.
#include <math.h>
#include <string.h>

#include <algorithm>
#include <vector>

#include <atomic>

#include "common.cuh"

thread_local std::string g_create_err;

int mcl_fail(mcl_handle *h, int code, const std::string &msg) {
    if (h) h->err = msg; else g_create_err = msg;
    return code;
}

extern "C" const char *mcl_version(void) { return "mcl-b200 0.1 (sm_100a)"; }

extern "C" const char *mcl_last_error(const mcl_handle *h) {
    return h ? h->err.c_str() : g_create_err.c_str();
}

extern "C" int mcl_create(mcl_handle **out, int device) {
    if (!out) return mcl_fail(nullptr, MCL_ERR_ARG, "mcl_create: out is NULL");
    *out = nullptr;
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0)
        return mcl_fail(nullptr, MCL_ERR_CUDA,
                        std::string("mcl_create: no CUDA device (") + cudaGetErrorString(e) +
                            "); libmcl has no CPU fallback");
    if (device < 0 || device >= count) return mcl_fail(nullptr, MCL_ERR_ARG, "mcl_create: bad device index");
    DeviceGuard g(device);
    cudaDeviceProp prop;
    e = cudaGetDeviceProperties(&prop, device);
    if (e != cudaSuccess) return mcl_fail(nullptr, MCL_ERR_CUDA, cudaGetErrorString(e));
    if (prop.major < 10)
        return mcl_fail(nullptr, MCL_ERR_CUDA,
                        "mcl_create: device is sm_" + std::to_string(prop.major * 10 + prop.minor) +
                            ", this library is built for sm_100a only");
    mcl_handle *h = new mcl_handle();
    h->device = device;
    h->sm_count = prop.multiProcessorCount;
    h->smem_optin = (int)prop.sharedMemPerBlockOptin;
    h->cc_major = prop.major;
    h->cc_minor = prop.minor;
    if (cudaMallocHost((void **)&h->h_pinned, 64 * sizeof(double)) != cudaSuccess) {
        delete h;
        return mcl_fail(nullptr, MCL_ERR_NOMEM, "mcl_create: cudaMallocHost failed");
    }
    if (cudaMalloc((void **)&h->d_est18, 32 * sizeof(double)) != cudaSuccess ||
        cudaEventCreateWithFlags(&h->ev_est, cudaEventDisableTiming) != cudaSuccess) {
        cudaFreeHost(h->h_pinned);
        delete h;
        return mcl_fail(nullptr, MCL_ERR_NOMEM, "mcl_create: device allocation failed");
    }
    *out = h;
    return MCL_OK;
}

extern "C" int mcl_destroy(mcl_handle *h) {
    if (!h) return MCL_OK;
    DeviceGuard g(h->device);
    cudaStreamSynchronize(h->stream);
    mcl_filter_forget(h);
    mcl_raycast_forget(h);
    cudaFree(h->d_est18); cudaFree(h->d_fused); cudaFree(h->d_code8); cudaFree(h->d_code8p); cudaFree(h->d_tiled);
    cudaFree(h->d_kld);
    cudaFree(h->d_seq);
    cudaFree(h->d_tail); cudaFree(h->d_tail_prof); cudaFree(h->d_motion_stats); cudaFree(h->d_retry_idx); cudaFree(h->d_retry_thr); cudaFree(h->d_retry_ctr); cudaFree(h->d_win_skew);
    if (h->ev_est) cudaEventDestroy(h->ev_est);
    if (h->ev_beams) cudaEventDestroy(h->ev_beams);
    cudaFree(h->d_occ); cudaFree(h->d_dist); cudaFree(h->d_logtab); cudaFree(h->d_win);
    cudaFree(h->d_beams); cudaFree(h->d_batch); cudaFree(h->d_scratch); cudaFree(h->d_win8); cudaFree(h->d_lut);
    cudaFreeHost(h->h_beams); cudaFreeHost(h->h_pinned);
    for (auto &p : h->lik_events) { cudaEventDestroy(p.first); cudaEventDestroy(p.second); }
    delete h;
    return MCL_OK;
}

// process-wide unique id of "what the active beam table holds" (likelihood.cu keys its constant-bank copy on it)
uint64_t mcl_next_scan_uid() {
    static std::atomic<uint64_t> next{1};
    return next.fetch_add(1);
}

extern "C" int mcl_set_stream(mcl_handle *h, void *s) {
    if (!h) return MCL_ERR_ARG;
    h->stream = (cudaStream_t)s;
    return MCL_OK;
}

extern "C" int mcl_sync(mcl_handle *h) {
    if (!h) return MCL_ERR_ARG;
    DeviceGuard g(h->device);
    MCL_CUDA(h, cudaStreamSynchronize(h->stream));
    return MCL_OK;
}

extern "C" int mcl_device_info(mcl_handle *h, int *sm, int *smem, int *maj, int *min) {
    if (!h) return MCL_ERR_ARG;
    if (sm) *sm = h->sm_count;
    if (smem) *smem = h->smem_optin;
    if (maj) *maj = h->cc_major;
    if (min) *min = h->cc_minor;
    return MCL_OK;
}

extern "C" int64_t mcl_launch_count(const mcl_handle *h) { return h ? h->launches : 0; }

int mcl_ensure_scratch(mcl_handle *h, size_t bytes) {
    if (bytes <= h->scratch_bytes) return MCL_OK;
    MCL_CUDA(h, cudaStreamSynchronize(h->stream));
    cudaFree(h->d_scratch);
    h->d_scratch = nullptr;
    h->scratch_bytes = 0;
    size_t want = std::max(bytes, (size_t)1 << 20);
    want = (want + 255) & ~(size_t)255;
    MCL_CUDA(h, cudaMalloc(&h->d_scratch, want));
    h->scratch_bytes = want;
    return MCL_OK;
}

// ------------------------------------------------------------------------------------------
// configuration
// ------------------------------------------------------------------------------------------
extern "C" int mcl_set_map(mcl_handle *h, const int8_t *h_occ, const float *h_dist, int W, int H,
                           double res, double ox, double oy) {
    if (!h) return MCL_ERR_ARG;
    if ((!h_occ && !h_dist) || W <= 0 || H <= 0 || !(res > 0))
        return mcl_fail(h, MCL_ERR_ARG, "mcl_set_map: bad argument");
    if ((int64_t)W * H > ((int64_t)1 << 31) - 1) return mcl_fail(h, MCL_ERR_CAPACITY, "mcl_set_map: map too large");
    DeviceGuard g(h->device);
    MCL_CUDA(h, cudaStreamSynchronize(h->stream));
    cudaFree(h->d_occ); cudaFree(h->d_dist); cudaFree(h->d_logtab); cudaFree(h->d_win);
    h->d_occ = nullptr; h->d_dist = nullptr; h->d_logtab = nullptr; h->d_win = nullptr;
    const size_t cells = (size_t)W * H;
    h->W = W; h->H = H; h->res = res; h->ox = ox; h->oy = oy;
    h->scan_set = false;
    if (h_occ) {   // occupancy only needed by predict / init (pu:388-396)
        MCL_CUDA(h, cudaMalloc((void **)&h->d_occ, cells));
        MCL_CUDA(h, cudaMemcpy(h->d_occ, h_occ, cells, cudaMemcpyHostToDevice));
    }
    if (!h_dist) { h->wx0 = h->wy0 = h->ww = h->wh = 0; h->win_bytes = 0; h->win_ok = false; return MCL_OK; }
    MCL_CUDA(h, cudaMalloc((void **)&h->d_dist, cells * sizeof(float)));
    MCL_CUDA(h, cudaMalloc((void **)&h->d_logtab, cells * sizeof(int32_t)));
    MCL_CUDA(h, cudaMemcpy(h->d_dist, h_dist, cells * sizeof(float), cudaMemcpyHostToDevice));
    // bounding box of cells with dist > 0 (free space): everywhere else the table is the constant c0
    int x0 = W, x1 = -1, y0 = H, y1 = -1;
    for (int y = 0; y < H; ++y) {
        const float *row = h_dist + (size_t)y * W;
        int rx0 = -1, rx1 = -1;
        for (int x = 0; x < W; ++x) if (row[x] != 0.0f) { rx0 = x; break; }
        if (rx0 < 0) continue;
        for (int x = W - 1; x >= 0; --x) if (row[x] != 0.0f) { rx1 = x; break; }
        x0 = std::min(x0, rx0); x1 = std::max(x1, rx1);
        y0 = std::min(y0, y); y1 = std::max(y1, y);
    }
    if (x1 < 0) { h->wx0 = 0; h->wy0 = 0; h->ww = 0; h->wh = 0; }
    else { h->wx0 = x0; h->wy0 = y0; h->ww = x1 - x0 + 1; h->wh = y1 - y0 + 1; }
    h->win_ok = false; h->win_bytes = 0;      // the window is laid out by mcl_prepare_table (needs the beam reach)
    h->tab_dirty = true;
    return MCL_OK;
}

extern "C" int mcl_set_sensor(mcl_handle *h, double sigma_hit, double z_hit, double z_rand,
                              double max_range, int step) {
    if (!h) return MCL_ERR_ARG;
    if (!(sigma_hit > 0) || !(max_range > 0) || step < 1)
        return mcl_fail(h, MCL_ERR_ARG, "mcl_set_sensor: need sigma_hit > 0, max_range > 0, step >= 1");
    h->sigma_hit = sigma_hit; h->z_hit = z_hit; h->z_rand = z_rand; h->max_range = max_range;
    h->step = step;
    h->sensor_set = true;
    h->tab_dirty = true;
    h->scan_set = false;   // validity of beams depends on max_range / step
    return MCL_OK;
}

extern "C" int mcl_set_motion(mcl_handle *h, const float alpha[4]) {
    if (!h || !alpha) return MCL_ERR_ARG;
    memcpy(h->alpha, alpha, sizeof(h->alpha));
    h->motion_set = true;
    return MCL_OK;
}

extern "C" int mcl_set_likelihood_path(mcl_handle *h, int path) {
    if (!h || path < 0 || path > 2) return MCL_ERR_ARG;
    h->lik_path = path;
    return MCL_OK;
}

__global__ void k_build_logtab(const float *__restrict__ dist, int32_t *__restrict__ logtab, int64_t cells,
                               double sigma_hit, double z_hit, double z_rand, double max_range) {
    for (int64_t c = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; c < cells;
         c += (int64_t)gridDim.x * blockDim.x)
        logtab[c] = quantise_logp(cell_logp(dist[c], sigma_hit, z_hit, z_rand, max_range, true));
}

// window cell (ix, iy) <-> map cell (ix + ofx, iy + ofy); see mcl_handle::win_edge for the EDGE sides
__global__ void k_pack_window(const int32_t *__restrict__ logtab, int32_t *__restrict__ win, int W, int H, int ofx, int ofy,
                              int edge, int rows, int tpose, int32_t c0, int32_t voff, int skew) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < rows * 256; i += gridDim.x * blockDim.x) {
        const int major = i >> 8, minor = i & 255;
        const int ix = tpose ? major : minor, iy = tpose ? minor : major;
        int mx = ix + ofx, my = iy + ofy;
        const bool outside = ((edge & 1) && mx <= -2) || ((edge & 2) && mx >= W) || ((edge & 4) && my <= -2) ||
                             ((edge & 8) && my >= H);
        if ((edge & 1) && mx == -1) mx = 0;          // int() sends (-1, 0) to cell 0
        if ((edge & 4) && my == -1) my = 0;
        int32_t v = c0;                               // in-map cells outside the free-space box
        if (outside) v = 0;                           // beyond the map: the beam is skipped (pu:131-132), adds 0
        else if (mx >= 0 && mx < W && my >= 0 && my < H) v = logtab[(size_t)my * W + mx];
        win[skew ? i + (i >> 5) : i] = v - voff;     // >= 0: see mcl_handle::voff; skew: rows of 264 words
    }
}

__global__ void k_c0(int32_t *out, double sigma_hit, double z_hit, double z_rand, double max_range) {
    out[0] = quantise_logp(cell_logp(0.0f, sigma_hit, z_hit, z_rand, max_range, true));
}

int mcl_prepare_table(mcl_handle *h) {
    if (!h->d_dist) return mcl_fail(h, MCL_ERR_STATE, "map not set (mcl_set_map)");
    if (!h->sensor_set) return mcl_fail(h, MCL_ERR_STATE, "sensor parameters not set (mcl_set_sensor)");
    if (!h->tab_dirty) return MCL_OK;
    const int64_t cells = (int64_t)h->W * h->H;
    int rc = mcl_ensure_scratch(h, 256);
    if (rc) return rc;
    const int blocks = (int)std::min<int64_t>((cells + 255) / 256, (int64_t)h->sm_count * 16);
    k_build_logtab<<<blocks, 256, 0, h->stream>>>(h->d_dist, h->d_logtab, cells, h->sigma_hit, h->z_hit,
                                                  h->z_rand, h->max_range);
    MCL_LAUNCH_CHECK(h);
    k_c0<<<1, 1, 0, h->stream>>>((int32_t *)h->d_scratch, h->sigma_hit, h->z_hit, h->z_rand, h->max_range);
    MCL_LAUNCH_CHECK(h);
    MCL_CUDA(h, cudaMemcpyAsync(h->h_pinned, h->d_scratch, sizeof(int32_t), cudaMemcpyDeviceToHost, h->stream));
    MCL_CUDA(h, cudaStreamSynchronize(h->stream));
    memcpy(&h->c0, h->h_pinned, sizeof(int32_t));
    // every table value is >= quantise(log 1e-6) (pu:141 clamps p at 1e-6); 4 units of slack for libm differences
    h->voff = (int32_t)llrint(log(1e-6) * MCL_LOGP_SCALE) - 4;
    h->acc_terms_ok = (uint64_t)((int64_t)h->c0 - h->voff) * MCL_ACC_TERMS < ((uint64_t)1 << 32);
    // cell arithmetic: 2^(E-1) must cover the map, the window offset and two beam lengths
    {
        const double rmax = ceil(h->max_range / h->res);
        const double need = (double)std::max(h->W, h->H) + 2.0 * rmax + 8.0;
        int E = 12;
        while (E < 40 && ldexp(1.0, E - 1) < need) ++E;
        if (E > 20) return mcl_fail(h, MCL_ERR_CAPACITY, "map extent + max_range/res exceeds 2^19 cells");
        h->cell_S = 20 - E;
        h->cell_M = ldexp(1.5, E);
        h->cell_K = (int)(((uint32_t)(1023 + E) << 20) + (1u << 19));
        h->cell_lim = ldexp(1.0, E - 1) - rmax - 4.0;
    }
    cudaFree(h->d_win8); cudaFree(h->d_lut); cudaFree(h->d_win);
    h->d_win8 = nullptr; h->d_lut = nullptr; h->d_win = nullptr; h->coded = false; h->win8_bytes = 0;
    {   // window layout: free-space box + border, extended to the map edge on the sides within beam reach of it
        const int reach = (int)ceil(h->max_range / h->res) + 2;
        const bool any = h->ww > 0 && h->wh > 0;
        const int bx0 = h->wx0, bx1 = h->wx0 + h->ww - 1, by0 = h->wy0, by1 = h->wy0 + h->wh - 1;
        int want = 0;
        if (any) {
            if (bx0 < reach) want |= 1;
            if (bx1 > h->W - 1 - reach) want |= 2;
            if (by0 < reach) want |= 4;
            if (by1 > h->H - 1 - reach) want |= 8;
        }
        // extending the window costs shared memory and one axis must stay within 256 cells: take the EDGE sides
        // of both axes if that fits, else those of one axis, else none
        const int tries[4] = {want, want & 3, want & 12, 0};
        const size_t limit = (size_t)h->smem_optin - 512;
        for (int k = 0; k < 4; ++k) {
            const int edge = tries[k];
            h->win_edge = edge;
            h->win_ofx = (edge & 1) ? -2 : h->wx0 - 1;
            h->win_ofy = (edge & 4) ? -2 : h->wy0 - 1;
            h->win_cx = ((edge & 2) ? h->W : (any ? bx1 + 1 : h->wx0)) - h->win_ofx;
            h->win_cy = ((edge & 8) ? h->H : (any ? by1 + 1 : h->wy0)) - h->win_ofy;
            // pitch-256 layout: the minor axis is the larger extent that still fits in 256 columns (fewest rows)
            const int ex = h->win_cx + 1, ey = h->win_cy + 1;
            h->win_ok = std::min(ex, ey) <= 256;
            h->win_tpose = h->win_ok && (ex > 256 || (ey <= 256 && ey > ex));
            h->win_rows = h->win_tpose ? ex : ey;
            // The 256 columns of a row exist whether the box uses them or not: centre the box in them (sides that
            // do not run to the map edge only).  A particle whose beams stay inside the 256 columns then needs no
            // clamp on that coordinate (likelihood.cu, g1_slices: one instruction in fifteen).
            if (h->win_ok) {
                const bool tp = h->win_tpose;
                const int ext = tp ? ey : ex, sides = tp ? (edge & 12) : (edge & 3);
                if (sides == 0 && ext < 256) {
                    const int pad = (256 - ext) / 2;
                    if (tp) { h->win_ofy -= pad; h->win_cy = 255; } else { h->win_ofx -= pad; h->win_cx = 255; }
                }
            }
            h->win_bytes = h->win_ok ? (size_t)h->win_rows * 256 * sizeof(int32_t) : 0;
            const bool fits = h->win_ok && (16 + h->win_bytes <= limit || (size_t)h->win_rows * 256 + 16 + 32768 + 64 <= limit);
            if (fits || edge == 0) break;
        }
        if (h->win_ok) MCL_CUDA(h, cudaMalloc((void **)&h->d_win, h->win_bytes));
    }
    if (h->win_ok) {
        const int n = h->win_rows * 256;
        k_pack_window<<<std::max(1, std::min((n + 255) / 256, h->sm_count * 8)), 256, 0, h->stream>>>(
            h->d_logtab, h->d_win, h->W, h->H, h->win_ofx, h->win_ofy, h->win_edge, h->win_rows, h->win_tpose ? 1 : 0,
            h->c0, h->voff, 0);
        MCL_LAUNCH_CHECK(h);
        // skewed copy (rows of 264 words: consecutive rows start 8 banks apart)
        cudaFree(h->d_win_skew); h->d_win_skew = nullptr;
        h->win_skew_bytes = ((size_t)n + ((size_t)n >> 5) + 8) * sizeof(int32_t);
        if (16 + h->win_skew_bytes <= (size_t)h->smem_optin - 512) {
            MCL_CUDA(h, cudaMalloc((void **)&h->d_win_skew, h->win_skew_bytes));
            MCL_CUDA(h, cudaMemsetAsync(h->d_win_skew, 0, h->win_skew_bytes, h->stream));
            k_pack_window<<<std::max(1, std::min((n + 255) / 256, h->sm_count * 8)), 256, 0, h->stream>>>(
                h->d_logtab, h->d_win_skew, h->W, h->H, h->win_ofx, h->win_ofy, h->win_edge, h->win_rows, h->win_tpose ? 1 : 0,
                h->c0, h->voff, 1);
            MCL_LAUNCH_CHECK(h);
        }
        // coded window (uint8 + table of distinct values) when the int32 window does not fit in shared memory
        const size_t limit = (size_t)h->smem_optin - 512;
        if (16 + h->win_bytes > limit && (size_t)n + 16 + 32768 + 64 <= limit) {
            std::vector<int32_t> win((size_t)n);
            MCL_CUDA(h, cudaMemcpyAsync(win.data(), h->d_win, (size_t)n * sizeof(int32_t), cudaMemcpyDeviceToHost, h->stream));
            MCL_CUDA(h, cudaStreamSynchronize(h->stream));
            std::vector<int32_t> uniq(win);
            std::sort(uniq.begin(), uniq.end());
            uniq.erase(std::unique(uniq.begin(), uniq.end()), uniq.end());
            if (uniq.size() <= 256) {
                std::vector<uint8_t> codes((size_t)n, 0);
                for (int i = 0; i < n; ++i)
                    codes[i] = (uint8_t)(std::lower_bound(uniq.begin(), uniq.end(), win[i]) - uniq.begin());
                uniq.resize(256, 0);
                MCL_CUDA(h, cudaMalloc((void **)&h->d_win8, codes.size()));
                MCL_CUDA(h, cudaMalloc((void **)&h->d_lut, 256 * sizeof(int32_t)));
                MCL_CUDA(h, cudaMemcpy(h->d_win8, codes.data(), codes.size(), cudaMemcpyHostToDevice));
                MCL_CUDA(h, cudaMemcpy(h->d_lut, uniq.data(), 256 * sizeof(int32_t), cudaMemcpyHostToDevice));
                h->win8_bytes = codes.size();
                h->coded = true;
            }
        }
    }
    // large maps: byte-coded copy of the whole table for the tiled kernel
    cudaFree(h->d_code8);
    h->d_code8 = nullptr; h->tiled_ok = false;
    const bool window_in_smem = h->win_ok && (16 + h->win_bytes <= (size_t)h->smem_optin - 512 || h->coded);
    if (!window_in_smem && h->cell_S >= 0) {
        const int margin = (((int)ceil(h->max_range / h->res) + 2) + 3) & ~3;       // multiple of 4: word-aligned staging
        // tile height: the largest multiple of 32 (at most 96) whose sub-window lets two CTAs share an SM
        int tw = (256 - 2 * margin) & ~15;            // multiple of 16: tile origins stay 16-byte aligned (bulk copies)
        if (tw < 16) tw = (256 - 2 * margin) & ~3;
        int th = 96;
        while (th > 32 && 2 * (16 + 32768 + (size_t)(th + 2 * margin) * 260 + 1536) > (size_t)h->smem_optin) th -= 32;
        const int tx = tw > 0 ? (h->W + tw - 1) / tw : 0, ty = (h->H + th - 1) / th;
        const size_t sub_bytes = (size_t)(th + 2 * margin) * 260;
        if (tw >= 16 && (int64_t)tx * ty <= 65536 && 16 + 32768 + sub_bytes + 64 <= (size_t)h->smem_optin - 512) {
            std::vector<int32_t> tab((size_t)cells);
            MCL_CUDA(h, cudaMemcpyAsync(tab.data(), h->d_logtab, (size_t)cells * sizeof(int32_t), cudaMemcpyDeviceToHost, h->stream));
            MCL_CUDA(h, cudaStreamSynchronize(h->stream));
            // distinct values (few: the value depends only on dist) through a small open-addressing table
            const int HS = 2048;
            std::vector<int32_t> hkey(HS, 0), hcode(HS, -1);
            std::vector<char> hused(HS, 0);
            std::vector<int32_t> uniq;
            auto slot_of = [&](int32_t v) { unsigned s0 = ((unsigned)v * 2654435761u) >> 21; while (hused[s0] && hkey[s0] != v) s0 = (s0 + 1) & (HS - 1); return (int)s0; };
            for (int64_t c = 0; c < cells && uniq.size() <= 255; ++c) {
                const int sl = slot_of(tab[c]);
                if (!hused[sl]) { hused[sl] = 1; hkey[sl] = tab[c]; uniq.push_back(tab[c]); }
            }
            if (uniq.size() <= 255) {
                std::sort(uniq.begin(), uniq.end());
                for (size_t k = 0; k < uniq.size(); ++k) hcode[slot_of(uniq[k])] = (int)k;
                std::vector<uint8_t> codes((size_t)cells);
                // code 0 = outside the map (adds 0): what a zero-filled staging area means; values are codes 1 .. 255
                for (int64_t c = 0; c < cells; ++c) codes[c] = (uint8_t)(hcode[slot_of(tab[c])] + 1);
                std::vector<int32_t> lut(256, -h->voff);
                for (size_t k = 0; k < uniq.size(); ++k) lut[k + 1] = uniq[k] - h->voff;
                MCL_CUDA(h, cudaMalloc((void **)&h->d_code8, (size_t)cells));
                MCL_CUDA(h, cudaMemcpy(h->d_code8, codes.data(), (size_t)cells, cudaMemcpyHostToDevice));
                // padded copy for the bulk-copy kernel: 16 columns in front of every row and one row in front of the
                // map, zero (= outside) except the cells at x = -1 / y = -1, which repeat column / row 0: int() sends a
                // coordinate in (-1, 0) to cell 0 (pu:128-129), so a staged sub-window then needs no in-map test at all
                {
                    const size_t Wp = (size_t)h->W + 16, Hp = (size_t)h->H + 1;
                    std::vector<uint8_t> pad(Wp * Hp, 0);
                    for (int y = -1; y < h->H; ++y) {
                        const uint8_t *src = codes.data() + (size_t)(y < 0 ? 0 : y) * h->W;
                        uint8_t *dst = pad.data() + (size_t)(y + 1) * Wp + 16;
                        memcpy(dst, src, (size_t)h->W);
                        dst[-1] = src[0];
                    }
                    cudaFree(h->d_code8p);
                    h->d_code8p = nullptr;
                    MCL_CUDA(h, cudaMalloc((void **)&h->d_code8p, Wp * Hp));
                    MCL_CUDA(h, cudaMemcpy(h->d_code8p, pad.data(), Wp * Hp, cudaMemcpyHostToDevice));
                }
                cudaFree(h->d_lut);
                h->d_lut = nullptr;
                MCL_CUDA(h, cudaMalloc((void **)&h->d_lut, 256 * sizeof(int32_t)));
                MCL_CUDA(h, cudaMemcpy(h->d_lut, lut.data(), 256 * sizeof(int32_t), cudaMemcpyHostToDevice));
                h->tile_w = tw; h->tile_h = th; h->tile_margin = margin; h->tiles_x = tx; h->tiles_y = ty;
                h->tiled_ok = true;
            }
        }
    }
    h->tab_dirty = false;
    return MCL_OK;
}

// pu:119-129: beams j = 0, step, 2 step, ...; valid iff isfinite(r) and r < max_range.
// endpoint offset r*(cos a, sin a) with a = (double)angles[j], scaled to cells.
// Valid beams with r >= 0 first ("0 <= r <= max_range", pu:139), then negative finite ranges (p_rand = 0).
// A finite range <= -max_range also passes the reference's test (pu:123 has no lower bound), but its endpoint lies
// further from the particle than the cell arithmetic is sized for (mcl_prepare_table: map extent + 2 max_range):
// such a scan is refused (returns false) instead of scored wrongly.  No LaserScan holds negative ranges.
static bool build_beam_table(const mcl_handle *h, const float *h_ranges, const float *h_angles, int M,
                             BeamTable *out, int &n_pos, int &n_neg, double &rmax) {
    n_pos = 0; n_neg = 0; rmax = 0;
    std::vector<BeamTable> neg;
    for (int j = 0; j < M; j += h->step) {
        const double r = (double)h_ranges[j];
        if (!(isfinite(r) && r < h->max_range)) continue;
        if (r <= -h->max_range) return false;
        const double a = (double)h_angles[j];
        BeamTable b;
        b.bx = r * cos(a) / h->res;
        b.by = r * sin(a) / h->res;
        rmax = std::max(rmax, fabs(r) / h->res);
        if (r >= 0) out[n_pos++] = b;
        else neg.push_back(b);
    }
    for (auto &b : neg) out[n_pos + n_neg++] = b;
    return true;
}

extern "C" int mcl_set_scan(mcl_handle *h, const float *h_ranges, const float *h_angles, int M) {
    if (!h) return MCL_ERR_ARG;
    if (M < 0 || (M > 0 && (!h_ranges || !h_angles))) return mcl_fail(h, MCL_ERR_ARG, "mcl_set_scan: bad argument");
    if (!h->sensor_set || h->W == 0) return mcl_fail(h, MCL_ERR_STATE, "mcl_set_scan: set map and sensor first");
    DeviceGuard g(h->device);
    if (M > h->beams_cap) {
        MCL_CUDA(h, cudaStreamSynchronize(h->stream));
        cudaFree(h->d_beams); cudaFreeHost(h->h_beams);
        h->d_beams = nullptr; h->h_beams = nullptr; h->beams_cap = 0;
        const int cap = std::max(M, 512);
        MCL_CUDA(h, cudaMalloc((void **)&h->d_beams, (size_t)cap * sizeof(BeamTable)));
        MCL_CUDA(h, cudaMallocHost((void **)&h->h_beams, (size_t)cap * sizeof(BeamTable)));
        h->beams_cap = cap;
    } else if (h->ev_beams) {
        // the pinned staging buffer may still be the source of an in-flight copy: wait for THAT copy only, so that
        // the table of the next scan is built while kernels enqueued since (the motion kernels of the step) run
        MCL_CUDA(h, cudaEventSynchronize(h->ev_beams));
    }
    if (!h->ev_beams) MCL_CUDA(h, cudaEventCreateWithFlags(&h->ev_beams, cudaEventDisableTiming));
    int n_pos = 0, n_neg = 0;
    double rmax = 0;
    if (!build_beam_table(h, h_ranges, h_angles, M, h->h_beams, n_pos, n_neg, rmax))
        return mcl_fail(h, MCL_ERR_ARG, "mcl_set_scan: a range <= -max_range (the cell arithmetic covers |r| < max_range)");
    h->n_pos = n_pos; h->n_neg = n_neg; h->rmax_cells = rmax;
    if (n_pos + n_neg > 0) {
        MCL_CUDA(h, cudaMemcpyAsync(h->d_beams, h->h_beams, (size_t)(n_pos + n_neg) * sizeof(BeamTable),
                                    cudaMemcpyHostToDevice, h->stream));
        MCL_CUDA(h, cudaEventRecord(h->ev_beams, h->stream));
    }
    h->d_beams_active = h->d_beams;
    h->scan_gen = mcl_next_scan_uid();
    h->scan_set = true;
    return MCL_OK;
}

extern "C" int mcl_set_scan_batch(mcl_handle *h, const float *h_ranges, const float *h_angles, int M, int K) {
    if (!h) return MCL_ERR_ARG;
    if (M <= 0 || K <= 0 || !h_ranges || !h_angles) return mcl_fail(h, MCL_ERR_ARG, "mcl_set_scan_batch: bad argument");
    if (!h->sensor_set || h->W == 0) return mcl_fail(h, MCL_ERR_STATE, "mcl_set_scan_batch: set map and sensor first");
    DeviceGuard g(h->device);
    MCL_CUDA(h, cudaStreamSynchronize(h->stream));
    cudaFree(h->d_batch);
    h->d_batch = nullptr;
    h->batch_meta.clear();
    h->batch_stride = M;
    std::vector<BeamTable> host((size_t)M * K);
    for (int k = 0; k < K; ++k) {
        mcl_handle::ScanMeta m;
        if (!build_beam_table(h, h_ranges + (size_t)k * M, h_angles, M, host.data() + (size_t)k * M, m.n_pos, m.n_neg,
                              m.rmax_cells))
            return mcl_fail(h, MCL_ERR_ARG, "mcl_set_scan_batch: a range <= -max_range (the cell arithmetic covers |r| < max_range)");
        h->batch_meta.push_back(m);
    }
    MCL_CUDA(h, cudaMalloc((void **)&h->d_batch, host.size() * sizeof(BeamTable)));
    MCL_CUDA(h, cudaMemcpy(h->d_batch, host.data(), host.size() * sizeof(BeamTable), cudaMemcpyHostToDevice));
    return MCL_OK;
}

extern "C" int mcl_use_scan(mcl_handle *h, int k) {
    if (!h) return MCL_ERR_ARG;
    if (k < 0 || k >= (int)h->batch_meta.size()) return mcl_fail(h, MCL_ERR_ARG, "mcl_use_scan: no such pre-staged scan");
    h->d_beams_active = h->d_batch + (size_t)k * h->batch_stride;
    h->scan_gen = mcl_next_scan_uid();
    h->n_pos = h->batch_meta[k].n_pos; h->n_neg = h->batch_meta[k].n_neg; h->rmax_cells = h->batch_meta[k].rmax_cells;
    h->scan_set = true;
    return MCL_OK;
}

extern "C" int mcl_scan_valid_count(mcl_handle *h, int *valid_count) {
    if (!h || !valid_count) return MCL_ERR_ARG;
    if (!h->scan_set) return mcl_fail(h, MCL_ERR_STATE, "scan not set");
    *valid_count = h->n_pos + h->n_neg;
    return MCL_OK;
}

// node:410-421 compute_motion
extern "C" int mcl_compute_motion(const double o1[3], const double o2[3], double delta[3]) {
    if (!o1 || !o2 || !delta) return MCL_ERR_ARG;
    const double dx = o2[0] - o1[0], dy = o2[1] - o1[1];
    double t = fmod((o2[2] - o1[2]) + MCL_PI, MCL_TWO_PI);
    if (t != 0.0 && t < 0.0) t += MCL_TWO_PI;
    const double dtheta = t - MCL_PI;
    const double rot1 = atan2(dy, dx) - o1[2];
    delta[0] = rot1;
    delta[1] = hypot(dx, dy);
    delta[2] = dtheta - rot1;
    return MCL_OK;
}

// ------------------------------------------------------------------------------------------
// timing of likelihood launches
// ------------------------------------------------------------------------------------------
extern "C" int mcl_timing_start(mcl_handle *h) {
    if (!h) return MCL_ERR_ARG;
    for (auto &p : h->lik_events) { cudaEventDestroy(p.first); cudaEventDestroy(p.second); }
    h->lik_events.clear();
    h->lik_sets_timed = 0;
    h->timing = true;
    return MCL_OK;
}

// particle sets evaluated by the likelihood launches of the last timing window (a pair launch counts two)
extern "C" int64_t mcl_timing_sets(const mcl_handle *h) { return h ? h->lik_sets_timed : 0; }

extern "C" int mcl_timing_stop(mcl_handle *h, double *ms, int64_t *launches) {
    if (!h) return MCL_ERR_ARG;
    DeviceGuard g(h->device);
    h->timing = false;
    MCL_CUDA(h, cudaStreamSynchronize(h->stream));
    double tot = 0;
    for (auto &p : h->lik_events) {
        float t = 0;
        MCL_CUDA(h, cudaEventElapsedTime(&t, p.first, p.second));
        tot += t;
    }
    if (ms) *ms = tot;
    if (launches) *launches = (int64_t)h->lik_events.size();
    for (auto &p : h->lik_events) { cudaEventDestroy(p.first); cudaEventDestroy(p.second); }
    h->lik_events.clear();
    return MCL_OK;
}
