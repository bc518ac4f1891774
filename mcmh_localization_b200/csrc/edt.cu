// edt.cu -- node:153-157: distance_map = scipy.ndimage.distance_transform_edt(map == 0) * resolution as f32,
// computed on the device (matters for 4096 x 4096 maps: SciPy needs ~3 s there).
// Exact Euclidean transform in integers, so the result is bit-identical to SciPy's:
//   pass 1 (thread per column): g[y][x] = cells to the nearest non-free cell in the same column (up or down)
//   pass 2 (thread per cell):   d2 = min over x' of (x - x')^2 + g[y][x']^2, searched outwards from x and stopped
//                               as soon as dx^2 >= best (distances on these maps are tens of cells)
//   dist = (float)(sqrt((double)d2) * resolution)     -- IEEE sqrt of an exact integer, like SciPy's
#include <string.h>

#include <algorithm>
#include <vector>

#include "common.cuh"

#define EDT_INF 0x3fffffff

__global__ void k_edt_columns(const int8_t *__restrict__ occ, int W, int H, int *__restrict__ g) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    if (x >= W) return;
    int d = EDT_INF;
    for (int y = 0; y < H; ++y) {                  // distance to the nearest non-free cell above (smaller y)
        d = occ[(size_t)y * W + x] != 0 ? 0 : (d == EDT_INF ? EDT_INF : d + 1);
        g[(size_t)y * W + x] = d;
    }
    d = EDT_INF;
    for (int y = H - 1; y >= 0; --y) {             // ... or below
        d = occ[(size_t)y * W + x] != 0 ? 0 : (d == EDT_INF ? EDT_INF : d + 1);
        const size_t i = (size_t)y * W + x;
        if (d < g[i]) g[i] = d;
    }
}

__global__ void k_edt_rows(const int *__restrict__ g, int W, int H, double res, float *__restrict__ dist) {
    const int64_t cells = (int64_t)W * H;
    for (int64_t c = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; c < cells; c += (int64_t)gridDim.x * blockDim.x) {
        const int y = (int)(c / W), x = (int)(c - (int64_t)y * W);
        const int *row = g + (size_t)y * W;
        const long long g0 = row[x];
        long long best = g0 >= EDT_INF ? 0x7fffffffffffffffll : g0 * g0;
        for (long long dx = 1; dx * dx < best && (x - dx >= 0 || x + dx < W); ++dx) {
            if (x - dx >= 0) { const long long gv = row[x - dx]; if (gv < EDT_INF) best = min(best, dx * dx + gv * gv); }
            if (x + dx < W) { const long long gv = row[x + dx]; if (gv < EDT_INF) best = min(best, dx * dx + gv * gv); }
        }
        // no non-free cell anywhere: SciPy's value is arbitrary there; report +inf
        dist[c] = best == 0x7fffffffffffffffll ? __int_as_float(0x7f800000) : (float)__dmul_rn(sqrt((double)best), res);
    }
}

// h_dist_out (nullable): receives the distance map (the caller keeps it, like node.distance_map)
extern "C" int mcl_set_map_edt(mcl_handle *h, const int8_t *h_occ, int W, int H, double res, double ox, double oy,
                               float *h_dist_out) {
    if (!h) return MCL_ERR_ARG;
    if (!h_occ || W <= 0 || H <= 0 || !(res > 0)) return mcl_fail(h, MCL_ERR_ARG, "mcl_set_map_edt: bad argument");
    DeviceGuard guard(h->device);
    const size_t cells = (size_t)W * H;
    int8_t *d_occ = nullptr;
    int *d_g = nullptr;
    float *d_dist = nullptr;
    MCL_CUDA(h, cudaMalloc((void **)&d_occ, cells));
    cudaError_t e = cudaMalloc((void **)&d_g, cells * sizeof(int));
    if (e == cudaSuccess) e = cudaMalloc((void **)&d_dist, cells * sizeof(float));
    if (e != cudaSuccess) { cudaFree(d_occ); cudaFree(d_g); return mcl_fail(h, MCL_ERR_NOMEM, "mcl_set_map_edt: out of device memory"); }
    std::vector<float> dist(cells);
    e = cudaMemcpyAsync(d_occ, h_occ, cells, cudaMemcpyHostToDevice, h->stream);
    if (e == cudaSuccess) {
        k_edt_columns<<<(W + 127) / 128, 128, 0, h->stream>>>(d_occ, W, H, d_g);
        k_edt_rows<<<std::min<int64_t>((int64_t)(cells + 255) / 256, (int64_t)h->sm_count * 32), 256, 0, h->stream>>>(d_g, W, H, res, d_dist);
        h->launches += 2;
        e = cudaMemcpyAsync(dist.data(), d_dist, cells * sizeof(float), cudaMemcpyDeviceToHost, h->stream);
    }
    if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
    cudaFree(d_occ); cudaFree(d_g); cudaFree(d_dist);
    if (e != cudaSuccess) return mcl_fail(h, MCL_ERR_CUDA, std::string("mcl_set_map_edt: ") + cudaGetErrorString(e));
    if (h_dist_out) memcpy(h_dist_out, dist.data(), cells * sizeof(float));
    return mcl_set_map(h, h_occ, dist.data(), W, H, res, ox, oy);
}
