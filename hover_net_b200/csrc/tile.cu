// Whole-image tile path on the device (reference infer/tile.py:46-143):
// reflect-pad + patch grid (`_prepare_patching`), the batch loop over patches, and the stitch + crop of
// `_post_process_patches` -- as index arithmetic.  The padded image never exists: a patch pixel reads the
// source image through numpy's "reflect" index map (no edge repeat, period 2*(L-1)), and every patch
// output pixel is written straight to its place in the [H,W,C] prediction map (pixels past H/W are the
// crop).  Geometry: step = patch_out; rows = ceil((H - step)/step) + 1 (tile.py:60-69);
// pad top/left = (patch_in - patch_out)/2; patch (r, c) covers padded rows [r*step, r*step + patch_in).
#include "cnn.h"
#include "tile.h"

namespace hvn {
namespace {

__device__ __forceinline__ int reflect_index(int i, int L) {
    if (L == 1) return 0;
    const int p = 2 * (L - 1);
    int t = i % p;
    if (t < 0) t += p;
    return t < L ? t : p - t;
}

// patches [n, win, win, 3] u8 for grid cells first .. first+n-1 (row-major over (row, col))
__global__ void k_tile_gather(const uint8_t *__restrict__ img, int H, int W, int win, int step, int pad, int cols,
                              int first, int n, uint8_t *__restrict__ patches) {
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long total = (long long)n * win * win;
    if (idx >= total) return;
    const int p = (int)(idx / (win * win));
    const int r = (int)(idx - (long long)p * win * win);
    const int py = r / win, px = r - py * win;
    const int cell = first + p;
    const int gr = cell / cols, gc = cell - gr * cols;
    const int sy = reflect_index(gr * step + py - pad, H), sx = reflect_index(gc * step + px - pad, W);
    const uint8_t *s = img + ((size_t)sy * W + sx) * 3;
    uint8_t *d = patches + (size_t)idx * 3;
    d[0] = s[0]; d[1] = s[1]; d[2] = s[2];
}

// outputs [n, step, step, C] f32 -> pred [H, W, C] at (row*step, col*step), cropped to H x W
__global__ void k_tile_scatter(const float *__restrict__ outs, int n, int step, int C, int cols, int first, int H, int W,
                               float *__restrict__ pred) {
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long total = (long long)n * step * step;
    if (idx >= total) return;
    const int p = (int)(idx / (step * step));
    const int r = (int)(idx - (long long)p * step * step);
    const int py = r / step, px = r - py * step;
    const int cell = first + p;
    const int gr = cell / cols, gc = cell - gr * cols;
    const int y = gr * step + py, x = gc * step + px;
    if (y >= H || x >= W) return;
    const float *s = outs + (size_t)idx * C;
    float *d = pred + ((size_t)y * W + x) * C;
    for (int c = 0; c < C; ++c) d[c] = s[c];
}

// Instance tables [n, max_rows, 10] (padded) -> packed rows [sum n_rows, 10] in map order + offs [n + 1].
// One block per map: its offset is the sum of the earlier maps' row counts (n is a batch size: a strided sum).
__global__ void k_pack_tables(const long long *__restrict__ table, const int *__restrict__ n_rows, int n, int max_rows,
                              long long *__restrict__ packed, long long cap, int *__restrict__ offs) {
    __shared__ int s_part[32];
    __shared__ int s_off;
    const int m = blockIdx.x;
    int part = 0;
    for (int i = threadIdx.x; i < m; i += blockDim.x) part += min(n_rows[i], max_rows);
    for (int o = 16; o; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
    if ((threadIdx.x & 31) == 0) s_part[threadIdx.x >> 5] = part;
    __syncthreads();
    if (threadIdx.x == 0) {
        int t = 0;
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += s_part[w];
        s_off = t;
        offs[m] = t;
        if (m == n - 1) offs[n] = t + min(n_rows[m], max_rows);
    }
    __syncthreads();
    const int off = s_off, nr = min(n_rows[m], max_rows);
    const long long *src = table + (size_t)m * max_rows * 10;
    for (int i = threadIdx.x; i < nr * 10; i += blockDim.x) {
        const long long row = off + i / 10;
        if (row < cap) packed[(size_t)off * 10 + i] = src[i];
    }
}

}  // namespace

int pack_tables(cudaStream_t s, const long long *table, const int *n_rows, int n, int max_rows, long long *packed,
                long long cap, int *offs) {
    k_pack_tables<<<n, 256, 0, s>>>(table, n_rows, n, max_rows, packed, cap, offs);
    HVN_CUDA(cudaGetLastError());
    return 1;
}

void tile_grid(int H, int W, int patch_out, int *rows, int *cols) {
    auto steps = [&](int L) { return (L <= patch_out ? 0 : (L - patch_out + patch_out - 1) / patch_out) + 1; };
    *rows = steps(H);
    *cols = steps(W);
}

size_t tile_workspace_bytes(int patch_in, int patch_out, int C, int batch) {
    return ((size_t)batch * patch_in * patch_in * 3 + 255) / 256 * 256 + ((size_t)batch * patch_out * patch_out * C * 4 + 255) / 256 * 256;
}

int tile_predict(Model &model, Arena &ws, cudaStream_t s, const uint8_t *img, int H, int W, int patch_in, int cell_lo,
                 int cell_hi, int batch, int chunk, float *pred) {
    int oh, ow, oc;
    model.out_shape(patch_in, patch_in, oh, ow, oc);
    HVN_CHECK(oh == ow && oh >= 1, -1, "unsupported patch size for this model mode");
    const int step = oh, pad = (patch_in - step) / 2;
    int rows, cols;
    tile_grid(H, W, step, &rows, &cols);
    HVN_CHECK(cell_lo >= 0 && cell_hi <= rows * cols && cell_lo <= cell_hi, -1, "patch range outside the grid");
    if (batch < 1) batch = 64;
    uint8_t *patches = ws.take<uint8_t>((size_t)batch * patch_in * patch_in * 3);
    float *outs = ws.take<float>((size_t)batch * step * step * oc);
    int launches = 0;
    for (int first = cell_lo; first < cell_hi; first += batch) {
        const int n = std::min(batch, cell_hi - first);
        const long long tg = (long long)n * patch_in * patch_in;
        k_tile_gather<<<(unsigned)((tg + 255) / 256), 256, 0, s>>>(img, H, W, patch_in, step, pad, cols, first, n, patches);
        model.forward(patches, n, patch_in, patch_in, outs, chunk, s);
        const long long ts = (long long)n * step * step;
        k_tile_scatter<<<(unsigned)((ts + 255) / 256), 256, 0, s>>>(outs, n, step, oc, cols, first, H, W, pred);
        launches += 2;
    }
    HVN_CUDA(cudaGetLastError());
    return launches;
}

}  // namespace hvn
