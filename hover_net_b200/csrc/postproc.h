// Host-side interface of the device post-processing path (postproc.cu).
#pragma once
#include <string>

#include "common.cuh"

namespace hvn {

struct PostprocBuffers {
    int max_blobs = 0, max_ids = 0;
    void *stats = nullptr;
    unsigned char *fg = nullptr, *mk0 = nullptr, *flag = nullptr, *filled = nullptr, *eroded = nullptr, *opened = nullptr;
    int *L1 = nullptr, *size1 = nullptr, *blob_of_root = nullptr, *blob_root = nullptr;
    void *bbox = nullptr, *lab_range = nullptr;
    double *sobh = nullptr, *sobv = nullptr, *din = nullptr, *dist = nullptr;
    int *L2 = nullptr, *size3 = nullptr, *id3 = nullptr, *rowcnt = nullptr;
    void *heap = nullptr, *acc = nullptr;
    int *tcnt = nullptr;
};

void postproc_set_flood_impl(int v);  // test hook: 0 auto, 1 shared-memory heap, 2 global-memory heap
size_t postproc_workspace_bytes(int n, int H, int W, int nr_types);

// pred [n,H,W,C] f32 (device) -> inst [n,H,W] i32, table [n,max_rows,10] i64, n_rows [n] i32 (device).
// Returns the number of kernels launched on `stream`.
int postproc_run(Arena &arena, cudaStream_t stream, const float *pred, int n, int H, int W, int C, int nr_types,
                 int *inst, long long *table, int max_rows, int *n_rows, std::string *prof = nullptr);

// Contours of every table row (contour.cu): offs [n*max_rows + 1] i32 exclusive prefix of the per-row point
// counts, pts [cap][2] i32 (x, y) in map coordinates.  All pointers are device pointers.  Returns the number
// of kernels launched.
int contours_run(cudaStream_t stream, const int *inst, const long long *table, const int *n_rows, int n_maps, int H,
                 int W, int max_rows, int *pts, long long cap, int *offs);

}  // namespace hvn
