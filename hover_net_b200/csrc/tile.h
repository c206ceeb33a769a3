// Host-side interface of the whole-image tile path (tile.cu).
#pragma once
#include "common.cuh"

namespace hvn {

struct Model;

// patch grid of reference infer/tile.py:60-69 for an H x W image and a patch_out step
void tile_grid(int H, int W, int patch_out, int *rows, int *cols);
size_t tile_workspace_bytes(int patch_in, int patch_out, int C, int batch);
// Runs grid cells [cell_lo, cell_hi) (row-major) of the image through the network and writes each patch's
// map into pred [H,W,C] (device).  Pixels of other cells are left untouched.  Returns its own kernel launches.
int tile_predict(Model &model, Arena &ws, cudaStream_t s, const uint8_t *img, int H, int W, int patch_in, int cell_lo,
                 int cell_hi, int batch, int chunk, float *pred);

// Instance tables [n,max_rows,10] -> packed [<= cap rows,10] in map order, offs [n+1] (offs[n] = total rows; rows past
// cap are not written).  Device pointers.  Returns its kernel launches.
int pack_tables(cudaStream_t s, const long long *table, const int *n_rows, int n, int max_rows, long long *packed,
                long long cap, int *offs);

}  // namespace hvn
