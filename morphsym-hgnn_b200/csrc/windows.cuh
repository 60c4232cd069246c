// Device-side window builder: the reference's per-sample dataset path, batched on the GPU.
//
//  Reference per graph (one dataset index `idx`):
//    LinTzuYaunDataset.load_data_at_dataset_seq      datasets_py/LinTzuYaunDataset.py:L66-88   rows [idx, idx+T) of every
//                                                                                             channel array; label = last row
//    ..._Morph.load_data_sorted_k4 / _c2             LinTzuYaunDataset_Morph.py:L156-347      URDF column order, base tiling,
//                                                                                             per-window z-score (Bessel std,
//                                                                                             NaN -> 0)
//    ..._Morph.get_helper_heterogeneous_gnn(_c2)     LinTzuYaunDataset_Morph.py:L555-697      x_t[node] = concat over variables of
//    FlexibleDataset.get_helper_heterogeneous_gnn    flexibleDataset.py:L537-607              column(s).flatten('F')  ([axis][time])
//  and torch_geometric's collate concatenates the graphs.
//
//  Here: the raw sequence lives on the device ONCE as seq[n_rows][C] (all channel arrays side by side, dataset column
//  order); a batch is a list of window start rows.  One CTA per graph: the T x C window is one contiguous run of
//  global memory (coalesced 128-bit-friendly loads), it is transposed into shared memory, every source column is
//  z-scored once (fp64 statistics, two-pass like torch.mean / torch.std(correction=1)), and each (node, variable, axis)
//  block of the three x tensors is written as a run of T consecutive floats (coalesced).  The column / sign tables are
//  compiled once on the host and travel as kernel parameters (constant bank).  No atomics, no index tensors.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <cfloat>

namespace mshgnn {

constexpr int WIN_MAX_BLOCKS = 160;     // sum over node types of nodes * blocks-per-node
constexpr int WIN_MAX_COLS = 96;        // source columns of the raw sequence
constexpr int WIN_MAX_LABELS = 32;
constexpr int WIN_THREADS = 256;

struct WindowTable {
    int T;                               // history_length
    int C;                               // columns of seq
    int CL;                              // columns of the label sequence
    int n_types;
    int nodes[4];                        // nodes per graph of each type
    int blocks[4];                       // blocks per node of each type (row width = blocks * blen)
    int blen[4];                         // block length: T, or 1 for a type without variables (constant 1.0 feature)
    int first_block[4];                  // index of the type's first block in col[] / sign[]
    int n_blocks;
    int normalize;
    int n_labels;
    int16_t col[WIN_MAX_BLOCKS];         // source column of block (type, node, k): first_block[t] + node * blocks[t] + k; -1: constant
    int8_t sign[WIN_MAX_BLOCKS];         // +-1 factor (dataset-level group action), or the constant's value when col < 0
    int16_t label_col[WIN_MAX_LABELS];
    int8_t label_sign[WIN_MAX_LABELS];
    uint8_t col_used[WIN_MAX_COLS];      // columns that feed at least one block (others are skipped by the statistics)
};

struct WindowPtrs {
    float* x[4];
};

template <typename TIn>
__global__ void __launch_bounds__(WIN_THREADS)
k_build_windows(const WindowTable tb, const TIn* __restrict__ seq, const TIn* __restrict__ label_seq, const int64_t n_rows,
                const int64_t* __restrict__ starts, const int64_t B, const WindowPtrs out, float* __restrict__ y) {
    extern __shared__ __align__(16) uint8_t win_smem[];
    const int T = tb.T, C = tb.C;
    const int Tp = T | 1;                                 // odd column pitch: conflict-free transposed writes
    TIn* w = reinterpret_cast<TIn*>(win_smem);            // [C][Tp] raw window
    float* nz = reinterpret_cast<float*>(w + (size_t)C * Tp);   // [C][Tp] z-scored (or plain fp32) window
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

    for (int64_t g = blockIdx.x; g < B; g += gridDim.x) {
        int64_t s = starts[g];
        s = s < 0 ? 0 : (s > n_rows - T ? n_rows - T : s);            // validated on the host; clamp keeps loads in bounds
        const TIn* src = seq + s * C;
        const int n = T * C;
        for (int i = tid; i < n; i += WIN_THREADS) {
            const int t = i / C, c = i - t * C;
            w[c * Tp + t] = __ldg(src + i);
        }
        __syncthreads();
        for (int c = warp; c < C; c += WIN_THREADS / 32) {
            if (!tb.col_used[c]) continue;
            const TIn* wc = w + c * Tp;
            float* nc = nz + c * Tp;
            if (tb.normalize) {
                double sum = 0.0;
                for (int t = lane; t < T; t += 32) sum += (double)wc[t];
#pragma unroll
                for (int o = 16; o; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
                const double mean = sum / (double)T;
                double ss = 0.0;
                for (int t = lane; t < T; t += 32) { const double d = (double)wc[t] - mean; ss += d * d; }
#pragma unroll
                for (int o = 16; o; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
                const double sd = sqrt(ss / (double)(T - 1));
                if (sd > 0.0) {
                    const double r = 1.0 / sd;
                    for (int t = lane; t < T; t += 32) nc[t] = (float)(((double)wc[t] - mean) * r);
                } else {
                    // (v - mean) / 0: 0/0 = NaN -> 0 (np.nan_to_num(nan=0.0)); +-x/0 = +-inf -> +-largest finite
                    for (int t = lane; t < T; t += 32) {
                        const double d = (double)wc[t] - mean;
                        nc[t] = d == 0.0 ? 0.f : (d > 0.0 ? FLT_MAX : -FLT_MAX);
                    }
                }
            } else {
                for (int t = lane; t < T; t += 32) nc[t] = (float)wc[t];
            }
        }
        __syncthreads();
        for (int b = warp; b < tb.n_blocks; b += WIN_THREADS / 32) {
            int ty = 0;
#pragma unroll
            for (int q = 1; q < 4; ++q) if (q < tb.n_types && b >= tb.first_block[q]) ty = q;
            const int local = b - tb.first_block[ty];
            const int node = local / tb.blocks[ty], k = local - node * tb.blocks[ty];
            const int len = tb.blen[ty];
            float* dst = out.x[ty] + ((g * tb.nodes[ty] + node) * (int64_t)tb.blocks[ty] + k) * len;
            const int c = tb.col[b];
            const float f = (float)tb.sign[b];
            if (c >= 0) {
                const float* nc = nz + c * Tp;
                for (int t = lane; t < len; t += 32) dst[t] = nc[t] * f;
            } else {
                for (int t = lane; t < len; t += 32) dst[t] = f;
            }
        }
        if (y != nullptr && tid < tb.n_labels) {
            const TIn* lrow = label_seq + (s + T - 1) * tb.CL;
            y[g * tb.n_labels + tid] = (float)lrow[tb.label_col[tid]] * (float)tb.label_sign[tid];
        }
        __syncthreads();
    }
}

}  // namespace mshgnn
