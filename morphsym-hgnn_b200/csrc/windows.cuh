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
//  global memory (one bulk async copy), every source column is z-scored once (fp64 statistics, two-pass like
//  torch.mean / torch.std(correction=1)), and each (node, variable, axis) block of the three x tensors is written as a
//  run of T consecutive floats (coalesced).  The column / sign tables are compiled once on the host and travel as
//  kernel parameters (constant bank).  No atomics, no index tensors.
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

__device__ __forceinline__ uint32_t win_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// Window pipeline of one CTA (grid-stride over graphs):
//   (0) the T x C window - one contiguous run of the row-major sequence - lands in shared memory as raw[T][C]: by ONE bulk
//       async copy (cp.async.bulk global -> shared, mbarrier completion) when rows are 16-byte multiples, else by
//       coalesced element loads;
//   (1)-(2) thread (column c, row quarter q) accumulates fp64 partial sums / squared deviations (conflict-free: the
//       threads of a warp read consecutive columns of one row), combined through shared memory;
//   (3) the same threads write the z-scored values transposed into nz[C][T|1] (odd pitch: conflict-free);
//       the raw buffer is free from here on, so the NEXT graph's bulk copy is issued now and overlaps (4);
//   (4) every (node, variable, axis) block is a run of T floats in nz and in the output row: coalesced stores.
template <typename TIn>
__global__ void __launch_bounds__(WIN_THREADS)
k_build_windows(const WindowTable tb, const TIn* __restrict__ seq, const TIn* __restrict__ label_seq, const int64_t n_rows,
                const int64_t* __restrict__ starts, const int64_t B, const WindowPtrs out, float* __restrict__ y, const int bulk) {
    extern __shared__ __align__(128) uint8_t win_smem[];
    __shared__ __align__(8) uint64_t bar;
    __shared__ double part[4][WIN_MAX_COLS];
    __shared__ double mean_s[WIN_MAX_COLS], rstd_s[WIN_MAX_COLS];
    const int T = tb.T, C = tb.C;
    const int Tp = T | 1;
    TIn* raw = reinterpret_cast<TIn*>(win_smem);                        // [T][C]
    float* nz = reinterpret_cast<float*>(raw + (size_t)T * C);          // [C][Tp]
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t bar_a = win_smem_u32(&bar);
    const uint32_t bytes = (uint32_t)((size_t)T * C * sizeof(TIn));
    // statistics threads: column c = tid % C4, row quarter q = tid / C4 (C4 = columns rounded so that 4 quarters fit)
    const int nq = (WIN_THREADS / C) < 4 ? (WIN_THREADS / C) : 4;       // 1..4 row ranges
    const int sc = tid % C, sq = tid / C;
    const bool stat = sq < nq;
    const int t0 = stat ? (int)((int64_t)T * sq / nq) : 0, t1 = stat ? (int)((int64_t)T * (sq + 1) / nq) : 0;

    auto window_start = [&](const int64_t g) {
        int64_t s = starts[g];
        return s < 0 ? (int64_t)0 : (s > n_rows - T ? n_rows - T : s);  // validated on the host; the clamp keeps loads in bounds
    };
    auto issue = [&](const int64_t g) {                                 // thread 0 only
        const TIn* src = seq + window_start(g) * C;
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");     // generic reads of raw (before the CTA barrier) -> async write
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar_a), "r"(bytes) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     ::"r"(win_smem_u32(raw)), "l"(src), "r"(bytes), "r"(bar_a) : "memory");
    };
    if (bulk) {
        if (tid == 0) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar_a) : "memory");
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        __syncthreads();
        if (tid == 0 && (int64_t)blockIdx.x < B) issue(blockIdx.x);
    }
    uint32_t phase = 0;
    for (int64_t g = blockIdx.x; g < B; g += gridDim.x) {
        const int64_t s = window_start(g);
        if (bulk) {
            uint32_t ok = 0;
            while (!ok) {
                asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}"
                             : "=r"(ok) : "r"(bar_a), "r"(phase) : "memory");
            }
            phase ^= 1u;
        } else {
            const TIn* src = seq + s * C;
            const int n = T * C;
            for (int i = tid; i < n; i += WIN_THREADS) raw[i] = __ldg(src + i);
            __syncthreads();
        }
        if (tb.normalize) {
            if (stat && tb.col_used[sc]) {
                double a = 0.0;
                for (int t = t0; t < t1; ++t) a += (double)raw[t * C + sc];
                part[sq][sc] = a;
            }
            __syncthreads();
            if (tid < C && tb.col_used[tid]) {
                double a = 0.0;
                for (int q = 0; q < nq; ++q) a += part[q][tid];
                mean_s[tid] = a / (double)T;
            }
            __syncthreads();
            if (stat && tb.col_used[sc]) {
                const double m = mean_s[sc];
                double a = 0.0;
                for (int t = t0; t < t1; ++t) { const double d = (double)raw[t * C + sc] - m; a += d * d; }
                part[sq][sc] = a;
            }
            __syncthreads();
            if (tid < C && tb.col_used[tid]) {
                double a = 0.0;
                for (int q = 0; q < nq; ++q) a += part[q][tid];
                const double sd = sqrt(a / (double)(T - 1));
                rstd_s[tid] = sd > 0.0 ? 1.0 / sd : 0.0;                // 0: the (v - mean) / 0 branch below
            }
            __syncthreads();
        }
        if (stat && tb.col_used[sc]) {
            float* nc = nz + sc * Tp;
            if (tb.normalize) {
                const double m = mean_s[sc], r = rstd_s[sc];
                if (r != 0.0) {
                    for (int t = t0; t < t1; ++t) nc[t] = (float)(((double)raw[t * C + sc] - m) * r);
                } else {
                    // (v - mean) / 0: 0/0 = NaN -> 0 (np.nan_to_num(nan=0.0)); +-x/0 = +-inf -> +-largest finite
                    for (int t = t0; t < t1; ++t) {
                        const double d = (double)raw[t * C + sc] - m;
                        nc[t] = d == 0.0 ? 0.f : (d > 0.0 ? FLT_MAX : -FLT_MAX);
                    }
                }
            } else {
                for (int t = t0; t < t1; ++t) nc[t] = (float)raw[t * C + sc];
            }
        }
        __syncthreads();                                               // nz complete, raw free
        if (bulk && tid == 0 && g + gridDim.x < B) issue(g + gridDim.x);
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
        __syncthreads();                                               // nz free for the next graph
    }
}

}  // namespace mshgnn
