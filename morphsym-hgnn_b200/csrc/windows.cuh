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
#include <cstring>

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
//   (0) the T x C window - one contiguous run of the row-major sequence - lands in shared memory as raw[T][Cp]: by ONE bulk
//       async copy (cp.async.bulk global -> shared, mbarrier completion) when rows are 16-byte multiples, else by
//       coalesced element loads;
//   (1) thread (column group of 16 bytes, row group) accumulates in ONE pass the fp64 sums of d = v - pivot and d^2 over its
//       rows (pivot = the column's first row: the shifted sums give mean and Bessel variance without cancellation), one
//       128-bit shared load per 4 (fp32) / 2 (fp64) columns; partials are combined through shared memory;
//   (2) the same threads write the z-scored values transposed into nz[C][T|1] (odd pitch); the raw buffer is free from here
//       on, so the NEXT graph's bulk copy is issued now and overlaps (3);
//   (3) every (node, variable, axis) block is a run of T floats in nz and in the output row: coalesced stores, block offsets
//       precomputed once per CTA.
//  (ncu on the first version: 18 K warp instructions per graph, 52 % issue-bound, three scalar passes with a conversion to
//  fp64 per element each; this version needs ~4 K.)
constexpr int WIN_MAX_RG = 16;           // row groups of the statistics pass

template <typename TIn> struct WinVec;
template <> struct WinVec<float> { static constexpr int W = 4; };
template <> struct WinVec<double> { static constexpr int W = 2; };

template <typename TIn>
__global__ void __launch_bounds__(WIN_THREADS)
k_build_windows(const WindowTable tb, const TIn* __restrict__ seq, const TIn* __restrict__ label_seq, const int64_t n_rows,
                const int64_t* __restrict__ starts, const int64_t B, const WindowPtrs out, float* __restrict__ y, const int bulk,
                const int nz_bytes) {
    extern __shared__ __align__(128) uint8_t win_smem[];
    __shared__ __align__(8) uint64_t bar;
    __shared__ float dm_s[WIN_MAX_COLS], rstd_s[WIN_MAX_COLS];       // mean - pivot, 1 / std (0: the std == 0 branch)
    __shared__ int boff_s[WIN_MAX_BLOCKS];                            // element offset of block b inside its graph's rows
    __shared__ int8_t bty_s[WIN_MAX_BLOCKS];
    constexpr int VW = WinVec<TIn>::W;
    const int T = tb.T, C = tb.C;
    const int Cp = (C + VW - 1) / VW * VW;                              // shared-memory row pitch (== C on the bulk path)
    const int Tp = T | 1;
    float* nz = reinterpret_cast<float*>(win_smem);                     // [C][Tp] z-scored window; doubles as the partial-sum scratch
    double* part = reinterpret_cast<double*>(win_smem);                 // [rg][Cp][2] during the statistics pass
    TIn* raw = reinterpret_cast<TIn*>(win_smem + nz_bytes);             // [T][Cp]
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t bar_a = win_smem_u32(&bar);
    const uint32_t bytes = (uint32_t)((size_t)T * C * sizeof(TIn));
    const int nquad = Cp / VW;
    const int nrg = (WIN_THREADS / nquad) < WIN_MAX_RG ? (WIN_THREADS / nquad) : WIN_MAX_RG;
    const int quad = tid % nquad, rg = tid / nquad;
    const bool stat = rg < nrg;
    const int c0 = quad * VW;

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
    for (int b = tid; b < tb.n_blocks; b += WIN_THREADS) {
        int ty = 0;
#pragma unroll
        for (int q = 1; q < 4; ++q) if (q < tb.n_types && b >= tb.first_block[q]) ty = q;
        bty_s[b] = (int8_t)ty;
        boff_s[b] = (b - tb.first_block[ty]) * tb.blen[ty];             // (node * blocks + k) * len
    }
    if (!bulk)
        for (int i = tid; i < T * (Cp - C); i += WIN_THREADS) raw[(i / (Cp - C)) * Cp + C + i % (Cp - C)] = (TIn)0;   // pad columns
    if (bulk && tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar_a) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (bulk && tid == 0 && (int64_t)blockIdx.x < B) issue(blockIdx.x);
    uint32_t phase = 0;
    for (int64_t g = blockIdx.x; g < B; g += gridDim.x) {
        const int64_t s = window_start(g);
        if (bulk) {
            if (warp == 0) {                                           // one warp polls; the others sleep in the CTA barrier
                uint32_t ok = 0;
                while (!ok) {
                    asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}"
                                 : "=r"(ok) : "r"(bar_a), "r"(phase) : "memory");
                }
            }
            phase ^= 1u;
            __syncthreads();
        } else {
            const TIn* src = seq + s * C;
            const int n = T * C;
            for (int i = tid; i < n; i += WIN_THREADS) raw[(i / C) * Cp + (i % C)] = __ldg(src + i);
            __syncthreads();
        }
        TIn piv[VW];
        if (stat) {
            const uint4 pv = *reinterpret_cast<const uint4*>(raw + c0);
            memcpy(piv, &pv, 16);
        }
        if (tb.normalize) {
            if (stat) {
                double sd[VW], sq[VW];
#pragma unroll
                for (int e = 0; e < VW; ++e) { sd[e] = 0.0; sq[e] = 0.0; }
                for (int t = rg; t < T; t += nrg) {
                    const uint4 pv = *reinterpret_cast<const uint4*>(raw + t * Cp + c0);
                    TIn v[VW];
                    memcpy(v, &pv, 16);
#pragma unroll
                    for (int e = 0; e < VW; ++e) {
                        const double d = (double)(v[e] - piv[e]);
                        sd[e] += d;
                        sq[e] = fma(d, d, sq[e]);
                    }
                }
#pragma unroll
                for (int e = 0; e < VW; ++e) { part[(rg * Cp + c0 + e) * 2] = sd[e]; part[(rg * Cp + c0 + e) * 2 + 1] = sq[e]; }
            }
            __syncthreads();
            if (tid < C) {
                double a = 0.0, q2 = 0.0;
                for (int r = 0; r < nrg; ++r) { a += part[(r * Cp + tid) * 2]; q2 += part[(r * Cp + tid) * 2 + 1]; }
                const double m = a / (double)T;                                   // mean - pivot
                double var = (q2 - a * m) / (double)(T - 1);                      // sum (d - m)^2 = sum d^2 - (sum d)^2 / T
                var = var > 0.0 ? var : 0.0;
                // a column whose values are all equal has d == 0 everywhere: var == 0 exactly, like the two-pass reference
                dm_s[tid] = (float)m;
                rstd_s[tid] = var > 0.0 ? (float)(1.0 / sqrt(var)) : 0.f;
            }
            __syncthreads();
        }
        if (stat) {
            float dm[VW], rs[VW];
#pragma unroll
            for (int e = 0; e < VW; ++e) {
                const int c = c0 + e < C ? c0 + e : C - 1;
                dm[e] = tb.normalize ? dm_s[c] : 0.f;
                rs[e] = tb.normalize ? rstd_s[c] : 1.f;
            }
            for (int t = rg; t < T; t += nrg) {
                const uint4 pv = *reinterpret_cast<const uint4*>(raw + t * Cp + c0);
                TIn v[VW];
                memcpy(v, &pv, 16);
#pragma unroll
                for (int e = 0; e < VW; ++e) {
                    if (c0 + e >= C) continue;
                    float z;
                    if (!tb.normalize) z = (float)v[e];
                    else {
                        const float d = (float)(v[e] - piv[e]) - dm[e];
                        // (v - mean) / 0: 0/0 = NaN -> 0 (np.nan_to_num(nan=0.0)); +-x/0 = +-inf -> +-largest finite
                        z = rs[e] != 0.f ? d * rs[e] : (d == 0.f ? 0.f : (d > 0.f ? FLT_MAX : -FLT_MAX));
                    }
                    nz[(c0 + e) * Tp + t] = z;
                }
            }
        }
        __syncthreads();                                               // nz complete, raw free
        if (bulk && tid == 0 && g + gridDim.x < B) issue(g + gridDim.x);
        for (int b = warp; b < tb.n_blocks; b += WIN_THREADS / 32) {
            const int ty = bty_s[b];
            const int len = tb.blen[ty];
            float* dst = out.x[ty] + g * ((int64_t)tb.nodes[ty] * tb.blocks[ty] * len) + boff_s[b];
            const int c = tb.col[b];
            const float f = (float)tb.sign[b];
            if (c >= 0) {
                const float* nc = nz + c * Tp;
                for (int t0 = lane; t0 < len; t0 += 160) {            // 5 shared loads in flight, then 5 coalesced stores
                    float vv[5];
#pragma unroll
                    for (int it = 0; it < 5; ++it) vv[it] = t0 + 32 * it < len ? nc[t0 + 32 * it] * f : 0.f;
#pragma unroll
                    for (int it = 0; it < 5; ++it) if (t0 + 32 * it < len) dst[t0 + 32 * it] = vv[it];
                }
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
