// tcgen05 tensor-core kernels of the MS-HGNN hot path (MSHGNN_MODE_TC / MSHGNN_MODE_TC_1X), sm_100a only.
//
//  k_tc_rowgemm : the same tile/chunk program as k_rowgemm (kernels_simt.cuh) for slab inputs:
//        D[128 graphs, 128] = sum_chunks A_c[128, 128] * W_c[128, 128]^T + fused epilogue.
//     * operands are fp16; every activation / weight value v is stored as a pair (hi, lo) with
//       hi = fp16(v), lo = fp16(v - hi)  (~22 significant bits).  MODE_TC issues three MMAs per chunk
//       (hi*hi + lo*hi + hi*lo, fp32 accumulation in TMEM) which keeps forward pre-activations at fp32
//       accuracy - required because ReLU makes the gradient discontinuous in the forward numerics
//       (DESIGN.md "precision").  MODE_TC_1X issues hi*hi only.
//     * TMA (cp.async.bulk.tensor, SWIZZLE_128B) stages the slot tile picked by the gather table and the
//       weight tile straight into the UMMA shared-memory layout: the morphology gather never touches a
//       register.  One elected thread issues tcgen05.mma; the accumulator lives in TMEM; four epilogue
//       warps read it back with tcgen05.ld (one thread per graph row) and apply bias / ReLU / mask /
//       residual, then write the fp32 slab plus its (hi, lo) fp16 images for the next layer's TMA.
//  warp roles: 0 = TMA producer, 1 = MMA issuer + TMEM owner, 2..5 = epilogue.
#pragma once
#include <cuda.h>
#include <cuda_fp16.h>

#include "common.cuh"

namespace mshgnn {

constexpr int TC_STAGES = 3;
constexpr int TC_TILE_BYTES = 128 * 128;           // 128 rows x 64 fp16 (one 128B-swizzled K block)
constexpr int TC_STAGE_BYTES = 4 * TC_TILE_BYTES;  // A_hi, A_lo, W_hi, W_lo
constexpr int TC_SMEM_BYTES = TC_STAGES * TC_STAGE_BYTES + 1024 /*align slack*/ + 256 /*barriers*/;
constexpr int TC_THREADS = 192;
constexpr float TC_W_SCALE = 256.f;                // weights are stored as fp16 pairs of (w * 2^8)
constexpr float TC_W_UNSCALE = 1.f / 256.f;
// The low halves are stored multiplied by 2^11: lo = fp16((v - hi) * 2048), which is always a NORMAL fp16 when hi
// is (|lo*2048| <= |v|), so the pair keeps ~22 significant bits over fp16's whole normal range instead of hitting the
// 2^-24 subnormal floor for |v| < 0.25.  The cross terms therefore accumulate in a second TMEM accumulator (D1) and the
// epilogue combines D0 + D1 * 2^-11.
constexpr float TC_LO_SCALE = 2048.f;
constexpr float TC_LO_UNSCALE = 1.f / 2048.f;
constexpr uint32_t TC_TMEM_COLS = 256;

struct alignas(64) TcMaps {
    CUtensorMap a_hi, a_lo, w_hi, w_lo;
};

// ------------------------------------------------------------------------------------------
// PTX wrappers
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_LOOP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra WAIT_DONE;\n"
        "bra WAIT_LOOP;\n"
        "WAIT_DONE:\n"
        "}\n" ::"r"(bar), "r"(parity)
        : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(dst),
        "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(c0), "r"(c1)
        : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tmem_alloc(uint32_t dst_smem, uint32_t cols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t cols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(cols) : "memory");
}
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
        "}\n" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
          "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
          "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
          "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr)
        : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// K-major, 128B-swizzled shared-memory operand: 8-row groups 1024 B apart (SBO), descriptor version 1.
__device__ __forceinline__ uint64_t smem_desc_sw128(uint32_t saddr) {
    const uint32_t lo = ((saddr & 0x3FFFFu) >> 4) | (1u << 16);
    const uint32_t hi = (1024u >> 4) | (1u << 14) | (2u << 29);
    return ((uint64_t)hi << 32) | lo;
}
// kind::f16, A = B = fp16, D = fp32, both K-major, M = 128, N = 128
constexpr uint32_t TC_IDESC = (1u << 4) | ((128u >> 3) << 17) | ((128u >> 4) << 24);

__device__ __forceinline__ void split_store(__half* hi, __half* lo, int64_t off, const float (&v)[32]) {
    // 32 consecutive columns of one row -> 4 x 16-byte stores per image
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        __half2 h[4], l[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const float a = v[q * 8 + 2 * j], b = v[q * 8 + 2 * j + 1];
            const __half ha = __float2half_rn(a), hb = __float2half_rn(b);
            h[j] = __halves2half2(ha, hb);
            l[j] = __halves2half2(__float2half_rn((a - __half2float(ha)) * TC_LO_SCALE), __float2half_rn((b - __half2float(hb)) * TC_LO_SCALE));
        }
        *reinterpret_cast<uint4*>(hi + off + q * 8) = *reinterpret_cast<uint4*>(h);
        *reinterpret_cast<uint4*>(lo + off + q * 8) = *reinterpret_cast<uint4*>(l);
    }
}

// Epilogue shared by the row-GEMM and the encoder: one thread per graph row reads the fp32 accumulators from TMEM
// (D0 + D1 * 2^-11), applies bias / ReLU / stored-mask / residual and writes the fp32 slab plus its (hi, lo) images.
__device__ __forceinline__ void tc_epilogue_rows(const Tile& t, const BufTable& bt, const BufTable16& bh, const uint32_t tmem_base,
                                                 const int row0, const int64_t B, const int64_t Bp, const int split, const int warp, const int lane) {
    const int q = warp & 3;                       // TMEM lane quarter this warp may access
    const int64_t row = row0 + q * 32 + lane;
    const bool live = row < B;
#pragma unroll 1
    for (int cc = 0; cc < 4; ++cc) {
        uint32_t raw[32], raw1[32];
        tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + cc * 32, raw);
        if (split) tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + 128 + cc * 32, raw1);
        const int col = cc * 32;
        if (!live) {
            // rows [B, Bp) of every fp16 image are kept at zero: the weight-gradient kernel reduces over whole
            // 64-row blocks (k_tc_reducegemm) and must not see stale data there
            if (row < Bp) {
                const uint4 z = make_uint4(0u, 0u, 0u, 0u);
                if (t.out_buf >= 0) {
                    const int64_t off = ((int64_t)t.out_slot * Bp + row) * H + col;
#pragma unroll
                    for (int j = 0; j < 4; ++j) { *reinterpret_cast<uint4*>(bh.hi[t.out_buf] + off + j * 8) = z; *reinterpret_cast<uint4*>(bh.lo[t.out_buf] + off + j * 8) = z; }
                }
                if (t.out2_buf >= 0) {
                    const int64_t off = ((int64_t)t.out2_slot * Bp + row) * H + col;
#pragma unroll
                    for (int j = 0; j < 4; ++j) { *reinterpret_cast<uint4*>(bh.hi[t.out2_buf] + off + j * 8) = z; *reinterpret_cast<uint4*>(bh.lo[t.out2_buf] + off + j * 8) = z; }
                }
            }
            continue;
        }
        float v[32];
        unsigned mask = 0;
#pragma unroll
        for (int j = 0; j < 32; ++j) {
            float x = __uint_as_float(raw[j]);
            if (split) x = fmaf(__uint_as_float(raw1[j]), TC_LO_UNSCALE, x);
            x *= TC_W_UNSCALE;
            if (t.bias_buf >= 0) x += __ldg((const float*)bt.p[t.bias_buf] + t.bias_off + col + j);
            if (x > 0.f) mask |= 1u << j;
            if (t.relu) x = fmaxf(x, 0.f);
            v[j] = x;
        }
        if (t.mask_out_buf >= 0)
            *((unsigned*)bt.p[t.mask_out_buf] + ((int64_t)t.mask_out_slot * Bp + row) * 4 + cc) = mask;
        if (t.posmask_buf >= 0) {
            const unsigned w = *((const unsigned*)bt.p[t.posmask_buf] + ((int64_t)t.posmask_slot * Bp + row) * 4 + cc);
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = ((w >> j) & 1u) ? v[j] : 0.f;
        }
        if (t.res_buf >= 0) {
            const float4* p = reinterpret_cast<const float4*>((const float*)bt.p[t.res_buf] + ((int64_t)t.res_slot * Bp + row) * H + col);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const float4 m = p[j];
                v[4 * j] += m.x; v[4 * j + 1] += m.y; v[4 * j + 2] += m.z; v[4 * j + 3] += m.w;
            }
        }
        if (t.out_buf >= 0) {
            const int64_t off = ((int64_t)t.out_slot * Bp + row) * H + col;
            float4* o = reinterpret_cast<float4*>((float*)bt.p[t.out_buf] + off);
#pragma unroll
            for (int j = 0; j < 8; ++j) o[j] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
            split_store(bh.hi[t.out_buf], bh.lo[t.out_buf], off, v);
        }
        if (t.out2_buf >= 0) {
            if (t.out2_mask_kind == MK_BITS) {
                const unsigned w = *((const unsigned*)bt.p[t.out2_mask_buf] + ((int64_t)t.out2_mask_slot * Bp + row) * 4 + cc);
#pragma unroll
                for (int j = 0; j < 32; ++j) v[j] = ((w >> j) & 1u) ? v[j] : 0.f;
            }
            const int64_t off = ((int64_t)t.out2_slot * Bp + row) * H + col;
            float4* o = reinterpret_cast<float4*>((float*)bt.p[t.out2_buf] + off);
#pragma unroll
            for (int j = 0; j < 8; ++j) o[j] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
            split_store(bh.hi[t.out2_buf], bh.lo[t.out2_buf], off, v);
        }
    }
}

// ------------------------------------------------------------------------------------------
// row-GEMM on tcgen05
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(TC_THREADS, 1)
k_tc_rowgemm(const __grid_constant__ TcMaps maps, const Tile* __restrict__ tiles, const BufTable bt, const BufTable16 bh,
             const int64_t B, const int64_t Bp, const int split) {
    extern __shared__ uint8_t smem_raw[];
    __shared__ Tile t;
    __shared__ uint32_t tmem_base_s;

    const int tid = threadIdx.x;
    const int warp = tid >> 5, lane = tid & 31;
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + TC_STAGES * TC_STAGE_BYTES);
    const uint32_t full0 = smem_u32(bars), empty0 = smem_u32(bars + TC_STAGES), accum_bar = smem_u32(bars + 2 * TC_STAGES);
    const uint32_t smem_base = smem_u32(smem);

    {
        const int* src = reinterpret_cast<const int*>(tiles + blockIdx.y);
        int* dst = reinterpret_cast<int*>(&t);
        for (int i = tid; i < (int)(sizeof(Tile) / 4); i += TC_THREADS) dst[i] = src[i];
    }
    if (tid == 0) {
        for (int s = 0; s < TC_STAGES; ++s) { mbar_init(full0 + 8 * s, 1); mbar_init(empty0 + 8 * s, 1); }
        mbar_init(accum_bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) tmem_alloc(smem_u32(&tmem_base_s), TC_TMEM_COLS);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_base_s;

    const int row0 = blockIdx.x * TILE_M;
    const int n_steps = t.n_chunks * 2;   // two 64-wide K blocks per 128-wide chunk

    if (warp == 0) {
        if (lane == 0) {
            const uint32_t tx_bytes = split ? 4 * TC_TILE_BYTES : 2 * TC_TILE_BYTES;
            for (int i = 0; i < n_steps; ++i) {
                const int s = i % TC_STAGES;
                mbar_wait(empty0 + 8 * s, ((i / TC_STAGES) & 1) ^ 1);
                const Chunk& ch = t.chunks[i >> 1];
                const int kcol = (i & 1) * 64;
                const int arow = (int)((int64_t)ch.a_slot * Bp + row0);
                const uint32_t st = smem_base + s * TC_STAGE_BYTES;
                const uint32_t fb = full0 + 8 * s;
                mbar_expect_tx(fb, tx_bytes);
                tma_load_2d(st, &maps.a_hi, fb, kcol, arow);
                tma_load_2d(st + 2 * TC_TILE_BYTES, &maps.w_hi, fb, kcol, ch.w16_row);
                if (split) {
                    tma_load_2d(st + TC_TILE_BYTES, &maps.a_lo, fb, kcol, arow);
                    tma_load_2d(st + 3 * TC_TILE_BYTES, &maps.w_lo, fb, kcol, ch.w16_row);
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            for (int i = 0; i < n_steps; ++i) {
                const int s = i % TC_STAGES;
                mbar_wait(full0 + 8 * s, (i / TC_STAGES) & 1);
                tc_fence_after();
                const uint32_t st = smem_base + s * TC_STAGE_BYTES;
                const uint64_t a_hi = smem_desc_sw128(st), a_lo = smem_desc_sw128(st + TC_TILE_BYTES);
                const uint64_t w_hi = smem_desc_sw128(st + 2 * TC_TILE_BYTES), w_lo = smem_desc_sw128(st + 3 * TC_TILE_BYTES);
#pragma unroll
                for (int ks = 0; ks < 4; ++ks) {
                    const uint64_t adv = (uint64_t)(ks * 2);      // +32 bytes (16 fp16) along K inside the swizzle atom
                    umma_f16(tmem_base, a_hi + adv, w_hi + adv, TC_IDESC, (i | ks) ? 1u : 0u);             // D0 += hi * hi
                    if (split) {
                        umma_f16(tmem_base + 128, a_lo + adv, w_hi + adv, TC_IDESC, (i | ks) ? 1u : 0u);   // D1 += lo * hi
                        umma_f16(tmem_base + 128, a_hi + adv, w_lo + adv, TC_IDESC, 1u);                   // D1 += hi * lo
                    }
                }
                umma_commit(empty0 + 8 * s);          // frees the stage once these MMAs have read it
            }
            umma_commit(accum_bar);                   // accumulator complete
        }
        __syncwarp();
    } else {
        // ---------------- epilogue: one thread per graph row ----------------
        mbar_wait(accum_bar, 0);
        tc_fence_after();
        tc_epilogue_rows(t, bt, bh, tmem_base, row0, B, Bp, split, warp, lane);
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, TC_TMEM_COLS);
    }
}

// ------------------------------------------------------------------------------------------
// reduce-over-rows GEMM on tcgen05 (weight gradients):  dW[128 out, 128 in] = sum_pairs dC[rows,128]^T * A[rows,128]
// ------------------------------------------------------------------------------------------
//  The reduction dimension is the graph-row dimension, so both MMA operands are "MN-major": the fp16 images are
//  [row][feature] and a TMA box of 64 features x 64 rows (SWIZZLE_128B) is exactly the canonical MN-major SW128 atom
//  sequence (8 rows x 128 B per atom, SBO = 1024 B between 8-row groups, LBO = 8192 B between the two 64-feature
//  halves).  No transposed copy of any activation is ever made.
//  Bias gradients (column sums of dC) ride on the same operand: one extra N=16 MMA against an all-ones B tile.
//  Accumulators (TMEM columns): D0 [0,128) hi*hi, D1 [128,256) cross terms (x 2^11), D2 [256,272) colsum hi,
//  D3 [288,304) colsum lo (x 2^11).  One CTA = one task (<= 4 pairs) x one row split; fp32 partials go to part_w /
//  part_b and are summed in double by k_reduce_partials (deterministic, no atomics).
constexpr int BUF_DC1_ID = 11;     // plan.cuh BUF_DC1: dpre of the encoder lives there after the layer-0 dX launch
struct BufRows {                 // first 256-byte row of each fp16 image, relative to the workspace base
    int hi[MAX_BUFS];
    int lo[MAX_BUFS];
};

constexpr int DW_STAGES = 3;
constexpr int DW_KB = 64;                              // graph rows (MMA K) per pipeline stage
constexpr int DW_IMG_BYTES = DW_KB * H * 2;            // 64 rows x 128 fp16 = two 64x64 boxes
constexpr int DW_STAGE_BYTES = 4 * DW_IMG_BYTES;       // dC_hi, dC_lo, A_hi, A_lo
constexpr int DW_ONES_BYTES = 2048;
constexpr int DW_SMEM_BYTES = DW_STAGES * DW_STAGE_BYTES + DW_ONES_BYTES + 1024 /*align slack*/ + 256 /*barriers*/;
constexpr uint32_t DW_TMEM_COLS = 512;
constexpr int DW_MAX_PAIRS = 4;

// MN-major, 128B-swizzled operand of 128 (MN) x 16 (K) fp16: LBO = 8192 B (next 64 MN elements), SBO = 1024 B (next 8 K rows)
__device__ __forceinline__ uint64_t smem_desc_mn_sw128(uint32_t saddr) {
    const uint32_t lo = ((saddr & 0x3FFFFu) >> 4) | ((uint32_t)(DW_IMG_BYTES / 2 >> 4) << 16);
    const uint32_t hi = (1024u >> 4) | (1u << 14) | (2u << 29);
    return ((uint64_t)hi << 32) | lo;
}
// kind::f16, fp16 x fp16 -> fp32, A and B MN-major, M = 128, N = 128
constexpr uint32_t DW_IDESC = (1u << 4) | (1u << 15) | (1u << 16) | ((128u >> 3) << 17) | ((128u >> 4) << 24);
// column sums: A MN-major (dC), B K-major (all ones), M = 128, N = 16
constexpr uint32_t DW_IDESC_CS = (1u << 4) | (1u << 15) | ((16u >> 3) << 17) | ((128u >> 4) << 24);

__global__ void __launch_bounds__(TC_THREADS, 1)
k_tc_reducegemm(const __grid_constant__ CUtensorMap map, const RTask* __restrict__ tasks, const RPair* __restrict__ pairs,
                const int task0, const BufRows br, const int64_t B, const int64_t Bp, const int rows_per, const int n_splits,
                const int split, float* __restrict__ part_w, float* __restrict__ part_b) {
    extern __shared__ uint8_t smem_raw[];
    __shared__ RTask t;
    __shared__ RPair prs[DW_MAX_PAIRS];
    __shared__ uint32_t tmem_base_s;

    const int tid = threadIdx.x;
    const int warp = tid >> 5, lane = tid & 31;
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint8_t* ones = smem + DW_STAGES * DW_STAGE_BYTES;
    uint64_t* bars = reinterpret_cast<uint64_t*>(ones + DW_ONES_BYTES);
    const uint32_t full0 = smem_u32(bars), empty0 = smem_u32(bars + DW_STAGES), accum_bar = smem_u32(bars + 2 * DW_STAGES);
    const uint32_t smem_base = smem_u32(smem);

    const int task = task0 + blockIdx.x;
    if (tid < (int)(sizeof(RTask) / 4)) reinterpret_cast<int*>(&t)[tid] = reinterpret_cast<const int*>(tasks + task)[tid];
    for (int i = tid; i < DW_ONES_BYTES / 4; i += TC_THREADS) reinterpret_cast<uint32_t*>(ones)[i] = 0x3C003C00u;   // fp16 1.0 pairs
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");        // the ones tile is read by the MMA (async proxy)
    if (tid == 0) {
        for (int s = 0; s < DW_STAGES; ++s) { mbar_init(full0 + 8 * s, 1); mbar_init(empty0 + 8 * s, 1); }
        mbar_init(accum_bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) tmem_alloc(smem_u32(&tmem_base_s), DW_TMEM_COLS);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    {
        const int np = t.n_pairs < DW_MAX_PAIRS ? t.n_pairs : DW_MAX_PAIRS;
        const int* src = reinterpret_cast<const int*>(pairs + t.pair_begin);
        int* dst = reinterpret_cast<int*>(prs);
        for (int i = tid; i < np * (int)(sizeof(RPair) / 4); i += TC_THREADS) dst[i] = src[i];
    }
    __syncthreads();
    const uint32_t tmem_base = tmem_base_s;

    const int sp = blockIdx.y;
    const int64_t r_begin = (int64_t)sp * rows_per;
    const int64_t r_end = (r_begin + rows_per < B) ? (r_begin + rows_per) : B;
    const int n_kb = r_end > r_begin ? (int)((r_end - r_begin + DW_KB - 1) / DW_KB) : 0;
    const int n_pairs = t.n_pairs < DW_MAX_PAIRS ? t.n_pairs : DW_MAX_PAIRS;
    const int n_steps = n_kb * n_pairs;
    const int want_cs = t.want_colsum;

    if (warp == 0) {
        if (lane == 0) {
            const uint32_t tx_bytes = split ? 4 * DW_IMG_BYTES : 2 * DW_IMG_BYTES;
            for (int i = 0; i < n_steps; ++i) {
                const int s = i % DW_STAGES;
                mbar_wait(empty0 + 8 * s, ((i / DW_STAGES) & 1) ^ 1);
                const RPair& pr = prs[i / n_kb];
                const int r0 = (int)(r_begin + (int64_t)(i % n_kb) * DW_KB);
                const int d_off = (int)((int64_t)pr.d_slot * Bp) + r0, a_off = (int)((int64_t)pr.a_slot * Bp) + r0;
                const uint32_t st = smem_base + s * DW_STAGE_BYTES;
                const uint32_t fb = full0 + 8 * s;
                mbar_expect_tx(fb, tx_bytes);
                tma_load_2d(st, &map, fb, 0, br.hi[pr.d_buf] + d_off);
                tma_load_2d(st + DW_IMG_BYTES / 2, &map, fb, 64, br.hi[pr.d_buf] + d_off);
                tma_load_2d(st + 2 * DW_IMG_BYTES, &map, fb, 0, br.hi[pr.a_buf] + a_off);
                tma_load_2d(st + 2 * DW_IMG_BYTES + DW_IMG_BYTES / 2, &map, fb, 64, br.hi[pr.a_buf] + a_off);
                if (split) {
                    tma_load_2d(st + DW_IMG_BYTES, &map, fb, 0, br.lo[pr.d_buf] + d_off);
                    tma_load_2d(st + DW_IMG_BYTES + DW_IMG_BYTES / 2, &map, fb, 64, br.lo[pr.d_buf] + d_off);
                    tma_load_2d(st + 3 * DW_IMG_BYTES, &map, fb, 0, br.lo[pr.a_buf] + a_off);
                    tma_load_2d(st + 3 * DW_IMG_BYTES + DW_IMG_BYTES / 2, &map, fb, 64, br.lo[pr.a_buf] + a_off);
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            const uint64_t ones_desc = smem_desc_sw128(smem_u32(ones));
            for (int i = 0; i < n_steps; ++i) {
                const int s = i % DW_STAGES;
                mbar_wait(full0 + 8 * s, (i / DW_STAGES) & 1);
                tc_fence_after();
                const uint32_t st = smem_base + s * DW_STAGE_BYTES;
                const uint64_t d_hi = smem_desc_mn_sw128(st), d_lo = smem_desc_mn_sw128(st + DW_IMG_BYTES);
                const uint64_t a_hi = smem_desc_mn_sw128(st + 2 * DW_IMG_BYTES), a_lo = smem_desc_mn_sw128(st + 3 * DW_IMG_BYTES);
#pragma unroll
                for (int ks = 0; ks < DW_KB / 16; ++ks) {
                    const uint64_t adv = (uint64_t)(ks * (2048 >> 4));          // 16 rows = two 8-row groups of 1024 B
                    const uint32_t acc = (i | ks) ? 1u : 0u;
                    umma_f16(tmem_base, d_hi + adv, a_hi + adv, DW_IDESC, acc);                       // D0 += dC_hi^T A_hi
                    if (split) {
                        umma_f16(tmem_base + 128, d_lo + adv, a_hi + adv, DW_IDESC, acc);             // D1 += dC_lo^T A_hi
                        umma_f16(tmem_base + 128, d_hi + adv, a_lo + adv, DW_IDESC, 1u);              // D1 += dC_hi^T A_lo
                    }
                    if (want_cs) {
                        umma_f16(tmem_base + 256, d_hi + adv, ones_desc, DW_IDESC_CS, acc);           // D2 += dC_hi^T 1
                        if (split) umma_f16(tmem_base + 288, d_lo + adv, ones_desc, DW_IDESC_CS, acc); // D3 += dC_lo^T 1
                    }
                }
                umma_commit(empty0 + 8 * s);
            }
            umma_commit(accum_bar);
        }
        __syncwarp();
    } else {
        // ---------------- epilogue: one thread per output feature (TMEM lane) ----------------
        mbar_wait(accum_bar, 0);
        tc_fence_after();
        const int q = warp & 3;
        const int o = q * 32 + lane;
        const int64_t slot = (int64_t)task * n_splits + sp;
        float* pw = part_w + slot * (H * H) + (int64_t)o * H;
#pragma unroll 1
        for (int cc = 0; cc < 4; ++cc) {
            uint32_t raw[32], raw1[32];
            tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + cc * 32, raw);
            if (split) tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + 128 + cc * 32, raw1);
            float v[32];
#pragma unroll
            for (int j = 0; j < 32; ++j) {
                float x = __uint_as_float(raw[j]);
                if (split) x = fmaf(__uint_as_float(raw1[j]), TC_LO_UNSCALE, x);
                v[j] = n_steps ? x : 0.f;
            }
            float4* dst = reinterpret_cast<float4*>(pw + cc * 32);
#pragma unroll
            for (int j = 0; j < 8; ++j) dst[j] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
        }
        if (want_cs) {
            uint32_t raw[32], raw1[32];
            tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + 256, raw);
            float x = __uint_as_float(raw[0]);
            if (split) { tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + 288, raw1); x = fmaf(__uint_as_float(raw1[0]), TC_LO_UNSCALE, x); }
            part_b[slot * H + o] = n_steps ? x : 0.f;
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, DW_TMEM_COLS);
    }
}

// ------------------------------------------------------------------------------------------
// encoder on tcgen05:  h0[slot] = relu((x[slot] * sign[slot]) W_enc[type]^T + b)          (hgnn_k4.py:L159-160, L198-237)
// ------------------------------------------------------------------------------------------
//  The caller's fp32 (or fp64) feature rows are the only HBM stream of the whole model.  Eight loader warps read them
//  with 128-bit coalesced loads, fold the +-1 symmetry signs in, split every value into the (hi, lo) fp16 pair and write
//  it straight into the 128B-swizzled K-major UMMA operand layout in shared memory; the weight tiles come by TMA from
//  the padded fp16 weight image [n_types*128][enc_kmax].  Two loader groups alternate K blocks so that two blocks of
//  loads are in flight per SM.  warp roles: 0 = TMA (weights), 1 = MMA issuer, 2..9 = loaders, 2..5 = epilogue.
constexpr int ENC_THREADS = 320;
constexpr int ENC_LOADER_WARPS = 4;               // per group

struct alignas(64) EncMaps {
    CUtensorMap w_hi, w_lo;
};

__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// 4 consecutive values of one row -> 8 bytes in the hi tile and 8 bytes in the lo tile
__device__ __forceinline__ void split_to_smem(uint32_t hi_addr, uint32_t lo_addr, const float4 v) {
    const __half h0 = __float2half_rn(v.x), h1 = __float2half_rn(v.y), h2 = __float2half_rn(v.z), h3 = __float2half_rn(v.w);
    const __half2 a = __halves2half2(h0, h1), b = __halves2half2(h2, h3);
    const __half2 c = __halves2half2(__float2half_rn((v.x - __half2float(h0)) * TC_LO_SCALE), __float2half_rn((v.y - __half2float(h1)) * TC_LO_SCALE));
    const __half2 d = __halves2half2(__float2half_rn((v.z - __half2float(h2)) * TC_LO_SCALE), __float2half_rn((v.w - __half2float(h3)) * TC_LO_SCALE));
    asm volatile("st.shared.v2.b32 [%0], {%1, %2};" ::"r"(hi_addr), "r"(*reinterpret_cast<const uint32_t*>(&a)), "r"(*reinterpret_cast<const uint32_t*>(&b)) : "memory");
    asm volatile("st.shared.v2.b32 [%0], {%1, %2};" ::"r"(lo_addr), "r"(*reinterpret_cast<const uint32_t*>(&c)), "r"(*reinterpret_cast<const uint32_t*>(&d)) : "memory");
}

// 4 consecutive feature values x[row][k .. k+3] (zero outside [0, K) / beyond the batch), times their signs
__device__ __forceinline__ float4 load_x4(const void* xb, const int x_f64, const int64_t idx, const int k, const int K, const bool row_ok,
                                          const float4 sg) {
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (row_ok && k < K) {
        if (!x_f64) {
            const float* pf = (const float*)xb + idx;
            if (k + 3 < K && ((reinterpret_cast<uintptr_t>(pf) & 15) == 0)) {
                v = __ldg(reinterpret_cast<const float4*>(pf));
            } else {
                v.x = __ldg(pf);
                if (k + 1 < K) v.y = __ldg(pf + 1);
                if (k + 2 < K) v.z = __ldg(pf + 2);
                if (k + 3 < K) v.w = __ldg(pf + 3);
            }
        } else {
            const double* pd = (const double*)xb + idx;
            v.x = (float)__ldg(pd);
            if (k + 1 < K) v.y = (float)__ldg(pd + 1);
            if (k + 2 < K) v.z = (float)__ldg(pd + 2);
            if (k + 3 < K) v.w = (float)__ldg(pd + 3);
        }
        v.x *= sg.x; v.y *= sg.y; v.z *= sg.z; v.w *= sg.w;
    }
    return v;
}

__device__ __forceinline__ float4 load_sign4(const float* signs, const int sign_off, const int k, const int K) {
    float4 sg = make_float4(1.f, 1.f, 1.f, 1.f);
    if (sign_off >= 0 && k < K) {
        const float* sp = signs + sign_off + k;
        sg.x = __ldg(sp);
        if (k + 1 < K) sg.y = __ldg(sp + 1);
        if (k + 2 < K) sg.z = __ldg(sp + 2);
        if (k + 3 < K) sg.w = __ldg(sp + 3);
    }
    return sg;
}

__global__ void __launch_bounds__(ENC_THREADS, 1)
k_tc_encoder(const __grid_constant__ EncMaps maps, const Tile* __restrict__ tiles, const BufTable bt, const BufTable16 bh,
             const int64_t B, const int64_t Bp, const int x_f64, const int split) {
    extern __shared__ uint8_t smem_raw[];
    __shared__ Tile t;
    __shared__ uint32_t tmem_base_s;

    const int tid = threadIdx.x;
    const int warp = tid >> 5, lane = tid & 31;
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + TC_STAGES * TC_STAGE_BYTES);
    const uint32_t full0 = smem_u32(bars), empty0 = smem_u32(bars + TC_STAGES), accum_bar = smem_u32(bars + 2 * TC_STAGES);
    const uint32_t smem_base = smem_u32(smem);
    {
        const int* src = reinterpret_cast<const int*>(tiles + blockIdx.y);
        int* dst = reinterpret_cast<int*>(&t);
        for (int i = tid; i < (int)(sizeof(Tile) / 4); i += ENC_THREADS) dst[i] = src[i];
    }
    if (tid == 0) {
        for (int s = 0; s < TC_STAGES; ++s) { mbar_init(full0 + 8 * s, 1 + ENC_LOADER_WARPS); mbar_init(empty0 + 8 * s, 1); }
        mbar_init(accum_bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) tmem_alloc(smem_u32(&tmem_base_s), TC_TMEM_COLS);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_base_s;

    const int row0 = blockIdx.x * TILE_M;
    const int K = t.chunks[0].K;
    const int n_kb = (K + 63) / 64;

    if (warp == 0) {
        if (lane == 0) {
            const uint32_t tx_bytes = split ? 2 * TC_TILE_BYTES : TC_TILE_BYTES;
            const int wrow = t.chunks[0].w16_row;
            for (int i = 0; i < n_kb; ++i) {
                const int s = i % TC_STAGES;
                mbar_wait(empty0 + 8 * s, ((i / TC_STAGES) & 1) ^ 1);
                const uint32_t st = smem_base + s * TC_STAGE_BYTES;
                const uint32_t fb = full0 + 8 * s;
                mbar_expect_tx(fb, tx_bytes);
                tma_load_2d(st + 2 * TC_TILE_BYTES, &maps.w_hi, fb, i * 64, wrow);
                if (split) tma_load_2d(st + 3 * TC_TILE_BYTES, &maps.w_lo, fb, i * 64, wrow);
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            for (int i = 0; i < n_kb; ++i) {
                const int s = i % TC_STAGES;
                mbar_wait(full0 + 8 * s, (i / TC_STAGES) & 1);
                tc_fence_after();
                const uint32_t st = smem_base + s * TC_STAGE_BYTES;
                const uint64_t a_hi = smem_desc_sw128(st), a_lo = smem_desc_sw128(st + TC_TILE_BYTES);
                const uint64_t w_hi = smem_desc_sw128(st + 2 * TC_TILE_BYTES), w_lo = smem_desc_sw128(st + 3 * TC_TILE_BYTES);
#pragma unroll
                for (int ks = 0; ks < 4; ++ks) {
                    const uint64_t adv = (uint64_t)(ks * 2);
                    umma_f16(tmem_base, a_hi + adv, w_hi + adv, TC_IDESC, (i | ks) ? 1u : 0u);
                    if (split) {
                        umma_f16(tmem_base + 128, a_lo + adv, w_hi + adv, TC_IDESC, (i | ks) ? 1u : 0u);
                        umma_f16(tmem_base + 128, a_hi + adv, w_lo + adv, TC_IDESC, 1u);
                    }
                }
                umma_commit(empty0 + 8 * s);
            }
            umma_commit(accum_bar);
        }
        __syncwarp();
    } else {
        // ---------------- loaders: group g takes the K blocks kb = g (mod 2) ----------------
        const int g = (warp - 2) / ENC_LOADER_WARPS;
        const int gt = tid - 64 - g * (ENC_LOADER_WARPS * 32);
        const Chunk& ch = t.chunks[0];
        const void* xb = bt.p[ch.a_buf];
        const float* signs = (const float*)bt.p[2];
        const int kq = gt & 15;                                   // which 4-column group of the 64-column block
        const int rsub = gt >> 4;                                 // 0..7
        for (int kb = g; kb < n_kb; kb += 2) {
            const int s = kb % TC_STAGES;
            const int k = kb * 64 + kq * 4;
            const float4 sg = load_sign4(signs, ch.sign_off, k, K);
            float4 v[16];
#pragma unroll
            for (int it = 0; it < 16; ++it) {
                const int64_t row = row0 + it * 8 + rsub;
                v[it] = load_x4(xb, x_f64, row * (int64_t)ch.lda + ch.a_off + k, k, K, row < B, sg);
            }
            mbar_wait(empty0 + 8 * s, ((kb / TC_STAGES) & 1) ^ 1);
            const uint32_t st = smem_base + s * TC_STAGE_BYTES;
#pragma unroll
            for (int it = 0; it < 16; ++it) {
                const int r = it * 8 + rsub;
                const uint32_t off = (uint32_t)(r * 128 + ((((kq >> 1) ^ (r & 7))) << 4) + ((kq & 1) << 3));
                split_to_smem(st + off, st + TC_TILE_BYTES + off, v[it]);
            }
            fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) mbar_arrive(full0 + 8 * s);
        }
        if (warp < 6) {
            // ---------------- epilogue (warps 2..5 = TMEM lane quarters 2, 3, 0, 1) ----------------
            mbar_wait(accum_bar, 0);
            tc_fence_after();
            tc_epilogue_rows(t, bt, bh, tmem_base, row0, B, Bp, split, warp, lane);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, TC_TMEM_COLS);
    }
}

// ------------------------------------------------------------------------------------------
// encoder weight gradient on tcgen05:  dW_enc[t][:, k0 : k0+64*nkb] = sum_slots dpre[slot]^T (x[slot] * sign[slot])
// ------------------------------------------------------------------------------------------
//  M = 128 output features (dpre images, MN-major via TMA like k_tc_reducegemm), N = 64*nkb input columns (the loaders
//  write the split x rows as 64-column MN-major blocks, LBO = 8192 B), K = graph rows.  x is read exactly once per step
//  over all units.  Accumulators: D0 [0,192) hi*hi, D1 [192,384) cross (x 2^11), D2 [384,400) / D3 [416,432) bias sums.
constexpr int EDW_STAGES = 2;
constexpr int EDW_DC_BYTES = DW_IMG_BYTES;                 // 64 rows x 128 features fp16
constexpr int EDW_X_BYTES = 3 * 64 * 128;                  // up to three 64-row x 64-column blocks
constexpr int EDW_STAGE_BYTES = 2 * EDW_DC_BYTES + 2 * EDW_X_BYTES;     // dC_hi, dC_lo, X_hi, X_lo = 80 KB
constexpr int EDW_SMEM_BYTES = EDW_STAGES * EDW_STAGE_BYTES + DW_ONES_BYTES + 1024 + 256;
constexpr int EDW_NMAX = 192;

__global__ void __launch_bounds__(ENC_THREADS, 1)
k_tc_encoder_dw(const __grid_constant__ CUtensorMap map, const EncDwUnit* __restrict__ units, const BufTable bt, const BufRows br,
                const int64_t B, const int64_t Bp, const int rows_per, const int n_splits, const int x_f64, const int split,
                float* __restrict__ part_w, float* __restrict__ part_b) {
    extern __shared__ uint8_t smem_raw[];
    __shared__ EncDwUnit u;
    __shared__ uint32_t tmem_base_s;

    const int tid = threadIdx.x;
    const int warp = tid >> 5, lane = tid & 31;
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint8_t* ones = smem + EDW_STAGES * EDW_STAGE_BYTES;
    uint64_t* bars = reinterpret_cast<uint64_t*>(ones + DW_ONES_BYTES);
    const uint32_t full0 = smem_u32(bars), empty0 = smem_u32(bars + EDW_STAGES), accum_bar = smem_u32(bars + 2 * EDW_STAGES);
    const uint32_t smem_base = smem_u32(smem);

    if (tid < (int)(sizeof(EncDwUnit) / 4)) reinterpret_cast<int*>(&u)[tid] = reinterpret_cast<const int*>(units + blockIdx.x)[tid];
    for (int i = tid; i < DW_ONES_BYTES / 4; i += ENC_THREADS) reinterpret_cast<uint32_t*>(ones)[i] = 0x3C003C00u;
    fence_proxy_async_smem();
    if (tid == 0) {
        for (int s = 0; s < EDW_STAGES; ++s) { mbar_init(full0 + 8 * s, 1 + ENC_LOADER_WARPS); mbar_init(empty0 + 8 * s, 1); }
        mbar_init(accum_bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) tmem_alloc(smem_u32(&tmem_base_s), DW_TMEM_COLS);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_base_s;

    const int sp = blockIdx.y;
    const int64_t r_begin = (int64_t)sp * rows_per;
    const int64_t r_end = (r_begin + rows_per < B) ? (r_begin + rows_per) : B;
    const int n_rb = r_end > r_begin ? (int)((r_end - r_begin + DW_KB - 1) / DW_KB) : 0;
    const int n_steps = n_rb * u.n_slots;
    const int nkb = u.nkb;
    const int want_cs = u.want_colsum;
    const uint32_t idesc = (1u << 4) | (1u << 15) | (1u << 16) | ((uint32_t)(nkb * 64 >> 3) << 17) | ((128u >> 4) << 24);

    if (warp == 0) {
        if (lane == 0) {
            const uint32_t tx_bytes = split ? 2 * EDW_DC_BYTES : EDW_DC_BYTES;
            for (int i = 0; i < n_steps; ++i) {
                const int s = i % EDW_STAGES;
                mbar_wait(empty0 + 8 * s, ((i / EDW_STAGES) & 1) ^ 1);
                const int r0 = (int)(r_begin + (int64_t)(i % n_rb) * DW_KB);
                const int d_off = (int)((int64_t)u.d_slot[i / n_rb] * Bp) + r0;
                const uint32_t st = smem_base + s * EDW_STAGE_BYTES;
                const uint32_t fb = full0 + 8 * s;
                mbar_expect_tx(fb, tx_bytes);
                tma_load_2d(st, &map, fb, 0, br.hi[BUF_DC1_ID] + d_off);
                tma_load_2d(st + EDW_DC_BYTES / 2, &map, fb, 64, br.hi[BUF_DC1_ID] + d_off);
                if (split) {
                    tma_load_2d(st + EDW_DC_BYTES, &map, fb, 0, br.lo[BUF_DC1_ID] + d_off);
                    tma_load_2d(st + EDW_DC_BYTES + EDW_DC_BYTES / 2, &map, fb, 64, br.lo[BUF_DC1_ID] + d_off);
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            const uint64_t ones_desc = smem_desc_sw128(smem_u32(ones));
            for (int i = 0; i < n_steps; ++i) {
                const int s = i % EDW_STAGES;
                mbar_wait(full0 + 8 * s, (i / EDW_STAGES) & 1);
                tc_fence_after();
                const uint32_t st = smem_base + s * EDW_STAGE_BYTES;
                const uint64_t d_hi = smem_desc_mn_sw128(st), d_lo = smem_desc_mn_sw128(st + EDW_DC_BYTES);
                const uint64_t x_hi = smem_desc_mn_sw128(st + 2 * EDW_DC_BYTES), x_lo = smem_desc_mn_sw128(st + 2 * EDW_DC_BYTES + EDW_X_BYTES);
#pragma unroll
                for (int ks = 0; ks < DW_KB / 16; ++ks) {
                    const uint64_t adv = (uint64_t)(ks * (2048 >> 4));
                    const uint32_t acc = (i | ks) ? 1u : 0u;
                    umma_f16(tmem_base, d_hi + adv, x_hi + adv, idesc, acc);
                    if (split) {
                        umma_f16(tmem_base + EDW_NMAX, d_lo + adv, x_hi + adv, idesc, acc);
                        umma_f16(tmem_base + EDW_NMAX, d_hi + adv, x_lo + adv, idesc, 1u);
                    }
                    if (want_cs) {
                        umma_f16(tmem_base + 384, d_hi + adv, ones_desc, DW_IDESC_CS, acc);
                        if (split) umma_f16(tmem_base + 416, d_lo + adv, ones_desc, DW_IDESC_CS, acc);
                    }
                }
                umma_commit(empty0 + 8 * s);
            }
            umma_commit(accum_bar);
        }
        __syncwarp();
    } else {
        // ---------------- loaders: group g takes the steps i = g (mod 2) ----------------
        const int g = (warp - 2) / ENC_LOADER_WARPS;
        const int gt = tid - 64 - g * (ENC_LOADER_WARPS * 32);
        const void* xb = bt.p[u.x_buf];
        const float* signs = (const float*)bt.p[2];
        const int kq = gt & 15, rsub = gt >> 4;
        const int K = u.K;
        for (int i = g; i < n_steps; i += 2) {
            const int s = i % EDW_STAGES;
            const int j = i / n_rb;
            const int64_t r0 = r_begin + (int64_t)(i % n_rb) * DW_KB;
            const uint32_t st = smem_base + s * EDW_STAGE_BYTES + 2 * EDW_DC_BYTES;
            bool waited = false;
            for (int jb = 0; jb < nkb; ++jb) {
                const int k = u.k0 + jb * 64 + kq * 4;
                const float4 sg = load_sign4(signs, u.sign_off[j], k, K);
                float4 v[8];
#pragma unroll
                for (int it = 0; it < 8; ++it) {
                    const int64_t row = r0 + it * 8 + rsub;
                    v[it] = load_x4(xb, x_f64, row * (int64_t)u.lda + u.a_off[j] + k, k, K, row < r_end, sg);
                }
                if (!waited) { mbar_wait(empty0 + 8 * s, ((i / EDW_STAGES) & 1) ^ 1); waited = true; }
#pragma unroll
                for (int it = 0; it < 8; ++it) {
                    const int r = it * 8 + rsub;
                    const uint32_t off = (uint32_t)(jb * 8192 + r * 128 + ((((kq >> 1) ^ (r & 7))) << 4) + ((kq & 1) << 3));
                    split_to_smem(st + off, st + EDW_X_BYTES + off, v[it]);
                }
            }
            fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) mbar_arrive(full0 + 8 * s);
        }
        if (warp < 6) {
            // ---------------- epilogue: one thread per output feature ----------------
            mbar_wait(accum_bar, 0);
            tc_fence_after();
            const int q = warp & 3;
            const int o = q * 32 + lane;
            const int64_t slot = (int64_t)blockIdx.x * n_splits + sp;
            float* pw = part_w + slot * (H * EDW_NMAX) + (int64_t)o * EDW_NMAX;
            for (int cc = 0; cc < nkb * 2; ++cc) {
                uint32_t raw[32], raw1[32];
                tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + cc * 32, raw);
                if (split) tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + EDW_NMAX + cc * 32, raw1);
                float4* dst = reinterpret_cast<float4*>(pw + cc * 32);
#pragma unroll
                for (int jj = 0; jj < 8; ++jj) {
                    float x[4];
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        float y = __uint_as_float(raw[4 * jj + e]);
                        if (split) y = fmaf(__uint_as_float(raw1[4 * jj + e]), TC_LO_UNSCALE, y);
                        x[e] = n_steps ? y : 0.f;
                    }
                    dst[jj] = make_float4(x[0], x[1], x[2], x[3]);
                }
            }
            if (want_cs) {
                uint32_t raw[32], raw1[32];
                tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + 384, raw);
                float x = __uint_as_float(raw[0]);
                if (split) { tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + 416, raw1); x = fmaf(__uint_as_float(raw1[0]), TC_LO_UNSCALE, x); }
                part_b[slot * H + o] = n_steps ? x : 0.f;
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, DW_TMEM_COLS);
    }
}

// sum the encoder partials of one (type, 192-column range) group into the flat gradient buffer (fixed order, double)
__global__ void __launch_bounds__(256)
k_reduce_enc(const EncDwGroup* __restrict__ groups, const float* __restrict__ part_w, const float* __restrict__ part_b,
             const int n_splits, float* __restrict__ grads, const float rscale) {
    const EncDwGroup g = groups[blockIdx.x];
    const int n = H * g.width;
    for (int e = blockIdx.y * 256 + threadIdx.x; e < n; e += gridDim.y * 256) {
        const int o = e / g.width, i = e % g.width;
        double sd = 0.0;
        const float* p = part_w + (int64_t)g.first * n_splits * (H * EDW_NMAX) + (int64_t)o * EDW_NMAX + i;
        for (int us = 0; us < g.count * n_splits; ++us) sd += (double)p[(int64_t)us * (H * EDW_NMAX)];
        grads[(int64_t)g.w_off + (int64_t)o * g.K + g.k0 + i] = (float)(sd * (double)rscale);
    }
    if (g.b_off >= 0 && blockIdx.y == 0)
        for (int o = threadIdx.x; o < H; o += 256) {
            double sd = 0.0;
            const float* p = part_b + (int64_t)g.first * n_splits * H + o;
            for (int us = 0; us < g.count * n_splits; ++us) sd += (double)p[(int64_t)us * H];
            grads[(int64_t)g.b_off + o] = (float)(sd * (double)rscale);
        }
}

// fp16 (hi, lo) image of the encoder weights: [n_types*128][kmax], zero padded beyond each type's in-width
__global__ void __launch_bounds__(256)
k_derive_enc16(const float* __restrict__ params, const int64_t w_off, const int K, const int kmax, const int row0,
               __half* __restrict__ w_hi, __half* __restrict__ w_lo) {
    for (int e = blockIdx.x * 256 + threadIdx.x; e < H * kmax; e += gridDim.x * 256) {
        const int o = e / kmax, k = e % kmax;
        const float s = k < K ? params[w_off + (int64_t)o * K + k] * TC_W_SCALE : 0.f;
        const __half h = __float2half_rn(s);
        w_hi[(int64_t)(row0 + o) * kmax + k] = h;
        w_lo[(int64_t)(row0 + o) * kmax + k] = __float2half_rn((s - __half2float(h)) * TC_LO_SCALE);
    }
}

// ------------------------------------------------------------------------------------------
// fp16 (hi, lo) images of the weights the tensor-core kernels read: W (forward) and W^T (backward dX)
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
k_derive16(const Derive16Op* __restrict__ ops, const float* __restrict__ params, __half* __restrict__ w_hi, __half* __restrict__ w_lo) {
    const Derive16Op op = ops[blockIdx.y];
    for (int e = blockIdx.x * 256 + threadIdx.x; e < H * H; e += gridDim.x * 256) {
        const int r = e / H, c = e % H;
        const int src = op.transpose ? (c * H + r) : e;
        float s = 0.f;
        for (int i = 0; i < op.n_src; ++i) s += params[(int64_t)op.src_off[i] + src];
        s *= TC_W_SCALE;
        const __half h = __float2half_rn(s);
        w_hi[(int64_t)op.dst_row * H + e] = h;
        w_lo[(int64_t)op.dst_row * H + e] = __float2half_rn((s - __half2float(h)) * TC_LO_SCALE);
    }
}

}  // namespace mshgnn
