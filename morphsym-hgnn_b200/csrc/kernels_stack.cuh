// Cross-layer "stack" kernel of the MS-HGNN layer loop (hgnn_k4.py:L170-186 and its autograd mirror), sm_100a only.
//
//  The per-layer launch sequence (kernels_tc.cuh) streams every activation slab through HBM once per layer: 8 x (conv,
//  Linear, Linear) launches forward and 8 x (2 Linear, dX) backward, each writing a [slots x graphs x 128] (hi, lo) slab
//  that the next launch reads back - 6.2 GB of the 11.3 GB a train step moved (profiles/r1_tc_v7_step_traffic.json),
//  where SURVEY 8d counts ZERO mandatory bytes for the layer stack.  A tile of 128 graphs x 20 node slots x (hi, lo)
//  images is 1.3 MB - it cannot be resident in one SM's 227 KB of shared memory (and five 128-row role tiles, the
//  smallest closed set under the morphology's edges, are 320 KB) - so this kernel keeps it resident one level up:
//
//   * ONE persistent launch walks all layers.  Work items are ordered (row chunk, layer, row tile, output tile); a chunk
//     is a few thousand graphs, sized so that the slabs a layer reads and writes for it stay in the 126 MB L2 until the
//     next layer has consumed them.  In inference the two ping-pong slabs of a chunk never reach HBM at all; in
//     training every h_l is written once (the backward pass needs it) and never read back from DRAM by the forward.
//   * No grid-wide barrier: layer l + 1 of row tile r starts as soon as every item of layer l of row tile r has
//     signalled a global counter (release / acquire at gpu scope, bounded spin).  Items are handed out round-robin in
//     that order, so the dependency distance ((rows per chunk - 1) x items per row tile) is kept above the number of
//     items in flight and the waits are normally already satisfied.
//   * base_transform (Linear -> ReLU -> Linear, + residual) is CHAINED behind the conv tile of its node inside one item:
//     the epilogue leaves the (hi, lo) result in its staging tiles, which have exactly the operand K-block layout
//     (128 rows x 32 fp16, SWIZZLE_64B), and the MMA warp issues the next GEMM straight from them.  The same chain runs
//     backward (dX of a base node -> dpre -> dc).  That removes 32 short launches per step and, in inference, every
//     byte of the intermediates.
//
//  Roles as in k_tc_rowgemm_persistent: warp 0 = TMA producer (4-stage ring: A_hi, A_lo, W_hi, W_lo K blocks), warp 1 =
//  tcgen05.mma issuer (two 128-column accumulators in TMEM), warps 2..17 = four epilogue groups (one 32-column quarter
//  each, private 16 KB staging pair).
#pragma once
#include "kernels_tc.cuh"

namespace mshgnn {

constexpr int SK_STAGES = 4;
constexpr int SK_THREADS = 576;
constexpr int SK_PIPE_BYTES = SK_STAGES * TC_STAGE_BYTES;                  // 128 KB operand ring
constexpr int SK_STG_BYTES = 65536;                                       // 4 groups x (hi | lo) x 8 KB
constexpr int SK_SMEM_BYTES = SK_PIPE_BYTES + SK_STG_BYTES + 1024 /*align*/ + 256 /*barriers*/;
constexpr uint32_t SK_TMEM_COLS = 256;
constexpr unsigned SK_SPIN_LIMIT = 1u << 22;                              // ~ seconds: a broken dependency must not hang the GPU

struct StackArgs {
    StackProg prog;
    int n_row_tiles;             // Bp / 128
    int rows_per_chunk;          // row tiles per L2-resident chunk
    int n_total;                 // n_row_tiles * prog.items_per_row
    int split;                   // 1: (hi, lo) operands, 3 MMAs per product; 0: hi only
    int64_t B, Bp;
    uint32_t* sync;              // [n_phases][n_row_tiles] completion counters
    uint32_t* err;               // error word (a dependency wait timed out)
    unsigned long long* timing;  // TIMING instantiation only: per CTA 8 cycle counters (see k_tc_stack)
};

struct ItemRef { int phase, row_tile, item; };

// item index -> (phase, row tile, item): chunks of rows_per_chunk row tiles, phase-major inside a chunk
__device__ __forceinline__ ItemRef stack_decode(const StackProg& pg, const int NT, const int RC, const int i) {
    const int chunk_items = RC * pg.items_per_row;
    const int c = i / chunk_items;
    int j = i - c * chunk_items;
    const int rows_c = min(RC, NT - c * RC);
    int p = 0;
    for (; p < pg.n_phases - 1; ++p) {
        const int blk = rows_c * pg.n_items[p];
        if (j < blk) break;
        j -= blk;
    }
    const int n = pg.n_items[p];
    ItemRef r;
    r.phase = p; r.row_tile = c * RC + j / n; r.item = pg.first_item[p] + j % n;
    return r;
}

// Bounded acquire-spin on a completion counter.  On timeout (or when another CTA already timed out) the error word is
// set and the wait gives up: the results are then garbage, but the kernel terminates and the host reports the error.
__device__ __forceinline__ void stack_wait(const uint32_t* ctr, const uint32_t target, uint32_t* err) {
    unsigned spins = 0;
    for (;;) {
        uint32_t v;
        asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(ctr) : "memory");
        if (v >= target) break;
        ++spins;
        if ((spins & 1023u) == 0) {
            uint32_t e;
            asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(e) : "l"(err) : "memory");
            if (e || spins >= SK_SPIN_LIMIT) { asm volatile("st.relaxed.gpu.global.u32 [%0], %1;" ::"l"(err), "r"(1u) : "memory"); break; }
        }
        __nanosleep(40);
    }
    // what the producers wrote with TMA stores (async proxy) is read here with TMA loads (async proxy)
    asm volatile("fence.proxy.async.global;" ::: "memory");
}

__device__ __forceinline__ void stack_signal(uint32_t* ctr) {
    asm volatile("fence.proxy.async.global;" ::: "memory");
    asm volatile("fence.acq_rel.gpu;" ::: "memory");
    asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(ctr), "r"(1u) : "memory");
}

__device__ __forceinline__ bool mbar_test(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
    return ok != 0;
}

template <int N>
__device__ __forceinline__ void tma_store_wait_pending() { asm volatile("cp.async.bulk.wait_group %0;" ::"n"(N) : "memory"); }

// header of a Tile (everything but the chunk list), fetched once per step into registers
struct TileHdr {
    int n_chunks, out_buf, out_slot, bias_buf, bias_off, relu, posmask_buf, posmask_slot, res_buf, res_slot, mask_out_buf,
        out2_buf, out2_slot, out2_mask_kind, out2_mask_buf, out2_mask_slot, mask_out_slot, a_stage, stage_out, pad_;
};
static_assert(sizeof(TileHdr) == offsetof(Tile, chunks), "TileHdr must mirror the head of Tile");
__device__ __forceinline__ TileHdr load_hdr(const Tile* t) {
    TileHdr h;
    const int4* src = reinterpret_cast<const int4*>(t);
    int4* dst = reinterpret_cast<int4*>(&h);
#pragma unroll
    for (int i = 0; i < (int)(sizeof(TileHdr) / 16); ++i) dst[i] = __ldg(src + i);
    return h;
}
static_assert(sizeof(TileHdr) % 16 == 0 && sizeof(Tile) % 16 == 0, "Tile entries are read with 128-bit loads");

struct StackEpi {
    uint32_t stg;            // this group's staging pair: hi tile (8 KB) | lo tile (8 KB)
    uint32_t bias;           // 128 B of shared memory: bias of this group's column quarter
    uint32_t res_bar, accum_bar, free_bar, stage_bar;
    uint32_t acc_parity;
};

// One step of an item through one epilogue group (cf. tc_epilogue_q).  `pend` (leader only) is the completion counter of
// the last item whose TMA stores this leader has committed but not yet published.
__device__ __forceinline__ void stack_epilogue(const TileHdr& t, const BufTable& bt, const BufRows& br, const CUtensorMap* map_k,
                                               const uint32_t tmem_acc, const int row0, const int64_t B, const int64_t Bp,
                                               const int warp, const int lane, const int grp, const StackEpi es, uint32_t& res_count,
                                               uint32_t*& pend, uint32_t* sig, const volatile int* dep_ok, const int dep_need,
                                               unsigned long long* t_wait_acc = nullptr) {
    const int q = warp & 3;                        // TMEM lane quarter this warp may access (hardware rule: warp index mod 4)
    const bool leader = q == 2 && lane == 0;       // first warp of the group (warps 2 + 4g .. 5 + 4g): warp index = 2 mod 4
    const int rl = q * 32 + lane;                  // row inside the tile
    const int64_t row = (int64_t)row0 + rl;
    const bool live = row < B;
    const uint32_t rsw = (uint32_t)((rl >> 1) & 3);     // SWIZZLE_64B: 16-byte chunk index ^= address bits [7, 9)
    const uint32_t tile = es.stg + (uint32_t)rl * 64u;
    const bool has_out = t.out_buf >= 0, has_out2 = t.out2_buf >= 0, has_res = t.res_buf >= 0;
    const bool want_mask = t.relu || t.mask_out_buf >= 0;
    const bool writes_stage = has_out || has_out2 || t.stage_out;
    const int col0 = grp * 32;

    auto fetch_residual = [&]() {
        const int r_hi = br.hi[t.res_buf] + (int)((int64_t)t.res_slot * Bp) + row0, r_lo = br.lo[t.res_buf] + (int)((int64_t)t.res_slot * Bp) + row0;
        mbar_expect_tx(es.res_bar, 2u * 8192u);
        tma_load_2d(es.stg, map_k, es.res_bar, col0, r_hi);
        tma_load_2d(es.stg + 8192, map_k, es.res_bar, col0, r_lo);
    };
    if (leader) {
        tma_store_wait_read();                     // the staging tiles may still feed this group's previous TMA stores
        if (has_res && !t.a_stage) {
            // early fetch (hidden behind the MMAs of this step).  The residual is an output of the previous phase: the
            // producer warp publishes, per item, that it has seen those outputs complete.
            if (dep_need > 0) {
                if (*dep_ok < dep_need) {
                    // about to wait on other CTAs: publish our own pending completion first (it may be what they wait for)
                    if (pend) { tma_store_wait_all(); stack_signal(pend); pend = nullptr; }
                    while (*dep_ok < dep_need) { }
                }
                asm volatile("fence.proxy.async.global;" ::: "memory");
            }
            fetch_residual();
        }
    }
    if (rl < 32) {
        const float bv = t.bias_buf >= 0 ? __ldg((const float*)bt.p[t.bias_buf] + t.bias_off + col0 + rl) : 0.f;
        asm volatile("st.shared.f32 [%0], %1;" ::"r"(es.bias + 4u * rl), "f"(bv) : "memory");
    }
    uint32_t pm = ~0u, m2 = ~0u;
    if (live && t.posmask_buf >= 0) pm = __ldg(reinterpret_cast<const uint32_t*>(bt.p[t.posmask_buf]) + ((int64_t)t.posmask_slot * Bp + row) * 4 + grp);
    if (live && has_out2 && t.out2_mask_kind == MK_BITS)
        m2 = __ldg(reinterpret_cast<const uint32_t*>(bt.p[t.out2_mask_buf]) + ((int64_t)t.out2_mask_slot * Bp + row) * 4 + grp);
    group_bar_sync(grp);

    if (leader && pend && !mbar_test(es.accum_bar, es.acc_parity)) {
        // about to block on the tensor pipe: publish what is still pending first (a CTA that waits never withholds a signal
        // another CTA's producer may be spinning on)
        tma_store_wait_all();
        stack_signal(pend);
        pend = nullptr;
    }
    if (t_wait_acc) { const long long t0 = clock64(); mbar_wait(es.accum_bar, es.acc_parity); *t_wait_acc += (unsigned long long)(clock64() - t0); }
    else mbar_wait(es.accum_bar, es.acc_parity);
    tc_fence_after();
    uint32_t raw[32];
    tmem_ld32(tmem_acc + ((uint32_t)(q * 32) << 16) + col0, raw);
    tmem_ld_wait();
    tc_fence_before();
    mbar_arrive(es.free_bar);                      // this thread's part of the accumulator is in registers
    if (has_res) {
        // chained step: the staging tiles were the A operand of THIS step's MMAs, which have completed by now
        if (t.a_stage && leader) fetch_residual();
        mbar_wait(es.res_bar, res_count & 1u);
        ++res_count;
    }
    float v[32];
    unsigned mask = 0;
#pragma unroll
    for (int j4 = 0; j4 < 8; ++j4) {
        float4 b4;
        asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(b4.x), "=f"(b4.y), "=f"(b4.z), "=f"(b4.w) : "r"(es.bias + 16u * j4));
        const float bb[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const int j = j4 * 4 + e;
            float x = fmaf(__uint_as_float(raw[j]), TC_W_UNSCALE, bb[e]);
            if (want_mask && x > 0.f) mask |= 1u << j;
            if (t.relu) x = fmaxf(x, 0.f);
            v[j] = x;
        }
    }
    if (t.posmask_buf >= 0) {
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = ((pm >> j) & 1u) ? v[j] : 0.f;
    }
#pragma unroll
    for (int g = 0; g < 4; ++g) {
        const uint32_t a = tile + ((((uint32_t)g) ^ rsw) << 4);
        if (has_res) join8_add(v + g * 8, lds128(a), lds128(a + 8192));
        if (writes_stage) {
            uint4 hi, lo;
            split8(v + g * 8, hi, lo);
            if (!has_out && has_out2) { hi = mask8(hi, m2 >> (g * 8)); lo = mask8(lo, m2 >> (g * 8)); }
            if (!live) { hi = make_uint4(0u, 0u, 0u, 0u); lo = hi; }     // rows [B, Bp) of every image stay zero
            sts128(a, hi);
            sts128(a + 8192, lo);
        }
    }
    fence_proxy_async_smem();
    group_bar_sync(grp);
    int committed = 0;
    if (leader) {
        const int ob = has_out ? t.out_buf : t.out2_buf, os = has_out ? t.out_slot : t.out2_slot;
        if (ob >= 0) {
            const int o = (int)((int64_t)os * Bp) + row0;
            tma_store_2d(map_k, es.stg, col0, br.hi[ob] + o);
            tma_store_2d(map_k, es.stg + 8192u, col0, br.lo[ob] + o);
            tma_store_commit();
            ++committed;
        }
        if (t.stage_out) mbar_arrive(es.stage_bar);        // this quarter of the next step's A operand is in place
    }
    if (has_out && has_out2) {
        // second output = first output with the masked-off lanes cleared, made in place once the first store has read the tile
        if (leader) tma_store_wait_read();
        group_bar_sync(grp);
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            const uint32_t a = tile + ((((uint32_t)c) ^ rsw) << 4);
            const uint32_t m = m2 >> (c * 8);
            sts128(a, mask8(lds128(a), m));
            sts128(a + 8192, mask8(lds128(a + 8192), m));
        }
        fence_proxy_async_smem();
        group_bar_sync(grp);
        if (leader) {
            const int o = (int)((int64_t)t.out2_slot * Bp) + row0;
            tma_store_2d(map_k, es.stg, col0, br.hi[t.out2_buf] + o);
            tma_store_2d(map_k, es.stg + 8192u, col0, br.lo[t.out2_buf] + o);
            tma_store_commit();
            ++committed;
        }
    }
    if (live && t.mask_out_buf >= 0)
        *(reinterpret_cast<uint32_t*>(bt.p[t.mask_out_buf]) + ((int64_t)t.mask_out_slot * Bp + row) * 4 + grp) = mask;
    if (leader) {
        if (pend) {
            // everything committed BEFORE this step has landed once at most this step's groups are still in flight
            if (committed == 0) tma_store_wait_pending<0>();
            else if (committed == 1) tma_store_wait_pending<1>();
            else tma_store_wait_pending<2>();
            stack_signal(pend);
            pend = nullptr;
        }
        if (sig) pend = sig;
    }
}

// TIMING = true is a diagnostic instantiation (MSHGNN_STACK_TIMING=1): the single-thread roles accumulate the cycles they
// spend in each kind of wait into args.timing[blockIdx.x * 8 + {0: producer/ring slot, 1: producer/dependency, 2: MMA/operands,
// 3: MMA/accumulator free, 4: MMA/staged operand, 5: epilogue group 0/accumulator, 6: kernel, 7: steps}].
#define SK_TIMED(slot, stmt) do { if (TIMING) { const long long t0_ = clock64(); stmt; tim[slot] += (unsigned long long)(clock64() - t0_); } else { stmt; } } while (0)
template <bool TIMING>
__global__ void __launch_bounds__(SK_THREADS, 1)
k_tc_stack(const __grid_constant__ TcMaps maps, const Tile* __restrict__ tiles, const StackItem* __restrict__ items,
           const __grid_constant__ StackArgs args, const BufTable bt, const BufRows br) {
    extern __shared__ uint8_t smem_raw[];
    __shared__ __align__(16) float bias_s[4][32];         // one quarter per epilogue group
    __shared__ uint32_t tmem_base_s;
    __shared__ volatile int dep_ok_s;                     // items whose input dependency the producer warp has seen satisfied

    const int tid = threadIdx.x;
    const int warp = tid >> 5, lane = tid & 31;
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + SK_PIPE_BYTES + SK_STG_BYTES);
    const uint32_t full0 = smem_u32(bars), empty0 = smem_u32(bars + SK_STAGES), acc_full0 = smem_u32(bars + 2 * SK_STAGES),
                   acc_free0 = smem_u32(bars + 2 * SK_STAGES + 2), res_bar = smem_u32(bars + 2 * SK_STAGES + 4),   // 4 residual barriers
                   stage_bar = smem_u32(bars + 2 * SK_STAGES + 8);
    const uint32_t smem_base = smem_u32(smem);
    const uint32_t stg_base = smem_base + SK_PIPE_BYTES;

    if (tid == 0) {
        for (int s = 0; s < SK_STAGES; ++s) { mbar_init(full0 + 8 * s, 1); mbar_init(empty0 + 8 * s, 1); }
        for (int a = 0; a < 2; ++a) { mbar_init(acc_full0 + 8 * a, 1); mbar_init(acc_free0 + 8 * a, 512); }
        for (int g = 0; g < 4; ++g) mbar_init(res_bar + 8 * g, 1);
        mbar_init(stage_bar, 4);
        dep_ok_s = 0;
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) tmem_alloc(smem_u32(&tmem_base_s), SK_TMEM_COLS);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_base_s;
    constexpr int SPC = H / TC_KB;                 // pipeline steps (K blocks) per chunk
    unsigned long long tim[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    const long long t_begin = TIMING ? clock64() : 0;
    const StackProg& pg = args.prog;
    const int NT = args.n_row_tiles, RC = args.rows_per_chunk, n_total = args.n_total, split = args.split;
    const int64_t B = args.B, Bp = args.Bp;
    uint32_t* const err = args.err;
    // Programmatic dependent launch: everything above ran while the previous kernel of the stream was still draining
    asm volatile("griddepcontrol.wait;" ::: "memory");
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");

    if (warp == 0) {
        if (lane == 0) {
            uint32_t g = 0;                        // pipeline step counter, runs across items
            int n_done = 0;
            // the completion counter of the NEXT item is polled one item ahead (relaxed load, consumed an item later), so
            // the L2 round trip of the dependency check hides behind the operand stream of the current item
            uint32_t pre = 0;
            bool have_pre = false;
            ItemRef ir = stack_decode(pg, NT, RC, blockIdx.x < n_total ? blockIdx.x : 0);
            for (int i = blockIdx.x; i < n_total; i += gridDim.x) {
                const StackItem it = items[ir.item];
                const int row0 = ir.row_tile * TILE_M;
                if (ir.phase > 0) {
                    const uint32_t target = 4u * (uint32_t)pg.n_items[ir.phase - 1];
                    if (have_pre && pre >= target) {
                        asm volatile("fence.acq_rel.gpu;" ::: "memory");
                        asm volatile("fence.proxy.async.global;" ::: "memory");
                    } else {
                        SK_TIMED(1, stack_wait(args.sync + (size_t)(ir.phase - 1) * NT + ir.row_tile, target, err));
                    }
                }
                __threadfence_block();
                dep_ok_s = ++n_done;
                have_pre = false;
                if (i + (int)gridDim.x < n_total) {
                    ir = stack_decode(pg, NT, RC, i + (int)gridDim.x);
                    if (ir.phase > 0) {
                        asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(pre) : "l"(args.sync + (size_t)(ir.phase - 1) * NT + ir.row_tile) : "memory");
                        have_pre = true;
                    }
                }
                for (int s = 0; s < it.n_steps; ++s) {
                    const Tile* t = tiles + it.tile + s;
                    const int n_chunks = __ldg(&t->n_chunks), a_stage = __ldg(&t->a_stage);
                    for (int c = 0; c < n_chunks; ++c) {
                        const bool from_stage = a_stage && c == 0;
                        const int a_buf = __ldg(&t->chunks[c].a_buf), a_slot = __ldg(&t->chunks[c].a_slot), w16_row = __ldg(&t->chunks[c].w16_row);
                        const int arow = (int)((int64_t)a_slot * Bp) + row0;
                        const uint32_t tx_bytes = (uint32_t)((from_stage ? 1 : 2) * (split ? 2 : 1) * TC_TILE_BYTES);
                        for (int kb = 0; kb < SPC; ++kb, ++g) {
                            const uint32_t s4 = g % SK_STAGES;
                            SK_TIMED(0, mbar_wait(empty0 + 8 * s4, ((g / SK_STAGES) & 1) ^ 1));
                            const int kcol = kb * TC_KB;
                            const uint32_t st = smem_base + s4 * TC_STAGE_BYTES;
                            const uint32_t fb = full0 + 8 * s4;
                            mbar_expect_tx(fb, tx_bytes);
                            if (!from_stage) tma_load_2d(st, &maps.k, fb, kcol, br.hi[a_buf] + arow);
                            tma_load_2d(st + 2 * TC_TILE_BYTES, &maps.k, fb, kcol, br.w_hi + w16_row);
                            if (split) {
                                if (!from_stage) tma_load_2d(st + TC_TILE_BYTES, &maps.k, fb, kcol, br.lo[a_buf] + arow);
                                tma_load_2d(st + 3 * TC_TILE_BYTES, &maps.k, fb, kcol, br.w_lo + w16_row);
                            }
                        }
                    }
                }
            }
            if (TIMING) { args.timing[(size_t)blockIdx.x * 8] = tim[0]; args.timing[(size_t)blockIdx.x * 8 + 1] = tim[1]; }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            uint32_t g = 0, k = 0, n_staged = 0;
            for (int i = blockIdx.x; i < n_total; i += gridDim.x) {
                const ItemRef ir = stack_decode(pg, NT, RC, i);
                const StackItem it = items[ir.item];
                for (int s = 0; s < it.n_steps; ++s, ++k) {
                    const Tile* t = tiles + it.tile + s;
                    const int n_chunks = __ldg(&t->n_chunks), a_stage = __ldg(&t->a_stage);
                    const uint32_t a = k & 1;
                    SK_TIMED(3, mbar_wait(acc_free0 + 8 * a, ((k >> 1) & 1) ^ 1));      // the epilogue has drained this accumulator
                    tc_fence_after();
                    if (TIMING) tim[7] += 1;
                    const uint32_t d0 = tmem_base + a * 128;
                    for (int c = 0; c < n_chunks; ++c) {
                        const bool from_stage = a_stage && c == 0;
                        if (from_stage) {
                            SK_TIMED(4, mbar_wait(stage_bar, n_staged & 1));   // the four quarters of the previous step's result are staged
                            ++n_staged;
                            tc_fence_after();
                        }
                        for (int kb = 0; kb < SPC; ++kb, ++g) {
                            const uint32_t s4 = g % SK_STAGES;
                            SK_TIMED(2, mbar_wait(full0 + 8 * s4, (g / SK_STAGES) & 1));
                            tc_fence_after();
                            const uint32_t st = smem_base + s4 * TC_STAGE_BYTES;
                            const uint32_t a_base = from_stage ? stg_base + (uint32_t)kb * 16384u : st;
                            const uint64_t a_hi = smem_desc_sw64(a_base), a_lo = smem_desc_sw64(a_base + (from_stage ? 8192u : (uint32_t)TC_TILE_BYTES));
                            const uint64_t w_hi = smem_desc_sw64(st + 2 * TC_TILE_BYTES), w_lo = smem_desc_sw64(st + 3 * TC_TILE_BYTES);
#pragma unroll
                            for (int ks = 0; ks < TC_KB / 16; ++ks) {
                                const uint64_t adv = (uint64_t)(ks * 2);
                                umma_f16(d0, a_hi + adv, w_hi + adv, TC_IDESC, (c | kb | ks) ? 1u : 0u);
                                if (split) {
                                    umma_f16(d0, a_lo + adv, w_hi + adv, TC_IDESC, 1u);
                                    umma_f16(d0, a_hi + adv, w_lo + adv, TC_IDESC, 1u);
                                }
                            }
                            umma_commit(empty0 + 8 * s4);
                        }
                    }
                    umma_commit(acc_full0 + 8 * a);
                }
            }
            if (TIMING) {
                tim[6] = (unsigned long long)(clock64() - t_begin);
                for (int j = 2; j < 8; ++j) if (j != 5) args.timing[(size_t)blockIdx.x * 8 + j] = tim[j];
            }
        }
        __syncwarp();
    } else {
        // group g (warps 2 + 4g .. 5 + 4g) drains column quarter g of every step
        const int grp = (warp - 2) >> 2;
        StackEpi es;
        es.stg = stg_base + grp * 16384; es.bias = smem_u32(bias_s[grp]); es.res_bar = res_bar + 8 * grp; es.stage_bar = stage_bar;
        uint32_t k = 0, n_res = 0;
        uint32_t* pend = nullptr;
        int n_item = 0;
        for (int i = blockIdx.x; i < n_total; i += gridDim.x) {
            const ItemRef ir = stack_decode(pg, NT, RC, i);
            const StackItem it = items[ir.item];
            ++n_item;
            uint32_t* const ctr = args.sync + (size_t)ir.phase * NT + ir.row_tile;
            for (int s = 0; s < it.n_steps; ++s, ++k) {
                const TileHdr t = load_hdr(tiles + it.tile + s);
                const uint32_t a = k & 1;
                es.accum_bar = acc_full0 + 8 * a; es.free_bar = acc_free0 + 8 * a;
                es.acc_parity = (k >> 1) & 1;
                stack_epilogue(t, bt, br, &maps.k, tmem_base + a * 128, ir.row_tile * TILE_M, B, Bp, warp, lane, grp, es, n_res, pend,
                               s == it.n_steps - 1 ? ctr : nullptr, &dep_ok_s, ir.phase > 0 ? n_item : 0, (TIMING && warp == 2 && lane == 0) ? &tim[5] : nullptr);
            }
        }
        if ((warp & 3) == 2 && lane == 0) {
            tma_store_wait_all();
            if (pend) stack_signal(pend);
        }
        if (TIMING && warp == 2 && lane == 0) args.timing[(size_t)blockIdx.x * 8 + 5] = tim[5];
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, SK_TMEM_COLS);
    }
}

}  // namespace mshgnn
