// ffn_tc.cu -- the block FFN on the 5th-generation tensor cores (tcgen05 + TMEM), sm_100a.
//
//   y = u + W2 relu(W1 LN(u) + b1) + b2         mssvt_backbone.py:337-340, 384-385
//   u = covered ? merged + x : 2 x  (two-window block)   |   u = merged  (compress block)
//
// One CTA of 256 threads owns a tile of 128 rows; threads t and t + 128 share row t = TMEM lane t
// (one half of the channels each; row sums for the LayerNorms are exchanged through shared memory).
//   1. each thread loads its row, builds u, LayerNorms it in registers and stores it, rounded to
//      TF32, as the A operand in the canonical K-major no-swizzle UMMA layout
//      (8-row x 16-byte core matrices: byte = chunk * LBO + (row / 8) * 128 + (row % 8) * 16);
//   2. one thread issues C/8 tcgen05.mma.kind::tf32 (M = 128, N = F) -> D1 in TMEM columns [0, F);
//   3. tcgen05.ld brings each row's D1 back, bias + ReLU, TF32 round, tcgen05.st writes the hidden row
//      back IN PLACE: TMEM columns [0, F) now hold the A operand of the second GEMM (lane = row,
//      column = k), so the 128 x F hidden tile never touches shared memory;
//   4. F/8 tcgen05.mma with A from TMEM (M = 128, N = C) -> D2 in TMEM columns [F, F + C);
//   5. tcgen05.ld, + b2 + u (still in registers), row store.
// W1 / W2 (nn.Linear [out][in] = N x K, K-major) sit in shared memory in the same canonical
// layout for the whole kernel.  ~98 KB of shared memory and 256 TMEM columns per CTA: two CTAs per
// SM, so one tile's loads and epilogue overlap the other's MMAs.  Completion of the async MMAs is tracked with tcgen05.commit on
// an mbarrier; generic-proxy shared stores are made visible with fence.proxy.async.
//
// Precision: TF32 operands (10-bit mantissa, round-to-nearest), fp32 accumulation in TMEM; the
// residual stream u and the LayerNorm stay fp32.  Selected with precision = "tf32"; the exact
// FFMA kernel in block.cu stays the default (tests state the tolerance for each).
#include <cstring>
#include "tma.cuh"

namespace mssvt {

#define TC_ROWS 128
// Which kernels move their dense row tiles by TMA (tma.cuh).  Measured per launch at 150 k rows, S0:
//   box STORES of y / the next LayerNorm rows: 51.0 -> 47.0 us (TF32 operands), 47.8 -> 43.4 us (bf16): the rows cross the LSU
//       data pipe once (STS) instead of three times (STS, LDS, STG) and leave asynchronously; with split operands
//       (one CTA per SM, loads software-pipelined) 65.3 -> 66.4 us: not used there;
//   box LOAD of x (mode 2) next to the gathers of the projected rows: 47.1 -> 51.9 us (TF32), 64.8 -> 68.9 us (split):
//       slower -- x then has to be held in registers next to both gathered row sets (spills), and the load phase is
//       bound by the gathers, not by x.  Kept selectable (-DFFN_TMA_LOAD), off by default.
#ifdef FFN_NO_TMA   // (A/B builds: rows through the LDG -> STS -> LDS staging everywhere)
#define FFN_TMA_SHAPE(C, TPR, TERMS) false
#else
#define FFN_TMA_SHAPE(C, TPR, TERMS) ((C) == 64 && (TPR) == 2 && (TERMS) != 3)
#endif
#ifdef MSSVT_TRACE
#define TRACE(i) do { if (tid == 0 && blockIdx.x == 0 && tile == (int)blockIdx.x + (int)gridDim.x) tr[i] = clock64(); } while (0)
#else
#define TRACE(i) do {} while (0)
#endif
// TPR threads share a row (= TMEM lane): warps w, w + 4, ... reach the same 32 TMEM lanes, each owns C / TPR channels
// and F / TPR hidden columns.  TPR = 2: 256 threads per tile; TPR = 4 (C = 64): 512 threads per tile -- the kernel is
// bound by the dependent-instruction latency of its per-row epilogues (a tile takes the same ~13 k clocks alone on
// an SM and next to a second CTA), so halving the per-thread chains and doubling the warps is what speeds it up.

struct FfnTcParams {
    int F, mode;            // hidden width; 0: u = merged, 1: u = covered ? merged + x : 2 x
    float eps;
    const float *w1, *b1;   // [F][C] packed (mssvt_pack_operand_tf32), [F]
    const float *w2, *b2;   // [C][F] packed, [C]
    const float *ln_g, *ln_b;
    const float *next_g, *next_b;  // optional: LayerNorm of the NEXT block (norm1), fused into the epilogue
    float next_eps;
    // mode 2: the three-NN feature interpolation + merge of the window attention (k_tca_merge) happens here,
    // on the way in: merged[row] = sum_i w_i * P[q_base[w] + nn_i] is never written to memory
    const int *vox_slot, *meta, *q_base;
    const unsigned char *nn_idx;
    const float *nn_w, *pbuf;
    int cap1;
};

// TERMS: 1 = TF32 operands, 3 = split operands (3xTF32), 0 = bf16 operands (kind::f16; half the operand bytes,
// the hidden tile packed two per TMEM column)
template <int C, int TERMS, int TPR>
__global__ void __launch_bounds__(TC_ROWS * TPR, TERMS == 3 ? 1 : 2)
k_ffn_tc(FfnTcParams P, int n_cap, const int *__restrict__ n_dev, const float *__restrict__ x,
         const float *__restrict__ merged, const unsigned char *__restrict__ covered,
         float *__restrict__ y, float *__restrict__ xn_next, const __grid_constant__ CUtensorMap tm_x,
         const __grid_constant__ CUtensorMap tm_y, const __grid_constant__ CUtensorMap tm_xn) {
    constexpr int TC_THREADS = TC_ROWS * TPR;
    constexpr int CH = C / TPR;  // channels per thread
    static_assert(CH == 32 || CH == 16, "a thread owns 16 or 32 channels of its row");
    extern __shared__ __align__(1024) char smem_raw[];
#ifdef MSSVT_TRACE
    unsigned long long gt[8] = {0};   // %globaltimer (ns) timeline of this CTA
    auto gtime = []() { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; };
    gt[0] = gtime();
#endif
    pdl_launch_dependents();
    const int F = P.F, FH = F / TPR;
    const int tid = threadIdx.x, warp = tid >> 5;
    const int r = tid & (TC_ROWS - 1), half = tid >> 7;   // half: which part of the row (0 .. TPR - 1)
    // carve shared memory
    constexpr int NT = TERMS == 3 ? 2 : 1;        // operand tiles: hi [, lo] (3xTF32, see tc_common.cuh)
    constexpr bool BF = TERMS == 0 || TERMS == 2; // bf16 operands (TERMS == 2: split in two, hi + mid: "bf16x3")
    constexpr int EB = BF ? 2 : 4;                // bytes per operand element
    constexpr int NW = TERMS == 3 || TERMS == 2 ? 2 : 1;   // tiles per weight matrix
    char *sA = smem_raw;                          // NT x [chunks][128][16 B] (at least the 32 KB staging area)
    char *sW1 = sA + (BF ? TC_ROWS * C * 4 : NT * TC_ROWS * C * 4);   // NT x [chunks][F][16 B]
    char *sW2 = sW1 + NW * F * C * EB;            // NW x [chunks][C][16 B]
    float *s_vec = (float *)(sW2 + NW * C * F * EB);   // ln_g[C], ln_b[C], b1[F], b2[C], next_g[C], next_b[C]
    float *s_red = s_vec + 5 * C + F;             // [4][TPR][128] partial row sums of the threads of a row
    uint64_t *s_bar = (uint64_t *)(s_red + 4 * TPR * TC_ROWS);  // 3 mbarriers (8-byte aligned: C, F even) + one per warp
    uint32_t *s_tmem = (uint32_t *)(s_bar + 3 + 4 * TPR);        //   (the warp's TMA row loads)
    // TMA: the dense row tiles (x in mode 2, y, the next LayerNorm rows) move as one box per warp (tma.cuh)
    constexpr bool TMA = FFN_TMA_SHAPE(C, TPR, TERMS);
#ifdef FFN_TMA_LOAD
    constexpr bool TMA_LD = TMA;
#else
    constexpr bool TMA_LD = false;
#endif
    const uint32_t xbar = smem_u32(s_bar + 3 + warp);
    uint32_t xphase = 0;
    if (TMA && (tid & 31) == 0) {
        mbar_init(xbar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    const uint32_t bar_w = smem_u32(s_bar + 2);           // weights landed (TMA bulk copies)
    if (tid == 0) {
        mbar_init(bar_w, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        // W1 and W2 (packed once on the host side) come in as two bulk copies of the TMA engine
        const uint32_t wbytes = (uint32_t)(NW * F * C * EB);
        bulk_expect(bar_w, 2u * wbytes);
        bulk_copy_g2s(P.w1, wbytes, sW1, bar_w);
        bulk_copy_g2s(P.w2, wbytes, sW2, bar_w);
    }
    for (int i = tid; i < C; i += TC_THREADS) {
        s_vec[i] = __ldg(P.ln_g + i);
        s_vec[C + i] = __ldg(P.ln_b + i);
        s_vec[2 * C + F + i] = __ldg(P.b2 + i);
        s_vec[3 * C + F + i] = P.next_g ? __ldg(P.next_g + i) : 1.f;
        s_vec[4 * C + F + i] = P.next_b ? __ldg(P.next_b + i) : 0.f;
    }
    for (int i = tid; i < F; i += TC_THREADS) s_vec[2 * C + i] = __ldg(P.b1 + i);
    const float *s_g = s_vec + half * CH, *s_b = s_vec + C + half * CH, *s_b1 = s_vec + 2 * C + half * FH;
    const float *s_b2 = s_vec + 2 * C + F + half * CH;
    const float *s_ng = s_vec + 3 * C + F + half * CH, *s_nb = s_vec + 4 * C + F + half * CH;

    const uint32_t bar1 = smem_u32(s_bar), bar2 = smem_u32(s_bar + 1);
    if (tid == 0) {
        mbar_init(bar1, 1);
        mbar_init(bar2, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    // TMEM: F columns for D1 / hidden + C columns for D2, rounded up to a power of two >= 32
    uint32_t tmem_cols = 32;
    while (tmem_cols < (uint32_t)(F + C + (TERMS == 3 ? F : BF ? F / 2 : 0))) tmem_cols <<= 1;   // (bf16x3: mid parts in place)
    if (warp == 0) tmem_alloc(smem_u32(s_tmem), tmem_cols);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // staged weights -> async proxy
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *s_tmem;
    const uint32_t tmem_d1 = tmem_base, tmem_d2 = tmem_base + (uint32_t)F;
    const uint32_t tmem_hlo = tmem_base + (uint32_t)(F + C);   // (TERMS == 3) low part of the hidden tile;
                                                               // (bf16) the packed hidden tile, F / 2 columns
    const uint32_t lane_off = (uint32_t)((warp & 3) * 32) << 16;  // a warp reaches TMEM lanes 32 (w % 4) ...

    const uint32_t idesc1 = BF ? umma_idesc_bf16(TC_ROWS, F) : umma_idesc_tf32(TC_ROWS, F);
    const uint32_t idesc2 = BF ? umma_idesc_bf16(TC_ROWS, C) : umma_idesc_tf32(TC_ROWS, C);
    const uint32_t a_lbo = TC_ROWS * 16, w1_lbo = (uint32_t)F * 16, w2_lbo = (uint32_t)C * 16;
    const UmmaDescBase dA = umma_desc_base(smem_u32(sA), a_lbo, 128), dW1 = umma_desc_base(smem_u32(sW1), w1_lbo, 128),
                       dW2 = umma_desc_base(smem_u32(sW2), w2_lbo, 128);

#ifdef MSSVT_TRACE
    gt[1] = gtime();
#endif
    pdl_wait();  // (everything above touched static parameters only)
#ifdef MSSVT_TRACE
    gt[2] = gtime();
#endif
    const int n = n_dev ? min(n_cap, __ldg(n_dev)) : n_cap;
    const int tiles = (n + TC_ROWS - 1) / TC_ROWS;
    uint32_t phase = 0;
    const uint32_t my_row_off = (uint32_t)(r >> 3) * 128u + (uint32_t)(r & 7) * 16u;
    float *red_mine = s_red + half * TC_ROWS + r;
    // sum of the row's TPR partials in slot `which` (the same order in every thread of the row)
    auto red_sum = [&](int which) {
        const float *p = s_red + which * TPR * TC_ROWS + r;
        float t = p[0];
#pragma unroll
        for (int i = 1; i < TPR; ++i) t += p[i * TC_ROWS];
        return t;
    };
    // warp-private staging inside the A tile (free between the second GEMM of a tile and the A stores of
    // the next): 32 half-rows of CH floats, 16-byte chunks XOR-swizzled by the row -> conflict-free both ways
    constexpr int CPR = CH / 4, RPI = 32 / CPR;  // chunks per part-row, rows per warp instruction
    constexpr int SWS = CH == 32 ? 0 : 1;        // swizzle key of a row: (row >> SWS) & (CPR - 1) -- 8 consecutive lanes
                                                 // (one 128-byte wavefront) always hit 8 different 16-byte bank groups
    const int lane = tid & 31;
    char *stg = sA + warp * (32 * CH * 4);
    // second box of the warp where the A region holds two tiles (split operands: the lo tile): x lands there while the
    // gathers use `stg`, the next LayerNorm rows leave from there while y leaves from `stg`
    char *stg2 = TERMS == 3 ? sA + TC_ROWS * C * 4 + warp * (32 * CH * 4) : stg;
    const int st_row = lane / CPR, st_ch = lane % CPR;
    int tile_row0 = 0;  // first row of the warp's 32 rows in the current tile
    auto stage_in = [&](const float *__restrict__ src, float4 *dst) {
        float4 v[CPR];
#pragma unroll
        for (int i = 0; i < CPR; ++i) {
            const int rr = RPI * i + st_row, grow = tile_row0 + rr;
            v[i] = grow < n ? __ldg((const float4 *)(src + (size_t)grow * C + half * CH) + st_ch)
                            : make_float4(0.f, 0.f, 0.f, 0.f);
        }
#pragma unroll
        for (int i = 0; i < CPR; ++i) {
            const int rr = RPI * i + st_row;
            *(float4 *)(stg + rr * (CH * 4) + ((st_ch ^ ((rr >> SWS) & (CPR - 1))) << 4)) = v[i];
        }
        __syncwarp();
#pragma unroll
        for (int q = 0; q < CPR; ++q) dst[q] = *(const float4 *)(stg + lane * (CH * 4) + ((q ^ ((lane >> SWS) & (CPR - 1))) << 4));
        __syncwarp();
    };
    // the same for rows named by a pointer per lane (my_src: this lane's CH-float segment, nullptr = a zero row)
    auto gather_in = [&](const float *my_src, float4 *dst) {
        float4 v[CPR];
#pragma unroll
        for (int i = 0; i < CPR; ++i) {
            const float4 *p = (const float4 *)__shfl_sync(0xffffffffu, (unsigned long long)my_src, RPI * i + st_row);
            v[i] = p ? __ldg(p + st_ch) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
#pragma unroll
        for (int i = 0; i < CPR; ++i) {
            const int rr = RPI * i + st_row;
            *(float4 *)(stg + rr * (CH * 4) + ((st_ch ^ ((rr >> SWS) & (CPR - 1))) << 4)) = v[i];
        }
        __syncwarp();
#pragma unroll
        for (int q = 0; q < CPR; ++q) dst[q] = *(const float4 *)(stg + lane * (CH * 4) + ((q ^ ((lane >> SWS) & (CPR - 1))) << 4));
        __syncwarp();
    };
    // two row sets per memory round trip: all global loads of both are issued before the first is parked in the
    // staging area (a tile's load phase is a chain of dependent round trips; this halves it).  Set a: rows named by a
    // pointer per lane; set b: rows named by a pointer per lane (b_src == nullptr) or the tile's rows of `b_src`
    auto gather_in2 = [&](const float *a_ptr, float4 *da, const float *b_ptr, const float *__restrict__ b_src, float4 *db) {
        float4 va[CPR], vb[CPR];
#pragma unroll
        for (int i = 0; i < CPR; ++i) {
            const float4 *p = (const float4 *)__shfl_sync(0xffffffffu, (unsigned long long)a_ptr, RPI * i + st_row);
            va[i] = p ? __ldg(p + st_ch) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
#pragma unroll
        for (int i = 0; i < CPR; ++i) {
            if (b_src) {
                const int grow = tile_row0 + RPI * i + st_row;
                vb[i] = grow < n ? __ldg((const float4 *)(b_src + (size_t)grow * C + half * CH) + st_ch)
                                 : make_float4(0.f, 0.f, 0.f, 0.f);
            } else {
                const float4 *p = (const float4 *)__shfl_sync(0xffffffffu, (unsigned long long)b_ptr, RPI * i + st_row);
                vb[i] = p ? __ldg(p + st_ch) : make_float4(0.f, 0.f, 0.f, 0.f);
            }
        }
#pragma unroll
        for (int i = 0; i < CPR; ++i) {
            const int rr = RPI * i + st_row;
            *(float4 *)(stg + rr * (CH * 4) + ((st_ch ^ ((rr >> SWS) & (CPR - 1))) << 4)) = va[i];
        }
        __syncwarp();
#pragma unroll
        for (int q = 0; q < CPR; ++q) da[q] = *(const float4 *)(stg + lane * (CH * 4) + ((q ^ ((lane >> SWS) & (CPR - 1))) << 4));
        __syncwarp();
#pragma unroll
        for (int i = 0; i < CPR; ++i) {
            const int rr = RPI * i + st_row;
            *(float4 *)(stg + rr * (CH * 4) + ((st_ch ^ ((rr >> SWS) & (CPR - 1))) << 4)) = vb[i];
        }
        __syncwarp();
#pragma unroll
        for (int q = 0; q < CPR; ++q) db[q] = *(const float4 *)(stg + lane * (CH * 4) + ((q ^ ((lane >> SWS) & (CPR - 1))) << 4));
        __syncwarp();
    };
    // the warp's 32 part-rows leave as one TMA box from `from` (the rows are written with the box swizzle)
    auto box_out = [&](const CUtensorMap *map, char *from, const float4 *src) {
#pragma unroll
        for (int q = 0; q < CPR; ++q) *(float4 *)(from + tma_swz(lane, q)) = src[q];
        fence_async_smem();
        __syncwarp();
        if (elect_one()) {
            tma_store_box(map, half * CH, tile_row0, from);
            tma_store_commit();
        }
        __syncwarp();
    };
    auto stage_out = [&](float *__restrict__ dst, const float4 *src) {
#pragma unroll
        for (int q = 0; q < CPR; ++q) *(float4 *)(stg + lane * (CH * 4) + ((q ^ ((lane >> SWS) & (CPR - 1))) << 4)) = src[q];
        __syncwarp();
#pragma unroll
        for (int i = 0; i < CPR; ++i) {
            const int rr = RPI * i + st_row, grow = tile_row0 + rr;
            if (grow < n)
                *((float4 *)(dst + (size_t)grow * C + half * CH) + st_ch) =
                    *(const float4 *)(stg + rr * (CH * 4) + ((st_ch ^ ((rr >> SWS) & (CPR - 1))) << 4));
        }
        __syncwarp();
    };

#ifdef MSSVT_TRACE
    long long tr[12] = {0};
#endif
    // ---- the load phase of a tile: this thread's half of the residual row u (registers).  Global rows move
    //      through a warp-private staging area so that every load / store instruction of a warp covers whole
    //      128-byte lines (RPI rows x CH floats), not 32 rows.  Called one tile AHEAD (software pipelining): the
    //      loads of tile i + 1 are issued while the second GEMM of tile i runs, so that the memory system and
    //      the tensor pipe work at the same time even with a single CTA on the SM (3xTF32).
    // mode 2, the index half of the chain: the voxel's win1 slot names its 3 nearest query slots -> pointers to
    // their projected rows + blend weights.  The chain slot -> window record -> rows is three dependent global
    // accesses; its links are issued in three STAGES spread over the phases of the previous tile (volatile loads:
    // they stay where they are written), each consumed a phase later, so that no warp ever waits for them:
    //     stage 0 (top of tile i)          vox_slot of the row in tile i + 1
    //     stage 1 (first GEMM issued)      #real queries, first query id, nn indices / weights of that slot
    //     stage 2 (second GEMM issued)     pointers -> load_tile gathers the rows
    struct MergeIdx { const float *p0, *p1, *p2; float w0, w1, w2; bool cov; int slot, nqr, q0; unsigned nn; };
    auto ldg_v = [](const int *p) { int v; asm volatile("ld.global.nc.s32 %0, [%1];" : "=r"(v) : "l"(p)); return v; };
    auto ldg_vf = [](const float *p) { float v; asm volatile("ld.global.nc.f32 %0, [%1];" : "=f"(v) : "l"(p)); return v; };
    auto ldg_vb = [](const unsigned char *p) { unsigned v; asm volatile("ld.global.nc.u8 %0, [%1];" : "=r"(v) : "l"(p)); return v; };
    auto index_stage0 = [&](int tile, MergeIdx &ix) {
        ix.slot = -1;
        if constexpr (C == 64) {
            const int row = tile * TC_ROWS + r;
            if (P.mode == 2 && row < n) ix.slot = ldg_v(P.vox_slot + row);
        }
    };
    auto index_stage1 = [&](MergeIdx &ix) {
        ix.nqr = ix.q0 = 0; ix.nn = 0u; ix.w0 = ix.w1 = ix.w2 = 0.f;
        if constexpr (C == 64) {
            asm volatile("" : "+r"(ix.slot));
            if (ix.slot >= 0) {
                const int w = ix.slot / P.cap1;
                ix.nqr = ldg_v(P.meta + 4 * (size_t)w);
                ix.q0 = ldg_v(P.q_base + w);
                const unsigned char *ni = P.nn_idx + (size_t)ix.slot * 3;
                ix.nn = ldg_vb(ni) | (ldg_vb(ni + 1) << 8) | (ldg_vb(ni + 2) << 16);
                const float *nw = P.nn_w + (size_t)ix.slot * 3;
                ix.w0 = ldg_vf(nw); ix.w1 = ldg_vf(nw + 1); ix.w2 = ldg_vf(nw + 2);
            }
        }
    };
    auto index_stage2 = [&](MergeIdx &ix) {
        ix.p0 = ix.p1 = ix.p2 = nullptr;
        ix.cov = ix.slot >= 0;
        if constexpr (C == 64) {
            asm volatile("" : "+r"(ix.nqr), "+r"(ix.q0), "+r"(ix.nn));
            if (ix.cov) {
                const int n0 = ix.nn & 0xff, n1 = (ix.nn >> 8) & 0xff, n2 = (ix.nn >> 16) & 0xff;
                // padded query slots (index >= #real queries) are zero rows in the reference
                if (n0 < ix.nqr) ix.p0 = P.pbuf + (size_t)(ix.q0 + n0) * 64 + half * CH;
                if (n1 < ix.nqr) ix.p1 = P.pbuf + (size_t)(ix.q0 + n1) * 64 + half * CH;
                if (n2 < ix.nqr) ix.p2 = P.pbuf + (size_t)(ix.q0 + n2) * 64 + half * CH;
            }
        }
    };
    auto load_tile = [&](int tile, const MergeIdx &ix, float *u) {
        const int row = tile * TC_ROWS + r;
        const bool live = row < n;
        tile_row0 = tile * TC_ROWS + (warp & 3) * 32;
        if (TMA) {   // the boxes of the previous tile's stores have been read out of the staging regions
            tma_store_wait_read();
            __syncwarp();
        }
        {
            float4 mv[CH / 4];
            bool cov = ix.cov;
            if constexpr (C == 64) {
                if (P.mode == 2) {
                    // the projected rows are fetched warp-cooperatively and blended in the reference's order (same
                    // arithmetic as k_tca_merge)
                    const float *p0 = ix.p0, *p1 = ix.p1, *p2 = ix.p2;
                    const float w0 = ix.w0, w1 = ix.w1, w2 = ix.w2;
                    float4 v[CPR];
                    if (TMA_LD) {
                        // x: one box load into the warp's region, in flight together with the gathers of p0 / p1
                        if (elect_one()) {
                            bulk_expect(xbar, TMA_BOX_BYTES);
                            tma_load_box(stg2, &tm_x, half * CH, tile_row0, xbar);
                        }
                        __syncwarp();
                        float4 va[CPR], vb[CPR], xv[CPR];
#pragma unroll
                        for (int i = 0; i < CPR; ++i) {
                            const float4 *pa = (const float4 *)__shfl_sync(0xffffffffu, (unsigned long long)p0, RPI * i + st_row);
                            va[i] = pa ? __ldg(pa + st_ch) : make_float4(0.f, 0.f, 0.f, 0.f);
                        }
#pragma unroll
                        for (int i = 0; i < CPR; ++i) {
                            const float4 *pb = (const float4 *)__shfl_sync(0xffffffffu, (unsigned long long)p1, RPI * i + st_row);
                            vb[i] = pb ? __ldg(pb + st_ch) : make_float4(0.f, 0.f, 0.f, 0.f);
                        }
                        mbar_wait(xbar, xphase);
                        xphase ^= 1u;
#pragma unroll
                        for (int q = 0; q < CPR; ++q) xv[q] = *(const float4 *)(stg2 + tma_swz(lane, q));
                        __syncwarp();
#pragma unroll
                        for (int i = 0; i < CPR; ++i) *(float4 *)(stg + tma_swz(RPI * i + st_row, st_ch)) = va[i];
                        __syncwarp();
#pragma unroll
                        for (int q = 0; q < CPR; ++q) mv[q] = *(const float4 *)(stg + tma_swz(lane, q));
                        __syncwarp();
#pragma unroll
                        for (int i = 0; i < CPR; ++i) *(float4 *)(stg + tma_swz(RPI * i + st_row, st_ch)) = vb[i];
                        __syncwarp();
#pragma unroll
                        for (int q = 0; q < CPR; ++q) v[q] = *(const float4 *)(stg + tma_swz(lane, q));
                        __syncwarp();
#pragma unroll
                        for (int q = 0; q < CPR; ++q)
                            mv[q] = make_float4(__fadd_rn(__fmul_rn(mv[q].x, w0), __fmul_rn(v[q].x, w1)), __fadd_rn(__fmul_rn(mv[q].y, w0), __fmul_rn(v[q].y, w1)),
                                                __fadd_rn(__fmul_rn(mv[q].z, w0), __fmul_rn(v[q].z, w1)), __fadd_rn(__fmul_rn(mv[q].w, w0), __fmul_rn(v[q].w, w1)));
                        gather_in(p2, v);
#pragma unroll
                        for (int c = 0; c < CPR; ++c) {
                            const float4 xx = xv[c];
                            float4 m = make_float4(__fadd_rn(mv[c].x, __fmul_rn(v[c].x, w2)), __fadd_rn(mv[c].y, __fmul_rn(v[c].y, w2)),
                                                   __fadd_rn(mv[c].z, __fmul_rn(v[c].z, w2)), __fadd_rn(mv[c].w, __fmul_rn(v[c].w, w2)));
                            if (!cov) m = xx;
                            u[4 * c] = m.x + xx.x; u[4 * c + 1] = m.y + xx.y; u[4 * c + 2] = m.z + xx.z; u[4 * c + 3] = m.w + xx.w;
                        }
                        return;
                    }
#ifndef FFN_BLEND_PER_ROW
                    // The blend happens in the lanes that FETCH the rows (RPI rows per instruction, CPR lanes per row):
                    // a lane holds the same 16-byte chunk of the row's three projected rows and of x, blends them in
                    // the reference's order and parks ONE chunk of u = (covered ? blend : x) + x -- one shared-memory
                    // store and load per part row instead of four of each.  Half of the row groups per memory round trip.
                    // (48.8 -> 45.5 us per launch with split bf16 operands, 44.2 -> 40.8 us with bf16; with split TF32
                    //  operands -- one CTA per SM, loads under the second GEMM -- 65.3 -> 66.3 us: the per-row form stays there.)
                    if constexpr (TERMS != 3) {
                    if (!live) cov = false;
#pragma unroll
                    for (int h2 = 0; h2 < 2; ++h2) {
                        float4 ua[CPR / 2];
                        {
                            float4 va[CPR / 2], vb[CPR / 2], vc[CPR / 2], vx[CPR / 2];
                            float wa[CPR / 2], wb[CPR / 2], wc[CPR / 2];
                            bool cv[CPR / 2];
#pragma unroll
                            for (int j = 0; j < CPR / 2; ++j) {
                                const int rr = RPI * (h2 * (CPR / 2) + j) + st_row, grow = tile_row0 + rr;
                                const float4 *pa = (const float4 *)__shfl_sync(0xffffffffu, (unsigned long long)p0, rr);
                                const float4 *pb = (const float4 *)__shfl_sync(0xffffffffu, (unsigned long long)p1, rr);
                                const float4 *pc = (const float4 *)__shfl_sync(0xffffffffu, (unsigned long long)p2, rr);
                                wa[j] = __shfl_sync(0xffffffffu, w0, rr); wb[j] = __shfl_sync(0xffffffffu, w1, rr);
                                wc[j] = __shfl_sync(0xffffffffu, w2, rr);
                                cv[j] = __shfl_sync(0xffffffffu, (int)cov, rr) != 0;
                                const float4 zero = make_float4(0.f, 0.f, 0.f, 0.f);
                                va[j] = pa ? __ldg(pa + st_ch) : zero;
                                vb[j] = pb ? __ldg(pb + st_ch) : zero;
                                vc[j] = pc ? __ldg(pc + st_ch) : zero;
                                vx[j] = grow < n ? __ldg((const float4 *)(x + (size_t)grow * C + half * CH) + st_ch) : zero;
                            }
#pragma unroll
                            for (int j = 0; j < CPR / 2; ++j) {
                                float4 m;
                                m.x = __fadd_rn(__fadd_rn(__fmul_rn(va[j].x, wa[j]), __fmul_rn(vb[j].x, wb[j])), __fmul_rn(vc[j].x, wc[j]));
                                m.y = __fadd_rn(__fadd_rn(__fmul_rn(va[j].y, wa[j]), __fmul_rn(vb[j].y, wb[j])), __fmul_rn(vc[j].y, wc[j]));
                                m.z = __fadd_rn(__fadd_rn(__fmul_rn(va[j].z, wa[j]), __fmul_rn(vb[j].z, wb[j])), __fmul_rn(vc[j].z, wc[j]));
                                m.w = __fadd_rn(__fadd_rn(__fmul_rn(va[j].w, wa[j]), __fmul_rn(vb[j].w, wb[j])), __fmul_rn(vc[j].w, wc[j]));
                                if (!cv[j]) m = vx[j];
                                ua[j] = make_float4(m.x + vx[j].x, m.y + vx[j].y, m.z + vx[j].z, m.w + vx[j].w);
                            }
                        }
#pragma unroll
                        for (int j = 0; j < CPR / 2; ++j) {
                            const int rr = RPI * (h2 * (CPR / 2) + j) + st_row;
                            *(float4 *)(stg + rr * (CH * 4) + ((st_ch ^ ((rr >> SWS) & (CPR - 1))) << 4)) = ua[j];
                        }
                    }
                    __syncwarp();
#pragma unroll
                    for (int q = 0; q < CPR; ++q) {
                        const float4 t = *(const float4 *)(stg + lane * (CH * 4) + ((q ^ ((lane >> SWS) & (CPR - 1))) << 4));
                        u[4 * q] = t.x; u[4 * q + 1] = t.y; u[4 * q + 2] = t.z; u[4 * q + 3] = t.w;
                    }
                    __syncwarp();
                    return;
                    }
#endif
                    gather_in2(p0, mv, p1, nullptr, v);
#pragma unroll
                    for (int q = 0; q < CPR; ++q)
                        mv[q] = make_float4(__fadd_rn(__fmul_rn(mv[q].x, w0), __fmul_rn(v[q].x, w1)), __fadd_rn(__fmul_rn(mv[q].y, w0), __fmul_rn(v[q].y, w1)),
                                            __fadd_rn(__fmul_rn(mv[q].z, w0), __fmul_rn(v[q].z, w1)), __fadd_rn(__fmul_rn(mv[q].w, w0), __fmul_rn(v[q].w, w1)));
                    float4 xv[CPR];
                    gather_in2(p2, v, nullptr, x, xv);
                    if (!live) cov = false;
#pragma unroll
                    for (int c = 0; c < CPR; ++c) {
                        const float4 xx = xv[c];
                        float4 m = make_float4(__fadd_rn(mv[c].x, __fmul_rn(v[c].x, w2)), __fadd_rn(mv[c].y, __fmul_rn(v[c].y, w2)),
                                               __fadd_rn(mv[c].z, __fmul_rn(v[c].z, w2)), __fadd_rn(mv[c].w, __fmul_rn(v[c].w, w2)));
                        if (!cov) m = xx;
                        u[4 * c] = m.x + xx.x; u[4 * c + 1] = m.y + xx.y; u[4 * c + 2] = m.z + xx.z; u[4 * c + 3] = m.w + xx.w;
                    }
                    return;
                }
            }
            if (P.mode != 2) stage_in(merged, mv);
            if (P.mode == 0) {
#pragma unroll
                for (int c = 0; c < CH / 4; ++c) {
                    u[4 * c] = mv[c].x; u[4 * c + 1] = mv[c].y; u[4 * c + 2] = mv[c].z; u[4 * c + 3] = mv[c].w;
                }
            } else {
                float4 xv[CH / 4];
                stage_in(x, xv);
                if (P.mode == 1) cov = live && covered[row] != 0;  // (uncovered rows of merged are never written)
#pragma unroll
                for (int c = 0; c < CH / 4; ++c) {
                    const float4 v = xv[c], m = cov ? mv[c] : v;
                    u[4 * c] = m.x + v.x; u[4 * c + 1] = m.y + v.y; u[4 * c + 2] = m.z + v.z; u[4 * c + 3] = m.w + v.w;
                }
            }
        }
    };

    // PIPE: the loads of tile i + 1 are software-pipelined under the GEMMs of tile i.  Pays with ONE CTA per SM
    // (3xTF32: 90 -> 73 us per launch); with two CTAs per SM the other CTA already fills those gaps and the 64 extra
    // live registers only cause spills, so the plain order is kept there.
    constexpr bool PIPE = TERMS == 3;
    float u[CH];
    MergeIdx ix;
    if ((int)blockIdx.x < tiles) {
        index_stage0(blockIdx.x, ix);
        index_stage1(ix);
        index_stage2(ix);
        if (PIPE) load_tile(blockIdx.x, ix, u);
    }
    for (int tile = blockIdx.x; tile < tiles; tile += gridDim.x, phase ^= 1u) {
        TRACE(0);
        const int row = tile * TC_ROWS + r;
        const bool live = row < n;
        (void)live;
        // (both orders: the index chain of the NEXT tile's rows is issued in stages during this tile, see above)
        const int next_tile = tile + (int)gridDim.x;
        if (!PIPE) load_tile(tile, ix, u);
        if (next_tile < tiles) index_stage0(next_tile, ix);
        {   // rows two tiles ahead: start them on their way from HBM to L2 now
            const int nrow = row + 2 * (int)gridDim.x * TC_ROWS;
            if (nrow < n) {
                if (P.mode != 2) asm volatile("prefetch.global.L2 [%0];" ::"l"(merged + (size_t)nrow * C + half * CH));
                if (P.mode != 0) asm volatile("prefetch.global.L2 [%0];" ::"l"(x + (size_t)nrow * C + half * CH));
            }
        }
        // ---- 1. LayerNorm of u, A operand
        if (TMA) tma_store_wait_read();   // (the A tiles host the boxes of the previous tile's stores; barriers follow)
        float part = 0.f;
#pragma unroll
        for (int c = 0; c < CH; ++c) part += u[c];
        TRACE(1);
        red_mine[0] = part;
        __syncthreads();
        const float mean = red_sum(0) * (1.0f / C);
        part = 0.f;
#pragma unroll
        for (int c = 0; c < CH; ++c) { const float d = u[c] - mean; part = fmaf(d, d, part); }
        red_mine[TPR * TC_ROWS] = part;
        __syncthreads();
        const float rstd = rsqrtf(red_sum(1) * (1.0f / C) + P.eps);
        if constexpr (BF) {
            // 16-byte chunks of 8 bf16: chunk = channel / 8
#pragma unroll
            for (int c = 0; c < CH / 8; ++c) {
                uint32_t w[4], wm[4];
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const int i = 8 * c + 2 * q;
                    const float v0 = (u[i] - mean) * rstd * s_g[i] + s_b[i], v1 = (u[i + 1] - mean) * rstd * s_g[i + 1] + s_b[i + 1];
                    if (TERMS == 2) split_bf16x2(v0, v1, w[q], wm[q]);
                    else w[q] = pack_bf16x2(v0, v1);
                }
                *(uint4 *)(sA + (uint32_t)(half * (CH / 8) + c) * a_lbo + my_row_off) = make_uint4(w[0], w[1], w[2], w[3]);
                if (TERMS == 2)   // second tile: the mid parts
                    *(uint4 *)(sA + TC_ROWS * C * 2 + (uint32_t)(half * (CH / 8) + c) * a_lbo + my_row_off) = make_uint4(wm[0], wm[1], wm[2], wm[3]);
            }
        } else {
#pragma unroll
            for (int c = 0; c < CH / 4; ++c) {
                float4 v, hi, lo;
                v.x = (u[4 * c] - mean) * rstd * s_g[4 * c] + s_b[4 * c];
                v.y = (u[4 * c + 1] - mean) * rstd * s_g[4 * c + 1] + s_b[4 * c + 1];
                v.z = (u[4 * c + 2] - mean) * rstd * s_g[4 * c + 2] + s_b[4 * c + 2];
                v.w = (u[4 * c + 3] - mean) * rstd * s_g[4 * c + 3] + s_b[4 * c + 3];
                split_tf32(v, hi, lo);
                *(float4 *)(sA + (uint32_t)(half * (CH / 4) + c) * a_lbo + my_row_off) = hi;
                if (TERMS == 3) *(float4 *)(sA + TC_ROWS * C * 4 + (uint32_t)(half * (CH / 4) + c) * a_lbo + my_row_off) = lo;
            }
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncthreads();
        if (tile == (int)blockIdx.x) mbar_wait(bar_w, 0);  // (first tile: the weight copies overlapped the loads and the LayerNorm)
        TRACE(2);
        // ---- 2. D1[128 x F] = A[128 x C] . W1^T, one K = 8 slice (two 16-byte chunks) per MMA
        if (issuer_elected()) {
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            if constexpr (BF) {
#pragma unroll
                for (int k = 0; k < C / 16; ++k) {     // K = 16 per MMA: two 16-byte chunks of 8 elements
                    const uint64_t ah = umma_desc_at(dA, (uint32_t)k * 2u * a_lbo), bh = umma_desc_at(dW1, (uint32_t)k * 2u * w1_lbo);
                    umma_bf16(tmem_d1, ah, bh, idesc1, k > 0 ? 1u : 0u);
                    if (TERMS == 2) {
                        umma_bf16(tmem_d1, umma_desc_at(dA, (uint32_t)(TC_ROWS * C * 2) + (uint32_t)k * 2u * a_lbo), bh, idesc1, 1u);
                        umma_bf16(tmem_d1, ah, umma_desc_at(dW1, (uint32_t)(F * C * 2) + (uint32_t)k * 2u * w1_lbo), idesc1, 1u);
                    }
                }
            } else {
#pragma unroll
                for (int k = 0; k < C / 8; ++k) {
                    const uint64_t ah = umma_desc_at(dA, (uint32_t)k * 2u * a_lbo), bh = umma_desc_at(dW1, (uint32_t)k * 2u * w1_lbo);
                    umma_tf32(tmem_d1, ah, bh, idesc1, k > 0 ? 1u : 0u);
                    if (TERMS == 3) {
                        umma_tf32(tmem_d1, umma_desc_at(dA, (uint32_t)(TC_ROWS * C * 4) + (uint32_t)k * 2u * a_lbo), bh, idesc1, 1u);
                        umma_tf32(tmem_d1, ah, umma_desc_at(dW1, (uint32_t)(F * C * 4) + (uint32_t)k * 2u * w1_lbo), idesc1, 1u);
                    }
                }
            }
            umma_commit(bar1);
        }
        if (next_tile < tiles) index_stage1(ix);
        TRACE(3);
        mbar_wait(bar1, phase);
        TRACE(4);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        // ---- 3. hidden = relu(D1 + b1), written back over D1: the A operand of the second GEMM
        for (int c0 = 0; c0 < FH; c0 += 32) {
            float d[32];
            const uint32_t col = tmem_d1 + lane_off + (uint32_t)(half * FH + c0);
            tmem_ld32(col, d);
            if constexpr (BF) {
                // relu(D1 + b1) as packed bf16 pairs: hidden element k lives in column k / 2 of the packed tile
                uint32_t w[16], wm[16];
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                    const float h0 = fmaxf(d[2 * i] + s_b1[c0 + 2 * i], 0.f), h1 = fmaxf(d[2 * i + 1] + s_b1[c0 + 2 * i + 1], 0.f);
                    if (TERMS == 2) split_bf16x2(h0, h1, w[i], wm[i]);
                    else w[i] = pack_bf16x2(h0, h1);
                }
                tmem_st16(tmem_hlo + lane_off + (uint32_t)((half * FH + c0) / 2), w);
                // the mid parts go back IN PLACE over the first half of this thread's own D1 columns (already read:
                // hidden element k = half * FH + j sits in column half * FH + j / 2)
                if (TERMS == 2) tmem_st16(tmem_d1 + lane_off + (uint32_t)(half * FH + c0 / 2), wm);
                continue;
            }
            if (TERMS == 3) {
                float lo[32];
#pragma unroll
                for (int i = 0; i < 32; ++i) split_tf32(fmaxf(d[i] + s_b1[c0 + i], 0.f), d[i], lo[i]);
                tmem_st32(tmem_hlo + lane_off + (uint32_t)(half * FH + c0), lo);
            } else {
#pragma unroll
                for (int i = 0; i < 32; ++i) d[i] = to_tf32(fmaxf(d[i] + s_b1[c0 + i], 0.f));
            }
            tmem_st32(col, d);
        }
        tmem_st_wait();
        TRACE(5);
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncthreads();
        TRACE(6);
        // ---- 4. D2[128 x C] = H[128 x F] . W2^T, H read from TMEM
        if (issuer_elected()) {
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            if constexpr (BF) {
                for (int k = 0; k < F / 16; ++k) {     // A: 8 packed columns per K = 16 step
                    const uint64_t db = umma_desc_at(dW2, (uint32_t)k * 2u * w2_lbo);
                    umma_bf16_ts(tmem_d2, tmem_hlo + (uint32_t)k * 8u, db, idesc2, k > 0 ? 1u : 0u);
                    if (TERMS == 2) {
                        const int hk = 16 * k, hh = hk / FH;       // (mid parts: per half of the hidden row, see above)
                        umma_bf16_ts(tmem_d2, tmem_d1 + (uint32_t)(hh * FH + (hk - hh * FH) / 2), db, idesc2, 1u);
                        umma_bf16_ts(tmem_d2, tmem_hlo + (uint32_t)k * 8u,
                                     umma_desc_at(dW2, (uint32_t)(C * F * 2) + (uint32_t)k * 2u * w2_lbo), idesc2, 1u);
                    }
                }
            } else
            for (int k = 0; k < F / 8; ++k) {
                const uint64_t db = umma_desc_at(dW2, (uint32_t)k * 2u * w2_lbo);
                umma_tf32_ts(tmem_d2, tmem_d1 + (uint32_t)k * 8u, db, idesc2, k > 0 ? 1u : 0u);
                if (TERMS == 3) {
                    umma_tf32_ts(tmem_d2, tmem_hlo + (uint32_t)k * 8u, db, idesc2, 1u);
                    umma_tf32_ts(tmem_d2, tmem_d1 + (uint32_t)k * 8u,
                                 umma_desc_at(dW2, (uint32_t)(C * F * 4) + (uint32_t)k * 2u * w2_lbo), idesc2, 1u);
                }
            }
            umma_commit(bar2);
        }
        TRACE(7);
        // ---- software pipeline: the next tile's loads run while the second GEMM does (the A tile, which hosts the
        //      staging areas, is free again: the first GEMM is complete, the second reads its A operand from TMEM)
        float un[PIPE ? CH : 1];
        if (next_tile < tiles) {
            index_stage2(ix);
            if (PIPE) load_tile(next_tile, ix, un);
        }
        tile_row0 = tile * TC_ROWS + (warp & 3) * 32;      // (back to this tile for the stores below)
        mbar_wait(bar2, phase);
        TRACE(8);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        // ---- 5. y = u + D2 + b2 (in place in u), optionally xn_next = LayerNorm_next(y)
        {
            float d[CH];
            if constexpr (CH == 32) tmem_ld32(tmem_d2 + lane_off + (uint32_t)(half * 32), d);
            else tmem_ld16(tmem_d2 + lane_off + (uint32_t)(half * 16), d);
#pragma unroll
            for (int i = 0; i < CH; ++i) u[i] += d[i] + s_b2[i];
        }
        {
            float4 o[CH / 4];
#pragma unroll
            for (int q = 0; q < CH / 4; ++q) o[q] = make_float4(u[4 * q], u[4 * q + 1], u[4 * q + 2], u[4 * q + 3]);
            if (TMA) box_out(&tm_y, stg, o);
            else stage_out(y, o);
        }
        if (xn_next) {  // (uniform over the CTA)
            part = 0.f;
#pragma unroll
            for (int c = 0; c < CH; ++c) part += u[c];
            red_mine[2 * TPR * TC_ROWS] = part;
            __syncthreads();
            const float m2 = red_sum(2) * (1.0f / C);
            part = 0.f;
#pragma unroll
            for (int c = 0; c < CH; ++c) { const float dd = u[c] - m2; part = fmaf(dd, dd, part); }
            red_mine[3 * TPR * TC_ROWS] = part;
            __syncthreads();
            const float r2 = rsqrtf(red_sum(3) * (1.0f / C) + P.next_eps);
            float4 o[CH / 4];
#pragma unroll
            for (int q = 0; q < CH / 4; ++q)
                o[q] = make_float4((u[4 * q] - m2) * r2 * s_ng[4 * q] + s_nb[4 * q],
                                   (u[4 * q + 1] - m2) * r2 * s_ng[4 * q + 1] + s_nb[4 * q + 1],
                                   (u[4 * q + 2] - m2) * r2 * s_ng[4 * q + 2] + s_nb[4 * q + 2],
                                   (u[4 * q + 3] - m2) * r2 * s_ng[4 * q + 3] + s_nb[4 * q + 3]);
            if (TMA) {
                if (TERMS != 3) {   // one box per warp: y's box has been read out of it by now
                    tma_store_wait_read();
                    __syncwarp();
                }
                box_out(&tm_xn, stg2, o);
            } else {
                stage_out(xn_next, o);
            }
        }
        TRACE(9);
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncthreads();  // TMEM and the operand tiles are free for the next tile
        TRACE(10);
#ifdef MSSVT_TRACE
        if (tile == (int)blockIdx.x) gt[3] = gtime();
        gt[4] = gtime();
        gt[5] += 1;
#endif
        if (PIPE && next_tile < tiles) {
#pragma unroll
            for (int c = 0; c < (PIPE ? CH : 1); ++c) u[c] = un[c];
        }
    }
#ifdef MSSVT_TRACE
    if (tid == 0 && (blockIdx.x % 41) == 0)
        printf("ffn cta %3d (sm %2u): entry %llu | prologue +%llu | pdl wait +%llu | first tile +%llu | last tile (%llu) +%llu ns\n",
               blockIdx.x, [] { unsigned s; asm("mov.u32 %0, %%smid;" : "=r"(s)); return s; }(), gt[0] % 100000000ull,
               gt[1] - gt[0], gt[2] - gt[0], gt[3] - gt[0], gt[5], gt[4] - gt[0]);
    if (tid == 0 && blockIdx.x == 0 && tr[10])
        printf("ffn tile: load+sum %lld | LN+A+sync %lld | issue1 %lld | wait1 %lld | epi1 %lld | sync %lld | issue2 %lld | wait2 %lld | epi2+store %lld | sync %lld | total %lld clk\n",
               tr[1] - tr[0], tr[2] - tr[1], tr[3] - tr[2], tr[4] - tr[3], tr[5] - tr[4], tr[6] - tr[5], tr[7] - tr[6], tr[8] - tr[7], tr[9] - tr[8], tr[10] - tr[9], tr[10] - tr[0]);
#endif
    if (TMA) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");   // the last tile's rows are on their way out
    if (warp == 0) tmem_dealloc(tmem_base, tmem_cols);
}

// row-major [n_rows][k] fp32 -> canonical K-major UMMA layout (8-row x 16-byte core matrices), TF32-rounded
__global__ void k_pack_operand_tf32(const float *__restrict__ src, int n_rows, int k, int terms, float *__restrict__ dst) {
    const int chunks = k >> 2;
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n_rows * chunks) return;
    const int n = e / chunks, c = e - n * chunks;
    const float4 v = __ldg((const float4 *)(src + (size_t)n * k) + c);
    float4 hi, lo;
    split_tf32(v, hi, lo);
    char *at = (char *)dst + (size_t)c * n_rows * 16 + (n >> 3) * 128 + (n & 7) * 16;
    *(float4 *)at = hi;
    if (terms == 3) *(float4 *)(at + (size_t)n_rows * k * 4) = lo;   // second tile: the low parts
}

// row-major [n_rows][k] fp32 -> canonical K-major UMMA layout of bf16 (8-row x 16-byte core matrices = 8 elements)
__global__ void k_pack_operand_bf16(const float *__restrict__ src, int n_rows, int k, int tiles, uint4 *__restrict__ dst) {
    const int chunks = k >> 3;
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n_rows * chunks) return;
    const int n = e / chunks, c = e - n * chunks;
    const float4 a = __ldg((const float4 *)(src + (size_t)n * k) + 2 * c), b = __ldg((const float4 *)(src + (size_t)n * k) + 2 * c + 1);
    char *at = (char *)dst + (size_t)c * n_rows * 16 + (n >> 3) * 128 + (n & 7) * 16;
    uint32_t h[4], m[4];
    split_bf16x2(a.x, a.y, h[0], m[0]); split_bf16x2(a.z, a.w, h[1], m[1]);
    split_bf16x2(b.x, b.y, h[2], m[2]); split_bf16x2(b.z, b.w, h[3], m[3]);
    *(uint4 *)at = make_uint4(h[0], h[1], h[2], h[3]);
    if (tiles == 2) *(uint4 *)(at + (size_t)n_rows * k * 2) = make_uint4(m[0], m[1], m[2], m[3]);   // second tile: the mid parts
}

// Self-check of the tensor-map row movement the FFN relies on: (rows, 64) fp32 src -> dst through TMA box loads,
// per-lane reads / writes of the swizzled boxes and TMA box stores, with the FFN's own thread -> row mapping.
__global__ void __launch_bounds__(256)
k_tma_copy_rows(const __grid_constant__ CUtensorMap src_map, const __grid_constant__ CUtensorMap dst_map, int rows) {
    extern __shared__ __align__(1024) char smem_raw[];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    char *box = smem_raw + warp * TMA_BOX_BYTES;
    uint64_t *bars = (uint64_t *)(smem_raw + 8 * TMA_BOX_BYTES);
    const uint32_t bar = smem_u32(bars + warp);
    if (lane == 0) {
        mbar_init(bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
    uint32_t phase = 0;
    const int tiles = (rows + TC_ROWS - 1) / TC_ROWS;
    for (int tile = blockIdx.x; tile < tiles; tile += gridDim.x, phase ^= 1u) {
        const int row0 = tile * TC_ROWS + (warp & 3) * 32, col0 = (warp >> 2) * 32;
        tma_store_wait_read();
        __syncwarp();
        if (elect_one()) {
            bulk_expect(bar, TMA_BOX_BYTES);
            tma_load_box(box, &src_map, col0, row0, bar);
        }
        __syncwarp();
        mbar_wait(bar, phase);
        float4 v[8];
#pragma unroll
        for (int c = 0; c < 8; ++c) v[c] = *(const float4 *)(box + tma_swz(lane, c));
        __syncwarp();
#pragma unroll
        for (int c = 0; c < 8; ++c) *(float4 *)(box + tma_swz(lane, c)) = v[c];
        fence_async_smem();
        __syncwarp();
        if (elect_one()) {
            tma_store_box(&dst_map, col0, row0, box);
            tma_store_commit();
        }
        __syncwarp();
    }
    tma_store_wait_read();
}

}  // namespace mssvt

using namespace mssvt;

extern "C" {

// Copies a row-major (num_rows, 64) fp32 matrix through the tensor-map (TMA) row movement of the tensor-core FFN:
// box loads, swizzled shared-memory boxes read and written by the owning lanes, box stores.  A self-check of that
// plumbing (driver entry point, tensor maps, swizzle, bounds clipping of the last tile); dst == src bit for bit.
int mssvt_tma_copy_rows(const float *src, float *dst, int num_rows, void *stream) {
    if (!src || !dst || num_rows < 0) return MSSVT_ERR_INVALID;
    if (num_rows == 0) return MSSVT_OK;
    CUtensorMap ms, md;
    if (!tma_rows_map(&ms, src, num_rows, 64) || !tma_rows_map(&md, dst, num_rows, 64)) return MSSVT_ERR_LAUNCH;
    const int tiles = (num_rows + TC_ROWS - 1) / TC_ROWS;
    const int grid = tiles < MSSVT_NUM_SMS * 4 ? tiles : MSSVT_NUM_SMS * 4;
    const size_t smem = 8 * TMA_BOX_BYTES + 64;
    cudaFuncSetAttribute(k_tma_copy_rows, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    ++g_launches;
    k_tma_copy_rows<<<grid, 256, smem, (cudaStream_t)stream>>>(ms, md, num_rows);
    return check_launch();
}

// The bf16 form of mssvt_pack_operand_tf32 (precision mode "bf16", tcgen05.mma.kind::f16): packed holds
// n_rows * k bf16 (2 bytes each).  n_rows % 8 == 0, k % 16 == 0.
int mssvt_pack_operand_bf16(const float *w, int n_rows, int k, void *packed, void *stream) {
    if (!w || !packed || n_rows <= 0 || k <= 0 || (n_rows & 7) || (k & 15)) return MSSVT_ERR_INVALID;
    const int n = n_rows * (k >> 3);
    ++g_launches;
    k_pack_operand_bf16<<<(n + 255) / 256, 256, 0, (cudaStream_t)stream>>>(w, n_rows, k, 1, (uint4 *)packed);
    return check_launch();
}

// The split form ("bf16x3", terms = 2 of the *_tc entry points): w = w_hi + w_mid, two bf16 each; packed = [hi | mid] =
// 2 * n_rows * k bf16.
int mssvt_pack_operand_bf16x2(const float *w, int n_rows, int k, void *packed, void *stream) {
    if (!w || !packed || n_rows <= 0 || k <= 0 || (n_rows & 7) || (k & 15)) return MSSVT_ERR_INVALID;
    const int n = n_rows * (k >> 3);
    ++g_launches;
    k_pack_operand_bf16<<<(n + 255) / 256, 256, 0, (cudaStream_t)stream>>>(w, n_rows, k, 2, (uint4 *)packed);
    return check_launch();
}

// Packs a weight matrix w [n_rows][k] (nn.Linear layout: out x in) for the tensor-core kernels: TF32
// rounding + the K-major core-matrix layout tcgen05.mma reads from shared memory.  Done once per
// weight; the *_tc entry points take the packed copies.  n_rows % 8 == 0, k % 8 == 0.
int mssvt_pack_operand_tf32(const float *w, int n_rows, int k, int terms, float *packed, void *stream) {
    if (!w || !packed || n_rows <= 0 || k <= 0 || (n_rows & 7) || (k & 7) || (terms != 1 && terms != 3))
        return MSSVT_ERR_INVALID;
    const int n = n_rows * (k >> 2);
    ++g_launches;
    k_pack_operand_tf32<<<(n + 255) / 256, 256, 0, (cudaStream_t)stream>>>(w, n_rows, k, terms, packed);
    return check_launch();
}

// Tensor-core FFN (fp32 accumulate; terms 1: TF32 operands, 3: split 3xTF32 operands, 0: bf16 operands).  w1 [F][C]
// and w2 [C][F] (nn.Linear layout) packed by mssvt_pack_operand_tf32 (terms 1 / 3) or mssvt_pack_operand_bf16 (terms 0).  Supported: C in {32, 64}, F a multiple of 64 with F + C <= 512 and the operand
// tiles fitting in shared memory; returns MSSVT_ERR_INVALID otherwise (callers then use mssvt_ffn).
int mssvt_ffn_tc(int C, int F, int mode, int terms, float eps, const float *ln_g, const float *ln_b, const float *w1,
                 const float *b1, const float *w2, const float *b2, int num_rows, const int *num_rows_dev,
                 const float *x, const float *merged, const unsigned char *covered, float *y,
                 const float *next_ln_g, const float *next_ln_b, float next_eps, float *xn_next,
                 const int *vox_slot, const int *meta, const int *q_base, const unsigned char *nn_idx,
                 const float *nn_w, const float *projected, int cap1, void *stream) {
    const bool bf = terms == 0 || terms == 2;
    if ((C != 32 && C != 64) || F <= 0 || (F & 63) || F + C + (terms == 3 ? F : bf ? F / 2 : 0) > 512 || num_rows < 0 ||
        (terms != 0 && terms != 1 && terms != 2 && terms != 3))
        return MSSVT_ERR_INVALID;
    if (num_rows == 0) return MSSVT_OK;
    if (mode < 0 || mode > 2 || !ln_g || !ln_b || !w1 || !b1 || !w2 || !b2 || !y || (mode != 2 && !merged) ||
        (mode == 1 && (!x || !covered)))
        return MSSVT_ERR_INVALID;
    if (mode == 2 && (C != 64 || !x || !vox_slot || !meta || !q_base || !nn_idx || !nn_w || !projected || cap1 <= 0))
        return MSSVT_ERR_INVALID;
    const int nt = terms == 3 ? 2 : 1;
    // threads per row: four where the shape allows it (16 channels and F / 4 hidden columns per thread)
    // (measured at 150 k rows: 54 -> 62 us with TF32 operands, 66 -> 70 us with split operands: the tile time is made of
    //  memory round trips and MMA round trips, not of the per-thread instruction chains -- the default stays two)
#ifdef FFN_TPR4
    const int tpr = (C == 64 && (F & 127) == 0 && terms != 2) ? 4 : 2;
#else
    const int tpr = 2;
#endif
    size_t smem = nt * ((size_t)TC_ROWS * C * 4 + 2 * (size_t)F * C * 4) + (size_t)(5 * C + F + 4 * tpr * TC_ROWS) * 4 + (3 + 4 * tpr) * 8 + 16 + 128;
    if (bf)  // bf16 operands: the A region keeps the size of the fp32 staging area, the weights halve (one or two tiles each)
        smem = (size_t)TC_ROWS * C * 4 + (terms == 2 ? 2 : 1) * 2 * (size_t)F * C * 2 + (size_t)(5 * C + F + 4 * tpr * TC_ROWS) * 4 + (3 + 4 * tpr) * 8 + 16 + 128;
    if (smem > 227 * 1024) return MSSVT_ERR_INVALID;
    if (xn_next && (!next_ln_g || !next_ln_b)) return MSSVT_ERR_INVALID;
    FfnTcParams P = {F, mode, eps, w1, b1, w2, b2, ln_g, ln_b, next_ln_g, next_ln_b, next_eps,
                     vox_slot, meta, q_base, nn_idx, nn_w, projected, cap1};
    // dense row tiles by TMA (tma.cuh): tensor maps over the capacity rows of x (mode 2), y and the next LayerNorm rows
    CUtensorMap tm_x, tm_y, tm_xn;
    memset(&tm_x, 0, sizeof(tm_x)); memset(&tm_y, 0, sizeof(tm_y)); memset(&tm_xn, 0, sizeof(tm_xn));
    if (FFN_TMA_SHAPE(C, tpr, terms) && !(tma_rows_map(&tm_y, y, num_rows, C) && (mode != 2 || tma_rows_map(&tm_x, x, num_rows, C)) &&
                                   (!xn_next || tma_rows_map(&tm_xn, xn_next, num_rows, C))))
        return MSSVT_ERR_LAUNCH;   // (no cuTensorMapEncodeTiled in this driver, or a misaligned buffer)
    int tiles = (num_rows + TC_ROWS - 1) / TC_ROWS;
    int tmem_cols = 32;
    while (tmem_cols < F + C + (terms == 3 ? F : bf ? F / 2 : 0)) tmem_cols <<= 1;
    int per_sm = (int)(227 * 1024 / (smem + 1024));
    if (per_sm > 512 / tmem_cols) per_sm = 512 / tmem_cols;
    per_sm = per_sm > 2 ? 2 : per_sm < 1 ? 1 : per_sm;
#ifdef FFN_MAX_PER_SM   // (occupancy experiments)
    if (per_sm > FFN_MAX_PER_SM) per_sm = FFN_MAX_PER_SM;
#endif
    int grid = tiles < MSSVT_NUM_SMS * per_sm ? tiles : MSSVT_NUM_SMS * per_sm;
    ++g_launches;
#define FFN_TC_LAUNCH(CC, TT, RR)                                                                          \
    cudaFuncSetAttribute(k_ffn_tc<CC, TT, RR>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);        \
    launch_pdl(k_ffn_tc<CC, TT, RR>, dim3(grid), dim3(TC_ROWS * RR), smem, (cudaStream_t)stream, P, num_rows,  \
               num_rows_dev, x, merged, covered, y, xn_next, tm_x, tm_y, tm_xn)
#ifdef FFN_TPR4   // (the four-threads-per-row kernels are only built for the experiment)
    if (tpr == 4) {
        if (terms == 0) { FFN_TC_LAUNCH(64, 0, 4); }
        else if (terms == 3) { FFN_TC_LAUNCH(64, 3, 4); }
        else { FFN_TC_LAUNCH(64, 1, 4); }
    } else
#endif
    if (C == 64 && terms == 2) { FFN_TC_LAUNCH(64, 2, 2); }
    else if (terms == 2) { FFN_TC_LAUNCH(32, 2, 2); }
    else if (C == 64 && terms == 0) { FFN_TC_LAUNCH(64, 0, 2); }
    else if (terms == 0) { FFN_TC_LAUNCH(32, 0, 2); }
    else if (C == 64 && terms == 3) { FFN_TC_LAUNCH(64, 3, 2); }
    else if (C == 64) { FFN_TC_LAUNCH(64, 1, 2); }
    else if (terms == 3) { FFN_TC_LAUNCH(32, 3, 2); }
    else { FFN_TC_LAUNCH(32, 1, 2); }
#undef FFN_TC_LAUNCH
    return check_launch();
}

}  // extern "C"
