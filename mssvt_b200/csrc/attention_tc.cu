// attention_tc.cu -- window attention of a two-window MsSVT block, task-parallel, with the K/V
// projection on the tcgen05 tensor cores (sm_100a).  Same mathematics as k_block_attention in
// attention.cu (mssvt_backbone.py:260-336, mssvt_utils.py:88-157), different mapping: instead of one
// warp walking through a window, every stage runs with one THREAD per task over the whole frame:
//
//   k_tc_linear   (tc_linear.cuh) q = (Wq (xn + posemb) + bq) * scale for every real query, on tcgen05
//   k_tca_plan    once per frame geometry: packs consecutive windows of one scale into tiles of <= 128
//                 distinct keys (windows without a real query are skipped), per-window offsets
//   k_tca_keys    thread = distinct key of a window, 128 keys of ONE scale per tile.  The thread
//                 gathers its 32-channel slice of the layer-normed row, adds the positional
//                 embedding and stores the row, TF32-rounded, as row t of the A operand; one thread
//                 issues 4 tcgen05.mma.kind::tf32 (M = 128, N = 64, K = 8): D = A Wkv^T into 64 TMEM
//                 columns; tcgen05.ld hands every thread the 64 K|V values of ITS key (TMEM lane =
//                 thread).  Scores against the window's queries, then softmax (with the multiplicity
//                 of the masked key) and AV with one thread per (query, head, quarter head).
//   k_tc_linear   output projection of the head outputs, on tcgen05
//   k_tca_merge   thread = (win1 voxel, 16 channels): 1/d blend of the 3 nearest query rows -> merged
//
// Queries are addressed by a compact id (q_base[w] + slot, an exclusive scan over the windows done
// with the geometry), so the three intermediates (q, head outputs, projected rows) are dense
// (#queries, 64) fp32 arrays that live in L2.  Compared with the warp-per-window kernel this executes
// ~5x fewer warp instructions, has no lane redundancy on the small per-window matrices, keeps five
// 128-thread CTAs per SM in flight, and the 2048 FMAs per key run on the tensor pipe.
//
// Supported shape (config S0 and relatives): C = 64, two head groups of 32 channels, 1/2/4 heads per
// group, nq <= 32, key_num_sample <= 63, max_num_win1 <= 128.  Everything else runs on
// k_block_attention.  Precision: TF32 operands for the Q, K/V and output projections; positional embedding,
// scores, softmax, AV and interpolation are fp32.
#include "tc_linear.cuh"

namespace mssvt {

#define TCA_THREADS 128
#ifdef MSSVT_TRACE
#define KTRACE(i) do { if (tid == 0 && blockIdx.x == gridDim.x - 1 && t == first + 2 * stride) tr[i] = clock64(); } while (0)
#else
#define KTRACE(i) do {} while (0)
#endif
#define TCA_TW 32        // windows per tile (at most)
#define TCA_SBUD 2048    // score slots per tile: sum over its windows of #queries x #keys x heads
#define TCA_PLAN_WB 64   // windows planned by one warp
#define TCA_C 64
#define TCA_SD 32
#define TCA_VPITCH 36    // V row pitch in floats: 16-byte aligned, conflict-free for quarter warps

struct TcAttnParams {
    int nq, K, cap1, interp, heads, smax;      // smax = nq * heads: score slots per key task
    float scale;
    float win_cell[3], lo[3];
    const float *pos_w, *pos_b;                // [64][6], [64]   (Conv1d weight (64, 6, 1))
    const float *wq, *bq[2];                   // blockdiag(Wq0, Wq1) [64][64] packed, [32] x 2
    const float *wkv[2], *bkv[2];              // [64][32] packed (mssvt_pack_operand_tf32), [64]
    const float *wp, *bp[2];                   // blockdiag(Wp0, Wp1) [64][64] packed, [32] x 2
};

// [64][8] per channel: w0..w5, bias, 0
__device__ __forceinline__ void stage_pos_weights(const TcAttnParams &P, float *sPos) {
    for (int i = threadIdx.x; i < 64 * 8; i += blockDim.x) {
        const int c = i >> 3, k = i & 7;
        sPos[i] = k < 6 ? __ldg(P.pos_w + c * 6 + k) : k == 6 ? __ldg(P.pos_b + c) : 0.f;
    }
}

__device__ __forceinline__ float pos_embed8(const float *sPos, int c, float rx, float ry, float rz, float cx,
                                            float cy, float cz) {
    const float4 wa = *(const float4 *)(sPos + c * 8), wb = *(const float4 *)(sPos + c * 8 + 4);
    float a = wb.z;
    a = fmaf(wa.x, rx, a); a = fmaf(wa.y, ry, a); a = fmaf(wa.z, rz, a);
    a = fmaf(wa.w, cx, a); a = fmaf(wb.x, cy, a); a = fmaf(wb.y, cz, a);
    return fmaxf(a, 0.f);
}

// ------------------------------------------------------------------------------- queries

// rows of k_tc_linear for the query projection: compact query id -> layer-normed row + positional embedding
struct TcaQueryRows {
    int nq, win_cap;
    float win_cell[3], lo[3];
    const float *pos_w, *pos_b;
    const int *win_count_total;
    const int4 *win_list;
    const float *xn, *xyz;
    const int *q_row, *q_base, *q_src;
    __device__ void init(float *sPos) const {
        for (int i = threadIdx.x; i < 64 * 8; i += blockDim.x) {
            const int c = i >> 3, k = i & 7;
            sPos[i] = k < 6 ? __ldg(pos_w + c * 6 + k) : k == 6 ? __ldg(pos_b + c) : 0.f;
        }
    }
    __device__ int rows() const { return __ldg(q_base + min(win_cap, __ldg(win_count_total))); }
    struct Ctx { float rx, ry, rz, cx, cy, cz; };
    __device__ const float4 *src(int qid, int half, Ctx &c) const {
        const int s = __ldg(q_src + qid), w = s / nq;
        const int row = __ldg(q_row + s);
        const int4 win = __ldg(win_list + w);
        c.cx = world_coord(win.w, win_cell[0], lo[0]);
        c.cy = world_coord(win.z, win_cell[1], lo[1]);
        c.cz = world_coord(win.y, win_cell[2], lo[2]);
        c.rx = __fsub_rn(__ldg(xyz + 3 * (size_t)row), c.cx);
        c.ry = __fsub_rn(__ldg(xyz + 3 * (size_t)row + 1), c.cy);
        c.rz = __fsub_rn(__ldg(xyz + 3 * (size_t)row + 2), c.cz);
        return (const float4 *)(xn + (size_t)row * TCA_C + half * TCA_SD);
    }
    __device__ void finish(const Ctx &c, int half, const float *sPos, float *in) const {
#pragma unroll
        for (int i = 0; i < TCA_SD; ++i) in[i] += pos_embed8(sPos, half * TCA_SD + i, c.rx, c.ry, c.rz, c.cx, c.cy, c.cz);
    }
};

// ------------------------------------------------------------------------------- tile plan

// A tile = consecutive windows of ONE scale whose distinct keys fill the 128 rows of an MMA.  The plan is
// a function of the geometry only, so it is made once per frame and shared by every block that uses the
// same window lists.  One warp plans TCA_PLAN_WB windows of one scale: lanes load 32 windows at a time,
// the greedy cut is a 32-step scan over shuffled values (every lane runs it, lane i keeps step i).
// Windows without a real query get no key tasks at all (nobody would read their K/V).
__global__ void __launch_bounds__(256)
k_tca_plan(int heads, int win_cap, const int *__restrict__ win_count_total, const int4 *__restrict__ win_list,
           const int4 *__restrict__ meta, const int *__restrict__ q_base, float3 win_cell, float3 lo,
           int2 *__restrict__ tiles, int *__restrict__ tile_count, int4 *__restrict__ win_rec,
           float4 *__restrict__ win_ctr) {
    const int num_wins = min(win_cap, __ldg(win_count_total));
    const int lane = threadIdx.x & 31;
    const int wid = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int g = wid & 1, w0 = (wid >> 1) * TCA_PLAN_WB;
    if (w0 >= num_wins) return;
    const int w1 = min(w0 + TCA_PLAN_WB, num_wins);
    int ts = w0, a = 0, aq = 0, as = 0, nw = 0;  // open tile: first window, key tasks, queries, score slots, windows
    for (int wb = w0; wb < w1; wb += 32) {
        const int w = wb + lane;
        int nqr = 0, r = 0, mult = 0, qb = 0;
        if (w < w1) {
            const int4 m = __ldg(meta + w);
            const int mm = g ? m.w : m.z;
            nqr = m.x; r = nqr ? mm & 0xff : 0; mult = mm >> 8; qb = __ldg(q_base + w);
            if (g == 0) {
                const int4 win = __ldg(win_list + w);
                win_ctr[w] = make_float4(world_coord(win.w, win_cell.x, lo.x), world_coord(win.z, win_cell.y, lo.y),
                                         world_coord(win.y, win_cell.z, lo.z), 0.f);
            }
        }
        int my_a = 0, my_aq = 0, my_as = 0;
        const int n = min(32, w1 - wb);
        for (int i = 0; i < n; ++i) {
            const int ri = __shfl_sync(0xffffffffu, r, i), qi = __shfl_sync(0xffffffffu, nqr, i);
            const int si = qi * ri * heads;
            if (nw > 0 && (a + ri > TCA_THREADS || aq + qi > TCA_THREADS || as + si > TCA_SBUD || nw == TCA_TW)) {
                if (lane == 0 && a > 0) tiles[(size_t)g * win_cap + atomicAdd(tile_count + g, 1)] = make_int2(ts, nw);
                ts = wb + i; a = aq = as = nw = 0;
            }
            if (i == lane) { my_a = a; my_aq = aq; my_as = as; }
            a += ri; aq += qi; as += si; ++nw;
        }
        if (w < w1)
            win_rec[(size_t)g * win_cap + w] = make_int4(qb, nqr | (r << 8) | (mult << 16), my_a | (my_aq << 8) | (my_as << 16), 0);
    }
    if (lane == 0 && a > 0) tiles[(size_t)g * win_cap + atomicAdd(tile_count + g, 1)] = make_int2(ts, nw);
}

// ------------------------------------------------------------------------------- keys + attention

// largest l in [0, n) with off[l] <= v (off[n] > v)
__device__ __forceinline__ int tile_window(const int *off, int n, int v) {
    int lo = 0, hi = n;
    while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (off[mid] <= v) lo = mid; else hi = mid;
    }
    return lo;
}

template <int HEADS, int TERMS>
__global__ void __launch_bounds__(TCA_THREADS, TERMS == 3 ? 3 : 5)
k_tca_keys(TcAttnParams P, int win_cap, const int2 *__restrict__ tiles, const int *__restrict__ tile_count,
           const int4 *__restrict__ win_rec, const float4 *__restrict__ win_ctr, const float *__restrict__ xn,
           const float *__restrict__ xyz, const int *__restrict__ rep_row, const float *__restrict__ Qbuf,
           float *__restrict__ Obuf) {
    constexpr int HD = TCA_SD / HEADS;
    constexpr int DPT = HD / 4;  // channels per thread in the AV phase
    extern __shared__ __align__(128) char smem_raw[];
    const int tid = threadIdx.x, warp = tid >> 5;
    const int K = P.K;
    pdl_launch_dependents();
    pdl_wait();  // (the scale of this CTA, hence its weights, depends on the tile counts: nothing to do before)

    // ---- every CTA works on one scale for its whole life (8 KB of weights instead of 16); the CTAs of the
    //      two scales are interleaved over the grid in proportion to the tile counts
    const int T0 = __ldg(tile_count), T1 = __ldg(tile_count + 1), G = gridDim.x, b = blockIdx.x;
    int G0 = T0 == 0 ? 0 : T1 == 0 ? G : (int)(((long long)G * T0 + (T0 + T1) / 2) / (T0 + T1));
    if (T0 > 0 && T1 > 0) G0 = min(max(G0, 1), G - 1);
    const int c0b = (int)((long long)b * G0 / G), c0n = (int)((long long)(b + 1) * G0 / G);
    const int g = c0n > c0b ? 0 : 1;
    const int first = g ? b - c0b : c0b, stride = g ? G - G0 : G0, T = g ? T1 : T0;
    tiles += (size_t)g * win_cap;
    win_rec += (size_t)g * win_cap;

    // ---- shared memory: 8 KB weights + 18 KB A/V (aliased) + 8 KB scores + bookkeeping = ~38 KB
    constexpr int NT = TERMS == 3 ? 2 : 1;                  // operand tiles: hi [, lo] (3xTF32, tc_common.cuh)
    constexpr int A_TILE = TCA_THREADS * TCA_SD * 4;        // 16 KB
    constexpr int A_REGION = NT * A_TILE > TCA_THREADS * TCA_VPITCH * 4 ? NT * A_TILE : TCA_THREADS * TCA_VPITCH * 4;
    char *sWkv = smem_raw;                                  // NT x [64 x 32] canonical, TF32      8 KB each
    char *sA = sWkv + NT * 64 * 32 * 4;                     // NT x [128 x 32] canonical (16 KB each) ...
    float *sV = (float *)sA;                                // ... reused as V [128][VPITCH] after the MMA
    float *sPos = (float *)(sA + A_REGION);                 // [32][8] (this scale's channels)
    float *sBkv = sPos + 32 * 8;                            // [64]
    float *sS = sBkv + 64;                                  // [SBUD] scores: window-major, [key][query][head]
    float4 *sCtr = (float4 *)(sS + TCA_SBUD);               // [TW] window centres
    int4 *sRec = (int4 *)(sCtr + TCA_TW);                   // [TW] {q_base, nqr | r << 8 | mult << 16, offsets}
    int *sToff = (int *)(sRec + TCA_TW);                    // [TW + 1] first key task of each window
    int *sQoff = sToff + TCA_TW + 1;                        // [TW + 1] first query of each window
    uint64_t *sBar = (uint64_t *)(sQoff + TCA_TW + 1);      // (2 * (TW + 1) ints: 8-byte aligned)
    uint32_t *sTmem = (uint32_t *)(sBar + 1);

    stage_packed(P.wkv[g], NT * 64 * 32, sWkv);
    for (int i = tid; i < 32 * 8; i += TCA_THREADS) {
        const int c = g * TCA_SD + (i >> 3), k = i & 7;
        sPos[i] = k < 6 ? __ldg(P.pos_w + c * 6 + k) : k == 6 ? __ldg(P.pos_b + c) : 0.f;
    }
    if (tid < 64) sBkv[tid] = __ldg(P.bkv[g] + tid);
    const uint32_t bar = smem_u32(sBar);
    if (tid == 0) {
        mbar_init(bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) tmem_alloc(smem_u32(sTmem), 64);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_d = *sTmem;
    const uint32_t lane_off = (uint32_t)(warp * 32) << 16;
    const uint32_t idesc = umma_idesc_tf32(128, 64);
    const uint32_t sA_u = smem_u32(sA), sWkv_u = smem_u32(sWkv);
    const uint32_t a_lbo = TCA_THREADS * 16, w_lbo = 64 * 16;
    const uint32_t my_row_off = (uint32_t)(tid >> 3) * 128u + (uint32_t)(tid & 7) * 16u;
    uint32_t phase = 0;

#ifdef MSSVT_TRACE
    long long tr[12] = {0};
#endif
    int2 tl_next = first < T ? __ldg(tiles + first) : make_int2(0, 0);
    for (int t = first; t < T; t += stride) {
        KTRACE(0);
        // the tile record is fetched one tile ahead; the next tile's window records and key lists are
        // pulled into L1 while this tile's MMA runs, so that the chain tile -> windows -> key rows ->
        // features starts from cached lines
        const int2 tl = tl_next;
        if (t + stride < T) tl_next = __ldg(tiles + t + stride);
        const int nwin = tl.y;
        if (tid < nwin) {
            const int4 rec = __ldg(win_rec + tl.x + tid);
            sRec[tid] = rec;
            sCtr[tid] = __ldg(win_ctr + tl.x + tid);
            sToff[tid] = rec.z & 0xff;
            sQoff[tid] = (rec.z >> 8) & 0xff;
            if (tid == nwin - 1) {
                sToff[nwin] = (rec.z & 0xff) + ((rec.y >> 8) & 0xff);
                sQoff[nwin] = ((rec.z >> 8) & 0xff) + (rec.y & 0xff);
            }
        }
        __syncthreads();
        KTRACE(1);
        const int nT = sToff[nwin], nQ = sQoff[nwin];
        // the tile's query rows (contiguous ids) are read after the MMA: pull their 128-byte halves into L1 now
        if (tid < nQ) prefetch_l1(Qbuf + (size_t)(sRec[0].x + tid) * TCA_C + g * TCA_SD);

        // ---- key task -> row t of the A operand
        const bool is_task = tid < nT;
        int j = 0, nqr = 0, q0 = 0, soff = 0;
        bool masked = false;
        float rx = 0.f, ry = 0.f, rz = 0.f;  // masked key: relative offset zeroed
        float4 ctr = make_float4(0.f, 0.f, 0.f, 0.f);
        const float4 *src = nullptr;
        if (is_task) {
            const int l = tile_window(sToff, nwin, tid);
            const int4 rec = sRec[l];
            j = tid - (rec.z & 0xff);
            nqr = rec.y & 0xff; q0 = rec.x; soff = rec.z >> 16;
            const int r = (rec.y >> 8) & 0xff, mult = rec.y >> 16;
            masked = mult > 0 && j == r - 1;  // last distinct key stands for all masked slots
            const int row = __ldg(rep_row + (size_t)(tl.x + l) * 2 * K + g * K + j);
            ctr = sCtr[l];
            if (!masked) {  // (coordinates and features travel together: the subtraction waits below)
                rx = __ldg(xyz + 3 * (size_t)row); ry = __ldg(xyz + 3 * (size_t)row + 1);
                rz = __ldg(xyz + 3 * (size_t)row + 2);
            }
            src = (const float4 *)(xn + (size_t)row * TCA_C + g * TCA_SD);
        }
        {
            // the warp gathers its 32 rows together (8 lanes per 128-byte slice) through the A tile's memory
            float4 xv[TCA_SD / 4];
            KTRACE(2);
            warp_rows_load<true>(sA + warp * 4096, src, xv);
            if (is_task && !masked) { rx = __fsub_rn(rx, ctr.x); ry = __fsub_rn(ry, ctr.y); rz = __fsub_rn(rz, ctr.z); }
            KTRACE(3);
            __syncthreads();
            KTRACE(4);  // every warp is done with its staging area: the A tile may be written
            if (is_task) {
#pragma unroll
                for (int c4 = 0; c4 < TCA_SD / 4; ++c4) {
                    const float4 v = xv[c4];
                    float4 o, hi, lo;
                    o.x = v.x + pos_embed8(sPos, 4 * c4, rx, ry, rz, ctr.x, ctr.y, ctr.z);
                    o.y = v.y + pos_embed8(sPos, 4 * c4 + 1, rx, ry, rz, ctr.x, ctr.y, ctr.z);
                    o.z = v.z + pos_embed8(sPos, 4 * c4 + 2, rx, ry, rz, ctr.x, ctr.y, ctr.z);
                    o.w = v.w + pos_embed8(sPos, 4 * c4 + 3, rx, ry, rz, ctr.x, ctr.y, ctr.z);
                    split_tf32(o, hi, lo);
                    *(float4 *)(sA + (uint32_t)c4 * a_lbo + my_row_off) = hi;
                    if (TERMS == 3) *(float4 *)(sA + A_TILE + (uint32_t)c4 * a_lbo + my_row_off) = lo;
                }
            }
        }
        KTRACE(5);
        stage_packed_wait();
        fence_async_smem();
        __syncthreads();
        KTRACE(6);
        // ---- D = A Wkv^T: K in TMEM columns 0..31, V in 32..63
        if (tid == 0) {
            tc_fence_after();
#pragma unroll
            for (int k = 0; k < TCA_SD / 8; ++k)
                umma_step<TERMS>(tmem_d, sA_u + (uint32_t)k * 2u * a_lbo, a_lbo, A_TILE,
                                 sWkv_u + (uint32_t)k * 2u * w_lbo, w_lbo, 64 * 32 * 4, idesc, k == 0);
            umma_commit(bar);
        }
        if (t + stride < T) {  // next tile: window records (16 B each), centres, key lists (K ints per window)
            if (tid < tl_next.y) prefetch_l1(rep_row + (size_t)(tl_next.x + tid) * 2 * K + g * K);
            else if (tid < tl_next.y + (tl_next.y + 7) / 8 + 1) prefetch_l1(win_rec + tl_next.x + 8 * (tid - tl_next.y));
            else if (tid < tl_next.y + 2 * ((tl_next.y + 7) / 8 + 1))
                prefetch_l1(win_ctr + tl_next.x + 8 * (tid - tl_next.y - (tl_next.y + 7) / 8 - 1));
        }
        mbar_wait(bar, phase);
        KTRACE(7);
        phase ^= 1u;
        tc_fence_after();
        // ---- this thread's key: K|V back from TMEM; scores against its window's queries
        {
            float kk[TCA_SD], vv[TCA_SD];
            tmem_ld32(tmem_d + lane_off, kk);
            tmem_ld32(tmem_d + lane_off + 32u, vv);
            tc_fence_before();
            __syncthreads();  // every thread has its K|V in registers: the A tile may become V
            if (is_task) {
                // Biases: q.(k + bk) shifts every score of a query by the same q.bk, which the softmax
                // cancels, so bk is dropped; sum_i p_i (v_i + bv) = sum_i p_i v_i + bv, so bv is added once
                // per output below instead of once per key here.
#pragma unroll
                for (int c4 = 0; c4 < TCA_SD / 4; ++c4)
                    *(float4 *)(sV + tid * TCA_VPITCH + 4 * c4) =
                        make_float4(vv[4 * c4], vv[4 * c4 + 1], vv[4 * c4 + 2], vv[4 * c4 + 3]);
                const float bias = masked ? -100.0f : 0.f;  // additive mask of the reference
                float *srow = sS + soff + j * nqr * HEADS;
                for (int s = 0; s < nqr; ++s) {
                    const float4 *qv = (const float4 *)(Qbuf + (size_t)(q0 + s) * TCA_C + g * TCA_SD);
#pragma unroll
                    for (int h = 0; h < HEADS; ++h) {
                        float a = 0.f;
#pragma unroll
                        for (int d4 = 0; d4 < HD / 4; ++d4) {
                            const float4 q4 = __ldg(qv + h * (HD / 4) + d4);
                            a = fmaf(q4.x, kk[h * HD + 4 * d4], a); a = fmaf(q4.y, kk[h * HD + 4 * d4 + 1], a);
                            a = fmaf(q4.z, kk[h * HD + 4 * d4 + 2], a); a = fmaf(q4.w, kk[h * HD + 4 * d4 + 3], a);
                        }
                        srow[s * HEADS + h] = a + bias;
                    }
                }
            }
        }
        KTRACE(8);
        __syncthreads();
        KTRACE(9);
        // ---- softmax over the window's distinct keys and AV, thread = (query, head, quarter of the head)
        for (int e = tid; e < nQ * HEADS * 4; e += TCA_THREADS) {
            const int dq = e & 3, qh = e >> 2;
            const int h = qh % HEADS, qt = qh / HEADS;
            const int lq = tile_window(sQoff, nwin, qt);
            const int4 rec = sRec[lq];
            const int s = qt - ((rec.z >> 8) & 0xff);
            const int wq = rec.y & 0xff, r = (rec.y >> 8) & 0xff, mult = rec.y >> 16;
            const float *sc = sS + (rec.z >> 16) + s * HEADS + h;  // + key * wq * HEADS
            const int step = wq * HEADS;
            float mx = -3.0e38f;
            for (int k = 0; k < r; ++k) mx = fmaxf(mx, sc[k * step]);
            float den = 0.f, acc[DPT];
#pragma unroll
            for (int d = 0; d < DPT; ++d) acc[d] = 0.f;
            const float *vp = sV + (rec.z & 0xff) * TCA_VPITCH + h * HD + dq * DPT;
            for (int k = 0; k < r; ++k) {
                float wgt = exp_neg(sc[k * step] - mx);
                if (k == r - 1 && mult > 0) wgt *= (float)mult;  // the masked key counts once per masked slot
                den += wgt;
#pragma unroll
                for (int d = 0; d < DPT; ++d) acc[d] = fmaf(wgt, vp[k * TCA_VPITCH + d], acc[d]);
            }
            const float inv = 1.0f / den;
            float *dst = Obuf + (size_t)(rec.x + s) * TCA_C + g * TCA_SD + h * HD + dq * DPT;
            const float *bv = sBkv + TCA_SD + h * HD + dq * DPT;
#pragma unroll
            for (int d = 0; d < DPT; ++d) dst[d] = fmaf(acc[d], inv, bv[d]);
        }
        KTRACE(10);
        __syncthreads();
        KTRACE(11);
    }
#ifdef MSSVT_TRACE
    if (tid == 0 && blockIdx.x == gridDim.x - 1 && tr[11])
        printf("keys tile g%d: header %lld | search+idx %lld | gather %lld | sync %lld | posemb+A %lld | sync %lld | mma %lld | tmem+scores %lld | sync %lld | softmax+AV %lld | sync %lld | total %lld clk\n", g,
               tr[1] - tr[0], tr[2] - tr[1], tr[3] - tr[2], tr[4] - tr[3], tr[5] - tr[4], tr[6] - tr[5], tr[7] - tr[6], tr[8] - tr[7], tr[9] - tr[8], tr[10] - tr[9], tr[11] - tr[10], tr[11] - tr[0]);
#endif
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem_d, 64);
}

// ------------------------------------------------------------------------------- merge

__global__ void __launch_bounds__(256)
k_tca_merge(TcAttnParams P, int win_cap, const int *__restrict__ win_count_total, int num_voxels,
            const int *__restrict__ meta, const int *__restrict__ q_base, const int *__restrict__ q_src,
            const int *__restrict__ q_row, const int *__restrict__ vox_slot,
            const unsigned char *__restrict__ nn_idx, const float *__restrict__ nn_w,
            const float *__restrict__ Pbuf, float *__restrict__ merged) {
    pdl_launch_dependents();
    pdl_wait();
    const int num_wins = min(win_cap, __ldg(win_count_total));
    if (P.interp) {
        // thread = (voxel row, 16 channels): the voxel's win1 slot names its 3 nearest query slots
        const long long total = (long long)num_voxels * 4;
        for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total;
             e += (long long)gridDim.x * blockDim.x) {
            const int cq = (int)(e & 3), row = (int)(e >> 2);
            const int slot = __ldg(vox_slot + row);
            if (slot < 0) continue;  // not in any win1 list: the FFN doubles the shortcut instead
            const int w = slot / P.cap1;
            const int nqr = __ldg(meta + 4 * (size_t)w), q0 = __ldg(q_base + w);
            const unsigned char *ni = nn_idx + (size_t)slot * 3;
            const float *nw = nn_w + (size_t)slot * 3;
            const int n0 = ni[0], n1 = ni[1], n2 = ni[2];
            // padded query slots (index >= #real queries) are zero rows in the reference
            const float4 *a0 = n0 < nqr ? (const float4 *)(Pbuf + (size_t)(q0 + n0) * TCA_C + 16 * cq) : nullptr;
            const float4 *a1 = n1 < nqr ? (const float4 *)(Pbuf + (size_t)(q0 + n1) * TCA_C + 16 * cq) : nullptr;
            const float4 *a2 = n2 < nqr ? (const float4 *)(Pbuf + (size_t)(q0 + n2) * TCA_C + 16 * cq) : nullptr;
            const float w0 = __ldg(nw), w1 = __ldg(nw + 1), w2 = __ldg(nw + 2);
            float4 *dst = (float4 *)(merged + (size_t)row * TCA_C + 16 * cq);
            const float4 zero = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
            for (int c4 = 0; c4 < 4; ++c4) {
                const float4 p0 = a0 ? __ldg(a0 + c4) : zero, p1 = a1 ? __ldg(a1 + c4) : zero;
                const float4 p2 = a2 ? __ldg(a2 + c4) : zero;
                float4 y;
                y.x = __fadd_rn(__fadd_rn(__fmul_rn(p0.x, w0), __fmul_rn(p1.x, w1)), __fmul_rn(p2.x, w2));
                y.y = __fadd_rn(__fadd_rn(__fmul_rn(p0.y, w0), __fmul_rn(p1.y, w1)), __fmul_rn(p2.y, w2));
                y.z = __fadd_rn(__fadd_rn(__fmul_rn(p0.z, w0), __fmul_rn(p1.z, w1)), __fmul_rn(p2.z, w2));
                y.w = __fadd_rn(__fadd_rn(__fmul_rn(p0.w, w0), __fmul_rn(p1.w, w1)), __fmul_rn(p2.w, w2));
                dst[c4] = y;
            }
        }
    } else {
        // thread = (query, 16 channels): the query voxel takes its own projected row
        const long long total = (long long)__ldg(q_base + num_wins) * 4;
        for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total;
             e += (long long)gridDim.x * blockDim.x) {
            const int cq = (int)(e & 3);
            const size_t qid = (size_t)(e >> 2);
            const int row = __ldg(q_row + __ldg(q_src + qid));
            const float4 *src = (const float4 *)(Pbuf + qid * TCA_C + 16 * cq);
            float4 *dst = (float4 *)(merged + (size_t)row * TCA_C + 16 * cq);
#pragma unroll
            for (int c4 = 0; c4 < 4; ++c4) dst[c4] = __ldg(src + c4);
        }
    }
}

static size_t tca_keys_smem_bytes(int terms) {
    const size_t nt = terms == 3 ? 2 : 1;
    size_t a_region = nt * TCA_THREADS * TCA_SD * 4;
    if (a_region < (size_t)TCA_THREADS * TCA_VPITCH * 4) a_region = (size_t)TCA_THREADS * TCA_VPITCH * 4;
    return nt * 64 * 32 * 4 + a_region + (size_t)(32 * 8 + 64 + TCA_SBUD) * 4 + TCA_TW * 32 +
           2 * (TCA_TW + 1) * 4 + 8 + 16 + 128;
}

}  // namespace mssvt

using namespace mssvt;

extern "C" {

/* Tile plan of the tensor-core window attention: a function of the geometry (meta, q_base, win_list)
 * only, made once and shared by every block / launch over the same window lists.
 * tiles (2, win_capacity, 2) int, tile_count (2) int, win_rec (2, win_capacity, 4) int, win_ctr
 * (win_capacity, 4) float: opaque to the caller. */
int mssvt_attention_tiles(int heads_per_group, int nq, int key_num_sample, int win_capacity,
                          const int *win_count_total, const int *win_list, const int *meta, const int *q_base,
                          const float *win_cell, const float *range_min, int *tiles, int *tile_count,
                          int *win_rec, float *win_ctr, void *stream) {
    if (heads_per_group <= 0 || nq <= 0 || nq > 32 || key_num_sample <= 0 || key_num_sample > 63 ||
        nq * (key_num_sample + 1) * heads_per_group > TCA_SBUD || win_capacity < 0)
        return MSSVT_ERR_INVALID;
    if (!win_count_total || !win_list || !meta || !q_base || !win_cell || !range_min || !tiles || !tile_count ||
        !win_rec || !win_ctr)
        return MSSVT_ERR_INVALID;
    cudaStream_t s = (cudaStream_t)stream;
    if (cudaMemsetAsync(tile_count, 0, 2 * sizeof(int), s) != cudaSuccess) return MSSVT_ERR_LAUNCH;
    if (win_capacity == 0) return MSSVT_OK;
    const int warps = 2 * ((win_capacity + TCA_PLAN_WB - 1) / TCA_PLAN_WB);
    ++g_launches;
    k_tca_plan<<<(warps + 7) / 8, 256, 0, s>>>(heads_per_group, win_capacity, win_count_total, (const int4 *)win_list,
                                               (const int4 *)meta, q_base,
                                               make_float3(win_cell[0], win_cell[1], win_cell[2]),
                                               make_float3(range_min[0], range_min[1], range_min[2]),
                                               (int2 *)tiles, tile_count, (int4 *)win_rec, (float4 *)win_ctr);
    return check_launch();
}

/* Tensor-core window attention of a two-window block (see the header of this file).  Weights in
 * their nn.Module layout: pos_w [64][6]; packed by mssvt_pack_operand_tf32: wkv [64][32] per head group,
 * wq_packed / wp_packed = the [64][64] block-diagonal matrix of the two groups' [32][32] weights.  rep_row / meta: compact key lists of mssvt_block_geometry; q_base:
 * mssvt_exclusive_scan of meta[:, 0] (win_capacity + 1 ints); tiles .. win_ctr: mssvt_attention_tiles.
 * scratch: 3 * num_voxels * 64 floats (q, head outputs, projected rows of every real query).  merged may be
 * NULL when interp is set: the blend of the projected rows (scratch + 2 * num_voxels * 64) is then left to
 * mssvt_ffn_tc in mode 2.
 * Returns MSSVT_ERR_INVALID for shapes outside C = 64 / 2 x 32 channels / nq <= 32 / K <= 63 /
 * cap1 <= 128 (callers then use mssvt_block_attention). */
int mssvt_block_attention_tc(int C, int heads_per_group, int nq, int key_num_sample, int cap1, int interp,
                             int terms, float scale, const float *win_cell, const float *range_min,
                             const float *pos_w, const float *pos_b, const float *wq_packed, const float *bq0,
                             const float *bq1, const float *wkv0, const float *bkv0, const float *wkv1,
                             const float *bkv1, const float *wp_packed, const float *bp0, const float *bp1,
                             int win_capacity, const int *win_count_total,
                             const int *win_list, const float *xn, const float *xyz, const int *q_row,
                             const int *rep_row, const int *meta, const int *q_base, const int *q_src,
                             const int *vox_slot, const int *win1_row, const unsigned char *nn_idx,
                             const float *nn_w, const int *tiles, const int *tile_count, const int *win_rec,
                             const float *win_ctr, int num_voxels, float *scratch, float *merged, void *stream) {
    if (C != 64 || (heads_per_group != 1 && heads_per_group != 2 && heads_per_group != 4) || nq <= 0 || nq > 32 ||
        key_num_sample <= 0 || key_num_sample > 63 || cap1 <= 0 || cap1 > 128 || win_capacity < 0 || num_voxels < 0 ||
        nq * (key_num_sample + 1) * heads_per_group > TCA_SBUD || (terms != 1 && terms != 3))
        return MSSVT_ERR_INVALID;
    if (win_capacity == 0) return MSSVT_OK;
    if (!win_cell || !range_min || !pos_w || !pos_b || !wq_packed || !bq0 || !wkv0 || !bkv0 || !wp_packed || !bp0 ||
        !bq1 || !wkv1 || !bkv1 || !bp1 || !win_count_total || !win_list || !xn || !xyz || !q_row ||
        !rep_row || !meta || !q_base || !q_src || !tiles || !tile_count || !win_rec || !win_ctr || !scratch ||
        (!merged && !interp))
        return MSSVT_ERR_INVALID;
    if (interp && (!vox_slot || !nn_idx || !nn_w)) return MSSVT_ERR_INVALID;
    (void)win1_row;
    TcAttnParams P;
    P.nq = nq; P.K = key_num_sample; P.cap1 = cap1; P.interp = interp ? 1 : 0;
    P.heads = heads_per_group; P.smax = nq * heads_per_group;
    P.scale = scale;
    for (int i = 0; i < 3; ++i) { P.win_cell[i] = win_cell[i]; P.lo[i] = range_min[i]; }
    P.pos_w = pos_w; P.pos_b = pos_b;
    P.wq = wq_packed; P.bq[0] = bq0; P.wkv[0] = wkv0; P.bkv[0] = bkv0; P.wp = wp_packed; P.bp[0] = bp0;
    P.bq[1] = bq1; P.wkv[1] = wkv1; P.bkv[1] = bkv1; P.bp[1] = bp1;
    const size_t smem = tca_keys_smem_bytes(terms);
    float *Qbuf = scratch, *Obuf = scratch + (size_t)num_voxels * 64, *Pbuf = scratch + 2 * (size_t)num_voxels * 64;
    cudaStream_t s = (cudaStream_t)stream;
    const int4 *wl = (const int4 *)win_list;
    const int wide = MSSVT_NUM_SMS * 8;  // grid-stride kernels: 8 CTAs of 256 threads per SM

    {
        TcaQueryRows rows;
        rows.nq = nq; rows.win_cap = win_capacity;
        for (int i = 0; i < 3; ++i) { rows.win_cell[i] = win_cell[i]; rows.lo[i] = range_min[i]; }
        rows.pos_w = pos_w; rows.pos_b = pos_b; rows.win_count_total = win_count_total; rows.win_list = wl;
        rows.xn = xn; rows.xyz = xyz; rows.q_row = q_row; rows.q_base = q_base; rows.q_src = q_src;
        const TclParams L = {wq_packed, bq0, bq1, scale};
        tcl_launch(L, rows, num_voxels, Qbuf, s, terms);
    }

    int per_sm = (int)(227 * 1024 / (smem + 1024));
    per_sm = per_sm > (terms == 3 ? 3 : 5) ? (terms == 3 ? 3 : 5) : per_sm < 1 ? 1 : per_sm;  // (64 TMEM columns each)
    const int grid = MSSVT_NUM_SMS * per_sm;
    ++g_launches;
#define TCA_LAUNCH(H, T)                                                                                   \
    cudaFuncSetAttribute(k_tca_keys<H, T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);         \
    launch_pdl(k_tca_keys<H, T>, dim3(grid), dim3(TCA_THREADS), smem, s, P, win_capacity, (const int2 *)tiles, \
               tile_count, (const int4 *)win_rec, (const float4 *)win_ctr, xn, xyz, rep_row, Qbuf, Obuf)
    if (terms == 3) {
        if (heads_per_group == 1) { TCA_LAUNCH(1, 3); }
        else if (heads_per_group == 2) { TCA_LAUNCH(2, 3); }
        else { TCA_LAUNCH(4, 3); }
    } else {
        if (heads_per_group == 1) { TCA_LAUNCH(1, 1); }
        else if (heads_per_group == 2) { TCA_LAUNCH(2, 1); }
        else { TCA_LAUNCH(4, 1); }
    }
#undef TCA_LAUNCH

    {
        const TclCopyRows rows = {Obuf, win_count_total, q_base, win_capacity};
        const TclParams L = {wp_packed, bp0, bp1, 1.0f};
        tcl_launch(L, rows, num_voxels, Pbuf, s, terms);
    }
    if (!merged) return check_launch();  // interpolation + merge left to mssvt_ffn_tc (mode 2)
    ++g_launches;
    launch_pdl(k_tca_merge, dim3(wide), dim3(256), 0, s, P, win_capacity, win_count_total, num_voxels, meta, q_base,
               q_src, q_row, vox_slot, nn_idx, nn_w, Pbuf, merged);
    return check_launch();
}

}  // extern "C"
