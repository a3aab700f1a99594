// attention_tc.cu -- window attention of a two-window MsSVT block, task-parallel, with the K/V
// projection on the tcgen05 tensor cores (sm_100a).  Same mathematics as k_block_attention in
// attention.cu (mssvt_backbone.py:260-336, mssvt_utils.py:88-157), different mapping: instead of one
// warp walking through a window, every stage runs with one THREAD per task over the whole frame:
//
//   k_tca_query   thread = (query, head group, 8 outputs): q = (Wq (xn + posemb) + bq) * scale
//   k_tca_keys    thread = distinct key of a window (both scales mixed, 128 per tile).  The thread
//                 gathers its 32-channel slice of the layer-normed row, adds the positional
//                 embedding and stores the row, TF32-rounded, as row t of the A operand; one thread
//                 issues 8 tcgen05.mma.kind::tf32 (M = 128, N = 64, K = 8): D0 = A Wkv0^T and
//                 D1 = A Wkv1^T into 128 TMEM columns; tcgen05.ld hands every thread the 64 K|V
//                 values of ITS key (TMEM lane = thread, columns of its scale).  Scores against the
//                 window's queries, then softmax (with the multiplicity of the masked key) and AV
//                 with one thread per (query, scale, head, quarter head).
//   k_tca_proj    thread = (query, head group, 8 outputs): output projection
//   k_tca_merge   thread = (win1 voxel, 16 channels): 1/d blend of the 3 nearest query rows -> merged
//
// Queries are addressed by a compact id (q_base[w] + slot, an exclusive scan over the windows done
// with the geometry), so the three intermediates (q, head outputs, projected rows) are dense
// (#queries, 64) fp32 arrays that live in L2.  Compared with the warp-per-window kernel this executes
// ~5x fewer warp instructions, has no lane redundancy on the small per-window matrices, keeps four
// 128-thread CTAs per SM in flight, and the 2048 FMAs per key run on the tensor pipe.
//
// Supported shape (config S0 and relatives): C = 64, two head groups of 32 channels, 1/2/4 heads per
// group, nq <= 32, key_num_sample <= 63, max_num_win1 <= 128.  Everything else runs on
// k_block_attention.  Precision: TF32 operands for K/V only; q, scores, softmax, AV, projections and
// interpolation are fp32.
#include "tc_common.cuh"

namespace mssvt {

#define TCA_THREADS 128
#define TCA_WB 16        // windows per batch (tile candidates)
#define TCA_C 64
#define TCA_SD 32
#define TCA_VPITCH 36    // V row pitch in floats: 16-byte aligned, conflict-free for quarter warps
#define TCA_WPITCH 36    // projection-weight row pitch (see stage_proj_weights)
#define TCA_WGRP (32 * TCA_WPITCH + 16)

struct TcAttnParams {
    int nq, K, cap1, interp, heads, smax;      // smax = nq * heads: score slots per key task
    float scale;
    float win_cell[3], lo[3];
    const float *pos_w, *pos_b;                // [64][6], [64]   (Conv1d weight (64, 6, 1))
    const float *wq[2], *bq[2];                // [32][32], [32]
    const float *wkv[2], *bkv[2];              // [64][32] packed (mssvt_pack_operand_tf32), [64]
    const float *wp[2], *bp[2];                // [32][32], [32]
};

// [64][8] per channel: w0..w5, bias, 0
__device__ __forceinline__ void stage_pos_weights(const TcAttnParams &P, float *sPos) {
    for (int i = threadIdx.x; i < 64 * 8; i += blockDim.x) {
        const int c = i >> 3, k = i & 7;
        sPos[i] = k < 6 ? __ldg(P.pos_w + c * 6 + k) : k == 6 ? __ldg(P.pos_b + c) : 0.f;
    }
}

// 32x32 projection weights of both groups: row pitch 36 floats, groups 1168 floats apart, so that the
// 8 (group, output-phase) rows read by the lanes of a warp fall into 8 different 16-byte bank groups
__device__ __forceinline__ void stage_proj_weights(const float *w0, const float *w1, float *sW) {
    for (int i = threadIdx.x; i < 2 * 32 * 32; i += blockDim.x) {
        const int g = i >> 10, o = (i >> 5) & 31, k = i & 31;
        sW[g * TCA_WGRP + o * TCA_WPITCH + k] = __ldg((g ? w1 : w0) + (i & 1023));
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

// 8 interleaved outputs (oq, oq+4, ...) of a 32 -> 32 projection for one input row held in registers;
// the 8 accumulators are independent FMA chains
__device__ __forceinline__ void proj8(const float *sW, const float *bias, const float *xin, int oq, float mul,
                                      float *dst) {
    float a[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) a[j] = bias[oq + 4 * j];
#pragma unroll
    for (int i4 = 0; i4 < TCA_SD / 4; ++i4) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const float4 wv = *(const float4 *)(sW + (oq + 4 * j) * TCA_WPITCH + 4 * i4);
            a[j] = fmaf(wv.x, xin[4 * i4], a[j]); a[j] = fmaf(wv.y, xin[4 * i4 + 1], a[j]);
            a[j] = fmaf(wv.z, xin[4 * i4 + 2], a[j]); a[j] = fmaf(wv.w, xin[4 * i4 + 3], a[j]);
        }
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) dst[oq + 4 * j] = a[j] * mul;
}

// ------------------------------------------------------------------------------- queries

__global__ void __launch_bounds__(256)
k_tca_query(TcAttnParams P, int win_cap, const int *__restrict__ win_count_total,
            const int4 *__restrict__ win_list, const float *__restrict__ xn, const float *__restrict__ xyz,
            const int *__restrict__ q_row, const int *__restrict__ q_base, const int *__restrict__ q_src,
            float *__restrict__ Qbuf) {
    __shared__ __align__(16) float sPos[64 * 8];
    __shared__ __align__(16) float sW[2 * TCA_WGRP];
    __shared__ float sB[64];
    stage_pos_weights(P, sPos);
    stage_proj_weights(P.wq[0], P.wq[1], sW);
    for (int i = threadIdx.x; i < 64; i += blockDim.x) sB[i] = __ldg(P.bq[i >> 5] + (i & 31));
    __syncthreads();
    const int num_wins = min(win_cap, __ldg(win_count_total));
    const long long total = (long long)__ldg(q_base + num_wins) * 8;  // #real queries of the frame
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total;
         e += (long long)gridDim.x * blockDim.x) {
        const int part = (int)(e & 7), g = part >> 2, oq = part & 3;
        const size_t qid = (size_t)(e >> 3);
        const int src = __ldg(q_src + qid), w = src / P.nq;
        const int row = __ldg(q_row + src);
        const int4 win = __ldg(win_list + w);
        const float cx = world_coord(win.w, P.win_cell[0], P.lo[0]);
        const float cy = world_coord(win.z, P.win_cell[1], P.lo[1]);
        const float cz = world_coord(win.y, P.win_cell[2], P.lo[2]);
        const float rx = __fsub_rn(__ldg(xyz + 3 * (size_t)row), cx);
        const float ry = __fsub_rn(__ldg(xyz + 3 * (size_t)row + 1), cy);
        const float rz = __fsub_rn(__ldg(xyz + 3 * (size_t)row + 2), cz);
        float xin[TCA_SD];
#pragma unroll
        for (int c4 = 0; c4 < TCA_SD / 4; ++c4) {
            const float4 v = __ldg((const float4 *)(xn + (size_t)row * TCA_C + g * TCA_SD) + c4);
            xin[4 * c4] = v.x + pos_embed8(sPos, g * TCA_SD + 4 * c4, rx, ry, rz, cx, cy, cz);
            xin[4 * c4 + 1] = v.y + pos_embed8(sPos, g * TCA_SD + 4 * c4 + 1, rx, ry, rz, cx, cy, cz);
            xin[4 * c4 + 2] = v.z + pos_embed8(sPos, g * TCA_SD + 4 * c4 + 2, rx, ry, rz, cx, cy, cz);
            xin[4 * c4 + 3] = v.w + pos_embed8(sPos, g * TCA_SD + 4 * c4 + 3, rx, ry, rz, cx, cy, cz);
        }
        proj8(sW + g * TCA_WGRP, sB + g * TCA_SD, xin, oq, P.scale, Qbuf + qid * TCA_C + g * TCA_SD);
    }
}

// ------------------------------------------------------------------------------- keys + attention

struct TcaTile {
    int ws, we, nQ, nT0, nT1;
};

template <int HEADS>
__global__ void __launch_bounds__(TCA_THREADS, 4)
k_tca_keys(TcAttnParams P, int win_cap, const int *__restrict__ win_count_total,
           const int4 *__restrict__ win_list, const float *__restrict__ xn, const float *__restrict__ xyz,
           const int *__restrict__ rep_row, const int *__restrict__ meta, const int *__restrict__ q_base,
           const float *__restrict__ Qbuf, float *__restrict__ Obuf) {
    constexpr int HD = TCA_SD / HEADS;
    constexpr int DPT = HD / 4;  // channels per thread in the AV phase
    extern __shared__ __align__(128) char smem_raw[];
    const int tid = threadIdx.x, warp = tid >> 5;
    const int K = P.K, smax = P.smax;

    // ---- shared memory: 16 KB weights + 18 KB A/V (aliased) + scores + bookkeeping = ~51 KB
    char *sWkv = smem_raw;                                  // 2 x [64 x 32] canonical, TF32   16 KB
    char *sA = sWkv + 2 * 64 * 32 * 4;                      // [128 x 32] canonical (16 KB) ...
    float *sV = (float *)sA;                                // ... reused as V [128][VPITCH] after the MMA
    float *sPos = sV + TCA_THREADS * TCA_VPITCH;            // [64][8]
    float *sBkv = sPos + 64 * 8;                            // [2][64]
    float *sS = sBkv + 128;                                 // [128][smax] scores of each key task
    float *sCtr = sS + TCA_THREADS * smax;                  // [WB][4] window centres
    int *sMeta = (int *)(sCtr + TCA_WB * 4);                // [WB][4] {nqr, q_base, rep0, rep1}
    int *sQoff = sMeta + TCA_WB * 4;                        // [WB + 1] prefix of real queries in the tile
    int *sToff = sQoff + TCA_WB + 1;                        // [2][WB + 1] prefix of key tasks per scale
    int *sQwin = sToff + 2 * (TCA_WB + 1);                  // [128] local window of each query of the tile
    int *sTwin = sQwin + TCA_THREADS;                       // [128] local window of each key task
    int *sTmult = sTwin + TCA_THREADS;                      // [128] multiplicity of each key task
    int *sTile = sTmult + TCA_THREADS;                      // TcaTile + pad
    uint64_t *sBar = (uint64_t *)(sTile + 8 + ((TCA_WB * 4 + 3 * (TCA_WB + 1) + 3 * TCA_THREADS + 8) & 1));
    uint32_t *sTmem = (uint32_t *)(sBar + 1);

    stage_packed(P.wkv[0], 64 * 32, sWkv);
    stage_packed(P.wkv[1], 64 * 32, sWkv + 64 * 32 * 4);
    stage_pos_weights(P, sPos);
    for (int i = tid; i < 128; i += TCA_THREADS) sBkv[i] = __ldg(P.bkv[i >> 6] + (i & 63));
    const uint32_t bar = smem_u32(sBar);
    if (tid == 0) {
        mbar_init(bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) tmem_alloc(smem_u32(sTmem), 128);
    fence_async_smem();
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

    const int num_wins = min(win_cap, __ldg(win_count_total));
    const int batches = (num_wins + TCA_WB - 1) / TCA_WB;
    TcaTile *tile = (TcaTile *)sTile;

    for (int batch = blockIdx.x; batch < batches; batch += gridDim.x) {
        const int wb0 = batch * TCA_WB, nb = min(TCA_WB, num_wins - wb0);
        __syncthreads();  // previous batch fully consumed
        if (tid < nb) {
            const int4 m = __ldg((const int4 *)meta + wb0 + tid);
            sMeta[4 * tid] = m.x; sMeta[4 * tid + 1] = __ldg(q_base + wb0 + tid);
            sMeta[4 * tid + 2] = m.z; sMeta[4 * tid + 3] = m.w;
            const int4 win = __ldg(win_list + wb0 + tid);
            sCtr[4 * tid] = world_coord(win.w, P.win_cell[0], P.lo[0]);
            sCtr[4 * tid + 1] = world_coord(win.z, P.win_cell[1], P.lo[1]);
            sCtr[4 * tid + 2] = world_coord(win.y, P.win_cell[2], P.lo[2]);
        }
        __syncthreads();
        int ws = 0;
        while (ws < nb) {
            // ---- tile = greedy prefix of the batch: <= 128 distinct keys (both scales), <= 128 queries
            if (tid == 0) {
                int we = ws, aq = 0, a0 = 0, a1 = 0;
                while (we < nb) {
                    const int nqr = sMeta[4 * we];
                    const int r0 = sMeta[4 * we + 2] & 0xff, r1 = sMeta[4 * we + 3] & 0xff;
                    // scale-1 tasks start at a warp boundary: tcgen05.ld takes ONE column address per warp
                    if (we > ws && (aq + nqr > TCA_THREADS || ((a0 + r0 + 31) & ~31) + a1 + r1 > TCA_THREADS)) break;
                    const int l = we - ws;
                    sQoff[l] = aq; sToff[l] = a0; sToff[TCA_WB + 1 + l] = a1;
                    aq += nqr; a0 += r0; a1 += r1;
                    ++we;
                }
                const int l = we - ws;
                sQoff[l] = aq; sToff[l] = a0; sToff[TCA_WB + 1 + l] = a1;
                tile->ws = ws; tile->we = we; tile->nQ = aq; tile->nT0 = a0; tile->nT1 = a1;
            }
            __syncthreads();
            const int t_ws = tile->ws, t_we = tile->we, nQ = tile->nQ, nT0 = tile->nT0, nT1 = tile->nT1;
            const int R1 = (nT0 + 31) & ~31;  // first row of the scale-1 tasks (warp aligned)
            const int nwin = t_we - t_ws;
            // task -> window maps; key tasks: scale 0 of every window first, then scale 1
            if (tid < nwin) {
                for (int i = sQoff[tid]; i < sQoff[tid + 1]; ++i) sQwin[i] = tid;
                for (int i = sToff[tid]; i < sToff[tid + 1]; ++i) sTwin[i] = tid;
                for (int i = sToff[TCA_WB + 1 + tid]; i < sToff[TCA_WB + 1 + tid + 1]; ++i) sTwin[R1 + i] = tid;
            }
            __syncthreads();

            // ---- key task -> row t of the A operand
            int l = 0, my_nqr = 0, my_q0 = 0;
            const int g = tid >= R1 ? 1 : 0;  // uniform within a warp
            const bool is_task = g ? tid - R1 < nT1 : tid < nT0;
            bool masked = false;
            if (is_task) {
                l = sTwin[tid];
                const int j = g ? tid - R1 - sToff[TCA_WB + 1 + l] : tid - sToff[l];
                const int w = wb0 + t_ws + l;
                const int m = sMeta[4 * (t_ws + l) + 2 + g];
                masked = (m >> 8) > 0 && j == (m & 0xff) - 1;  // last distinct key stands for all masked slots
                sTmult[tid] = masked ? (m >> 8) : 1;
                my_nqr = sMeta[4 * (t_ws + l)];
                my_q0 = sMeta[4 * (t_ws + l) + 1];
                const int row = __ldg(rep_row + (size_t)w * 2 * K + g * K + j);
                const float cx = sCtr[4 * (t_ws + l)], cy = sCtr[4 * (t_ws + l) + 1], cz = sCtr[4 * (t_ws + l) + 2];
                float rx = 0.f, ry = 0.f, rz = 0.f;  // masked key: relative offset zeroed
                if (!masked) {
                    rx = __fsub_rn(__ldg(xyz + 3 * (size_t)row), cx);
                    ry = __fsub_rn(__ldg(xyz + 3 * (size_t)row + 1), cy);
                    rz = __fsub_rn(__ldg(xyz + 3 * (size_t)row + 2), cz);
                }
                const float4 *src = (const float4 *)(xn + (size_t)row * TCA_C + g * TCA_SD);
                float4 xv[TCA_SD / 4];  // the whole 128-byte slice in flight at once
#pragma unroll
                for (int c4 = 0; c4 < TCA_SD / 4; ++c4) xv[c4] = __ldg(src + c4);
#pragma unroll
                for (int c4 = 0; c4 < TCA_SD / 4; ++c4) {
                    const float4 v = xv[c4];
                    const int c = g * TCA_SD + 4 * c4;
                    float4 o;
                    o.x = to_tf32(v.x + pos_embed8(sPos, c, rx, ry, rz, cx, cy, cz));
                    o.y = to_tf32(v.y + pos_embed8(sPos, c + 1, rx, ry, rz, cx, cy, cz));
                    o.z = to_tf32(v.z + pos_embed8(sPos, c + 2, rx, ry, rz, cx, cy, cz));
                    o.w = to_tf32(v.w + pos_embed8(sPos, c + 3, rx, ry, rz, cx, cy, cz));
                    *(float4 *)(sA + (uint32_t)c4 * a_lbo + my_row_off) = o;
                }
            }
            stage_packed_wait();
            fence_async_smem();
            __syncthreads();
            // ---- D0 = A Wkv0^T (columns 0..63), D1 = A Wkv1^T (columns 64..127) on the tensor cores
            if (tid == 0) {
                tc_fence_after();
#pragma unroll
                for (int gg = 0; gg < 2; ++gg)
#pragma unroll
                    for (int k = 0; k < TCA_SD / 8; ++k) {
                        const uint64_t da = umma_smem_desc(sA_u + (uint32_t)k * 2u * a_lbo, a_lbo, 128);
                        const uint64_t db = umma_smem_desc(sWkv_u + (uint32_t)gg * 64u * 32u * 4u + (uint32_t)k * 2u * w_lbo,
                                                           w_lbo, 128);
                        umma_tf32(tmem_d + (uint32_t)gg * 64u, da, db, idesc, k > 0 ? 1u : 0u);
                    }
                umma_commit(bar);
            }
            mbar_wait(bar, phase);
            phase ^= 1u;
            tc_fence_after();
            // ---- this thread's key: K|V of its scale back from TMEM; scores against its window's queries
            {
                float kk[TCA_SD], vv[TCA_SD];
                tmem_ld32(tmem_d + lane_off + (uint32_t)g * 64u, kk);
                tmem_ld32(tmem_d + lane_off + (uint32_t)g * 64u + 32u, vv);
                tc_fence_before();
                __syncthreads();  // every thread has its K|V in registers: the A tile may become V
                if (is_task) {
                    const float *bk = sBkv + g * 64;
#pragma unroll
                    for (int i = 0; i < TCA_SD; ++i) { kk[i] += bk[i]; vv[i] += bk[TCA_SD + i]; }
#pragma unroll
                    for (int c4 = 0; c4 < TCA_SD / 4; ++c4)
                        *(float4 *)(sV + tid * TCA_VPITCH + 4 * c4) =
                            make_float4(vv[4 * c4], vv[4 * c4 + 1], vv[4 * c4 + 2], vv[4 * c4 + 3]);
                    const float bias = masked ? -100.0f : 0.f;  // additive mask of the reference
                    for (int s = 0; s < my_nqr; ++s) {
                        const float4 *qv = (const float4 *)(Qbuf + (size_t)(my_q0 + s) * TCA_C + g * TCA_SD);
#pragma unroll
                        for (int h = 0; h < HEADS; ++h) {
                            float a = 0.f;
#pragma unroll
                            for (int d4 = 0; d4 < HD / 4; ++d4) {
                                const float4 q4 = __ldg(qv + h * (HD / 4) + d4);
                                a = fmaf(q4.x, kk[h * HD + 4 * d4], a); a = fmaf(q4.y, kk[h * HD + 4 * d4 + 1], a);
                                a = fmaf(q4.z, kk[h * HD + 4 * d4 + 2], a); a = fmaf(q4.w, kk[h * HD + 4 * d4 + 3], a);
                            }
                            sS[tid * smax + s * HEADS + h] = a + bias;
                        }
                    }
                }
            }
            __syncthreads();
            // ---- softmax over the window's distinct keys of a scale and AV,
            //      thread = (query, scale, head, quarter of the head)
            for (int e = tid; e < nQ * 2 * HEADS * 4; e += TCA_THREADS) {
                const int dq = e & 3, qgh = e >> 2;
                const int h = qgh % HEADS, qg = qgh / HEADS, gg = qg & 1, qt = qg >> 1;
                const int lq = sQwin[qt], s = qt - sQoff[lq];
                const int t0 = gg ? R1 + sToff[TCA_WB + 1 + lq] : sToff[lq];
                const int t1 = gg ? R1 + sToff[TCA_WB + 1 + lq + 1] : sToff[lq + 1];
                float mx = -3.0e38f;
                for (int t = t0; t < t1; ++t) mx = fmaxf(mx, sS[t * smax + s * HEADS + h]);
                float den = 0.f, acc[DPT];
#pragma unroll
                for (int d = 0; d < DPT; ++d) acc[d] = 0.f;
                for (int t = t0; t < t1; ++t) {
                    const float wgt = exp_neg(sS[t * smax + s * HEADS + h] - mx) * (float)sTmult[t];
                    den += wgt;
                    const float *vp = sV + t * TCA_VPITCH + h * HD + dq * DPT;
#pragma unroll
                    for (int d = 0; d < DPT; ++d) acc[d] = fmaf(wgt, vp[d], acc[d]);
                }
                const float inv = 1.0f / den;
                float *dst = Obuf + (size_t)(sMeta[4 * (t_ws + lq) + 1] + s) * TCA_C + gg * TCA_SD + h * HD + dq * DPT;
#pragma unroll
                for (int d = 0; d < DPT; ++d) dst[d] = acc[d] * inv;
            }
            __syncthreads();
            ws = t_we;
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem_d, 128);
}

// ------------------------------------------------------------------------------- output projection

__global__ void __launch_bounds__(256)
k_tca_proj(TcAttnParams P, int win_cap, const int *__restrict__ win_count_total,
           const int *__restrict__ q_base, const float *__restrict__ Obuf, float *__restrict__ Pbuf) {
    __shared__ __align__(16) float sW[2 * TCA_WGRP];
    __shared__ float sB[64];
    stage_proj_weights(P.wp[0], P.wp[1], sW);
    for (int i = threadIdx.x; i < 64; i += blockDim.x) sB[i] = __ldg(P.bp[i >> 5] + (i & 31));
    __syncthreads();
    const int num_wins = min(win_cap, __ldg(win_count_total));
    const long long total = (long long)__ldg(q_base + num_wins) * 8;  // #real queries of the frame
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total;
         e += (long long)gridDim.x * blockDim.x) {
        const int part = (int)(e & 7), g = part >> 2, oq = part & 3;
        const size_t qid = (size_t)(e >> 3);
        float xin[TCA_SD];
#pragma unroll
        for (int i4 = 0; i4 < TCA_SD / 4; ++i4) {
            const float4 v = __ldg((const float4 *)(Obuf + qid * TCA_C + g * TCA_SD) + i4);
            xin[4 * i4] = v.x; xin[4 * i4 + 1] = v.y; xin[4 * i4 + 2] = v.z; xin[4 * i4 + 3] = v.w;
        }
        proj8(sW + g * TCA_WGRP, sB + g * TCA_SD, xin, oq, 1.0f, Pbuf + qid * TCA_C + g * TCA_SD);
    }
}

// ------------------------------------------------------------------------------- merge

__global__ void __launch_bounds__(256)
k_tca_merge(TcAttnParams P, int win_cap, const int *__restrict__ win_count_total, int num_voxels,
            const int *__restrict__ meta, const int *__restrict__ q_base, const int *__restrict__ q_src,
            const int *__restrict__ q_row, const int *__restrict__ vox_slot,
            const unsigned char *__restrict__ nn_idx, const float *__restrict__ nn_w,
            const float *__restrict__ Pbuf, float *__restrict__ merged) {
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

static size_t tca_keys_smem_bytes(int smax) {
    size_t floats = TCA_THREADS * TCA_VPITCH + 64 * 8 + 128 + (size_t)TCA_THREADS * smax + TCA_WB * 4;
    size_t ints = TCA_WB * 4 + 3 * (TCA_WB + 1) + 3 * TCA_THREADS + 8;
    ints += ints & 1;  // keep the mbarrier 8-byte aligned
    return 2 * 64 * 32 * 4 + (floats + ints) * 4 + 8 + 16 + 128;
}

}  // namespace mssvt

using namespace mssvt;

extern "C" {

/* Tensor-core window attention of a two-window block (see the header of this file).  Weights in
 * their nn.Module layout: pos_w [64][6], wq/wp [32][32] per head group; wkv [64][32] packed by
 * mssvt_pack_operand_tf32.  rep_row / meta:
 * compact key lists of mssvt_block_geometry; q_base: mssvt_exclusive_scan of meta[:, 0] (win_capacity + 1
 * ints).  scratch: 3 * num_voxels * 64 floats (q, head outputs, projected rows of every real query).
 * Returns MSSVT_ERR_INVALID for shapes outside C = 64 / 2 x 32 channels / nq <= 32 / K <= 63 /
 * cap1 <= 128 (callers then use mssvt_block_attention). */
int mssvt_block_attention_tc(int C, int heads_per_group, int nq, int key_num_sample, int cap1, int interp,
                             float scale, const float *win_cell, const float *range_min,
                             const float *pos_w, const float *pos_b, const float *wq0, const float *bq0,
                             const float *wkv0, const float *bkv0, const float *wp0, const float *bp0,
                             const float *wq1, const float *bq1, const float *wkv1, const float *bkv1,
                             const float *wp1, const float *bp1, int win_capacity, const int *win_count_total,
                             const int *win_list, const float *xn, const float *xyz, const int *q_row,
                             const int *rep_row, const int *meta, const int *q_base, const int *q_src,
                             const int *vox_slot, const int *win1_row, const unsigned char *nn_idx,
                             const float *nn_w, int num_voxels, float *scratch, float *merged, void *stream) {
    if (C != 64 || (heads_per_group != 1 && heads_per_group != 2 && heads_per_group != 4) || nq <= 0 || nq > 32 ||
        key_num_sample <= 0 || key_num_sample > 63 || cap1 <= 0 || cap1 > 128 || win_capacity < 0 || num_voxels < 0)
        return MSSVT_ERR_INVALID;
    if (win_capacity == 0) return MSSVT_OK;
    if (!win_cell || !range_min || !pos_w || !pos_b || !wq0 || !bq0 || !wkv0 || !bkv0 || !wp0 || !bp0 || !wq1 ||
        !bq1 || !wkv1 || !bkv1 || !wp1 || !bp1 || !win_count_total || !win_list || !xn || !xyz || !q_row ||
        !rep_row || !meta || !q_base || !q_src || !scratch || !merged)
        return MSSVT_ERR_INVALID;
    if (interp && (!vox_slot || !nn_idx || !nn_w)) return MSSVT_ERR_INVALID;
    (void)win1_row;
    TcAttnParams P;
    P.nq = nq; P.K = key_num_sample; P.cap1 = cap1; P.interp = interp ? 1 : 0;
    P.heads = heads_per_group; P.smax = nq * heads_per_group;
    P.scale = scale;
    for (int i = 0; i < 3; ++i) { P.win_cell[i] = win_cell[i]; P.lo[i] = range_min[i]; }
    P.pos_w = pos_w; P.pos_b = pos_b;
    P.wq[0] = wq0; P.bq[0] = bq0; P.wkv[0] = wkv0; P.bkv[0] = bkv0; P.wp[0] = wp0; P.bp[0] = bp0;
    P.wq[1] = wq1; P.bq[1] = bq1; P.wkv[1] = wkv1; P.bkv[1] = bkv1; P.wp[1] = wp1; P.bp[1] = bp1;
    const size_t smem = tca_keys_smem_bytes(P.smax);
    if (smem > 227 * 1024) return MSSVT_ERR_INVALID;
    float *Qbuf = scratch, *Obuf = scratch + (size_t)num_voxels * 64, *Pbuf = scratch + 2 * (size_t)num_voxels * 64;
    cudaStream_t s = (cudaStream_t)stream;
    const int4 *wl = (const int4 *)win_list;
    const int wide = MSSVT_NUM_SMS * 8;  // grid-stride kernels: 8 CTAs of 256 threads per SM

    ++g_launches;
    k_tca_query<<<wide, 256, 0, s>>>(P, win_capacity, win_count_total, wl, xn, xyz, q_row, q_base, q_src, Qbuf);

    int per_sm = (int)(227 * 1024 / (smem + 1024));
    per_sm = per_sm > 4 ? 4 : per_sm < 1 ? 1 : per_sm;  // 4 x 128 TMEM columns = all 512
    const int batches = (win_capacity + TCA_WB - 1) / TCA_WB;
    int grid = MSSVT_NUM_SMS * per_sm;
    if (grid > batches) grid = batches;
    ++g_launches;
#define TCA_LAUNCH(H)                                                                                      \
    cudaFuncSetAttribute(k_tca_keys<H>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);            \
    k_tca_keys<H><<<grid, TCA_THREADS, smem, s>>>(P, win_capacity, win_count_total, wl, xn, xyz, rep_row,  \
                                                  meta, q_base, Qbuf, Obuf)
    if (heads_per_group == 1) { TCA_LAUNCH(1); }
    else if (heads_per_group == 2) { TCA_LAUNCH(2); }
    else { TCA_LAUNCH(4); }
#undef TCA_LAUNCH

    ++g_launches;
    k_tca_proj<<<wide, 256, 0, s>>>(P, win_capacity, win_count_total, q_base, Obuf, Pbuf);
    ++g_launches;
    k_tca_merge<<<wide, 256, 0, s>>>(P, win_capacity, win_count_total, num_voxels, meta, q_base, q_src, q_row,
                                     vox_slot, nn_idx, nn_w, Pbuf, merged);
    return check_launch();
}

}  // extern "C"
