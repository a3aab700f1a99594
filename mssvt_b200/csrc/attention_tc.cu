// attention_tc.cu -- window attention of a two-window MsSVT block, task-parallel, with the K/V
// projection on the tcgen05 tensor cores (sm_100a).  Same mathematics as k_block_attention in
// attention.cu (mssvt_backbone.py:260-336, mssvt_utils.py:88-157), different mapping:
//
//   * a CTA (128 threads) takes a batch of 32 consecutive windows and cuts it greedily into tiles
//     whose work fits one pass of 128 threads per phase: <= 32 real queries, <= 128 distinct keys
//     per scale, <= 128 win1 voxels;
//   * thread = task.  A key task gathers its 32-channel slice of the layer-normed row, adds the
//     positional embedding and stores the row, TF32-rounded, as row t of the A operand;
//     one thread issues 4 tcgen05.mma.kind::tf32 (M = 128, N = 64, K = 8 each) against W_kv of
//     the scale, which sits in shared memory in the canonical UMMA layout for the whole kernel;
//     tcgen05.ld hands every thread the 64 K|V values of ITS key (TMEM lane = thread);
//   * scores, softmax (with the multiplicity of the masked key), AV, the q / output projections
//     and the three-NN blend are thread-per-(query, head) / thread-per-voxel fp32 FFMA.
// Compared with the warp-per-window kernel this executes ~10x fewer warp instructions per window:
// no lane redundancy on small matrices, and the 2048 FMAs per key move to the tensor pipe.
//
// Supported shape (config S0 and relatives): C = 64, two head groups of 32 channels, 1-4 heads
// per group, nq <= 32, key_num_sample <= 127, max_num_win1 <= 128.  Everything else runs on
// k_block_attention.  Precision: TF32 operands for K/V only; q, scores, softmax, AV, projections,
// interpolation in fp32.
#include "tc_common.cuh"

namespace mssvt {

#define TCA_THREADS 128
#define TCA_WB 32        // windows per batch
#define TCA_QCAP 32      // real queries per tile
#define TCA_C 64
#define TCA_SD 32
#define TCA_VPITCH 36    // V row pitch in floats: 16-byte aligned, conflict-free for quarter warps
#define TCA_WPITCH 36    // projection-weight row pitch
#define TCA_WGRP (32 * TCA_WPITCH + 16)
#define TCA_WSZ (2 * TCA_WGRP)

struct TcAttnParams {
    int nq, K, cap1, interp, heads, hd, smax;  // smax = nq * heads: score slots per key task
    float scale;
    float win_cell[3], lo[3];
    const float *pos_w, *pos_b;                // [64][6], [64]   (Conv1d weight (64, 6, 1))
    const float *wq[2], *bq[2];                // [32][32], [32]
    const float *wkv[2], *bkv[2];              // [64][32], [64]
    const float *wp[2], *bp[2];                // [32][32], [32]
};

struct TcaTile {
    int ws, we, nQ, nT[2], nV;
};

template <int HEADS>
__global__ void __launch_bounds__(TCA_THREADS, 2)
k_block_attention_tc(TcAttnParams P, int win_cap, const int *__restrict__ win_count_total,
                     const int4 *__restrict__ win_list, const float *__restrict__ xn,
                     const float *__restrict__ xyz, const int *__restrict__ q_row,
                     const int *__restrict__ rep_row, const int *__restrict__ meta,
                     const int *__restrict__ win1_row, const unsigned char *__restrict__ nn_idx,
                     const float *__restrict__ nn_w, float *__restrict__ merged) {
    constexpr int HD = TCA_SD / HEADS;
    extern __shared__ __align__(128) char smem_raw[];
    const int tid = threadIdx.x, warp = tid >> 5;
    const int nq = P.nq, K = P.K, cap1 = P.cap1, smax = P.smax;

    // ---- shared memory carve-up
    char *sWkv = smem_raw;                                  // 2 x [64 x 32] canonical, TF32   16 KB
    char *sA = sWkv + 2 * 64 * 32 * 4;                      // [128 x 32] canonical            16 KB
    float *sPos = (float *)(sA + TCA_THREADS * 32 * 4);     // [64][8]: w0..w5, bias, 0         2 KB
    // q / output projection weights: row pitch 36 floats, head groups 1168 floats apart, so that the
    // 8 (group, output-phase) rows read by the lanes of a warp fall into 8 different 16-byte bank groups
    float *sWq = sPos + 64 * 8;                             // [2][32] rows                    9.2 KB
    float *sWp = sWq + TCA_WSZ;                             // [2][32] rows                    9.2 KB
    float *sBq = sWp + TCA_WSZ;                             // [64]
    float *sBkv = sBq + 64;                                 // [2][64]
    float *sBp = sBkv + 128;                                // [64]
    float *sQ = sBp + 64;                                   // [QCAP][64] scaled q
    float *sO = sQ + TCA_QCAP * 64;                         // [QCAP][64] head outputs -> attention rows
    float *sV = sO + TCA_QCAP * 64;                         // [128][VPITCH]
    float *sS = sV + TCA_THREADS * TCA_VPITCH;              // [128][smax] scores of each key task
    float *sCtr = sS + TCA_THREADS * smax;                  // [WB][4] window centres
    int *sMeta = (int *)(sCtr + TCA_WB * 4);                // [WB][4]
    int *sQoff = sMeta + TCA_WB * 4;                        // [WB + 1] prefix of real queries in the tile
    int *sToff = sQoff + TCA_WB + 1;                        // [2][WB + 1] prefix of key tasks per scale
    int *sVoff = sToff + 2 * (TCA_WB + 1);                  // [WB + 1] prefix of win1 voxels
    int *sQwin = sVoff + TCA_WB + 1;                        // [QCAP] local window of each query task
    int *sTwin = sQwin + TCA_QCAP;                          // [2][128] local window of each key task
    int *sVwin = sTwin + 2 * TCA_THREADS;                   // [128] local window of each voxel task
    int *sTmult = sVwin + TCA_THREADS;                      // [128] multiplicity of each key task
    int *sTile = sTmult + TCA_THREADS;                      // TcaTile (6 ints) + pad to 8
    uint64_t *sBar = (uint64_t *)(sTile + 8);
    uint32_t *sTmem = (uint32_t *)(sBar + 1);

    // ---- one-time setup: weights, barrier, TMEM
    stage_operand(P.wkv[0], 64, 32, sWkv);
    stage_operand(P.wkv[1], 64, 32, sWkv + 64 * 32 * 4);
    for (int i = tid; i < 64 * 8; i += TCA_THREADS) {
        const int c = i >> 3, k = i & 7;
        sPos[i] = k < 6 ? __ldg(P.pos_w + c * 6 + k) : k == 6 ? __ldg(P.pos_b + c) : 0.f;
    }
    for (int i = tid; i < 2 * 32 * 32; i += TCA_THREADS) {
        const int g = i >> 10, o = (i >> 5) & 31, k = i & 31;
        sWq[g * TCA_WGRP + o * TCA_WPITCH + k] = __ldg(P.wq[g] + (i & 1023));
        sWp[g * TCA_WGRP + o * TCA_WPITCH + k] = __ldg(P.wp[g] + (i & 1023));
    }
    for (int i = tid; i < 64; i += TCA_THREADS) {
        sBq[i] = __ldg(P.bq[i >> 5] + (i & 31));
        sBp[i] = __ldg(P.bp[i >> 5] + (i & 31));
    }
    for (int i = tid; i < 128; i += TCA_THREADS) sBkv[i] = __ldg(P.bkv[i >> 6] + (i & 63));
    const uint32_t bar = smem_u32(sBar);
    if (tid == 0) {
        mbar_init(bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) tmem_alloc(smem_u32(sTmem), 64);
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
            sMeta[4 * tid] = m.x; sMeta[4 * tid + 1] = P.interp ? m.y : 0;
            sMeta[4 * tid + 2] = m.z; sMeta[4 * tid + 3] = m.w;
            const int4 win = __ldg(win_list + wb0 + tid);
            sCtr[4 * tid] = world_coord(win.w, P.win_cell[0], P.lo[0]);
            sCtr[4 * tid + 1] = world_coord(win.z, P.win_cell[1], P.lo[1]);
            sCtr[4 * tid + 2] = world_coord(win.y, P.win_cell[2], P.lo[2]);
        }
        __syncthreads();
        int ws = 0;
        while (ws < nb) {
            // ---- tile formation: greedy prefix of the batch that fits one pass per phase
            if (tid == 0) {
                int we = ws, aq = 0, a0 = 0, a1 = 0, av = 0;
                while (we < nb) {
                    const int nqr = sMeta[4 * we], cnt1 = sMeta[4 * we + 1];
                    const int r0 = sMeta[4 * we + 2] & 0xff, r1 = sMeta[4 * we + 3] & 0xff;
                    if (we > ws && (aq + nqr > TCA_QCAP || a0 + r0 > TCA_THREADS || a1 + r1 > TCA_THREADS ||
                                    av + cnt1 > TCA_THREADS))
                        break;
                    const int l = we - ws;
                    sQoff[l] = aq; sToff[l] = a0; sToff[TCA_WB + 1 + l] = a1; sVoff[l] = av;
                    aq += nqr; a0 += r0; a1 += r1; av += cnt1;
                    ++we;
                }
                const int l = we - ws;
                sQoff[l] = aq; sToff[l] = a0; sToff[TCA_WB + 1 + l] = a1; sVoff[l] = av;
                tile->ws = ws; tile->we = we; tile->nQ = aq; tile->nT[0] = a0; tile->nT[1] = a1; tile->nV = av;
            }
            __syncthreads();
            const int t_ws = tile->ws, t_we = tile->we, nQ = tile->nQ, nV = tile->nV;
            const int nwin = t_we - t_ws;
            // task -> window maps, one thread per window of the tile
            if (tid < nwin) {
                for (int i = sQoff[tid]; i < sQoff[tid + 1]; ++i) sQwin[i] = tid;
                for (int i = sVoff[tid]; i < sVoff[tid + 1]; ++i) sVwin[i] = tid;
                for (int g = 0; g < 2; ++g)
                    for (int i = sToff[g * (TCA_WB + 1) + tid]; i < sToff[g * (TCA_WB + 1) + tid + 1]; ++i)
                        sTwin[g * TCA_THREADS + i] = tid;
            }
            __syncthreads();

            // ---- phase 1a: query inputs xn + posemb, thread = (query, 4-channel chunk) -> sO (as scratch)
            for (int e = tid; e < nQ * 16; e += TCA_THREADS) {
                const int qt = e >> 4, c4 = e & 15;
                const int l = sQwin[qt], s = qt - sQoff[l];
                const int row = __ldg(q_row + (size_t)(wb0 + t_ws + l) * nq + s);
                const float cx = sCtr[4 * (t_ws + l)], cy = sCtr[4 * (t_ws + l) + 1], cz = sCtr[4 * (t_ws + l) + 2];
                const float rx = __fsub_rn(__ldg(xyz + 3 * (size_t)row), cx);
                const float ry = __fsub_rn(__ldg(xyz + 3 * (size_t)row + 1), cy);
                const float rz = __fsub_rn(__ldg(xyz + 3 * (size_t)row + 2), cz);
                const float4 v = __ldg((const float4 *)(xn + (size_t)row * TCA_C) + c4);
                const float f[4] = {v.x, v.y, v.z, v.w};
                float o[4];
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const float4 wa = *(const float4 *)(sPos + (4 * c4 + k) * 8);
                    const float4 wb = *(const float4 *)(sPos + (4 * c4 + k) * 8 + 4);
                    float a = wb.z;
                    a = fmaf(wa.x, rx, a); a = fmaf(wa.y, ry, a); a = fmaf(wa.z, rz, a);
                    a = fmaf(wa.w, cx, a); a = fmaf(wb.x, cy, a); a = fmaf(wb.y, cz, a);
                    o[k] = f[k] + fmaxf(a, 0.f);
                }
                *(float4 *)(sO + qt * TCA_C + 4 * c4) = make_float4(o[0], o[1], o[2], o[3]);
            }
            __syncthreads();
            // ---- phase 1b: q = (Wq x + bq) * scale, thread = (query, head group, 8 outputs)
            for (int e = tid; e < nQ * 8; e += TCA_THREADS) {
                const int qt = e >> 3, g = (e >> 2) & 1, oq = e & 3;
                float xin[TCA_SD];
#pragma unroll
                for (int i4 = 0; i4 < TCA_SD / 4; ++i4) {
                    const float4 v = *(const float4 *)(sO + qt * TCA_C + g * TCA_SD + 4 * i4);
                    xin[4 * i4] = v.x; xin[4 * i4 + 1] = v.y; xin[4 * i4 + 2] = v.z; xin[4 * i4 + 3] = v.w;
                }
                const float *wq = sWq + g * TCA_WGRP;
#pragma unroll 2
                for (int j = 0; j < 8; ++j) {
                    const int o = oq + 4 * j;  // outputs interleaved over the 4 threads of a (query, group)
                    float a = sBq[g * TCA_SD + o];
#pragma unroll
                    for (int i4 = 0; i4 < TCA_SD / 4; ++i4) {
                        const float4 wv = *(const float4 *)(wq + o * TCA_WPITCH + 4 * i4);
                        a = fmaf(wv.x, xin[4 * i4], a); a = fmaf(wv.y, xin[4 * i4 + 1], a);
                        a = fmaf(wv.z, xin[4 * i4 + 2], a); a = fmaf(wv.w, xin[4 * i4 + 3], a);
                    }
                    sQ[qt * TCA_C + g * TCA_SD + o] = a * P.scale;
                }
            }
            __syncthreads();

            // ---- phase 2 + 3 per scale / head group
#pragma unroll 1
            for (int g = 0; g < 2; ++g) {
                const int nT = tile->nT[g];
                const int *toff = sToff + g * (TCA_WB + 1);
                // 2a: key task -> row t of the A operand
                int l = 0, my_nqr = 0, my_q0 = 0;
                bool masked = false;
                if (tid < nT) {
                    l = sTwin[g * TCA_THREADS + tid];
                    const int j = tid - toff[l];
                    const int w = wb0 + t_ws + l;
                    const int m = sMeta[4 * (t_ws + l) + 2 + g];
                    const int nrep = m & 0xff, nmask = m >> 8;
                    masked = nmask > 0 && j == nrep - 1;
                    sTmult[tid] = masked ? nmask : 1;
                    my_nqr = sMeta[4 * (t_ws + l)];
                    my_q0 = sQoff[l];
                    const int row = __ldg(rep_row + (size_t)w * 2 * K + g * K + j);
                    const float cx = sCtr[4 * (t_ws + l)], cy = sCtr[4 * (t_ws + l) + 1], cz = sCtr[4 * (t_ws + l) + 2];
                    float rx = 0.f, ry = 0.f, rz = 0.f;  // masked key: relative offset zeroed
                    if (!masked) {
                        rx = __fsub_rn(__ldg(xyz + 3 * (size_t)row), cx);
                        ry = __fsub_rn(__ldg(xyz + 3 * (size_t)row + 1), cy);
                        rz = __fsub_rn(__ldg(xyz + 3 * (size_t)row + 2), cz);
                    }
                    const float4 *src = (const float4 *)(xn + (size_t)row * TCA_C + g * TCA_SD);
#pragma unroll 2
                    for (int c4 = 0; c4 < TCA_SD / 4; ++c4) {
                        const float4 v = __ldg(src + c4);
                        const float f[4] = {v.x, v.y, v.z, v.w};
                        float o[4];
#pragma unroll
                        for (int k = 0; k < 4; ++k) {
                            const float4 wa = *(const float4 *)(sPos + (g * TCA_SD + 4 * c4 + k) * 8);
                            const float4 wb = *(const float4 *)(sPos + (g * TCA_SD + 4 * c4 + k) * 8 + 4);
                            float a = wb.z;
                            a = fmaf(wa.x, rx, a); a = fmaf(wa.y, ry, a); a = fmaf(wa.z, rz, a);
                            a = fmaf(wa.w, cx, a); a = fmaf(wb.x, cy, a); a = fmaf(wb.y, cz, a);
                            o[k] = to_tf32(f[k] + fmaxf(a, 0.f));
                        }
                        *(float4 *)(sA + (uint32_t)c4 * a_lbo + my_row_off) = make_float4(o[0], o[1], o[2], o[3]);
                    }
                }
                fence_async_smem();
                __syncthreads();
                // 2b: D[128 x 64] = A[128 x 32] . Wkv_g^T on the tensor cores
                if (tid == 0) {
                    tc_fence_after();
#pragma unroll
                    for (int k = 0; k < TCA_SD / 8; ++k) {
                        const uint64_t da = umma_smem_desc(sA_u + (uint32_t)k * 2u * a_lbo, a_lbo, 128);
                        const uint64_t db = umma_smem_desc(sWkv_u + (uint32_t)g * 64u * 32u * 4u + (uint32_t)k * 2u * w_lbo,
                                                           w_lbo, 128);
                        umma_tf32(tmem_d, da, db, idesc, k > 0 ? 1u : 0u);
                    }
                    umma_commit(bar);
                }
                mbar_wait(bar, phase);
                phase ^= 1u;
                tc_fence_after();
                // 2c: this thread's key: K|V back from TMEM, scores against its window's queries
                {
                    float kk[TCA_SD], vv[TCA_SD];
                    tmem_ld32(tmem_d + lane_off, kk);
                    tmem_ld32(tmem_d + lane_off + 32u, vv);
                    if (tid < nT) {
                        const float *bk = sBkv + g * 64;
#pragma unroll
                        for (int i = 0; i < TCA_SD; ++i) { kk[i] += bk[i]; vv[i] += bk[TCA_SD + i]; }
#pragma unroll
                        for (int c4 = 0; c4 < TCA_SD / 4; ++c4)
                            *(float4 *)(sV + tid * TCA_VPITCH + 4 * c4) =
                                make_float4(vv[4 * c4], vv[4 * c4 + 1], vv[4 * c4 + 2], vv[4 * c4 + 3]);
                        const float bias = masked ? -100.0f : 0.f;  // additive mask of the reference
                        for (int s = 0; s < my_nqr; ++s) {
                            const float *qv = sQ + (my_q0 + s) * TCA_C + g * TCA_SD;
#pragma unroll
                            for (int h = 0; h < HEADS; ++h) {
                                float a = 0.f;
#pragma unroll
                                for (int d4 = 0; d4 < HD / 4; ++d4) {
                                    const float4 q4 = *(const float4 *)(qv + h * HD + 4 * d4);
                                    a = fmaf(q4.x, kk[h * HD + 4 * d4], a); a = fmaf(q4.y, kk[h * HD + 4 * d4 + 1], a);
                                    a = fmaf(q4.z, kk[h * HD + 4 * d4 + 2], a); a = fmaf(q4.w, kk[h * HD + 4 * d4 + 3], a);
                                }
                                sS[tid * smax + s * HEADS + h] = a + bias;
                            }
                        }
                    }
                }
                tc_fence_before();
                __syncthreads();
                // 3: softmax over the window's distinct keys and AV, thread = (query, head, quarter of the head)
                for (int e = tid; e < nQ * HEADS * 4; e += TCA_THREADS) {
                    constexpr int DPT = HD / 4;  // channels per thread
                    const int qh = e >> 2, dq = e & 3;
                    const int qt = qh / HEADS, h = qh - qt * HEADS;
                    const int lq = sQwin[qt], s = qt - sQoff[lq];
                    const int t0 = toff[lq], t1 = toff[lq + 1];
                    float mx = -3.0e38f;
                    for (int t = t0; t < t1; ++t) mx = fmaxf(mx, sS[t * smax + s * HEADS + h]);
                    float den = 0.f, acc[DPT];
#pragma unroll
                    for (int d = 0; d < DPT; ++d) acc[d] = 0.f;
                    for (int t = t0; t < t1; ++t) {
                        const float w = exp_neg(sS[t * smax + s * HEADS + h] - mx) * (float)sTmult[t];
                        den += w;
                        const float *vp = sV + t * TCA_VPITCH + h * HD + dq * DPT;
#pragma unroll
                        for (int d = 0; d < DPT; ++d) acc[d] = fmaf(w, vp[d], acc[d]);
                    }
                    const float inv = 1.0f / den;
#pragma unroll
                    for (int d = 0; d < DPT; ++d) sO[qt * TCA_C + g * TCA_SD + h * HD + dq * DPT + d] = acc[d] * inv;
                }
                __syncthreads();
            }

            // ---- phase 4: output projection, thread = (query, head group, 8 outputs): sO -> sQ
            for (int e = tid; e < nQ * 8; e += TCA_THREADS) {
                const int qt = e >> 3, g = (e >> 2) & 1, oq = e & 3;
                float xin[TCA_SD];
#pragma unroll
                for (int i4 = 0; i4 < TCA_SD / 4; ++i4) {
                    const float4 v = *(const float4 *)(sO + qt * TCA_C + g * TCA_SD + 4 * i4);
                    xin[4 * i4] = v.x; xin[4 * i4 + 1] = v.y; xin[4 * i4 + 2] = v.z; xin[4 * i4 + 3] = v.w;
                }
                const float *wp = sWp + g * TCA_WGRP;
                float *dst = sQ + qt * TCA_C + g * TCA_SD;  // q is no longer needed
#pragma unroll 2
                for (int j = 0; j < 8; ++j) {
                    const int o = oq + 4 * j;
                    float a = sBp[g * TCA_SD + o];
#pragma unroll
                    for (int i4 = 0; i4 < TCA_SD / 4; ++i4) {
                        const float4 wv = *(const float4 *)(wp + o * TCA_WPITCH + 4 * i4);
                        a = fmaf(wv.x, xin[4 * i4], a); a = fmaf(wv.y, xin[4 * i4 + 1], a);
                        a = fmaf(wv.z, xin[4 * i4 + 2], a); a = fmaf(wv.w, xin[4 * i4 + 3], a);
                    }
                    dst[o] = a;
                }
            }
            __syncthreads();

            // ---- phase 5: merge.  interp: thread = (win1 voxel, 16-channel quarter), 1/d blend of the
            //      voxel's 3 nearest query rows; otherwise the query voxels take their own rows
            if (P.interp) {
                for (int e = tid; e < nV * 4; e += TCA_THREADS) {
                    const int vt = e >> 2, cq = e & 3;
                    const int lv = sVwin[vt], i = vt - sVoff[lv];
                    const int w = wb0 + t_ws + lv;
                    const int row = __ldg(win1_row + (size_t)w * cap1 + i);
                    const unsigned char *ni = nn_idx + ((size_t)w * cap1 + i) * 3;
                    const float *nw = nn_w + ((size_t)w * cap1 + i) * 3;
                    const int nqr = sMeta[4 * (t_ws + lv)], q0 = sQoff[lv];
                    const int n0 = ni[0], n1 = ni[1], n2 = ni[2];
                    // padded query slots (index >= #real queries) are zero rows in the reference
                    const float *a0 = n0 < nqr ? sQ + (q0 + n0) * TCA_C + 16 * cq : nullptr;
                    const float *a1 = n1 < nqr ? sQ + (q0 + n1) * TCA_C + 16 * cq : nullptr;
                    const float *a2 = n2 < nqr ? sQ + (q0 + n2) * TCA_C + 16 * cq : nullptr;
                    const float w0 = __ldg(nw), w1 = __ldg(nw + 1), w2 = __ldg(nw + 2);
                    float4 *dst = (float4 *)(merged + (size_t)row * TCA_C + 16 * cq);
                    const float4 zero = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
                    for (int c4 = 0; c4 < 4; ++c4) {
                        const float4 p0 = a0 ? *(const float4 *)(a0 + 4 * c4) : zero;
                        const float4 p1 = a1 ? *(const float4 *)(a1 + 4 * c4) : zero;
                        const float4 p2 = a2 ? *(const float4 *)(a2 + 4 * c4) : zero;
                        float4 y;
                        y.x = __fadd_rn(__fadd_rn(__fmul_rn(p0.x, w0), __fmul_rn(p1.x, w1)), __fmul_rn(p2.x, w2));
                        y.y = __fadd_rn(__fadd_rn(__fmul_rn(p0.y, w0), __fmul_rn(p1.y, w1)), __fmul_rn(p2.y, w2));
                        y.z = __fadd_rn(__fadd_rn(__fmul_rn(p0.z, w0), __fmul_rn(p1.z, w1)), __fmul_rn(p2.z, w2));
                        y.w = __fadd_rn(__fadd_rn(__fmul_rn(p0.w, w0), __fmul_rn(p1.w, w1)), __fmul_rn(p2.w, w2));
                        dst[c4] = y;
                    }
                }
            } else {
                for (int e = tid; e < nQ * 4; e += TCA_THREADS) {
                    const int qt = e >> 2, cq = e & 3;
                    const int lq = sQwin[qt], s = qt - sQoff[lq];
                    const int row = __ldg(q_row + (size_t)(wb0 + t_ws + lq) * nq + s);
                    float4 *dst = (float4 *)(merged + (size_t)row * TCA_C + 16 * cq);
#pragma unroll
                    for (int c4 = 0; c4 < 4; ++c4) dst[c4] = *(const float4 *)(sQ + qt * TCA_C + 16 * cq + 4 * c4);
                }
            }
            __syncthreads();
            ws = t_we;
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem_d, 64);
}

static size_t tca_smem_bytes(int smax) {
    size_t floats = 64 * 8 + 2 * TCA_WSZ + 64 + 128 + 64 + 2 * TCA_QCAP * 64 + TCA_THREADS * TCA_VPITCH +
                    (size_t)TCA_THREADS * smax + TCA_WB * 4;
    size_t ints = TCA_WB * 4 + 4 * (TCA_WB + 1) + TCA_QCAP + 2 * TCA_THREADS + TCA_THREADS + TCA_THREADS + 8;
    return 2 * 64 * 32 * 4 + TCA_THREADS * 32 * 4 + (floats + ints) * 4 + 8 + 16 + 128;
}

}  // namespace mssvt

using namespace mssvt;

extern "C" {

/* Tensor-core window attention of a two-window block (see the header of this file).  Weights in
 * their nn.Module layout: pos_w [64][6], wq/wp [32][32], wkv [64][32] per head group.
 * rep_row / meta: the compact key lists of mssvt_block_geometry.  Returns MSSVT_ERR_INVALID for
 * shapes outside C = 64 / 2 x 32 channels / nq <= 32 / K <= 127 / cap1 <= 128 (callers then use
 * mssvt_block_attention). */
int mssvt_block_attention_tc(int C, int heads_per_group, int nq, int key_num_sample, int cap1, int interp,
                             float scale, const float *win_cell, const float *range_min,
                             const float *pos_w, const float *pos_b, const float *wq0, const float *bq0,
                             const float *wkv0, const float *bkv0, const float *wp0, const float *bp0,
                             const float *wq1, const float *bq1, const float *wkv1, const float *bkv1,
                             const float *wp1, const float *bp1, int win_capacity, const int *win_count_total,
                             const int *win_list, const float *xn, const float *xyz, const int *q_row,
                             const int *rep_row, const int *meta, const int *win1_row,
                             const unsigned char *nn_idx, const float *nn_w, float *merged, void *stream) {
    if (C != 64 || (heads_per_group != 1 && heads_per_group != 2 && heads_per_group != 4) || nq <= 0 ||
        nq > TCA_QCAP || key_num_sample <= 0 || key_num_sample > 127 || cap1 <= 0 || cap1 > TCA_THREADS ||
        win_capacity < 0)
        return MSSVT_ERR_INVALID;
    if (win_capacity == 0) return MSSVT_OK;
    if (!win_cell || !range_min || !pos_w || !pos_b || !wq0 || !bq0 || !wkv0 || !bkv0 || !wp0 || !bp0 || !wq1 ||
        !bq1 || !wkv1 || !bkv1 || !wp1 || !bp1 || !win_count_total || !win_list || !xn || !xyz || !q_row ||
        !rep_row || !meta || !merged)
        return MSSVT_ERR_INVALID;
    if (interp && (!win1_row || !nn_idx || !nn_w)) return MSSVT_ERR_INVALID;
    TcAttnParams P;
    P.nq = nq; P.K = key_num_sample; P.cap1 = cap1; P.interp = interp ? 1 : 0;
    P.heads = heads_per_group; P.hd = TCA_SD / heads_per_group; P.smax = nq * heads_per_group;
    P.scale = scale;
    for (int i = 0; i < 3; ++i) { P.win_cell[i] = win_cell[i]; P.lo[i] = range_min[i]; }
    P.pos_w = pos_w; P.pos_b = pos_b;
    P.wq[0] = wq0; P.bq[0] = bq0; P.wkv[0] = wkv0; P.bkv[0] = bkv0; P.wp[0] = wp0; P.bp[0] = bp0;
    P.wq[1] = wq1; P.bq[1] = bq1; P.wkv[1] = wkv1; P.bkv[1] = bkv1; P.wp[1] = wp1; P.bp[1] = bp1;
    const size_t smem = tca_smem_bytes(P.smax);
    if (smem > 227 * 1024) return MSSVT_ERR_INVALID;
    const int per_sm = smem <= 110 * 1024 ? 2 : 1;
    const int batches = (win_capacity + TCA_WB - 1) / TCA_WB;
    int grid = MSSVT_NUM_SMS * per_sm;
    if (grid > batches) grid = batches;
    cudaStream_t s = (cudaStream_t)stream;
    ++g_launches;
#define TCA_LAUNCH(H)                                                                                        \
    cudaFuncSetAttribute(k_block_attention_tc<H>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);    \
    k_block_attention_tc<H><<<grid, TCA_THREADS, smem, s>>>(P, win_capacity, win_count_total,                \
                                                            (const int4 *)win_list, xn, xyz, q_row, rep_row, \
                                                            meta, win1_row, nn_idx, nn_w, merged)
    if (heads_per_group == 1) { TCA_LAUNCH(1); }
    else if (heads_per_group == 2) { TCA_LAUNCH(2); }
    else { TCA_LAUNCH(4); }
#undef TCA_LAUNCH
    return check_launch();
}

}  // extern "C"
