// compress_tc.cu -- attention of the one-window (compress) block, task-parallel, with the second
// positional-embedding layer and the K/V projection on the tcgen05 tensor cores (sm_100a).
// Same mathematics as k_compress_attention in attention.cu
// (MixedScaleSparseTransformerCompressBlock.forward, mssvt_backbone.py:361-383):
//
//   k_tc_linear   (tc_linear.cuh) q = (Wq maxpool(window rows incl. zero padding) + bq) * scale, on tcgen05
//   k_tcc_keys    thread = key of a window: one per voxel + one "pad key" per window that has padded
//                 slots (all padded slots carry the same key: zero feature, offset 0 - centre, mask -100,
//                 multiplicity = #padded slots).  128 keys per tile:
//                   A1 = relu(pos layer 1)               -> 8 x tcgen05.mma  D1 = A1 W2^T        (N = 64)
//                   A2 = xn + relu(D1 + b2)              -> 8 x tcgen05.mma  D2 = A2 Wkv^T       (N = 128)
//                 K|V come back per thread through tcgen05.ld; scores against the window's query,
//                 then softmax + AV with one thread per (window, head, quarter head).
//   k_tc_linear   output projection -> one row per window, on tcgen05
//
// Supported shape: C = 64, one head group (1, 2, 4 or 8 heads), two-layer pos_proj, max_num_win1 <= 127.
// Everything else runs on k_compress_attention.  TF32 operands for the two tensor-core GEMMs only.
#include "tc_linear.cuh"

namespace mssvt {

#define TCC_THREADS 128
#define TCC_WB 64        // windows per batch (tile candidates)
#define TCC_C 64
#define TCC_VPITCH 68    // V row pitch in floats (16-byte aligned, conflict-free for quarter warps)

struct TccParams {
    int n1, heads;
    float scale;
    float win_cell[3], lo[3];
    const float *pos_w, *pos_b;    // [64][6], [64]
    const float *pos2_w, *pos2_b;  // [64][64] packed (mssvt_pack_operand_tf32), [64]
    const float *wq, *bq;          // [64][64] packed, [64]
    const float *wkv, *bkv;        // [128][64] packed, [128]
    const float *wp, *bp;          // [64][64] packed, [64]
};

// ------------------------------------------------------------------------------- query

// rows of k_tc_linear for the query projection: channel-wise max over the window's n1 slots; padded
// slots contribute zeros (Q6)
struct TccQueryRows {
    int n1, win_cap;
    const int *win_count_total, *k_row;
    const float *xn;
    __device__ void init(float *) const {}
    __device__ int rows() const { return min(win_cap, __ldg(win_count_total)); }
    __device__ void load(int w, int half, const float *, float *in) const {
        const int *kr = k_row + (size_t)w * n1;
        const bool full = __ldg(kr + n1 - 1) >= 0;
#pragma unroll
        for (int c = 0; c < 32; ++c) in[c] = full ? -3.0e38f : 0.f;
        for (int t = 0; t < n1; ++t) {
            const int row = __ldg(kr + t);
            if (row < 0) break;  // real slots are compacted at the front
            const float4 *src = (const float4 *)(xn + (size_t)row * TCC_C + half * 32);
#pragma unroll
            for (int c4 = 0; c4 < 8; ++c4) {
                const float4 v = __ldg(src + c4);
                in[4 * c4] = fmaxf(in[4 * c4], v.x); in[4 * c4 + 1] = fmaxf(in[4 * c4 + 1], v.y);
                in[4 * c4 + 2] = fmaxf(in[4 * c4 + 2], v.z); in[4 * c4 + 3] = fmaxf(in[4 * c4 + 3], v.w);
            }
        }
    }
};

// ------------------------------------------------------------------------------- keys + attention

struct TccTile {
    int ws, we, nT;
};

template <int HEADS>
__global__ void __launch_bounds__(TCC_THREADS, 2)
k_tcc_keys(TccParams P, int win_cap, const int *__restrict__ win_count_total, const int4 *__restrict__ win_list,
           const float *__restrict__ xn, const float *__restrict__ xyz, const int *__restrict__ k_row,
           const float *__restrict__ Qc, float *__restrict__ Oc) {
    constexpr int HD = TCC_C / HEADS;
    constexpr int DPT = HD / 4;
    extern __shared__ __align__(128) char smem_raw[];
    const int tid = threadIdx.x, warp = tid >> 5;
    const int n1 = P.n1;

    char *sW2 = smem_raw;                                   // [64 x 64] canonical TF32    16 KB
    char *sWkv = sW2 + 64 * 64 * 4;                         // [128 x 64] canonical TF32   32 KB
    char *sA = sWkv + 128 * 64 * 4;                         // [128 x 64] canonical (32 KB) ...
    float *sV = (float *)sA;                                // ... reused as V [128][VPITCH] (34 KB)
    float *sPos = sV + TCC_THREADS * TCC_VPITCH;            // [64][8]
    float *sB2 = sPos + 64 * 8;                             // [64]
    float *sBkv = sB2 + 64;                                 // [128]
    float *sS = sBkv + 128;                                 // [128][HEADS] scores
    float *sCtr = sS + TCC_THREADS * HEADS;                 // [WB][4]
    int *sCnt = (int *)(sCtr + TCC_WB * 4);                 // [WB] real keys per window
    int *sToff = sCnt + TCC_WB;                             // [WB + 1] prefix of key tasks in the tile
    int *sTwin = sToff + TCC_WB + 1;                        // [128] local window of each key task
    int *sTmult = sTwin + TCC_THREADS;                      // [128] multiplicity (pad key: #padded slots)
    int *sTile = sTmult + TCC_THREADS;                      // TccTile + pad (4 ints)
    uint64_t *sBar = (uint64_t *)(sTile + 4 + ((2 * TCC_WB + 1 + 2 * TCC_THREADS + 4) & 1));
    uint32_t *sTmem = (uint32_t *)(sBar + 1);

    stage_packed(P.pos2_w, 64 * 64, sW2);
    stage_packed(P.wkv, 128 * 64, sWkv);
    for (int i = tid; i < 64 * 8; i += TCC_THREADS) {
        const int c = i >> 3, k = i & 7;
        sPos[i] = k < 6 ? __ldg(P.pos_w + c * 6 + k) : k == 6 ? __ldg(P.pos_b + c) : 0.f;
    }
    for (int i = tid; i < 64; i += TCC_THREADS) sB2[i] = __ldg(P.pos2_b + i);
    for (int i = tid; i < 128; i += TCC_THREADS) sBkv[i] = __ldg(P.bkv + i);
    const uint32_t bar = smem_u32(sBar);
    if (tid == 0) {
        mbar_init(bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) tmem_alloc(smem_u32(sTmem), 256);
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *sTmem;
    const uint32_t tmem_d1 = tmem_base, tmem_d2 = tmem_base + 64u;  // D1: 64 columns, D2: 128 columns
    const uint32_t lane_off = (uint32_t)(warp * 32) << 16;
    const uint32_t idesc1 = umma_idesc_tf32(128, 64), idesc2 = umma_idesc_tf32(128, 128);
    const uint32_t sA_u = smem_u32(sA), sW2_u = smem_u32(sW2), sWkv_u = smem_u32(sWkv);
    const uint32_t a_lbo = TCC_THREADS * 16, w2_lbo = 64 * 16, wkv_lbo = 128 * 16;
    const uint32_t my_row_off = (uint32_t)(tid >> 3) * 128u + (uint32_t)(tid & 7) * 16u;
    uint32_t phase = 0;

    const int num_wins = min(win_cap, __ldg(win_count_total));
    const int batches = (num_wins + TCC_WB - 1) / TCC_WB;
    TccTile *tile = (TccTile *)sTile;

    for (int batch = blockIdx.x; batch < batches; batch += gridDim.x) {
        const int wb0 = batch * TCC_WB, nb = min(TCC_WB, num_wins - wb0);
        __syncthreads();
        if (tid < nb) {
            const int *kr = k_row + (size_t)(wb0 + tid) * n1;
            int cnt = 0;  // real slots are compacted at the front: binary search for the first -1
            int lo = 0, hi = n1;
            while (lo < hi) { const int mid = (lo + hi) >> 1; if (__ldg(kr + mid) >= 0) lo = mid + 1; else hi = mid; }
            cnt = lo;
            sCnt[tid] = cnt;
            const int4 win = __ldg(win_list + wb0 + tid);
            sCtr[4 * tid] = world_coord(win.w, P.win_cell[0], P.lo[0]);
            sCtr[4 * tid + 1] = world_coord(win.z, P.win_cell[1], P.lo[1]);
            sCtr[4 * tid + 2] = world_coord(win.y, P.win_cell[2], P.lo[2]);
        }
        __syncthreads();
        int ws = 0;
        while (ws < nb) {
            if (tid == 0) {  // tile = greedy prefix of the batch with <= 128 keys (real + one pad key per window)
                int we = ws, at = 0;
                while (we < nb) {
                    const int k = sCnt[we] + (sCnt[we] < n1 ? 1 : 0);
                    if (we > ws && at + k > TCC_THREADS) break;
                    sToff[we - ws] = at;
                    at += k;
                    ++we;
                }
                sToff[we - ws] = at;
                tile->ws = ws; tile->we = we; tile->nT = at;
            }
            __syncthreads();
            const int t_ws = tile->ws, t_we = tile->we, nT = tile->nT;
            const int nwin = t_we - t_ws;
            if (tid < nwin)
                for (int i = sToff[tid]; i < sToff[tid + 1]; ++i) sTwin[i] = tid;
            __syncthreads();

            // ---- A1 = relu(pos layer 1 (offset to the window centre, centre))
            const bool is_task = tid < nT;
            int l = 0, row = -1;
            bool pad = false;
            if (is_task) {
                l = sTwin[tid];
                const int j = tid - sToff[l];
                const int cnt = sCnt[t_ws + l];
                pad = j >= cnt;
                sTmult[tid] = pad ? n1 - cnt : 1;
                const float cx = sCtr[4 * (t_ws + l)], cy = sCtr[4 * (t_ws + l) + 1], cz = sCtr[4 * (t_ws + l) + 2];
                float px = 0.f, py = 0.f, pz = 0.f;  // padded slots: grouped coordinate 0 -> offset 0 - centre
                if (!pad) {
                    row = __ldg(k_row + (size_t)(wb0 + t_ws + l) * n1 + j);
                    px = __ldg(xyz + 3 * (size_t)row); py = __ldg(xyz + 3 * (size_t)row + 1);
                    pz = __ldg(xyz + 3 * (size_t)row + 2);
                }
                const float rx = __fsub_rn(px, cx), ry = __fsub_rn(py, cy), rz = __fsub_rn(pz, cz);
#pragma unroll 4
                for (int c4 = 0; c4 < TCC_C / 4; ++c4) {
                    float o[4];
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        const float4 wa = *(const float4 *)(sPos + (4 * c4 + k) * 8);
                        const float4 wb = *(const float4 *)(sPos + (4 * c4 + k) * 8 + 4);
                        float a = wb.z;
                        a = fmaf(wa.x, rx, a); a = fmaf(wa.y, ry, a); a = fmaf(wa.z, rz, a);
                        a = fmaf(wa.w, cx, a); a = fmaf(wb.x, cy, a); a = fmaf(wb.y, cz, a);
                        o[k] = to_tf32(fmaxf(a, 0.f));
                    }
                    *(float4 *)(sA + (uint32_t)c4 * a_lbo + my_row_off) = make_float4(o[0], o[1], o[2], o[3]);
                }
            }
            stage_packed_wait();
            fence_async_smem();
            __syncthreads();
            if (tid == 0) {  // D1 = A1 W2^T
                tc_fence_after();
#pragma unroll
                for (int k = 0; k < TCC_C / 8; ++k)
                    umma_tf32(tmem_d1, umma_smem_desc(sA_u + (uint32_t)k * 2u * a_lbo, a_lbo, 128),
                              umma_smem_desc(sW2_u + (uint32_t)k * 2u * w2_lbo, w2_lbo, 128), idesc1, k > 0 ? 1u : 0u);
                umma_commit(bar);
            }
            mbar_wait(bar, phase);
            phase ^= 1u;
            tc_fence_after();
            // ---- A2 = xn + relu(D1 + b2)   (pad key: zero feature)
            for (int c0 = 0; c0 < TCC_C; c0 += 32) {
                float d[32];
                tmem_ld32(tmem_d1 + lane_off + (uint32_t)c0, d);
                if (is_task) {
#pragma unroll
                    for (int q = 0; q < 8; ++q) {
                        float4 f = make_float4(0.f, 0.f, 0.f, 0.f);
                        if (!pad) f = __ldg((const float4 *)(xn + (size_t)row * TCC_C + c0) + q);
                        float4 v;
                        v.x = to_tf32(f.x + fmaxf(d[4 * q] + sB2[c0 + 4 * q], 0.f));
                        v.y = to_tf32(f.y + fmaxf(d[4 * q + 1] + sB2[c0 + 4 * q + 1], 0.f));
                        v.z = to_tf32(f.z + fmaxf(d[4 * q + 2] + sB2[c0 + 4 * q + 2], 0.f));
                        v.w = to_tf32(f.w + fmaxf(d[4 * q + 3] + sB2[c0 + 4 * q + 3], 0.f));
                        *(float4 *)(sA + (uint32_t)(c0 / 4 + q) * a_lbo + my_row_off) = v;
                    }
                }
            }
            fence_async_smem();
            tc_fence_before();
            __syncthreads();
            if (tid == 0) {  // D2 = A2 Wkv^T
                tc_fence_after();
#pragma unroll
                for (int k = 0; k < TCC_C / 8; ++k)
                    umma_tf32(tmem_d2, umma_smem_desc(sA_u + (uint32_t)k * 2u * a_lbo, a_lbo, 128),
                              umma_smem_desc(sWkv_u + (uint32_t)k * 2u * wkv_lbo, wkv_lbo, 128), idesc2, k > 0 ? 1u : 0u);
                umma_commit(bar);
            }
            mbar_wait(bar, phase);
            phase ^= 1u;
            tc_fence_after();
            // ---- scores against the window's query (K = columns 0..63 of D2), V = columns 64..127
            {
                float sc[HEADS];
#pragma unroll
                for (int h = 0; h < HEADS; ++h) sc[h] = 0.f;
                const float4 *qv = (const float4 *)(Qc + (size_t)(wb0 + t_ws + l) * TCC_C);
                for (int c0 = 0; c0 < TCC_C; c0 += 32) {
                    float d[32];
                    tmem_ld32(tmem_d2 + lane_off + (uint32_t)c0, d);
                    if (is_task) {
#pragma unroll
                        for (int q = 0; q < 8; ++q) {
                            const float4 q4 = __ldg(qv + c0 / 4 + q);
                            const int h = (c0 + 4 * q) / HD;
                            sc[h] = fmaf(q4.x, d[4 * q] + sBkv[c0 + 4 * q], sc[h]);
                            sc[h] = fmaf(q4.y, d[4 * q + 1] + sBkv[c0 + 4 * q + 1], sc[h]);
                            sc[h] = fmaf(q4.z, d[4 * q + 2] + sBkv[c0 + 4 * q + 2], sc[h]);
                            sc[h] = fmaf(q4.w, d[4 * q + 3] + sBkv[c0 + 4 * q + 3], sc[h]);
                        }
                    }
                }
                float vv[TCC_C];
                tmem_ld32(tmem_d2 + lane_off + 64u, vv);
                tmem_ld32(tmem_d2 + lane_off + 96u, vv + 32);
                tc_fence_before();
                __syncthreads();  // all K|V are in registers: the A tile may become V
                if (is_task) {
#pragma unroll
                    for (int h = 0; h < HEADS; ++h) sS[tid * HEADS + h] = sc[h] + (pad ? -100.0f : 0.f);
#pragma unroll
                    for (int c4 = 0; c4 < TCC_C / 4; ++c4)
                        *(float4 *)(sV + tid * TCC_VPITCH + 4 * c4) =
                            make_float4(vv[4 * c4] + sBkv[64 + 4 * c4], vv[4 * c4 + 1] + sBkv[64 + 4 * c4 + 1],
                                        vv[4 * c4 + 2] + sBkv[64 + 4 * c4 + 2], vv[4 * c4 + 3] + sBkv[64 + 4 * c4 + 3]);
                }
            }
            __syncthreads();
            // ---- softmax over the window's keys and AV, thread = (window, head, quarter of the head)
            for (int e = tid; e < nwin * HEADS * 4; e += TCC_THREADS) {
                const int dq = e & 3, lh = e >> 2, h = lh % HEADS, lw = lh / HEADS;
                const int t0 = sToff[lw], t1 = sToff[lw + 1];
                float mx = -3.0e38f;
                for (int t = t0; t < t1; ++t) mx = fmaxf(mx, sS[t * HEADS + h]);
                float den = 0.f, acc[DPT];
#pragma unroll
                for (int d = 0; d < DPT; ++d) acc[d] = 0.f;
                for (int t = t0; t < t1; ++t) {
                    const float wgt = exp_neg(sS[t * HEADS + h] - mx) * (float)sTmult[t];
                    den += wgt;
                    const float *vp = sV + t * TCC_VPITCH + h * HD + dq * DPT;
#pragma unroll
                    for (int d = 0; d < DPT; ++d) acc[d] = fmaf(wgt, vp[d], acc[d]);
                }
                const float inv = 1.0f / den;
                float *dst = Oc + (size_t)(wb0 + t_ws + lw) * TCC_C + h * HD + dq * DPT;
#pragma unroll
                for (int d = 0; d < DPT; ++d) dst[d] = acc[d] * inv;
            }
            __syncthreads();
            ws = t_we;
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem_base, 256);
}

static size_t tcc_keys_smem_bytes(int heads) {
    size_t floats = TCC_THREADS * TCC_VPITCH + 64 * 8 + 64 + 128 + (size_t)TCC_THREADS * heads + TCC_WB * 4;
    size_t ints = 2 * TCC_WB + 1 + 2 * TCC_THREADS + 4;
    ints += ints & 1;
    return 64 * 64 * 4 + 128 * 64 * 4 + (floats + ints) * 4 + 8 + 16 + 128;
}

}  // namespace mssvt

using namespace mssvt;

extern "C" {

/* Tensor-core attention of a one-window (compress) block (see the header of this file).  Weights in
 * nn.Module layout: pos_w [64][6]; packed by mssvt_pack_operand_tf32: wq / wp / pos2_w [64][64] and
 * wkv [128][64].  k_row: (cap, n1)
 * global rows from mssvt_window_rows.  scratch: 2 * win_capacity * 64 floats.  out: (cap, 64).
 * Supported: C = 64, one head group with 1, 2, 4 or 8 heads, n1 <= 127; -1 otherwise. */
int mssvt_compress_attention_tc(int C, int heads, int n1, float scale, const float *win_cell,
                                const float *range_min, const float *pos_w, const float *pos_b,
                                const float *pos2_w, const float *pos2_b, const float *wq, const float *bq,
                                const float *wkv, const float *bkv, const float *wp, const float *bp,
                                int win_capacity, const int *win_count_total, const int *win_list,
                                const float *xn, const float *xyz, const int *k_row, float *scratch, float *out,
                                void *stream) {
    if (C != 64 || (heads != 1 && heads != 2 && heads != 4 && heads != 8) || n1 <= 0 || n1 > 127 || win_capacity < 0)
        return MSSVT_ERR_INVALID;
    if (win_capacity == 0) return MSSVT_OK;
    if (!win_cell || !range_min || !pos_w || !pos_b || !pos2_w || !pos2_b || !wq || !bq || !wkv || !bkv || !wp || !bp ||
        !win_count_total || !win_list || !xn || !xyz || !k_row || !scratch || !out)
        return MSSVT_ERR_INVALID;
    TccParams P;
    P.n1 = n1; P.heads = heads; P.scale = scale;
    for (int i = 0; i < 3; ++i) { P.win_cell[i] = win_cell[i]; P.lo[i] = range_min[i]; }
    P.pos_w = pos_w; P.pos_b = pos_b; P.pos2_w = pos2_w; P.pos2_b = pos2_b;
    P.wq = wq; P.bq = bq; P.wkv = wkv; P.bkv = bkv; P.wp = wp; P.bp = bp;
    float *Qc = scratch, *Oc = scratch + (size_t)win_capacity * 64;
    cudaStream_t s = (cudaStream_t)stream;
    {
        const TccQueryRows rows = {n1, win_capacity, win_count_total, k_row, xn};
        const TclParams L = {wq, bq, nullptr, scale};
        tcl_launch(L, rows, win_capacity, Qc, s);
    }

    const size_t smem = tcc_keys_smem_bytes(heads);
    if (smem > 227 * 1024) return MSSVT_ERR_INVALID;
    const int per_sm = smem <= 110 * 1024 ? 2 : 1;  // 2 x 256 TMEM columns = all 512
    const int batches = (win_capacity + TCC_WB - 1) / TCC_WB;
    int grid = MSSVT_NUM_SMS * per_sm;
    if (grid > batches) grid = batches;
    ++g_launches;
#define TCC_LAUNCH(H)                                                                                     \
    cudaFuncSetAttribute(k_tcc_keys<H>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);           \
    k_tcc_keys<H><<<grid, TCC_THREADS, smem, s>>>(P, win_capacity, win_count_total, (const int4 *)win_list, \
                                                  xn, xyz, k_row, Qc, Oc)
    if (heads == 1) { TCC_LAUNCH(1); }
    else if (heads == 2) { TCC_LAUNCH(2); }
    else if (heads == 4) { TCC_LAUNCH(4); }
    else { TCC_LAUNCH(8); }
#undef TCC_LAUNCH

    {
        const TclCopyRows rows = {Oc, win_count_total, nullptr, win_capacity};
        const TclParams L = {wp, bp, nullptr, 1.0f};
        tcl_launch(L, rows, win_capacity, out, s);
    }
    return check_launch();
}

}  // extern "C"
