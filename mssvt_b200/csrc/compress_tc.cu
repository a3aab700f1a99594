// compress_tc.cu -- attention of the one-window (compress) block, task-parallel, with the second
// positional-embedding layer and the K/V projection on the tcgen05 tensor cores (sm_100a).
// Same mathematics as k_compress_attention in attention.cu
// (MixedScaleSparseTransformerCompressBlock.forward, mssvt_backbone.py:361-383):
//
//   k_tc_linear   (tc_linear.cuh) q = (Wq pooled + bq) * scale, on tcgen05
//   k_tcc_plan    #real slots per window, tiles of <= 128 key tasks, window centres
//   k_tcc_pool    channel-wise max over the window's rows (zero padding included)
//   k_tcc_keys    task = key of a window: one per voxel + one "pad key" per window that has padded
//                 slots (all padded slots carry the same key: zero feature, offset 0 - centre, mask -100,
//                 multiplicity = #padded slots).  128 keys per tile:
//                   A1 = relu(pos layer 1)               -> 8 x tcgen05.mma  D1 = A1 W2^T        (N = 64)
//                   A2 = xn + relu(D1 + b2)              -> 8 x tcgen05.mma  D2 = A2 Wkv^T       (N = 128)
//                 K|V come back per thread through tcgen05.ld; scores against the window's query,
//                 then softmax + AV with one thread per (window, head, quarter head).
//   k_tc_linear   output projection -> one row per window, on tcgen05
//
// Supported shape: C = 64, one head group (2, 4 or 8 heads), two-layer pos_proj, max_num_win1 <= 127.
// Everything else runs on k_compress_attention.  TF32 operands for the two tensor-core GEMMs only.
#include "tc_linear.cuh"

namespace mssvt {

#define TCC_THREADS 256
#define TCC_ROWS 128     // key tasks per tile (TMEM lanes); two threads per task
#define TCC_TW 64        // windows per tile (at most)
#define TCC_PLAN_WB 128  // windows planned by one warp
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

// ------------------------------------------------------------------------------- plan, query pool

// One warp plans TCC_PLAN_WB consecutive windows: #real slots per window (binary search: real slots are
// compacted at the front of k_row), the greedy cut into tiles of <= 128 key tasks (real keys + one pad
// key per window that has padded slots) and <= TCC_TW windows, window centres.
__global__ void __launch_bounds__(256)
k_tcc_plan(int n1, int win_cap, const int *__restrict__ win_count_total, const int4 *__restrict__ win_list,
           const int *__restrict__ k_row, float3 win_cell, float3 lo, int2 *__restrict__ tiles,
           int *__restrict__ tile_count, int *__restrict__ win_rec, float4 *__restrict__ win_ctr) {
    const int num_wins = min(win_cap, __ldg(win_count_total));
    const int lane = threadIdx.x & 31;
    const int w0 = ((blockIdx.x * blockDim.x + threadIdx.x) >> 5) * TCC_PLAN_WB;
    if (w0 >= num_wins) return;
    const int w1 = min(w0 + TCC_PLAN_WB, num_wins);
    int ts = w0, a = 0, nw = 0;
    for (int wb = w0; wb < w1; wb += 32) {
        const int w = wb + lane;
        int cnt = 0, k = 0;
        if (w < w1) {
            const int *kr = k_row + (size_t)w * n1;
            int l = 0, h = n1;
            while (l < h) { const int mid = (l + h) >> 1; if (__ldg(kr + mid) >= 0) l = mid + 1; else h = mid; }
            cnt = l;
            k = cnt + (cnt < n1 ? 1 : 0);
            const int4 win = __ldg(win_list + w);
            win_ctr[w] = make_float4(world_coord(win.w, win_cell.x, lo.x), world_coord(win.z, win_cell.y, lo.y),
                                     world_coord(win.y, win_cell.z, lo.z), 0.f);
        }
        int my_a = 0;
        const int n = min(32, w1 - wb);
        for (int i = 0; i < n; ++i) {
            const int ki = __shfl_sync(0xffffffffu, k, i);
            if (nw > 0 && (a + ki > TCC_ROWS || nw == TCC_TW)) {
                if (lane == 0) tiles[atomicAdd(tile_count, 1)] = make_int2(ts, nw);
                ts = wb + i; a = nw = 0;
            }
            if (i == lane) my_a = a;
            a += ki; ++nw;
        }
        if (w < w1) win_rec[w] = cnt | (my_a << 8);
    }
    if (lane == 0 && nw > 0) tiles[atomicAdd(tile_count, 1)] = make_int2(ts, nw);
}

// channel-wise max over the window's n1 slots; padded slots contribute zeros (Q6).
// thread = (window, 4 channels): rows of a window are read as coalesced 256-byte segments
__global__ void __launch_bounds__(256)
k_tcc_pool(int n1, int win_cap, const int *__restrict__ win_count_total, const int *__restrict__ win_rec,
           const int *__restrict__ k_row, const float *__restrict__ xn, float *__restrict__ pooled) {
    pdl_launch_dependents();
    pdl_wait();
    const int num_wins = min(win_cap, __ldg(win_count_total));
    const long long total = (long long)num_wins * 16;
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total;
         e += (long long)gridDim.x * blockDim.x) {
        const int c4 = (int)(e & 15);
        const size_t w = (size_t)(e >> 4);
        const int cnt = __ldg(win_rec + w) & 0xff;
        const int *kr = k_row + w * n1;
        const float init = cnt < n1 ? 0.f : -3.0e38f;
        float4 acc = make_float4(init, init, init, init);
        int t = 0;
        for (; t + 4 <= cnt; t += 4) {  // four independent row loads in flight
            const int r0 = __ldg(kr + t), r1 = __ldg(kr + t + 1), r2 = __ldg(kr + t + 2), r3 = __ldg(kr + t + 3);
            const float4 v0 = __ldg((const float4 *)(xn + (size_t)r0 * TCC_C) + c4);
            const float4 v1 = __ldg((const float4 *)(xn + (size_t)r1 * TCC_C) + c4);
            const float4 v2 = __ldg((const float4 *)(xn + (size_t)r2 * TCC_C) + c4);
            const float4 v3 = __ldg((const float4 *)(xn + (size_t)r3 * TCC_C) + c4);
            acc.x = fmaxf(fmaxf(fmaxf(acc.x, v0.x), fmaxf(v1.x, v2.x)), v3.x);
            acc.y = fmaxf(fmaxf(fmaxf(acc.y, v0.y), fmaxf(v1.y, v2.y)), v3.y);
            acc.z = fmaxf(fmaxf(fmaxf(acc.z, v0.z), fmaxf(v1.z, v2.z)), v3.z);
            acc.w = fmaxf(fmaxf(fmaxf(acc.w, v0.w), fmaxf(v1.w, v2.w)), v3.w);
        }
        for (; t < cnt; ++t) {
            const float4 v = __ldg((const float4 *)(xn + (size_t)__ldg(kr + t) * TCC_C) + c4);
            acc.x = fmaxf(acc.x, v.x); acc.y = fmaxf(acc.y, v.y); acc.z = fmaxf(acc.z, v.z); acc.w = fmaxf(acc.w, v.w);
        }
        *((float4 *)(pooled + w * TCC_C) + c4) = acc;
    }
}

// ------------------------------------------------------------------------------- keys + attention

// largest l in [0, n) with off[l] <= v (off[n] > v)
__device__ __forceinline__ int tcc_tile_window(const int *off, int n, int v) {
    int lo = 0, hi = n;
    while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (off[mid] <= v) lo = mid; else hi = mid;
    }
    return lo;
}

// 256 threads per tile of 128 key tasks: threads t and t + 128 share task row t = TMEM lane t and own one
// half of the channels / of the heads each (HEADS >= 2)
template <int HEADS, int TERMS>
__global__ void __launch_bounds__(TCC_THREADS, TERMS == 3 ? 1 : 2)
k_tcc_keys(TccParams P, const int2 *__restrict__ tiles, const int *__restrict__ tile_count,
           const int *__restrict__ win_rec, const float4 *__restrict__ win_ctr, const float *__restrict__ xn,
           const float *__restrict__ xyz, const int *__restrict__ k_row, const float *__restrict__ Qc,
           float *__restrict__ Oc) {
    constexpr int HD = TCC_C / HEADS;
    constexpr int DPT = HD / 4;
    constexpr int HH = HEADS / 2;  // heads per thread
    pdl_launch_dependents();
    extern __shared__ __align__(128) char smem_raw[];
    const int tid = threadIdx.x, warp = tid >> 5;
    const int r = tid & (TCC_ROWS - 1), half = tid >> 7;
    const int n1 = P.n1;

    constexpr int NT = TERMS == 3 ? 2 : 1;                  // operand tiles: hi [, lo] (3xTF32, tc_common.cuh)
    constexpr int A_TILE = TCC_ROWS * TCC_C * 4;            // 32 KB
    constexpr int A_REGION = NT * A_TILE > TCC_ROWS * TCC_VPITCH * 4 ? NT * A_TILE : TCC_ROWS * TCC_VPITCH * 4;
    char *sW2 = smem_raw;                                   // NT x [64 x 64] canonical TF32    16 KB each
    char *sWkv = sW2 + NT * 64 * 64 * 4;                    // NT x [128 x 64] canonical TF32   32 KB each
    char *sA = sWkv + NT * 128 * 64 * 4;                    // NT x [128 x 64] canonical (32 KB each) ...
    float *sV = (float *)sA;                                // ... reused as V [128][VPITCH] (34 KB)
    float *sPos = (float *)(sA + A_REGION);                 // [64][8]
    float *sB2 = sPos + 64 * 8;                             // [64]
    float *sBkv = sB2 + 64;                                 // [128]
    float *sS = sBkv + 128;                                 // [128][HEADS] scores
    float4 *sCtr = (float4 *)(sS + TCC_ROWS * HEADS);       // [TW] window centres
    int *sCnt = (int *)(sCtr + TCC_TW);                     // [TW] real keys per window
    int *sToff = sCnt + TCC_TW;                             // [TW + 1] first key task of each window (+ end)
    uint64_t *sBar = (uint64_t *)(sToff + TCC_TW + 2);
    uint32_t *sTmem = (uint32_t *)(sBar + 1);

    stage_packed(P.pos2_w, NT * 64 * 64, sW2);
    stage_packed(P.wkv, NT * 128 * 64, sWkv);
    for (int i = tid; i < 64 * 8; i += TCC_THREADS) {
        const int c = i >> 3, k = i & 7;
        sPos[i] = k < 6 ? __ldg(P.pos_w + c * 6 + k) : k == 6 ? __ldg(P.pos_b + c) : 0.f;
    }
    if (tid < 64) sB2[tid] = __ldg(P.pos2_b + tid);
    if (tid < 128) sBkv[tid] = __ldg(P.bkv + tid);
    const uint32_t bar = smem_u32(sBar);
    if (tid == 0) {
        mbar_init(bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) tmem_alloc(smem_u32(sTmem), 256);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *sTmem;
    const uint32_t tmem_d1 = tmem_base, tmem_d2 = tmem_base + 64u;  // D1: 64 columns, D2: 128 columns
    const uint32_t lane_off = (uint32_t)((warp & 3) * 32) << 16;
    const uint32_t idesc1 = umma_idesc_tf32(128, 64), idesc2 = umma_idesc_tf32(128, 128);
    const uint32_t sA_u = smem_u32(sA), sW2_u = smem_u32(sW2), sWkv_u = smem_u32(sWkv);
    const uint32_t a_lbo = TCC_ROWS * 16, w2_lbo = 64 * 16, wkv_lbo = 128 * 16;
    const uint32_t my_row_off = (uint32_t)(r >> 3) * 128u + (uint32_t)(r & 7) * 16u;
    uint32_t phase = 0;
    pdl_wait();  // (everything above touched static parameters only)
    const int T = __ldg(tile_count);

    for (int t = blockIdx.x; t < T; t += gridDim.x) {
        const int2 tl = __ldg(tiles + t);
        const int nwin = tl.y;
        if (tid < nwin) {
            const int rec = __ldg(win_rec + tl.x + tid);
            const int cnt = rec & 0xff;
            sCnt[tid] = cnt;
            sToff[tid] = rec >> 8;
            sCtr[tid] = __ldg(win_ctr + tl.x + tid);
            if (tid == nwin - 1) sToff[nwin] = (rec >> 8) + cnt + (cnt < n1 ? 1 : 0);
        }
        __syncthreads();
        const int nT = sToff[nwin];
        // the windows' query rows are read after the second MMA: pull them into L1 now
        if (tid < 2 * nwin)
            asm volatile("prefetch.global.L1 [%0];" ::"l"(Qc + (size_t)tl.x * TCC_C + (size_t)tid * 32));

        // ---- A1 = relu(pos layer 1 (offset to the window centre, centre)), this thread's 32 channels
        const bool is_task = r < nT;
        int l = 0, row = -1;
        bool pad = false;
        float rx = 0.f, ry = 0.f, rz = 0.f;
        float4 ctr = make_float4(0.f, 0.f, 0.f, 0.f);
        if (is_task) {
            l = tcc_tile_window(sToff, nwin, r);
            const int j = r - sToff[l];
            pad = j >= sCnt[l];
            ctr = sCtr[l];
            float px = 0.f, py = 0.f, pz = 0.f;  // padded slots: grouped coordinate 0 -> offset 0 - centre
            if (!pad) {
                row = __ldg(k_row + (size_t)(tl.x + l) * n1 + j);
                px = __ldg(xyz + 3 * (size_t)row); py = __ldg(xyz + 3 * (size_t)row + 1);
                pz = __ldg(xyz + 3 * (size_t)row + 2);
            }
            rx = __fsub_rn(px, ctr.x); ry = __fsub_rn(py, ctr.y); rz = __fsub_rn(pz, ctr.z);
        }
        // the feature rows (needed after the first MMA) are gathered now, 8 lanes per 128-byte half row,
        // through the A tile's memory (pad key: zero feature)
        float4 f[8];
        warp_rows_load(sA + warp * 4096, is_task && !pad ? (const float4 *)(xn + (size_t)row * TCC_C + 32 * half) : nullptr, f);
        __syncthreads();  // every warp is done with its staging area: the A tile may be written
        if (is_task) {
#pragma unroll 4
            for (int c4 = 0; c4 < 8; ++c4) {
                float o[4];
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const float4 wa = *(const float4 *)(sPos + (32 * half + 4 * c4 + k) * 8);
                    const float4 wb = *(const float4 *)(sPos + (32 * half + 4 * c4 + k) * 8 + 4);
                    float a = wb.z;
                    a = fmaf(wa.x, rx, a); a = fmaf(wa.y, ry, a); a = fmaf(wa.z, rz, a);
                    a = fmaf(wa.w, ctr.x, a); a = fmaf(wb.x, ctr.y, a); a = fmaf(wb.y, ctr.z, a);
                    o[k] = fmaxf(a, 0.f);
                }
                float4 hi, lo;
                split_tf32(make_float4(o[0], o[1], o[2], o[3]), hi, lo);
                *(float4 *)(sA + (uint32_t)(8 * half + c4) * a_lbo + my_row_off) = hi;
                if (TERMS == 3) *(float4 *)(sA + A_TILE + (uint32_t)(8 * half + c4) * a_lbo + my_row_off) = lo;
            }
        }
        stage_packed_wait();
        fence_async_smem();
        __syncthreads();
        if (issuer_elected()) {  // D1 = A1 W2^T
            tc_fence_after();
#pragma unroll
            for (int k = 0; k < TCC_C / 8; ++k)
                umma_step<TERMS>(tmem_d1, sA_u + (uint32_t)k * 2u * a_lbo, a_lbo, A_TILE,
                                 sW2_u + (uint32_t)k * 2u * w2_lbo, w2_lbo, 64 * 64 * 4, idesc1, k == 0);
            umma_commit(bar);
        }
        mbar_wait(bar, phase);
        phase ^= 1u;
        tc_fence_after();
        // ---- A2 = xn + relu(D1 + b2)
        {
            float d[32];
            tmem_ld32(tmem_d1 + lane_off + (uint32_t)(32 * half), d);
            if (is_task) {
                const float *b2 = sB2 + 32 * half;
#pragma unroll
                for (int q = 0; q < 8; ++q) {
                    float4 v, hi, lo;
                    v.x = f[q].x + fmaxf(d[4 * q] + b2[4 * q], 0.f);
                    v.y = f[q].y + fmaxf(d[4 * q + 1] + b2[4 * q + 1], 0.f);
                    v.z = f[q].z + fmaxf(d[4 * q + 2] + b2[4 * q + 2], 0.f);
                    v.w = f[q].w + fmaxf(d[4 * q + 3] + b2[4 * q + 3], 0.f);
                    split_tf32(v, hi, lo);
                    *(float4 *)(sA + (uint32_t)(8 * half + q) * a_lbo + my_row_off) = hi;
                    if (TERMS == 3) *(float4 *)(sA + A_TILE + (uint32_t)(8 * half + q) * a_lbo + my_row_off) = lo;
                }
            }
        }
        fence_async_smem();
        tc_fence_before();
        __syncthreads();
        if (issuer_elected()) {  // D2 = A2 Wkv^T
            tc_fence_after();
#pragma unroll
            for (int k = 0; k < TCC_C / 8; ++k)
                umma_step<TERMS>(tmem_d2, sA_u + (uint32_t)k * 2u * a_lbo, a_lbo, A_TILE,
                                 sWkv_u + (uint32_t)k * 2u * wkv_lbo, wkv_lbo, 128 * 64 * 4, idesc2, k == 0);
            umma_commit(bar);
        }
        mbar_wait(bar, phase);
        phase ^= 1u;
        tc_fence_after();
        // ---- scores of this thread's heads against the window's query (K = columns 0..63 of D2),
        //      V = columns 64..127
        {
            float sc[HH];
#pragma unroll
            for (int h = 0; h < HH; ++h) sc[h] = 0.f;
            float d[32], vv[32];
            tmem_ld32(tmem_d2 + lane_off + (uint32_t)(32 * half), d);
            tmem_ld32(tmem_d2 + lane_off + 64u + (uint32_t)(32 * half), vv);
            tc_fence_before();
            __syncthreads();  // all K|V are in registers: the A tile may become V
            if (is_task) {
                const float4 *qv = (const float4 *)(Qc + (size_t)(tl.x + l) * TCC_C + 32 * half);
                // (biases: the K bias shifts all scores of the query equally -> cancelled by the softmax;
                //  the V bias is added once per output below)
#pragma unroll
                for (int q = 0; q < 8; ++q) {
                    const float4 q4 = __ldg(qv + q);
                    const int h = (4 * q) / HD;
                    sc[h] = fmaf(q4.x, d[4 * q], sc[h]);
                    sc[h] = fmaf(q4.y, d[4 * q + 1], sc[h]);
                    sc[h] = fmaf(q4.z, d[4 * q + 2], sc[h]);
                    sc[h] = fmaf(q4.w, d[4 * q + 3], sc[h]);
                }
#pragma unroll
                for (int h = 0; h < HH; ++h) sS[r * HEADS + half * HH + h] = sc[h] + (pad ? -100.0f : 0.f);
#pragma unroll
                for (int c4 = 0; c4 < 8; ++c4)
                    *(float4 *)(sV + r * TCC_VPITCH + 32 * half + 4 * c4) =
                        make_float4(vv[4 * c4], vv[4 * c4 + 1], vv[4 * c4 + 2], vv[4 * c4 + 3]);
            }
        }
        __syncthreads();
        // ---- softmax over the window's keys and AV, thread = (window, head, quarter of the head)
        for (int e = tid; e < nwin * HEADS * 4; e += TCC_THREADS) {
            const int dq = e & 3, lh = e >> 2, h = lh % HEADS, lw = lh / HEADS;
            const int t0 = sToff[lw], cnt = sCnt[lw];
            const int nk = cnt + (cnt < n1 ? 1 : 0);
            const float *sc = sS + t0 * HEADS + h;
            float mx = -3.0e38f;
            for (int k = 0; k < nk; ++k) mx = fmaxf(mx, sc[k * HEADS]);
            float den = 0.f, acc[DPT];
#pragma unroll
            for (int d = 0; d < DPT; ++d) acc[d] = 0.f;
            const float *vp = sV + t0 * TCC_VPITCH + h * HD + dq * DPT;
            for (int k = 0; k < nk; ++k) {
                float wgt = exp_neg(sc[k * HEADS] - mx);
                if (k >= cnt) wgt *= (float)(n1 - cnt);  // the pad key counts once per padded slot
                den += wgt;
#pragma unroll
                for (int d = 0; d < DPT; ++d) acc[d] = fmaf(wgt, vp[k * TCC_VPITCH + d], acc[d]);
            }
            const float inv = 1.0f / den;
            float *dst = Oc + (size_t)(tl.x + lw) * TCC_C + h * HD + dq * DPT;
            const float *bv = sBkv + 64 + h * HD + dq * DPT;
#pragma unroll
            for (int d = 0; d < DPT; ++d) dst[d] = fmaf(acc[d], inv, bv[d]);
        }
        __syncthreads();
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem_base, 256);
}

static size_t tcc_keys_smem_bytes(int heads, int terms) {
    const size_t nt = terms == 3 ? 2 : 1;
    size_t a_region = nt * TCC_ROWS * TCC_C * 4;
    if (a_region < (size_t)TCC_ROWS * TCC_VPITCH * 4) a_region = (size_t)TCC_ROWS * TCC_VPITCH * 4;
    return nt * (64 * 64 * 4 + 128 * 64 * 4) + a_region + (size_t)(64 * 8 + 64 + 128 + TCC_ROWS * heads) * 4 +
           TCC_TW * 16 + (2 * TCC_TW + 2) * 4 + 8 + 16 + 128;
}

}  // namespace mssvt

using namespace mssvt;

extern "C" {

/* Tile plan of mssvt_compress_attention_tc: #real slots per window, tiles of <= 128 key tasks, window
 * centres.  A function of the window rows only (coordinates), so it can be made ahead of / concurrently
 * with the feature kernels.  tiles (win_capacity, 2) int, tile_count (1) int, win_rec (win_capacity) int,
 * win_ctr (win_capacity, 4) float: opaque, caller-allocated. */
int mssvt_compress_tiles(int n1, int win_capacity, const int *win_count_total, const int *win_list,
                         const int *k_row, const float *win_cell, const float *range_min, int *tiles,
                         int *tile_count, int *win_rec, float *win_ctr, void *stream) {
    if (n1 <= 0 || n1 > 127 || win_capacity < 0) return MSSVT_ERR_INVALID;
    if (!win_count_total || !win_list || !k_row || !win_cell || !range_min || !tiles || !tile_count || !win_rec ||
        !win_ctr)
        return MSSVT_ERR_INVALID;
    cudaStream_t s = (cudaStream_t)stream;
    if (cudaMemsetAsync(tile_count, 0, sizeof(int), s) != cudaSuccess) return MSSVT_ERR_LAUNCH;
    if (win_capacity == 0) return MSSVT_OK;
    const int plan_warps = (win_capacity + TCC_PLAN_WB - 1) / TCC_PLAN_WB;
    ++g_launches;
    k_tcc_plan<<<(plan_warps + 7) / 8, 256, 0, s>>>(n1, win_capacity, win_count_total, (const int4 *)win_list, k_row,
                                                    make_float3(win_cell[0], win_cell[1], win_cell[2]),
                                                    make_float3(range_min[0], range_min[1], range_min[2]),
                                                    (int2 *)tiles, tile_count, win_rec, (float4 *)win_ctr);
    return check_launch();
}

/* Tensor-core attention of a one-window (compress) block (see the header of this file).  Weights in
 * nn.Module layout: pos_w [64][6]; packed by mssvt_pack_operand_tf32: wq / wp / pos2_w [64][64] and
 * wkv [128][64].  k_row: (cap, n1) global rows from mssvt_window_rows; tiles .. win_ctr: mssvt_compress_tiles.
 * scratch: 3 * win_capacity * 64 floats.  out: (cap, 64).
 * Supported: C = 64, one head group with 2, 4 or 8 heads, n1 <= 127; -1 otherwise. */
int mssvt_compress_attention_tc(int C, int heads, int n1, int terms, float scale, const float *win_cell,
                                const float *range_min, const float *pos_w, const float *pos_b,
                                const float *pos2_w, const float *pos2_b, const float *wq, const float *bq,
                                const float *wkv, const float *bkv, const float *wp, const float *bp,
                                int win_capacity, const int *win_count_total, const int *win_list,
                                const float *xn, const float *xyz, const int *k_row, const int *tiles_in,
                                const int *tile_count, const int *win_rec, const float *win_ctr_in, float *scratch,
                                float *out, void *stream) {
    if (C != 64 || (heads != 2 && heads != 4 && heads != 8) || n1 <= 0 || n1 > 127 || win_capacity < 0 ||
        (terms != 1 && terms != 3))
        return MSSVT_ERR_INVALID;
    if (win_capacity == 0) return MSSVT_OK;
    if (!win_cell || !range_min || !pos_w || !pos_b || !pos2_w || !pos2_b || !wq || !bq || !wkv || !bkv || !wp || !bp ||
        !win_count_total || !win_list || !xn || !xyz || !k_row || !tiles_in || !tile_count || !win_rec || !win_ctr_in ||
        !scratch || !out)
        return MSSVT_ERR_INVALID;
    TccParams P;
    P.n1 = n1; P.heads = heads; P.scale = scale;
    for (int i = 0; i < 3; ++i) { P.win_cell[i] = win_cell[i]; P.lo[i] = range_min[i]; }
    P.pos_w = pos_w; P.pos_b = pos_b; P.pos2_w = pos2_w; P.pos2_b = pos2_b;
    P.wq = wq; P.bq = bq; P.wkv = wkv; P.bkv = bkv; P.wp = wp; P.bp = bp;
    // scratch: Qc | Oc | pooled
    float *Qc = scratch, *Oc = scratch + (size_t)win_capacity * 64, *pooled = scratch + 2 * (size_t)win_capacity * 64;
    const float4 *win_ctr = (const float4 *)win_ctr_in;
    const int2 *tiles = (const int2 *)tiles_in;
    cudaStream_t s = (cudaStream_t)stream;
    ++g_launches;
    launch_pdl(k_tcc_pool, dim3(MSSVT_NUM_SMS * 8), dim3(256), 0, s, n1, win_capacity, win_count_total, (const int *)win_rec,
               k_row, xn, pooled);
    {
        const TclCopyRows rows = {pooled, win_count_total, nullptr, win_capacity};
        const TclParams L = {wq, bq, nullptr, scale};
        tcl_launch(L, rows, win_capacity, Qc, s, terms);
    }

    const size_t smem = tcc_keys_smem_bytes(heads, terms);
    if (smem > 227 * 1024) return MSSVT_ERR_INVALID;
    const int per_sm = smem <= 110 * 1024 ? 2 : 1;  // 2 x 256 TMEM columns = all 512
    const int grid = MSSVT_NUM_SMS * per_sm;
    ++g_launches;
#define TCC_LAUNCH(H, T)                                                                                  \
    cudaFuncSetAttribute(k_tcc_keys<H, T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);        \
    launch_pdl(k_tcc_keys<H, T>, dim3(grid), dim3(TCC_THREADS), smem, s, P, tiles, tile_count, win_rec, win_ctr, xn, \
               xyz, k_row, (const float *)Qc, Oc)
    if (terms == 3) {
        if (heads == 2) { TCC_LAUNCH(2, 3); }
        else if (heads == 4) { TCC_LAUNCH(4, 3); }
        else { TCC_LAUNCH(8, 3); }
    } else {
        if (heads == 2) { TCC_LAUNCH(2, 1); }
        else if (heads == 4) { TCC_LAUNCH(4, 1); }
        else { TCC_LAUNCH(8, 1); }
    }
#undef TCC_LAUNCH

    {
        const TclCopyRows rows = {Oc, win_count_total, nullptr, win_capacity};
        const TclParams L = {wp, bp, nullptr, 1.0f};
        tcl_launch(L, rows, win_capacity, out, s, terms);
    }
    return check_launch();
}

}  // extern "C"
