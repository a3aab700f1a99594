// compress_tc.cu -- attention of the one-window (compress) block as a tile kernel on the tcgen05 tensor cores
// (sm_100a).  Same mathematics as k_compress_attention in attention.cu
// (MixedScaleSparseTransformerCompressBlock.forward, mssvt_backbone.py:361-383):
//
//   k_tcc_plan    #real slots per window, tiles of <= 128 key tasks, window centres
//   k_tcc_pool    channel-wise max over the window's rows (zero padding included)
//   k_tc_linear   (tc_linear.cuh) q = (Wq pooled + bq) * scale, on tcgen05
//   k_tcc_tile    task = key of a window: one per voxel + one "pad key" per window that has padded slots (all padded
//                 slots carry the same key: zero feature, offset 0 - centre, mask -100, multiplicity = #padded slots).
//                 128 tasks per tile, three chained tcgen05 GEMMs with every A operand in TMEM:
//                   pos [128 x 8]                         -> MMA 0 (K = 8):  first positional layer        (N = 64)
//                   A1 = relu(D0)             (in place)  -> MMA 1:          D1 = A1 W2^T                  (N = 64)
//                   A2 = xn + relu(D1 + b2)   (in place)  -> MMA 2:          K | V = A2 Wkv^T              (N = 128)
//                 K | V come back per thread through tcgen05.ld; scores against the window's query, then softmax + AV
//                 with one thread per (window, head, quarter head).  Loads of the next tile software-pipelined under
//                 the current one; tile groups share the CTA's weights (see the kernel).
//   k_tc_linear   output projection -> one row per window, on tcgen05
//
// Supported shape: C = 64, one head group (2, 4 or 8 heads), two-layer pos_proj, max_num_win1 <= 127.
// Everything else runs on k_compress_attention.  TF32 (or split 3xTF32) operands for the GEMMs; the bf16 mode uses
// the TF32 kernel here.
#include "tc_linear.cuh"

namespace mssvt {

#define TCC_THREADS 256
#define TCC_ROWS 128     // key tasks per tile (TMEM lanes); two threads per task
#define TCC_TW 64        // windows per tile (at most)
#define TCC_PLAN_WB 128  // windows planned by one warp
#define TCC_C 64
#define TCC_VPITCH 68    // V row pitch in floats (16-byte aligned, conflict-free for quarter warps)
#define TCC_HDR_CTR (2 * TCC_TW * 4)                 // header: records [TW] int (+ pad), then centres [TW] float4
#define TCC_HDR_BYTES (TCC_HDR_CTR + TCC_TW * 16)

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

// shared-memory carve-up of k_tcc_tile (bytes).  Common to the CTA: the weights; per tile group: operands, rows, header
struct TccSmem {
    int wpos, w2, wkv, bias, group0, group_bytes, apos, rows, hdr, bar, total;
    // nt = operand tiles per weight matrix in units of ONE TF32 tile (split TF32: 2; TF32 and split bf16: 1)
    __host__ __device__ TccSmem(int nt, int heads, int groups) {
        wpos = 0;                                     // [hi | lo] x [2 chunks][64][16 B]                          4 KB
        w2 = wpos + 4096;                             // nt x [16 chunks][64][16 B]                         16 KB each
        wkv = w2 + nt * 64 * 64 * 4;                  // nt x [16 chunks][128][16 B]                        32 KB each
        bias = wkv + nt * 128 * 64 * 4;               // b2 [64], bkv [128]
        group0 = bias + (64 + 128) * 4;
        apos = 0;                                     // [hi | lo] x [2 chunks][128][16 B]; later the scores       8 KB
        rows = apos + 8192;                           // gather staging [8 warps][32 rows][128 B], later V [128][VPITCH]
        hdr = rows + TCC_ROWS * TCC_VPITCH * 4;       // 2 x {cnt [TW], toff [TW + 1], ctr [TW] float4}
        bar = hdr + 2 * TCC_HDR_BYTES;                // mbarrier of the group
        group_bytes = (bar + 16 + 127) & ~127;
        (void)heads;
        total = group0 + groups * group_bytes + 16;   // + TMEM base
    }
};

// One tile = 128 key tasks of consecutive pillar windows: one per voxel + one "pad key" per window that has padded
// slots.  A tile GROUP of 256 threads owns a tile: threads t and t + 128 share task row t = TMEM lane t and own one
// half of the channels / heads each.  A CTA hosts G groups that share one copy of the weights and run their tile loops
// independently (named barriers): G = 1 and two CTAs per SM with TF32 operands, G = 2 and one CTA per SM with split
// operands (the doubled weight tiles then leave room for one copy per SM only).  Per tile, four chained GEMMs with M = 128:
//   1. pos = [offset to the window centre | centre | 1 | 0] (8 values, hi / lo) -> shared memory;  MMA 0 (K = 8, always
//      3xTF32): the first positional layer of all rows -> TMEM [0, 64)
//   2. tcgen05.ld, ReLU, rounded, tcgen05.st IN PLACE: the A operand of MMA 1 = second positional layer (W2, N = 64)
//      -> TMEM [128, 192)
//   3. A2 = xn + relu(D1 + b2) -> TMEM [0, 64) (+ lo [64, 128)): the A operand of MMA 2 = K | V projection (N = 128)
//      -> TMEM [128, 256)
//   4. K | V back per thread; scores against the window's (single, max-pooled) query; softmax over the window's keys with
//      the multiplicity of the pad key; AV with one thread per (window, head, quarter head) -> one row per window.
// No operand ever sits in shared memory except the 8-wide positional input; the loads of tile i + 1 (window records ->
// row ids -> coordinates + feature rows) are issued a phase ahead of their use under tile i, and MMA 0 of tile i + 1 is
// issued at the end of tile i.
template <int HEADS, int TERMS, int G>
__global__ void __launch_bounds__(256 * G, G == 1 ? 2 : 1)
k_tcc_tile(TccParams P, const int2 *__restrict__ tiles, const int *__restrict__ tile_count,
           const int *__restrict__ win_rec, const float4 *__restrict__ win_ctr, const float *__restrict__ xn,
           const float *__restrict__ xyz, const int *__restrict__ k_row, const float *__restrict__ Qc,
           float *__restrict__ Oc) {
    constexpr int HD = TCC_C / HEADS;
    constexpr int DPT = HD / 4;
    constexpr int HH = HEADS / 2;  // heads per thread
    constexpr int NT = TERMS == 3 ? 2 : 1;         // (split bf16: two bf16 tiles = the bytes of one TF32 tile)
    constexpr bool BF = TERMS == 2;                // split bf16 operands (hi + mid, kind::f16) for W2 and Wkv
    pdl_launch_dependents();
    extern __shared__ __align__(128) char smem_raw[];
    const int tid = threadIdx.x, grp = tid >> 8, gtid = tid & 255;
    const int gw = (tid >> 5) & 7;                 // warp within the group
    const int r = gtid & (TCC_ROWS - 1), half = gtid >> 7, lane = tid & 31;
    const int n1 = P.n1;
    const TccSmem L(NT, HEADS, G);
    char *sWpos = smem_raw + L.wpos, *sW2 = smem_raw + L.w2, *sWkv = smem_raw + L.wkv;
    float *sB2 = (float *)(smem_raw + L.bias), *sBkv = sB2 + 64;
    char *gbase = smem_raw + L.group0 + grp * L.group_bytes;
    char *sApos = gbase + L.apos, *sRows = gbase + L.rows, *sHdr = gbase + L.hdr;
    float *sS = (float *)sApos;                    // [128][HEADS] scores (the positional operand is consumed by then)
    float *sV = (float *)sRows;                    // [128][VPITCH]
    uint64_t *sBar = (uint64_t *)(gbase + L.bar);
    uint32_t *sTmem = (uint32_t *)(smem_raw + L.group0 + G * L.group_bytes);
    char *stg = sRows + gw * 4096;                 // this warp's staging area (32 half rows of 128 bytes)
    const uint32_t group_barrier = 1u + (uint32_t)grp;
    auto group_sync = [&]() { asm volatile("bar.sync %0, 256;" ::"r"(group_barrier) : "memory"); };

    // ---- weights (static parameters: before the dependency wait)
    stage_packed(P.pos2_w, NT * 64 * 64, sW2);
    stage_packed(P.wkv, NT * 128 * 64, sWkv);
    for (int i = tid; i < 128; i += 256 * G) {   // first positional layer [64][8] = [pos_w | pos_b | 0], hi / lo, canonical
        const int n = i & 63, ch = i >> 6;
        float4 w, hi, lo;
        if (ch == 0) w = make_float4(__ldg(P.pos_w + n * 6), __ldg(P.pos_w + n * 6 + 1), __ldg(P.pos_w + n * 6 + 2), __ldg(P.pos_w + n * 6 + 3));
        else w = make_float4(__ldg(P.pos_w + n * 6 + 4), __ldg(P.pos_w + n * 6 + 5), __ldg(P.pos_b + n), 0.f);
        split_tf32(w, hi, lo);
        *(float4 *)(sWpos + ch * 1024 + n * 16) = hi;
        *(float4 *)(sWpos + 2048 + ch * 1024 + n * 16) = lo;
    }
    if (tid < 64) sB2[tid] = __ldg(P.pos2_b + tid);
    if (tid < 128) sBkv[tid] = __ldg(P.bkv + tid);
    const uint32_t bar = smem_u32(sBar);
    if (gtid == 0) {
        mbar_init(bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (tid < 32) tmem_alloc(smem_u32(sTmem), 256 * G);
    stage_packed_wait();
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tb = *sTmem + (uint32_t)(grp * 256);
    const uint32_t tm_a = tb, tm_alo = tb + 64u, tm_d = tb + 128u;   // A hi / A lo (split operands) / D1, then K | V
    // split bf16: operands packed two per column.  A1 = [hi 64..95 | mid 96..127] (MMA 0's accumulator in 0..63 is still
    // being read by the row's other thread), A2 = [hi 0..31 | mid 32..63]
    const uint32_t tm_a1 = BF ? tb + 64u : tb, tm_a1m = tb + 96u, tm_a2m = tb + 32u;
    const uint32_t lane_off = (uint32_t)((gw & 3) * 32) << 16;
    const uint32_t id_pos = umma_idesc_tf32(128, 64);
    const uint32_t id_n64 = BF ? umma_idesc_bf16(128, 64) : umma_idesc_tf32(128, 64);
    const uint32_t id_n128 = BF ? umma_idesc_bf16(128, 128) : umma_idesc_tf32(128, 128);
    const UmmaDescBase dWpos = umma_desc_base(smem_u32(sWpos), 1024, 128), dApos = umma_desc_base(smem_u32(sApos), 2048, 128),
                       dW2 = umma_desc_base(smem_u32(sW2), 64 * 16, 128), dWkv = umma_desc_base(smem_u32(sWkv), 128 * 16, 128);
    const uint32_t row_off = (uint32_t)(r >> 3) * 128u + (uint32_t)(r & 7) * 16u;   // canonical 8-row groups
    uint32_t phase = 0;
    pdl_wait();  // (everything above touched static parameters only)
    const int T = __ldg(tile_count);
    const int first = blockIdx.x * G + grp, stride = gridDim.x * G;

    // header of a tile (per window: #real keys, first key task, centre) -> buffer `buf`, asynchronously
    auto header_async = [&](int2 tl, int buf) {
        char *h = sHdr + buf * TCC_HDR_BYTES;
        if (gtid < tl.y) {
            asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(smem_u32(h + gtid * 4)), "l"(win_rec + tl.x + gtid) : "memory");
            asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(smem_u32(h + TCC_HDR_CTR + gtid * 16)), "l"(win_ctr + tl.x + gtid) : "memory");
        }
    };
    // role of this thread's task row in a tile whose header has landed: window l, key j, pad key?; -> feature row (issued load)
    auto role = [&](int2 tl, const char *h, int &l, int &row, bool &is_task, bool &pad) {
        const int *rec = (const int *)h;
        // (records: cnt | first task << 8; the window of a task by binary search over the first tasks)
        int lo = 0, hi = tl.y;
        while (hi - lo > 1) {
            const int mid = (lo + hi) >> 1;
            if ((rec[mid] >> 8) <= r) lo = mid; else hi = mid;
        }
        l = lo;
        const int last = rec[tl.y - 1], cl = last & 0xff;
        const int nT = (last >> 8) + cl + (cl < n1 ? 1 : 0);
        is_task = r < nT;
        const int j = r - (rec[l] >> 8);
        pad = j >= (rec[l] & 0xff);
        row = -1;
        if (is_task && !pad) row = __ldg(k_row + (size_t)(tl.x + l) * n1 + j);
    };
    // positional input of this thread's row -> the A operand of MMA 0 (one writer per row: half 0)
    auto pos_operand = [&](const char *h, int l, bool is_task, float x, float y, float z) {
        if (half) return;
        float4 p0 = make_float4(0.f, 0.f, 0.f, 0.f), p1 = p0;
        if (is_task) {
            const float4 ctr = ((const float4 *)(h + TCC_HDR_CTR))[l];
            // (padded slots: grouped coordinate 0 -> offset 0 - centre, quirk Q6: x = y = z = 0 for the pad key)
            p0 = make_float4(__fsub_rn(x, ctr.x), __fsub_rn(y, ctr.y), __fsub_rn(z, ctr.z), ctr.x);
            p1 = make_float4(ctr.y, ctr.z, 1.f, 0.f);
        }
        float4 hi, lo;
        split_tf32(p0, hi, lo);
        *(float4 *)(sApos + row_off) = hi;
        *(float4 *)(sApos + 4096 + row_off) = lo;
        split_tf32(p1, hi, lo);
        *(float4 *)(sApos + 2048 + row_off) = hi;
        *(float4 *)(sApos + 4096 + 2048 + row_off) = lo;
    };
    auto issue_pos = [&]() {   // MMA 0 of the tile whose positional operand was just built
        if (((warp_uniform() & 7) == 0) && elect_one()) {
            tc_fence_after();
            const uint64_t ah = umma_desc_at(dApos, 0u), al = umma_desc_at(dApos, 4096u);
            const uint64_t wh = umma_desc_at(dWpos, 0u), wl = umma_desc_at(dWpos, 2048u);
            umma_tf32(tm_a, ah, wh, id_pos, 0u);
            umma_tf32(tm_a, al, wh, id_pos, 1u);
            umma_tf32(tm_a, ah, wl, id_pos, 1u);
            umma_commit(bar);
        }
        __syncwarp();
    };

    int2 tl = first < T ? __ldg(tiles + first) : make_int2(0, 0);
    int2 tl_next = first + stride < T ? __ldg(tiles + first + stride) : make_int2(0, 0);
    int l = 0, row = -1;
    bool is_task = false, pad = false;
    if (first < T) {   // pipeline prologue: header, role, coordinates, feature rows, positional operand, MMA 0 of the first tile
        header_async(tl, 0);
        stage_packed_wait();
        group_sync();
        role(tl, sHdr, l, row, is_task, pad);
        float px = 0.f, py = 0.f, pz = 0.f;
        if (row >= 0) { px = __ldg(xyz + 3 * (size_t)row); py = __ldg(xyz + 3 * (size_t)row + 1); pz = __ldg(xyz + 3 * (size_t)row + 2); }
        warp_rows_copy_async(stg, xn + 32 * half, row);
        cp_async_commit();
        pos_operand(sHdr, l, is_task, px, py, pz);
        fence_async_smem();
        tc_fence_before();
        group_sync();
        issue_pos();
    }
    int buf = 0;
    for (int t = first; t < T; t += stride, buf ^= 1) {
        const char *hdr = sHdr + buf * TCC_HDR_BYTES, *hdr_next = sHdr + (buf ^ 1) * TCC_HDR_BYTES;
        const int *sRec = (const int *)hdr;
        const int nwin = tl.y;
        const bool more = t + stride < T;
        if (more) header_async(tl_next, buf ^ 1);   // the next tile's window records: on their way now
        cp_async_commit();
        const int2 tl_after = t + 2 * stride < T ? __ldg(tiles + t + 2 * stride) : make_int2(0, 0);
        // the windows' query rows are read after the last MMA: pull them into L1 now
        if (gtid < 2 * nwin) prefetch_l1(Qc + (size_t)tl.x * TCC_C + (size_t)gtid * 32);
        // the warp's 32 half rows (copied by its own lanes at the end of the previous tile) have landed
        float4 f[8];
        cp_async_wait_group<1>();
        __syncwarp();
#pragma unroll
        for (int q = 0; q < 8; ++q)
            f[q] = row >= 0 ? *(const float4 *)(stg + lane * 128 + ((q ^ (lane & 7)) << 4)) : make_float4(0.f, 0.f, 0.f, 0.f);
        mbar_wait(bar, phase);   // MMA 0 of this tile
        phase ^= 1u;
        tc_fence_after();
        {
            // ---- 2. A1 = relu(first positional layer), written back over the accumulator: the A operand of MMA 1
            float d[32];
            tmem_ld32(tm_a + lane_off + (uint32_t)(32 * half), d);
            if constexpr (BF) {
                uint32_t w[16], wm[16];
#pragma unroll
                for (int i = 0; i < 16; ++i) split_bf16x2(fmaxf(d[2 * i], 0.f), fmaxf(d[2 * i + 1], 0.f), w[i], wm[i]);
                tmem_st16(tm_a1 + lane_off + (uint32_t)(16 * half), w);
                tmem_st16(tm_a1m + lane_off + (uint32_t)(16 * half), wm);
            } else {
                if (TERMS == 3) {
                    float lo[32];
#pragma unroll
                    for (int i = 0; i < 32; ++i) split_tf32(fmaxf(d[i], 0.f), d[i], lo[i]);
                    tmem_st32(tm_alo + lane_off + (uint32_t)(32 * half), lo);
                } else {
#pragma unroll
                    for (int i = 0; i < 32; ++i) d[i] = to_tf32(fmaxf(d[i], 0.f));
                }
                tmem_st32(tm_a + lane_off + (uint32_t)(32 * half), d);
            }
            tmem_st_wait();
        }
        cp_async_wait_group<0>();   // this thread's part of the next header
        tc_fence_before();
        group_sync();               // (also: the next header is visible to the group)
        if (((warp_uniform() & 7) == 0) && elect_one()) {   // D1 = A1 W2^T -> [128, 192)
            tc_fence_after();
            if constexpr (BF) {
#pragma unroll
                for (int k = 0; k < TCC_C / 16; ++k) {   // K = 16 per MMA: 8 packed A columns, two 16-byte B chunks
                    const uint64_t wh = umma_desc_at(dW2, (uint32_t)k * 2u * 64u * 16u);
                    umma_bf16_ts(tm_d, tm_a1 + (uint32_t)k * 8u, wh, id_n64, k > 0 ? 1u : 0u);
                    umma_bf16_ts(tm_d, tm_a1m + (uint32_t)k * 8u, wh, id_n64, 1u);
                    umma_bf16_ts(tm_d, tm_a1 + (uint32_t)k * 8u, umma_desc_at(dW2, 64u * 64u * 2u + (uint32_t)k * 2u * 64u * 16u), id_n64, 1u);
                }
            } else
#pragma unroll
            for (int k = 0; k < TCC_C / 8; ++k) {
                const uint64_t wh = umma_desc_at(dW2, (uint32_t)k * 2u * 64u * 16u);
                umma_tf32_ts(tm_d, tm_a + (uint32_t)k * 8u, wh, id_n64, k > 0 ? 1u : 0u);
                if (TERMS == 3) {
                    umma_tf32_ts(tm_d, tm_alo + (uint32_t)k * 8u, wh, id_n64, 1u);
                    umma_tf32_ts(tm_d, tm_a + (uint32_t)k * 8u, umma_desc_at(dW2, 64u * 64u * 4u + (uint32_t)k * 2u * 64u * 16u), id_n64, 1u);
                }
            }
            umma_commit(bar);
        }
        __syncwarp();
        // while the MMA runs: role and feature row id of this thread in the NEXT tile
        int n_l = 0, n_row = -1;
        bool n_task = false, n_pad = false;
        if (more) role(tl_next, hdr_next, n_l, n_row, n_task, n_pad);
        mbar_wait(bar, phase);
        phase ^= 1u;
        tc_fence_after();
        {
            // ---- 3. A2 = xn + relu(D1 + b2): the A operand of MMA 2
            float d[32];
            tmem_ld32(tm_d + lane_off + (uint32_t)(32 * half), d);
            const float *b2 = sB2 + 32 * half;
            if constexpr (BF) {
                uint32_t w[16], wm[16];
#pragma unroll
                for (int q = 0; q < 8; ++q) {
                    split_bf16x2(f[q].x + fmaxf(d[4 * q] + b2[4 * q], 0.f), f[q].y + fmaxf(d[4 * q + 1] + b2[4 * q + 1], 0.f), w[2 * q], wm[2 * q]);
                    split_bf16x2(f[q].z + fmaxf(d[4 * q + 2] + b2[4 * q + 2], 0.f), f[q].w + fmaxf(d[4 * q + 3] + b2[4 * q + 3], 0.f), w[2 * q + 1], wm[2 * q + 1]);
                }
                tmem_st16(tm_a + lane_off + (uint32_t)(16 * half), w);
                tmem_st16(tm_a2m + lane_off + (uint32_t)(16 * half), wm);
            } else if (TERMS == 3) {
                float lo[32];
#pragma unroll
                for (int q = 0; q < 8; ++q) {
                    split_tf32(f[q].x + fmaxf(d[4 * q] + b2[4 * q], 0.f), d[4 * q], lo[4 * q]);
                    split_tf32(f[q].y + fmaxf(d[4 * q + 1] + b2[4 * q + 1], 0.f), d[4 * q + 1], lo[4 * q + 1]);
                    split_tf32(f[q].z + fmaxf(d[4 * q + 2] + b2[4 * q + 2], 0.f), d[4 * q + 2], lo[4 * q + 2]);
                    split_tf32(f[q].w + fmaxf(d[4 * q + 3] + b2[4 * q + 3], 0.f), d[4 * q + 3], lo[4 * q + 3]);
                }
                tmem_st32(tm_alo + lane_off + (uint32_t)(32 * half), lo);
            } else {
#pragma unroll
                for (int q = 0; q < 8; ++q) {
                    d[4 * q] = to_tf32(f[q].x + fmaxf(d[4 * q] + b2[4 * q], 0.f));
                    d[4 * q + 1] = to_tf32(f[q].y + fmaxf(d[4 * q + 1] + b2[4 * q + 1], 0.f));
                    d[4 * q + 2] = to_tf32(f[q].z + fmaxf(d[4 * q + 2] + b2[4 * q + 2], 0.f));
                    d[4 * q + 3] = to_tf32(f[q].w + fmaxf(d[4 * q + 3] + b2[4 * q + 3], 0.f));
                }
            }
            if constexpr (!BF) tmem_st32(tm_a + lane_off + (uint32_t)(32 * half), d);
            tmem_st_wait();
        }
        tc_fence_before();
        group_sync();
        if (((warp_uniform() & 7) == 0) && elect_one()) {   // K | V = A2 Wkv^T -> [128, 256)
            tc_fence_after();
            if constexpr (BF) {
#pragma unroll
                for (int k = 0; k < TCC_C / 16; ++k) {
                    const uint64_t wh = umma_desc_at(dWkv, (uint32_t)k * 2u * 128u * 16u);
                    umma_bf16_ts(tm_d, tm_a + (uint32_t)k * 8u, wh, id_n128, k > 0 ? 1u : 0u);
                    umma_bf16_ts(tm_d, tm_a2m + (uint32_t)k * 8u, wh, id_n128, 1u);
                    umma_bf16_ts(tm_d, tm_a + (uint32_t)k * 8u, umma_desc_at(dWkv, 128u * 64u * 2u + (uint32_t)k * 2u * 128u * 16u), id_n128, 1u);
                }
            } else
#pragma unroll
            for (int k = 0; k < TCC_C / 8; ++k) {
                const uint64_t wh = umma_desc_at(dWkv, (uint32_t)k * 2u * 128u * 16u);
                umma_tf32_ts(tm_d, tm_a + (uint32_t)k * 8u, wh, id_n128, k > 0 ? 1u : 0u);
                if (TERMS == 3) {
                    umma_tf32_ts(tm_d, tm_alo + (uint32_t)k * 8u, wh, id_n128, 1u);
                    umma_tf32_ts(tm_d, tm_a + (uint32_t)k * 8u, umma_desc_at(dWkv, 128u * 64u * 4u + (uint32_t)k * 2u * 128u * 16u), id_n128, 1u);
                }
            }
            umma_commit(bar);
        }
        __syncwarp();
        // the next tile's coordinates: on their way during the last MMA (the row id was loaded a phase ago)
        float nx = 0.f, ny = 0.f, nz = 0.f;
        asm volatile("" : "+r"(n_row));
        if (n_row >= 0) { nx = __ldg(xyz + 3 * (size_t)n_row); ny = __ldg(xyz + 3 * (size_t)n_row + 1); nz = __ldg(xyz + 3 * (size_t)n_row + 2); }
        mbar_wait(bar, phase);
        phase ^= 1u;
        tc_fence_after();
        // ---- 4. scores of this thread's heads against the window's query (K = columns 0..63 of D2), V = columns 64..127
        {
            float sc[HH];
#pragma unroll
            for (int h = 0; h < HH; ++h) sc[h] = 0.f;
            float d[32], vv[32];
            tmem_ld32(tm_d + lane_off + (uint32_t)(32 * half), d);
            tmem_ld32(tm_d + lane_off + 64u + (uint32_t)(32 * half), vv);
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
        tc_fence_before();
        group_sync();
        // ---- softmax over the window's keys and AV, thread = (window, head, quarter of the head)
        for (int e = gtid; e < nwin * HEADS * 4; e += 256) {
            const int dq = e & 3, lh = e >> 2, h = lh % HEADS, lw = lh / HEADS;
            const int rec = sRec[lw], t0 = rec >> 8, cnt = rec & 0xff;
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
        group_sync();   // scores and V rows are consumed: their memory takes the next tile's operands
        if (more) {
            warp_rows_copy_async(stg, xn + 32 * half, n_row);
            pos_operand(hdr_next, n_l, n_task, nx, ny, nz);
        }
        cp_async_commit();
        fence_async_smem();
        tc_fence_before();
        group_sync();
        if (more) issue_pos();
        l = n_l; row = n_row; is_task = n_task; pad = n_pad;
        tl = tl_next; tl_next = tl_after;
    }
    stage_packed_wait();
    tc_fence_before();
    __syncthreads();
    if (tid < 32) tmem_dealloc(*sTmem, 256 * G);
}

static size_t tcc_tile_smem_bytes(int heads, int terms, int groups) {
    return (size_t)TccSmem(terms == 3 ? 2 : 1, heads, groups).total + 128;
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
        (terms != 1 && terms != 2 && terms != 3))
        return MSSVT_ERR_INVALID;
    if (win_capacity == 0) return MSSVT_OK;
    if (!win_cell || !range_min || !pos_w || !pos_b || !pos2_w || !pos2_b || !wq || !bq || !wkv || !bkv || !wp || !bp ||
        !win_count_total || !win_list || !xn || !xyz || !k_row || !tiles_in || !tile_count || !win_rec || !win_ctr_in ||
        !scratch || !out)
        return MSSVT_ERR_INVALID;
    // terms = 2 (split bf16): pos2_w / wkv packed by mssvt_pack_operand_bf16x2; the two small row-wise projections
    // (query, output) have no bf16 form and take wq / wp packed with terms = 3
    const int lin_terms = terms == 2 ? 3 : terms;
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
        tcl_launch(L, rows, win_capacity, Qc, s, lin_terms);
    }

    // TF32 operands: one tile group per CTA, two CTAs per SM; split operands: two groups share the CTA's weights
    const int groups = terms == 3 ? 2 : 1;
    const size_t smem = tcc_tile_smem_bytes(heads, terms, groups);
    if (smem > 227 * 1024) return MSSVT_ERR_INVALID;
    const int grid = MSSVT_NUM_SMS * (groups == 1 ? 2 : 1);
    ++g_launches;
#define TCC_LAUNCH(H, T, GG)                                                                               \
    cudaFuncSetAttribute(k_tcc_tile<H, T, GG>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);    \
    launch_pdl(k_tcc_tile<H, T, GG>, dim3(grid), dim3(256 * GG), smem, s, P, tiles, tile_count, win_rec, win_ctr, xn, \
               xyz, k_row, (const float *)Qc, Oc)
    if (terms == 2) {
        if (heads == 2) { TCC_LAUNCH(2, 2, 1); }
        else if (heads == 4) { TCC_LAUNCH(4, 2, 1); }
        else { TCC_LAUNCH(8, 2, 1); }
    } else if (terms == 3) {
        if (heads == 2) { TCC_LAUNCH(2, 3, 2); }
        else if (heads == 4) { TCC_LAUNCH(4, 3, 2); }
        else { TCC_LAUNCH(8, 3, 2); }
    } else {
        if (heads == 2) { TCC_LAUNCH(2, 1, 1); }
        else if (heads == 4) { TCC_LAUNCH(4, 1, 1); }
        else { TCC_LAUNCH(8, 1, 1); }
    }
#undef TCC_LAUNCH

    {
        const TclCopyRows rows = {Oc, win_count_total, nullptr, win_capacity};
        const TclParams L = {wp, bp, nullptr, 1.0f};
        tcl_launch(L, rows, win_capacity, out, s, lin_terms);
    }
    return check_launch();
}

}  // extern "C"
