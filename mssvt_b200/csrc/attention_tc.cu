// attention_tc.cu -- window attention of a two-window MsSVT block as ONE tile kernel on the tcgen05 tensor
// cores (sm_100a).  Same mathematics as k_block_attention in attention.cu (mssvt_backbone.py:260-336,
// mssvt_utils.py:88-157), different mapping: every stage runs with one THREAD per row over the whole frame.
//
//   k_tca_plan    once per frame geometry: packs consecutive windows of one scale into tiles of <= 128 ROWS
//                 (the windows' distinct keys, then their real queries), per-window offsets
//   k_tca_tile    thread = row of a tile of ONE scale (head group).  Per tile, five steps, four of them GEMMs
//                 with M = 128 (lane = row) on tcgen05:
//                 1. pos = [rel | centre | 1 | 0] (8 values, split hi / lo) -> shared memory; MMA 1 (K = 8,
//                    always 3xTF32): the positional-embedding layer Conv1d(6 -> C) of every row -> TMEM
//                 2. tcgen05.ld, ReLU, + the gathered 32-channel slice of the layer-normed row, TF32 round,
//                    tcgen05.st IN PLACE: TMEM now holds the A operand of MMA 2 (A from TMEM, N = 96):
//                    [K | V | Q * scale] = A [Wk; Wv; scale Wq]^T -- key rows use K | V, query rows use Q
//                 3. scores q.k of every (key, query of its window, head) on the FP32 pipe from registers /
//                    shared memory, additive -100 for the key that stands for the masked slots
//                 4. softmax over the window's distinct keys (with the multiplicity of the masked key) and AV
//                    with one thread per (query, head, quarter head); head outputs -> A operand in shared
//                    memory; MMA 3: the output projection of the group (N = 32)
//                 5. tcgen05.ld, + bias, projected query rows -> (#queries, 64) array
//   k_tca_merge   (only without interpolation; with it the blend lives in mssvt_ffn_tc mode 2)
//
// The per-window matrices are tiny (2.3 real queries x 6.5 / 16 distinct keys on a 150 k-voxel frame, 3.9 M
// score triples per block): QK^T and AV stay on the FP32 pipe, where they cost ~8 us per block when packed
// without divergence.  As tcgen05 tiles (M = 128 lanes = queries of ~10 windows, block-diagonal mask) they need
// 128 x HEADS more TMEM columns per tile for the 4 x lane redundancy of the diagonal blocks -> one CTA per SM
// instead of four, two more MMA round trips per tile, and the same number of exp / mask instructions; see
// DESIGN.md section 5.  What DOES pay on the tensor cores are the three per-row linear maps above.
//
// Queries are addressed by a compact id (q_base[w] + slot, an exclusive scan over the windows done with the
// geometry); the projected rows form a dense (#queries, 64) fp32 array that lives in L2.
//
// Supported shape (config S0 and relatives): C = 64, two head groups of 32 channels, 1/2/4 heads per
// group, nq <= 32, key_num_sample <= 63, max_num_win1 <= 128.  Everything else runs on
// k_block_attention.  Precision: TF32 (or split 3xTF32) operands for the positional embedding, the K / V / Q
// and output projections; scores, softmax, AV and interpolation are fp32.
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
#define TCA_QMAX 48      // queries per tile (rows of the output-projection operand)
#define TCA_PLAN_WB 64   // windows planned by one warp
#define TCA_HDR_BYTES (TCA_TW * 32 + TCA_THREADS)   // window records + centres + row table of one tile
#define TCA_C 64
#define TCA_SD 32

struct TcAttnParams {
    int nq, K, cap1, interp, heads;
    const float *wpos;                         // [64][8] = [pos_w | pos_b | 0], packed with terms = 3 ([hi | lo])
    const float *wkvq[2];                      // per group [96][32] = [Wk; Wv; scale * Wq], packed (terms)
    const float *wp[2];                        // per group [32][32] output projection, packed (terms)
    const float *bq[2], *bkv[2], *bp[2];       // [32], [64], [32] per group
    float scale;
};

// ------------------------------------------------------------------------------- tile plan

// A tile = consecutive windows of ONE scale whose rows -- distinct keys + real queries -- fill the 128 rows of
// an MMA.  The plan is a function of the geometry only, so it is made once per frame and shared by every block
// that uses the same window lists.  One warp plans TCA_PLAN_WB windows of one scale: lanes load 32 windows at
// a time, the greedy cut is a 32-step scan over shuffled values (every lane runs it, lane i keeps step i).
// Windows without a real query get no rows at all (nobody would read their K/V).
__global__ void __launch_bounds__(256)
k_tca_plan(int heads, int win_cap, const int *__restrict__ win_count_total, const int4 *__restrict__ win_list,
           const int4 *__restrict__ meta, const int *__restrict__ q_base, float3 win_cell, float3 lo,
           int2 *__restrict__ tiles, int *__restrict__ tile_count, int4 *__restrict__ win_rec,
           float4 *__restrict__ win_ctr) {
    __shared__ int2 s_tiles[8][TCA_PLAN_WB];   // the tiles of a warp's chunk (at most one per window)
    const int num_wins = min(win_cap, __ldg(win_count_total));
    const int lane = threadIdx.x & 31;
    const int wid = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int g = wid & 1, w0 = (wid >> 1) * TCA_PLAN_WB;
    if (w0 >= num_wins) return;
    int2 *mine = s_tiles[threadIdx.x >> 5];
    const int w1 = min(w0 + TCA_PLAN_WB, num_wins);
    int ts = w0, a = 0, aq = 0, as = 0, nw = 0;  // open tile: first window, key rows, query rows, score slots, windows
    int nt = 0;                                  // tiles closed so far
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
            if (nw > 0 && (a + aq + ri + qi > TCA_THREADS || aq + qi > TCA_QMAX || as + si > TCA_SBUD || nw == TCA_TW)) {
                if (a > 0) {   // (a tile without key rows has no real query either: nobody would read it)
                    if (lane == 0) mine[nt] = make_int2(ts, nw);
                    ++nt;
                }
                ts = wb + i; a = aq = as = nw = 0;
            }
            if (i == lane) { my_a = a; my_aq = aq; my_as = as; }
            a += ri; aq += qi; as += si; ++nw;
        }
        if (w < w1)
            win_rec[(size_t)g * win_cap + w] = make_int4(qb, nqr | (r << 8) | (mult << 16), my_a | (my_aq << 8) | (my_as << 16), 0);
    }
    if (a > 0) {
        if (lane == 0) mine[nt] = make_int2(ts, nw);
        ++nt;
    }
    // one slot reservation per chunk (a same-address atomic per TILE made the 8 000 tiles of a frame queue up at the L2)
    int base = 0;
    if (lane == 0 && nt > 0) base = atomicAdd(tile_count + g, nt);
    base = __shfl_sync(0xffffffffu, base, 0);
    __syncwarp();
    for (int i = lane; i < nt; i += 32) tiles[(size_t)g * win_cap + base + i] = mine[i];
}

// Row table of every tile: byte t of a tile = window (within the tile, 6 bits) | kind << 6 (1 = distinct key,
// 2 = real query, 0 = idle row).  One warp per tile, lane = window: every lane marks the rows of its window in a
// 128-byte line of shared memory, the warp writes the line out.  Made once per frame geometry with the plan; the tile
// kernel then finds the role of a row with one byte load instead of two binary searches over the window offsets
// (every instruction on the critical path of a tile costs ~20 clocks at 16 warps per SM).
__global__ void __launch_bounds__(256)
k_tca_rows(int win_cap, const int2 *__restrict__ tiles, const int *__restrict__ tile_count,
           const int4 *__restrict__ win_rec, unsigned char *__restrict__ tile_rows) {
    __shared__ __align__(16) unsigned char s_line[8][TCA_THREADS];
    const int lane = threadIdx.x & 31;
    unsigned char *line = s_line[threadIdx.x >> 5];
    const int warps = (gridDim.x * blockDim.x) >> 5;
    for (int g = 0; g < 2; ++g) {
        const int T = min(win_cap, __ldg(tile_count + g));
        for (int t = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; t < T; t += warps) {
            const int2 tl = __ldg(tiles + (size_t)g * win_cap + t);
            int4 rec = make_int4(0, 0, 0, 0);
            if (lane < tl.y) rec = __ldg(win_rec + (size_t)g * win_cap + tl.x + lane);
            const int last_z = __shfl_sync(0xffffffffu, rec.z, tl.y - 1), last_y = __shfl_sync(0xffffffffu, rec.y, tl.y - 1);
            const int nT = (last_z & 0xff) + ((last_y >> 8) & 0xff);      // key rows of the tile; the queries follow
            *(unsigned *)(line + 4 * lane) = 0u;
            __syncwarp();
            if (lane < tl.y) {
                const int k0 = rec.z & 0xff, nk = (rec.y >> 8) & 0xff;           // key rows of this lane's window
                const int q0 = nT + ((rec.z >> 8) & 0xff), nqw = rec.y & 0xff;    // its query rows
                for (int i = 0; i < nk; ++i) line[k0 + i] = (unsigned char)(lane | 0x40);
                for (int i = 0; i < nqw; ++i) line[q0 + i] = (unsigned char)(lane | 0x80);
            }
            __syncwarp();
            *(unsigned *)(tile_rows + ((size_t)g * win_cap + t) * TCA_THREADS + 4 * lane) = *(const unsigned *)(line + 4 * lane);
            __syncwarp();
        }
    }
}

// ------------------------------------------------------------------------------- the tile kernel

// shared-memory carve-up of k_tca_tile (bytes); NT = operand tiles per matrix: hi [, lo]
struct TcaSmem {
    int wpos, wkvq, wp, apos, ao, rows, hdr, misc, total;
    // nt = operand tiles per matrix (hi [, lo]); eb = bytes per operand element (4: TF32, 2: bf16)
    __host__ __device__ explicit TcaSmem(int nt, int eb = 4) {
        wpos = 0;                                  // [hi | lo] x [2 chunks][32][16 B]                    2 KB
        wkvq = wpos + 2048;                        // nt x [chunks][96][16 B]                    12 KB each (TF32)
        wp = wkvq + nt * 96 * 32 * eb;             // nt x [chunks][32][16 B]                     4 KB each (TF32)
        apos = wp + nt * 32 * 32 * eb;             // [hi | lo] x [2 chunks][128][16 B]; later the scores  8 KB
        ao = apos + 8192;                          // nt x [QMAX / 8][chunks][8][16 B]            6 KB each (TF32)
        rows = ao + nt * TCA_QMAX * TCA_SD * eb;   // gather staging, later V / Q rows [128][32]          16 KB
        hdr = rows + TCA_THREADS * TCA_SD * 4;     // (the MMA reads 128 rows of `ao`: it runs into `rows`)
        misc = hdr + 2 * TCA_HDR_BYTES;            // 2 x {window records [TW] int4, centres [TW] float4, row table [128]}
        total = misc + 128 * 4 + 16 + 16 + 128;    // biases, barriers
    }
};

// window records of a tile: offsets of the key rows / query rows are packed in rec.z (see k_tca_plan)
__device__ __forceinline__ int rec_koff(const int4 &r) { return r.z & 0xff; }
__device__ __forceinline__ int rec_qoff(const int4 &r) { return (r.z >> 8) & 0xff; }
__device__ __forceinline__ float ex2_fast(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

// Software pipeline over the tiles of a CTA.  Two things are pipelined ACROSS tiles:
//   * the loads of a tile form a chain of dependent global accesses (tile record -> window records -> row ids
//     -> coordinates + feature rows, ~3 L2 latencies); every link of tile i + 1 is issued one phase of tile i
//     ahead of its use and lands in shared memory (cp.async groups) or in a few registers:
//         window records of i + 1: cp.async at the top of tile i
//         row ids of i + 1:        loaded while MMA 2 of tile i runs
//         coordinates of i + 1:    loaded before the softmax of tile i
//         feature rows of i + 1:   cp.async into the staging area once tile i is done with it (after softmax / AV)
//   * the tensor-core round trips: the positional operand of tile i + 1 is built at the end of tile i, and MMA 1
//     of tile i + 1 is issued TOGETHER with MMA 3 (output projection) of tile i under one commit; tile i + 1
//     starts by reading both results.  A tile therefore costs two MMA round trips and five CTA barriers
//     (before: three and nine); the projected rows of tile i leave the CTA at the top of tile i + 1.
// TMEM columns: [0,32) pos -> A1 (MMA 1 of the NEXT tile may overwrite it as soon as MMA 2 is complete),
// [32,64) K, [64,96) V, [96,128) Q -- and, once Q has been unloaded, the output projection (MMA 3).
template <int HEADS, int TERMS>
__global__ void __launch_bounds__(TCA_THREADS, TERMS == 3 ? 3 : 4)
k_tca_tile(TcAttnParams P, int win_cap, const int2 *__restrict__ tiles, const int *__restrict__ tile_count,
           const int4 *__restrict__ win_rec, const float4 *__restrict__ win_ctr,
           const unsigned char *__restrict__ tile_rows, const float *__restrict__ xn,
           const float *__restrict__ xyz, const int *__restrict__ rep_row, const int *__restrict__ q_row,
           float *__restrict__ Pbuf) {
    constexpr int HD = TCA_SD / HEADS;
    constexpr int NT = TERMS == 3 || TERMS == 2 ? 2 : 1;   // operand tiles: hi [, lo] (3xTF32) / hi [, mid] (bf16x3), tc_common.cuh
    constexpr bool BF = TERMS == 0 || TERMS == 2;          // bf16 operands (kind::f16) for the K|V|Q and output projections
    constexpr int EB = BF ? 2 : 4;
    // TMEM: 128 columns; the 3xTF32 kernel takes 32 more (low half of the A operand) as a SECOND allocation:
    // 160 columns per CTA keep three CTAs on an SM, one 256-column allocation would allow two
    extern __shared__ __align__(128) char smem_raw[];
    const int tid = threadIdx.x, warp = tid >> 5;
    const int K = P.K, nq = P.nq;
#ifdef MSSVT_TRACE
    const long long t_entry = clock64();
    auto gtime = []() { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; };
    const unsigned long long gt0 = gtime();
#endif
    pdl_launch_dependents();
    pdl_wait();  // (the scale of this CTA, hence its weights, depends on the tile counts: nothing to do before)
#ifdef MSSVT_TRACE
    const long long t_pdl = clock64();
    const unsigned long long gt1 = gtime();
#endif

    // ---- every CTA works on one scale (head group) for its whole life; the CTAs of the two scales are
    //      interleaved over the grid in proportion to the tile counts
    const int T0 = __ldg(tile_count), T1 = __ldg(tile_count + 1), G = gridDim.x, b = blockIdx.x;
    int G0 = T0 == 0 ? 0 : T1 == 0 ? G : (int)(((long long)G * T0 + (T0 + T1) / 2) / (T0 + T1));
    if (T0 > 0 && T1 > 0) G0 = min(max(G0, 1), G - 1);
    const int c0b = (int)((long long)b * G0 / G), c0n = (int)((long long)(b + 1) * G0 / G);
    const int g = c0n > c0b ? 0 : 1;
    const int first = g ? b - c0b : c0b, stride = g ? G - G0 : G0, T = g ? T1 : T0;
    tiles += (size_t)g * win_cap;
    win_rec += (size_t)g * win_cap;
    tile_rows += (size_t)g * win_cap * TCA_THREADS;

    const TcaSmem L(NT, EB);
    char *sWpos = smem_raw + L.wpos, *sWkvq = smem_raw + L.wkvq, *sWp = smem_raw + L.wp;
    char *sApos = smem_raw + L.apos, *sAO = smem_raw + L.ao, *sRows = smem_raw + L.rows;
    float *sS = (float *)sApos;                             // [SBUD] scores: window-major, [key][query][head]
    char *sHdr = smem_raw + L.hdr;                          // 2 x {int4 rec[TW], float4 ctr[TW], uint8 rows[128]}
    float *sBias = (float *)(smem_raw + L.misc);            // [32] bq * scale, [32] bv, [32] bp (+ pad)
    uint64_t *sBar = (uint64_t *)(sBias + 128);             // [0] MMA completion, [1] weights landed (TMA bulk copies)
    uint32_t *sTmem = (uint32_t *)(sBar + 2);
    char *stg = sRows + warp * 4096;                        // this warp's staging area

    // weights of this scale: asynchronous 16-byte copies, waited for before the first MMA
    {
        const uint32_t d = smem_u32(sWpos);
        // packed [64][8] hi | lo: plane = chunk * 1024 + row * 16 (+ 2048 for lo); this scale's rows g*32 ..
        for (int i = tid; i < 128; i += TCA_THREADS) {   // i = (hl, chunk, row)
            const int hl = i >> 6, ch = (i >> 5) & 1, n = i & 31;
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d + (uint32_t)(hl * 1024 + ch * 512 + n * 16)),
                         "l"((const char *)P.wpos + hl * 2048 + ch * 1024 + (g * 32 + n) * 16) : "memory");
        }
    }
    const uint32_t bar_w = smem_u32(sBar + 1);
    if (tid == 0) {
        mbar_init(bar_w, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        // the group's projection matrices (packed once on the host side): two bulk copies of the TMA engine
        bulk_expect(bar_w, (uint32_t)(NT * (96 + 32) * 32 * EB));
        bulk_copy_g2s(P.wkvq[g], (uint32_t)(NT * 96 * 32 * EB), sWkvq, bar_w);
        bulk_copy_g2s(P.wp[g], (uint32_t)(NT * 32 * 32 * EB), sWp, bar_w);
    }
    if (tid < 32) {
        sBias[tid] = __ldg(P.bq[g] + tid) * P.scale;
        sBias[32 + tid] = __ldg(P.bkv[g] + 32 + tid);
        sBias[64 + tid] = __ldg(P.bp[g] + tid);
    }
    const uint32_t bar = smem_u32(sBar);
    if (tid == 0) {
        mbar_init(bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
#ifdef MSSVT_TRACE
    const long long t_stage = clock64();
#endif
    if (warp == 0) {
        tmem_alloc(smem_u32(sTmem), 128, TERMS != 3);
        if (TERMS == 3) tmem_alloc(smem_u32(sTmem + 1), 32);
    }
#ifdef MSSVT_TRACE
    const long long t_alloc = clock64();
#endif

    // header of tile number t (window records + centres + row table) -> buffer `buf`, asynchronously
    auto header_async = [&](int t, int2 tl, int buf) {
        const uint32_t d = smem_u32(sHdr + buf * TCA_HDR_BYTES);
        if (tid < tl.y) {
            asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(d + (uint32_t)tid * 16u), "l"(win_rec + tl.x + tid) : "memory");
            asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(d + (uint32_t)(TCA_TW * 16 + tid * 16)), "l"(win_ctr + tl.x + tid) : "memory");
        } else if (tid >= TCA_THREADS - 8) {
            const int c = tid - (TCA_THREADS - 8);
            asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(d + (uint32_t)(TCA_TW * 32 + c * 16)),
                         "l"(tile_rows + (size_t)t * TCA_THREADS + c * 16) : "memory");
        }
    };
    // role of this thread's row in a tile whose header is in shared memory: kind 1 = distinct key of window l,
    // 2 = real query of window l, 0 = idle; -> global feature row (issued load)
    auto role = [&](int2 tl, const char *hdr, int &kind, int &l, int &row) {
        const int4 *rec = (const int4 *)hdr;
        const unsigned code = ((const unsigned char *)(hdr + TCA_TW * 32))[tid];
        kind = (int)(code >> 6);
        l = (int)(code & 63u);
        row = 0;
        if (kind == 1) {
            row = __ldg(rep_row + (size_t)(tl.x + l) * 2 * K + g * K + (tid - rec_koff(rec[l])));
        } else if (kind == 2) {
            const int4 last = rec[tl.y - 1];
            const int nT = rec_koff(last) + ((last.y >> 8) & 0xff);
            row = __ldg(q_row + (size_t)(tl.x + l) * nq + (tid - nT - rec_qoff(rec[l])));
        }
    };
    const uint32_t row_off = (uint32_t)(tid >> 3) * 128u + (uint32_t)(tid & 7) * 16u;   // canonical 8-row groups
    // positional-embedding input of this thread's row, [rel | centre | 1 | 0] split hi / lo, canonical K-major
    // [128][8] -> the A operand of MMA 1.  hdr: header of the row's tile; (x, y, z): centre of the row's voxel
    auto pos_operand = [&](const char *hdr, int kind, int l, float x, float y, float z) {
        float4 p0 = make_float4(0.f, 0.f, 0.f, 0.f), p1 = p0;
        if (kind) {
            const int4 rec = ((const int4 *)hdr)[l];
            const float4 ctr = ((const float4 *)(hdr + TCA_TW * 16))[l];
            // the last distinct key of a window stands for all masked slots: relative offset zeroed
            const bool masked = kind == 1 && (rec.y >> 16) > 0 && tid - rec_koff(rec) == ((rec.y >> 8) & 0xff) - 1;
            p0 = masked ? make_float4(0.f, 0.f, 0.f, ctr.x)
                        : make_float4(__fsub_rn(x, ctr.x), __fsub_rn(y, ctr.y), __fsub_rn(z, ctr.z), ctr.x);
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

    int2 tl = first < T ? __ldg(tiles + first) : make_int2(0, 0);
    int2 tl_next = first + stride < T ? __ldg(tiles + first + stride) : make_int2(0, 0);
    if (first < T) header_async(first, tl, 0);
    stage_packed_wait();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    mbar_wait(bar_w, 0);
#ifdef MSSVT_TRACE
    if (tid == 0 && (blockIdx.x % 97) == 0)
        printf("tile kernel prologue (block %d): pdl wait %lld | stage issue %lld | tmem alloc %lld | copies + sync %lld clk\n",
               blockIdx.x, t_pdl - t_entry, t_stage - t_pdl, t_alloc - t_stage, clock64() - t_alloc);
#endif
    const uint32_t tm = *sTmem;
    const uint32_t lane_off = (uint32_t)(warp * 32) << 16;
    const uint32_t tm_a = tm, tm_k = tm + 32u, tm_v = tm + 64u, tm_q = tm + 96u;
    const uint32_t tm_alo = TERMS == 3 ? sTmem[1] : 0u;
    const uint32_t id_n32 = umma_idesc_tf32(128, 32), id_n96 = BF ? umma_idesc_bf16(128, 96) : umma_idesc_tf32(128, 96);
    const uint32_t id_p32 = BF ? umma_idesc_bf16(128, 32) : id_n32;   // output projection
    // operand regions as descriptor bases (tile / K chunk = one add on the address field)
    const UmmaDescBase dWpos = umma_desc_base(smem_u32(sWpos), 512, 128), dWkvq = umma_desc_base(smem_u32(sWkvq), 1536, 128),
                       dWp = umma_desc_base(smem_u32(sWp), 512, 128), dApos = umma_desc_base(smem_u32(sApos), 2048, 128),
                       dAO = umma_desc_base(smem_u32(sAO), 128, BF ? 512 : 1024);
    uint32_t phase = 0;

    // One commit for: MMA 3 of the tile that just finished its softmax / AV (output projection of its queries,
    // operand `ao` -> the Q columns, free since the unload) and MMA 1 of the tile whose positional operand was
    // just built (-> columns [0,32), always 3xTF32)
    auto issue_proj_and_pos = [&](bool with_proj, bool with_pos) {
        if (issuer_elected()) {
            tc_fence_after();
            if (with_proj) {
                if constexpr (BF) {
#pragma unroll
                    for (int k = 0; k < TCA_SD / 16; ++k) {
                        const uint64_t ah = umma_desc_at(dAO, (uint32_t)k * 256u), wh = umma_desc_at(dWp, (uint32_t)k * 1024u);
                        umma_bf16(tm_q, ah, wh, id_p32, k > 0 ? 1u : 0u);
                        if (TERMS == 2) {
                            umma_bf16(tm_q, umma_desc_at(dAO, TCA_QMAX * TCA_SD * 2 + (uint32_t)k * 256u), wh, id_p32, 1u);
                            umma_bf16(tm_q, ah, umma_desc_at(dWp, 32u * 32u * 2u + (uint32_t)k * 1024u), id_p32, 1u);
                        }
                    }
                } else {
#pragma unroll
                    for (int k = 0; k < TCA_SD / 8; ++k) {
                        const uint64_t ah = umma_desc_at(dAO, (uint32_t)k * 256u), wh = umma_desc_at(dWp, (uint32_t)k * 1024u);
                        umma_tf32(tm_q, ah, wh, id_n32, k > 0 ? 1u : 0u);
                        if (TERMS == 3) {
                            umma_tf32(tm_q, umma_desc_at(dAO, TCA_QMAX * TCA_SD * 4 + (uint32_t)k * 256u), wh, id_n32, 1u);
                            umma_tf32(tm_q, ah, umma_desc_at(dWp, 32u * 32u * 4u + (uint32_t)k * 1024u), id_n32, 1u);
                        }
                    }
                }
            }
            if (with_pos) {
                const uint64_t ah = umma_desc_at(dApos, 0u), al = umma_desc_at(dApos, 4096u);
                const uint64_t wh = umma_desc_at(dWpos, 0u), wl = umma_desc_at(dWpos, 1024u);
                umma_tf32(tm_a, ah, wh, id_n32, 0u);
                umma_tf32(tm_a, al, wh, id_n32, 1u);
                umma_tf32(tm_a, ah, wl, id_n32, 1u);
            }
            umma_commit(bar);
        }
        __syncwarp();
    };

    // pipeline prologue: role, coordinates, feature rows and positional operand of the first tile; MMA 1
    int kind = 0, l = 0, row = 0;
    if (first < T) {
        role(tl, sHdr, kind, l, row);
        float px = 0.f, py = 0.f, pz = 0.f;
        if (kind) { px = __ldg(xyz + 3 * (size_t)row); py = __ldg(xyz + 3 * (size_t)row + 1); pz = __ldg(xyz + 3 * (size_t)row + 2); }
        warp_rows_copy_async(stg, xn + g * TCA_SD, kind ? row : -1);
        cp_async_commit();
        pos_operand(sHdr, kind, l, px, py, pz);
        fence_async_smem();
        __syncthreads();
        issue_proj_and_pos(false, true);
    }

#ifdef MSSVT_TRACE
    long long tr[16] = {0};
#endif
    int buf = 0, prev_nQ = 0, prev_q0 = 0;
    for (int t = first; t < T; t += stride, buf ^= 1) {
        KTRACE(0);
        const char *hdr = sHdr + buf * TCA_HDR_BYTES, *hdr_next = sHdr + (buf ^ 1) * TCA_HDR_BYTES;
        const int4 *sRec = (const int4 *)hdr;
        const unsigned char *sRowTab = (const unsigned char *)(hdr + TCA_TW * 32);
        const int nwin = tl.y;
        const bool more = t + stride < T;
        if (more) header_async(t + stride, tl_next, buf ^ 1);   // the next tile's window records: on their way now
        cp_async_commit();
        const int2 tl_after = t + 2 * stride < T ? __ldg(tiles + t + 2 * stride) : make_int2(0, 0);
        // the warp's 32 feature rows (copied by its own lanes at the end of the previous tile) have landed
        float4 xv[TCA_SD / 4];
        cp_async_wait_group<1>();
        __syncwarp();
        {
            const int lane = tid & 31;
#pragma unroll
            for (int q = 0; q < 8; ++q)
                xv[q] = kind ? *(const float4 *)(stg + lane * 128 + ((q ^ (lane & 7)) << 4)) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
        KTRACE(1);
        mbar_wait(bar, phase);   // MMA 1 of this tile (+ MMA 3 of the previous one)
        phase ^= 1u;
        tc_fence_after();
        KTRACE(2);
        // ---- 5 (of the previous tile). projected rows of its queries (contiguous compact ids) -> (#queries, 64) array
        if (warp * 32 < prev_nQ) {   // (warp-uniform)
            float d[TCA_SD];
            tmem_ld32(tm_q + lane_off, d);
            if (tid < prev_nQ) {
                float4 *dst = (float4 *)(Pbuf + (size_t)(prev_q0 + tid) * TCA_C + g * TCA_SD);
#pragma unroll
                for (int c4 = 0; c4 < TCA_SD / 4; ++c4)
                    dst[c4] = make_float4(d[4 * c4] + sBias[64 + 4 * c4], d[4 * c4 + 1] + sBias[64 + 4 * c4 + 1],
                                          d[4 * c4 + 2] + sBias[64 + 4 * c4 + 2], d[4 * c4 + 3] + sBias[64 + 4 * c4 + 3]);
            }
        }
        {
            // ---- 2. A1 = xn slice + relu(pos), TF32, written back over the accumulator (A operand from TMEM)
            float d[TCA_SD];
            tmem_ld32(tm_a + lane_off, d);
            if constexpr (BF) {
                // packed bf16 pairs: channel k in column k / 2 (16 columns, in place over the accumulator)
                uint32_t w[TCA_SD / 2], wm[TCA_SD / 2];
#pragma unroll
                for (int c4 = 0; c4 < TCA_SD / 4; ++c4) {
                    const float a0 = xv[c4].x + fmaxf(d[4 * c4], 0.f), a1 = xv[c4].y + fmaxf(d[4 * c4 + 1], 0.f);
                    const float a2 = xv[c4].z + fmaxf(d[4 * c4 + 2], 0.f), a3 = xv[c4].w + fmaxf(d[4 * c4 + 3], 0.f);
                    if (TERMS == 2) {
                        split_bf16x2(a0, a1, w[2 * c4], wm[2 * c4]);
                        split_bf16x2(a2, a3, w[2 * c4 + 1], wm[2 * c4 + 1]);
                    } else {
                        w[2 * c4] = pack_bf16x2(a0, a1);
                        w[2 * c4 + 1] = pack_bf16x2(a2, a3);
                    }
                }
                tmem_st16(tm_a + lane_off, w);
                if (TERMS == 2) tmem_st16(tm_a + 16u + lane_off, wm);   // (mid parts: columns 16..31 of the same accumulator)
            } else if (TERMS == 3) {
                float lo[TCA_SD];
#pragma unroll
                for (int c4 = 0; c4 < TCA_SD / 4; ++c4) {
                    split_tf32(xv[c4].x + fmaxf(d[4 * c4], 0.f), d[4 * c4], lo[4 * c4]);
                    split_tf32(xv[c4].y + fmaxf(d[4 * c4 + 1], 0.f), d[4 * c4 + 1], lo[4 * c4 + 1]);
                    split_tf32(xv[c4].z + fmaxf(d[4 * c4 + 2], 0.f), d[4 * c4 + 2], lo[4 * c4 + 2]);
                    split_tf32(xv[c4].w + fmaxf(d[4 * c4 + 3], 0.f), d[4 * c4 + 3], lo[4 * c4 + 3]);
                }
                tmem_st32(tm_alo + lane_off, lo);
            } else {
#pragma unroll
                for (int c4 = 0; c4 < TCA_SD / 4; ++c4) {
                    d[4 * c4] = to_tf32(xv[c4].x + fmaxf(d[4 * c4], 0.f));
                    d[4 * c4 + 1] = to_tf32(xv[c4].y + fmaxf(d[4 * c4 + 1], 0.f));
                    d[4 * c4 + 2] = to_tf32(xv[c4].z + fmaxf(d[4 * c4 + 2], 0.f));
                    d[4 * c4 + 3] = to_tf32(xv[c4].w + fmaxf(d[4 * c4 + 3], 0.f));
                }
            }
            if constexpr (!BF) tmem_st32(tm_a + lane_off, d);
            tmem_st_wait();
        }
        cp_async_wait_group<0>();   // this thread's part of the next header
        tc_fence_before();
        __syncthreads();            // (also: every thread's header copies are complete -> the next header is visible)
        KTRACE(3);
        // ---- [K | V | Q] = A1 [Wk; Wv; scale Wq]^T: K in TMEM columns 32..63, V in 64..95, Q in 96..127
        if (issuer_elected()) {
            tc_fence_after();
            if constexpr (BF) {
#pragma unroll
                for (int k = 0; k < TCA_SD / 16; ++k) {   // K = 16 per MMA: 8 packed A columns, two 16-byte B chunks
                    const uint64_t wh = umma_desc_at(dWkvq, (uint32_t)k * 2u * 1536u);
                    umma_bf16_ts(tm_k, tm_a + (uint32_t)k * 8u, wh, id_n96, k > 0 ? 1u : 0u);
                    if (TERMS == 2) {
                        umma_bf16_ts(tm_k, tm_a + 16u + (uint32_t)k * 8u, wh, id_n96, 1u);
                        umma_bf16_ts(tm_k, tm_a + (uint32_t)k * 8u, umma_desc_at(dWkvq, 96u * 32u * 2u + (uint32_t)k * 2u * 1536u), id_n96, 1u);
                    }
                }
            } else
#pragma unroll
            for (int k = 0; k < TCA_SD / 8; ++k) {
                const uint64_t wh = umma_desc_at(dWkvq, (uint32_t)k * 2u * 1536u);
                umma_tf32_ts(tm_k, tm_a + (uint32_t)k * 8u, wh, id_n96, k > 0 ? 1u : 0u);
                if (TERMS == 3) {
                    umma_tf32_ts(tm_k, tm_alo + (uint32_t)k * 8u, wh, id_n96, 1u);
                    umma_tf32_ts(tm_k, tm_a + (uint32_t)k * 8u, umma_desc_at(dWkvq, 96u * 32u * 4u + (uint32_t)k * 2u * 1536u), id_n96, 1u);
                }
            }
            umma_commit(bar);
        }
        __syncwarp();
        KTRACE(4);
        // while the MMA runs: role and feature row id of this thread in the NEXT tile, this tile's bookkeeping
        int n_kind = 0, n_l = 0, n_row = 0;
        if (more) role(tl_next, hdr_next, n_kind, n_l, n_row);
        const int4 last = sRec[nwin - 1];
        const int nT = rec_koff(last) + ((last.y >> 8) & 0xff), nQ = rec_qoff(last) + (last.y & 0xff);
        const bool is_key = kind == 1, is_query = kind == 2;
        int j = 0, nqr = 0, soff = 0, qrow0 = 0;
        bool masked = false;
        if (is_key) {
            const int4 rec = sRec[l];
            j = tid - rec_koff(rec);
            nqr = rec.y & 0xff; soff = rec.z >> 16; qrow0 = nT + rec_qoff(rec);
            masked = (rec.y >> 16) > 0 && j == ((rec.y >> 8) & 0xff) - 1;  // last distinct key stands for all masked slots
        }
        const int q0_tile = sRec[0].x;
        KTRACE(5);
        mbar_wait(bar, phase);
        phase ^= 1u;
        tc_fence_after();
        KTRACE(6);
        // ---- rows back from TMEM: V of the key rows and Q (+ bias) of the query rows -> shared memory (16-byte
        //      chunks XOR-swizzled by the row), K of this thread's key stays in registers.  Biases: q.(k + bk)
        //      shifts every score of a query by the same q.bk, which the softmax cancels, so bk is dropped;
        //      sum_i p_i (v_i + bv) = sum_i p_i v_i + bv, so bv is added once per output below.
        float kk[TCA_SD];
        {
            const int w_lo = warp * 32, w_hi = w_lo + 32;
            const bool warp_keys = w_lo < nT, warp_queries = w_hi > nT && w_lo < nT + nQ;   // (warp-uniform)
            float4 *my = (float4 *)(sRows + tid * 128);
            const int sw = tid & 7;
            if (warp_keys) {
                float vv[TCA_SD];
                tmem_ld32(tm_v + lane_off, vv);
                if (is_key) {
#pragma unroll
                    for (int c4 = 0; c4 < TCA_SD / 4; ++c4)
                        my[c4 ^ sw] = make_float4(vv[4 * c4], vv[4 * c4 + 1], vv[4 * c4 + 2], vv[4 * c4 + 3]);
                }
            }
            if (warp_queries) {
                float qq[TCA_SD];
                tmem_ld32(tm_q + lane_off, qq);
                if (is_query) {
#pragma unroll
                    for (int c4 = 0; c4 < TCA_SD / 4; ++c4)
                        my[c4 ^ sw] = make_float4(qq[4 * c4] + sBias[4 * c4], qq[4 * c4 + 1] + sBias[4 * c4 + 1],
                                                  qq[4 * c4 + 2] + sBias[4 * c4 + 2], qq[4 * c4 + 3] + sBias[4 * c4 + 3]);
                }
            }
            if (warp_keys) tmem_ld32(tm_k + lane_off, kk);
        }
        tc_fence_before();
        __syncthreads();
        KTRACE(7);
        // ---- 3. scores of this thread's key against the queries of its window
        if (is_key) {
            const float bias = masked ? -100.0f : 0.f;  // additive mask of the reference
            float *srow = sS + soff + j * nqr * HEADS;
            for (int s = 0; s < nqr; ++s) {
                const int qr = qrow0 + s;
                const float4 *qv = (const float4 *)(sRows + qr * 128);
                const int qsw = qr & 7;
#pragma unroll
                for (int h = 0; h < HEADS; ++h) {
                    float a0 = 0.f, a1 = 0.f;
#pragma unroll
                    for (int d4 = 0; d4 < HD / 4; ++d4) {
                        const float4 q4 = qv[(h * (HD / 4) + d4) ^ qsw];
                        a0 = fmaf(q4.x, kk[h * HD + 4 * d4], a0); a1 = fmaf(q4.y, kk[h * HD + 4 * d4 + 1], a1);
                        a0 = fmaf(q4.z, kk[h * HD + 4 * d4 + 2], a0); a1 = fmaf(q4.w, kk[h * HD + 4 * d4 + 3], a1);
                    }
                    srow[s * HEADS + h] = a0 + a1 + bias;
                }
            }
        }
        __syncthreads();
        KTRACE(8);
        // the next tile's coordinates: on their way during the softmax (the row id was loaded a phase ago; its use --
        // address arithmetic included -- is kept down here so that its load is not waited for up there)
        float nx = 0.f, ny = 0.f, nz = 0.f;
        asm volatile("" : "+r"(n_row));
        if (n_kind) { nx = __ldg(xyz + 3 * (size_t)n_row); ny = __ldg(xyz + 3 * (size_t)n_row + 1); nz = __ldg(xyz + 3 * (size_t)n_row + 2); }
        // ---- 4. softmax over the window's distinct keys and AV, thread = (query, head, half of the head);
        //      head outputs (+ bv) -> rows of the output-projection operand (canonical K-major, 8-row groups of
        //      8 chunks: byte = (q / 8) * 1024 + chunk * 128 + (q % 8) * 16)
        for (int e = tid; e < nQ * HEADS * 2; e += TCA_THREADS) {
            constexpr int DPT = HD / 2;   // channels per thread (16 / 8 / 4)
            const int half = e & 1, qh = e >> 1;
            const int h = qh % HEADS, qs = qh / HEADS;
            const int4 rec = sRec[sRowTab[nT + qs] & 63];
            const int s = qs - rec_qoff(rec);
            const int wq = rec.y & 0xff, r = (rec.y >> 8) & 0xff, mult = rec.y >> 16;
            const float *sc = sS + (rec.z >> 16) + s * HEADS + h;  // + key * wq * HEADS
            const int step = wq * HEADS;
            float mx = -3.0e38f;
            for (int k = 0; k < r; ++k) mx = fmaxf(mx, sc[k * step]);
            mx *= 1.4426950408889634f;
            float den = 0.f, acc[DPT];
#pragma unroll
            for (int d = 0; d < DPT; ++d) acc[d] = 0.f;
            const int ch0 = (h * HD + half * DPT) >> 2, k0 = rec_koff(rec);   // first 16-byte chunk, first key row
            for (int k = 0; k < r; ++k) {
                float wgt = ex2_fast(fmaf(sc[k * step], 1.4426950408889634f, -mx));
                if (k == r - 1 && mult > 0) wgt *= (float)mult;  // the masked key counts once per masked slot
                den += wgt;
                const int kr = k0 + k;
                const float4 *vp = (const float4 *)(sRows + kr * 128);
#pragma unroll
                for (int d4 = 0; d4 < DPT / 4; ++d4) {
                    const float4 v = vp[(ch0 + d4) ^ (kr & 7)];
                    acc[4 * d4] = fmaf(wgt, v.x, acc[4 * d4]); acc[4 * d4 + 1] = fmaf(wgt, v.y, acc[4 * d4 + 1]);
                    acc[4 * d4 + 2] = fmaf(wgt, v.z, acc[4 * d4 + 2]); acc[4 * d4 + 3] = fmaf(wgt, v.w, acc[4 * d4 + 3]);
                }
            }
            const float inv = 1.0f / den;
            const float *bv = sBias + 32 + 4 * ch0;
            if constexpr (BF) {
                // bf16 operand rows: 8-row groups of 4 chunks of 8 elements, byte = (q / 8) * 512 + chunk * 128 + (q % 8) * 16
                const int c0 = 4 * ch0;
                char *dst = sAO + (qs >> 3) * 512 + (qs & 7) * 16 + (c0 >> 3) * 128 + (c0 & 7) * 2;
#pragma unroll
                for (int d = 0; d < DPT; d += 2) {
                    const int c = c0 + d;      // (pairs never straddle a chunk: c0 and DPT are multiples of 4)
                    char *at = dst + ((c >> 3) - (c0 >> 3)) * 128 + ((c & 7) - (c0 & 7)) * 2;
                    const float o0 = fmaf(acc[d], inv, bv[d]), o1 = fmaf(acc[d + 1], inv, bv[d + 1]);
                    if (TERMS == 2) {
                        uint32_t hi, mid;
                        split_bf16x2(o0, o1, hi, mid);
                        *(uint32_t *)at = hi;
                        *(uint32_t *)(at + TCA_QMAX * TCA_SD * 2) = mid;   // second tile: the mid parts
                    } else {
                        *(uint32_t *)at = pack_bf16x2(o0, o1);
                    }
                }
                continue;
            }
            char *dst = sAO + (qs >> 3) * 1024 + (qs & 7) * 16 + ch0 * 128;
#pragma unroll
            for (int d4 = 0; d4 < DPT / 4; ++d4) {
                float4 hi, lo;
                split_tf32(make_float4(fmaf(acc[4 * d4], inv, bv[4 * d4]), fmaf(acc[4 * d4 + 1], inv, bv[4 * d4 + 1]),
                                       fmaf(acc[4 * d4 + 2], inv, bv[4 * d4 + 2]), fmaf(acc[4 * d4 + 3], inv, bv[4 * d4 + 3])), hi, lo);
                *(float4 *)(dst + d4 * 128) = hi;
                if (TERMS == 3) *(float4 *)(dst + TCA_QMAX * TCA_SD * 4 + d4 * 128) = lo;
            }
        }
        __syncthreads();   // scores and V / Q rows are consumed: their memory takes the next tile's operands
        KTRACE(9);
        if (more) {
            warp_rows_copy_async(stg, xn + g * TCA_SD, n_kind ? n_row : -1);
            pos_operand(hdr_next, n_kind, n_l, nx, ny, nz);
        }
        cp_async_commit();
        fence_async_smem();
        __syncthreads();
        KTRACE(10);
        issue_proj_and_pos(true, more);
        KTRACE(11);
        prev_nQ = nQ; prev_q0 = q0_tile;
        kind = n_kind; l = n_l; row = n_row;
        tl = tl_next; tl_next = tl_after;
    }
    // the projected rows of the CTA's last tile
    if (first < T) {
        mbar_wait(bar, phase);
        tc_fence_after();
        if (warp * 32 < prev_nQ) {
            float d[TCA_SD];
            tmem_ld32(tm_q + lane_off, d);
            if (tid < prev_nQ) {
                float4 *dst = (float4 *)(Pbuf + (size_t)(prev_q0 + tid) * TCA_C + g * TCA_SD);
#pragma unroll
                for (int c4 = 0; c4 < TCA_SD / 4; ++c4)
                    dst[c4] = make_float4(d[4 * c4] + sBias[64 + 4 * c4], d[4 * c4 + 1] + sBias[64 + 4 * c4 + 1],
                                          d[4 * c4 + 2] + sBias[64 + 4 * c4 + 2], d[4 * c4 + 3] + sBias[64 + 4 * c4 + 3]);
            }
        }
    }
#ifdef MSSVT_TRACE
    if (tid == 0 && blockIdx.x == gridDim.x - 1 && tr[11])
        printf("tile g%d: rows in %lld | wait A %lld | proj(prev)+A1+sync %lld | mma2 issue %lld | roles(next) %lld | wait %lld | unload+sync %lld | scores+sync %lld | softmax+AV+sync %lld | gather+pos(next)+sync %lld | issue %lld | total %lld clk\n", g,
               tr[1] - tr[0], tr[2] - tr[1], tr[3] - tr[2], tr[4] - tr[3], tr[5] - tr[4], tr[6] - tr[5], tr[7] - tr[6], tr[8] - tr[7], tr[9] - tr[8], tr[10] - tr[9], tr[11] - tr[10], tr[11] - tr[0]);
#endif
#ifdef MSSVT_TRACE
    if (tid == 0 && (blockIdx.x % 59) == 0)
        printf("tile cta %3d (sm %2u, scale %d, %d tiles): entry %llu | pdl wait +%llu | exit +%llu ns\n", blockIdx.x,
               [] { unsigned s; asm("mov.u32 %0, %%smid;" : "=r"(s)); return s; }(), g, first < T ? (T - first + stride - 1) / stride : 0,
               gt0 % 100000000ull, gt1 - gt0, gtime() - gt0);
#endif
    stage_packed_wait();
    tc_fence_before();
    __syncthreads();
    if (warp == 0) {
        tmem_dealloc(tm, 128);
        if (TERMS == 3) tmem_dealloc(tm_alo, 32);
    }
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

}  // namespace mssvt

using namespace mssvt;

extern "C" {

/* Tile plan of the tensor-core window attention: a function of the geometry (meta, q_base, win_list)
 * only, made once and shared by every block / launch over the same window lists.
 * tiles (2, win_capacity, 2) int, tile_count (2) int, win_rec (2, win_capacity, 4) int, win_ctr
 * (win_capacity, 4) float, tile_rows (2, win_capacity, 128) bytes: opaque to the caller. */
int mssvt_attention_tiles(int heads_per_group, int nq, int key_num_sample, int win_capacity,
                          const int *win_count_total, const int *win_list, const int *meta, const int *q_base,
                          const float *win_cell, const float *range_min, int *tiles, int *tile_count,
                          int *win_rec, float *win_ctr, unsigned char *tile_rows, void *stream) {
    if (heads_per_group <= 0 || nq <= 0 || nq > 32 || key_num_sample <= 0 || key_num_sample > 63 ||
        nq * (key_num_sample + 1) * heads_per_group > TCA_SBUD || win_capacity < 0)
        return MSSVT_ERR_INVALID;
    if (!win_count_total || !win_list || !meta || !q_base || !win_cell || !range_min || !tiles || !tile_count ||
        !win_rec || !win_ctr || !tile_rows)
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
    ++g_launches;
    k_tca_rows<<<MSSVT_NUM_SMS * 2, 256, 0, s>>>(win_capacity, (const int2 *)tiles, tile_count, (const int4 *)win_rec,
                                                 tile_rows);
    return check_launch();
}

/* Tensor-core window attention of a two-window block (see the header of this file).  Weights packed by
 * mssvt_pack_operand_tf32: wpos_packed = [pos_w | pos_b | 0] as a [64][8] matrix, ALWAYS packed with terms = 3;
 * wkvq0 / wkvq1 = per head group the [96][32] matrix [Wk; Wv; scale * Wq]; wp0 / wp1 = the group's [32][32]
 * output projection (packed with `terms`).  bq / bkv / bp: the nn.Linear biases ([32], [64], [32] per group).
 * rep_row / meta: compact key lists of mssvt_block_geometry; q_base: mssvt_exclusive_scan of meta[:, 0]
 * (win_capacity + 1 ints); tiles .. win_ctr: mssvt_attention_tiles.  scratch: 3 * num_voxels * 64 floats; the
 * projected row of every real query lands in scratch + 2 * num_voxels * 64.  merged may be NULL when interp is
 * set: the blend of the projected rows is then left to mssvt_ffn_tc in mode 2.
 * Returns MSSVT_ERR_INVALID for shapes outside C = 64 / 2 x 32 channels / nq <= 32 / K <= 63 /
 * cap1 <= 128 (callers then use mssvt_block_attention). */
int mssvt_block_attention_tc(int C, int heads_per_group, int nq, int key_num_sample, int cap1, int interp,
                             int terms, float scale, const float *wpos_packed, const float *wkvq0,
                             const float *wkvq1, const float *wp0, const float *wp1, const float *bq0,
                             const float *bq1, const float *bkv0, const float *bkv1, const float *bp0,
                             const float *bp1, int win_capacity, const int *win_count_total,
                             const float *xn, const float *xyz, const int *q_row,
                             const int *rep_row, const int *meta, const int *q_base, const int *q_src,
                             const int *vox_slot, const unsigned char *nn_idx,
                             const float *nn_w, const int *tiles, const int *tile_count, const int *win_rec,
                             const float *win_ctr, const unsigned char *tile_rows, int num_voxels, float *scratch,
                             float *merged, void *stream) {
    if (C != 64 || (heads_per_group != 1 && heads_per_group != 2 && heads_per_group != 4) || nq <= 0 || nq > 32 ||
        key_num_sample <= 0 || key_num_sample > 63 || cap1 <= 0 || cap1 > 128 || win_capacity < 0 || num_voxels < 0 ||
        nq * (key_num_sample + 1) * heads_per_group > TCA_SBUD || (terms != 0 && terms != 1 && terms != 2 && terms != 3))
        return MSSVT_ERR_INVALID;
    if (win_capacity == 0) return MSSVT_OK;
    if (!wpos_packed || !wkvq0 || !wkvq1 || !wp0 || !wp1 || !bq0 || !bq1 || !bkv0 || !bkv1 || !bp0 || !bp1 ||
        !win_count_total || !xn || !xyz || !q_row || !rep_row || !meta || !q_base || !q_src || !tiles ||
        !tile_count || !win_rec || !win_ctr || !tile_rows || !scratch || (!merged && !interp))
        return MSSVT_ERR_INVALID;
    if (interp && (!vox_slot || !nn_idx || !nn_w)) return MSSVT_ERR_INVALID;
    TcAttnParams P;
    P.nq = nq; P.K = key_num_sample; P.cap1 = cap1; P.interp = interp ? 1 : 0; P.heads = heads_per_group;
    P.scale = scale;
    P.wpos = wpos_packed;
    P.wkvq[0] = wkvq0; P.wkvq[1] = wkvq1; P.wp[0] = wp0; P.wp[1] = wp1;
    P.bq[0] = bq0; P.bq[1] = bq1; P.bkv[0] = bkv0; P.bkv[1] = bkv1; P.bp[0] = bp0; P.bp[1] = bp1;
    const size_t smem = (size_t)TcaSmem(terms == 3 || terms == 2 ? 2 : 1, terms == 0 || terms == 2 ? 2 : 4).total;
    float *Pbuf = scratch + 2 * (size_t)num_voxels * 64;
    cudaStream_t s = (cudaStream_t)stream;
    const int wide = MSSVT_NUM_SMS * 8;  // grid-stride kernels: 8 CTAs of 256 threads per SM

    int per_sm = (int)(227 * 1024 / (smem + 1024));
    const int tmem_limit = terms == 3 ? 3 : 4;   // 160 / 128 TMEM columns each
    per_sm = per_sm > tmem_limit ? tmem_limit : per_sm < 1 ? 1 : per_sm;
#ifdef TCA_MAX_PER_SM   // (occupancy experiments)
    if (per_sm > TCA_MAX_PER_SM) per_sm = TCA_MAX_PER_SM;
#endif
    const int grid = MSSVT_NUM_SMS * per_sm;
    ++g_launches;
#define TCA_LAUNCH(H, T)                                                                                   \
    cudaFuncSetAttribute(k_tca_tile<H, T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);         \
    launch_pdl(k_tca_tile<H, T>, dim3(grid), dim3(TCA_THREADS), smem, s, P, win_capacity, (const int2 *)tiles, \
               tile_count, (const int4 *)win_rec, (const float4 *)win_ctr, tile_rows, xn, xyz, rep_row, q_row, Pbuf)
    if (terms == 2) {
        if (heads_per_group == 1) { TCA_LAUNCH(1, 2); }
        else if (heads_per_group == 2) { TCA_LAUNCH(2, 2); }
        else { TCA_LAUNCH(4, 2); }
    } else if (terms == 0) {
        if (heads_per_group == 1) { TCA_LAUNCH(1, 0); }
        else if (heads_per_group == 2) { TCA_LAUNCH(2, 0); }
        else { TCA_LAUNCH(4, 0); }
    } else if (terms == 3) {
        if (heads_per_group == 1) { TCA_LAUNCH(1, 3); }
        else if (heads_per_group == 2) { TCA_LAUNCH(2, 3); }
        else { TCA_LAUNCH(4, 3); }
    } else {
        if (heads_per_group == 1) { TCA_LAUNCH(1, 1); }
        else if (heads_per_group == 2) { TCA_LAUNCH(2, 1); }
        else { TCA_LAUNCH(4, 1); }
    }
#undef TCA_LAUNCH
    if (!merged) return check_launch();  // interpolation + merge left to mssvt_ffn_tc (mode 2)
    ++g_launches;
    launch_pdl(k_tca_merge, dim3(wide), dim3(256), 0, s, P, win_capacity, win_count_total, num_voxels, meta, q_base,
               q_src, q_row, vox_slot, nn_idx, nn_w, Pbuf, merged);
    return check_launch();
}

}  // extern "C"
