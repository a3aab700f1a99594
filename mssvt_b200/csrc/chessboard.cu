// chessboard.cu -- chessboard sampling of query/key voxels per window (sm_100a).
//
// Op-level entry points (reference contract, padded outputs):
//   mssvt_gather_two_window  <- gather_two_window_voxels_with_hash_kernel
//                               pcdet/ops/mssvt/src/ms_sparse_attention_gpu.cu:193-381
//   mssvt_gather_one_window  <- gather_one_window_voxels_with_hash_kernel   ...:383-458
// Fused entry point used by the backbone module:
//   mssvt_block_geometry     <- everything in MixedScaleSparseTransformerBlock.forward that depends
//                               on coordinates only: the two-window gather, both FPS passes
//                               (sampling_gpu.cu:100-260), the float round trip that picks the key
//                               indices (mssvt_backbone.py:247-258, quirk Q1), three_nn
//                               (interpolate_gpu.cu:16-59) and the interpolation weights
//                               (mssvt_backbone.py:304-307).  Lists never leave shared memory;
//                               only ~0.7 KB per window of compact maps is written instead of
//                               the 2.7 KB of padded lists + FPS temporaries of the reference.
//
// One warp per window.  The reference walks the offset tables with one thread per window and T
// dependent probes; here the 32 lanes probe 32 offsets at once and a ballot + prefix popcount
// appends the hits in table order, so every list is identical to the sequential result.
#include "common.cuh"

namespace mssvt {

#define GEO_WARPS 8

struct GatherShape {
    int x_max, y_max, z_max, x_ws, y_ws, z_ws, hash_size;
    int seg[4];   // sizes of the offset tables odd, even, win1-rest, win2-rest (0 if absent)
    int cap[4];   // caps of the lists   odd, even, win1, win2 (0 if absent)
};

// membership of a hit from table segment c in the four lists (bit L = list L)
__device__ __forceinline__ unsigned list_membership(int c) {
    // odd -> {odd, win1, win2}; even -> {even, win1, win2}; win1 -> {win1, win2}; win2 -> {win2}
    return c == 0 ? 0xDu : c == 1 ? 0xEu : c == 2 ? 0xCu : 0x8u;
}

__device__ __forceinline__ int pack_off(int x, int y, int z) {
    return (x & 0xff) | ((y & 0xff) << 8) | ((z & 0xff) << 16);
}
__device__ __forceinline__ int off_x(int p) { return (int)(signed char)(p & 0xff); }
__device__ __forceinline__ int off_y(int p) { return (int)(signed char)((p >> 8) & 0xff); }
__device__ __forceinline__ int off_z(int p) { return (int)(signed char)((p >> 16) & 0xff); }

// Probe every offset for one window; lists land in shared memory in table order.
// s_tab: total x 3 ints (concatenated tables); s_ind/s_off: per-warp list storage, list L at
// list_at[L].  Returns the (capped) list lengths in cnt[].
template <typename IDX>
__device__ __forceinline__ void probe_window(const GatherShape &g, int4 win, const int *s_tab,
                                             const IDX &index, int *s_ind, int *s_off,
                                             const int list_at[4], int cnt[4], int &cx, int &cy,
                                             int &cz) {
    const int lane = threadIdx.x & 31;
    cz = win.y * g.z_ws + g.z_ws / 2;
    cy = win.z * g.y_ws + g.y_ws / 2;
    cx = win.w * g.x_ws + g.x_ws / 2;
    const int e0 = g.seg[0], e1 = e0 + g.seg[1], e2 = e1 + g.seg[2], total = e2 + g.seg[3];
    cnt[0] = cnt[1] = cnt[2] = cnt[3] = 0;
    for (int base = 0; base < total; base += 32) {
        int q = base + lane;
        int v = MSSVT_EMPTY, packed = 0;
        unsigned member = 0;
        if (q < total) {
            int ox = s_tab[3 * q], oy = s_tab[3 * q + 1], oz = s_tab[3 * q + 2];
            int sx = cx + ox, sy = cy + oy, sz = cz + oz;
            if (!(sx >= g.x_max || sx < 0 || sy >= g.y_max || sy < 0 || sz >= g.z_max || sz < 0)) {
                v = index.find(win.x, sx, sy, sz);
                packed = pack_off(ox, oy, oz);
                member = list_membership(q < e0 ? 0 : q < e1 ? 1 : q < e2 ? 2 : 3);
            }
        }
        bool hit = v != MSSVT_EMPTY;
#pragma unroll
        for (int L = 0; L < 4; ++L) {
            if (g.cap[L] == 0) continue;
            // (table segments that feed list L: odd [0, e0), even [e0, e1), win1 [0, e2), win2 everything; a pass
            //  of 32 offsets that does not touch them appends nothing to L -- three of four passes for win1 and its parts)
            const int src_lo = L == 1 ? e0 : 0, src_hi = L == 0 ? e0 : L == 1 ? e1 : L == 2 ? e2 : total;
            if (base >= src_hi || base + 32 <= src_lo) continue;
            bool mine = hit && ((member >> L) & 1u);
            unsigned m = __ballot_sync(0xffffffffu, mine);
            int pos = cnt[L] + __popc(m & lanemask_lt());
            if (mine && pos < g.cap[L]) {
                s_ind[list_at[L] + pos] = v;
                s_off[list_at[L] + pos] = packed;
            }
            cnt[L] += __popc(m);
        }
    }
#pragma unroll
    for (int L = 0; L < 4; ++L) cnt[L] = min(cnt[L], g.cap[L]);
    __syncwarp();
}

__device__ __forceinline__ void load_tables(const GatherShape &g, const int *const q_tab[4],
                                            int *s_tab) {
    int at = 0;
    for (int c = 0; c < 4; ++c) {
        for (int i = threadIdx.x; i < g.seg[c] * 3; i += blockDim.x) s_tab[at + i] = q_tab[c][i];
        at += g.seg[c] * 3;
    }
}

struct GatherOut {
    int *ind[4];    // (W, cap[L])     -1 padded
    int *coord[4];  // (W, cap[L], 3)   0 padded
};

struct TablePtrs {
    const int *p[4];
};

__global__ void __launch_bounds__(GEO_WARPS * 32)
k_gather_lists(GatherShape g, TablePtrs tabs, int num_wins, const int4 *__restrict__ win_list,
               const int2 *__restrict__ table, GatherOut out) {
    extern __shared__ int smem[];
    const int total = g.seg[0] + g.seg[1] + g.seg[2] + g.seg[3];
    const int caps = g.cap[0] + g.cap[1] + g.cap[2] + g.cap[3];
    int *s_tab = smem;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    int *s_ind = smem + total * 3 + warp * caps * 2;
    int *s_off = s_ind + caps;
    load_tables(g, tabs.p, s_tab);
    __syncthreads();
    int list_at[4] = {0, g.cap[0], g.cap[0] + g.cap[1], g.cap[0] + g.cap[1] + g.cap[2]};
    const HashIdx index = {table, g.hash_size, g.y_max, g.z_max};
    for (int w = blockIdx.x * GEO_WARPS + warp; w < num_wins; w += gridDim.x * GEO_WARPS) {
        int cnt[4], cx, cy, cz;
        probe_window(g, __ldg(win_list + w), s_tab, index, s_ind, s_off, list_at, cnt, cx, cy, cz);
#pragma unroll
        for (int L = 0; L < 4; ++L) {
            const int cap = g.cap[L];
            if (cap == 0) continue;
            int *ind = out.ind[L] + (size_t)w * cap;
            int *coord = out.coord[L] + (size_t)w * cap * 3;
            for (int i = lane; i < cap; i += 32) ind[i] = i < cnt[L] ? s_ind[list_at[L] + i] : -1;
            for (int j = lane; j < cap * 3; j += 32) {
                int i = j / 3, a = j - 3 * i;
                int p = i < cnt[L] ? s_off[list_at[L] + i] : 0;
                coord[j] = a == 0 ? off_x(p) : a == 1 ? off_y(p) : off_z(p);
            }
        }
        __syncwarp();
    }
}

// ------------------------------------------------------------------------------- fused geometry

struct GeoParams {
    GatherShape g;
    int K;          // key_num_sample per scale
    int pattern;    // cbs_pattern: 0 even, 1 odd, 2 win1 list as queries
    int interp;     // use_feature_interpolation
    int log2b[2];   // log2 of the reference's FPS block size for n = cap1 / cap2
    float cell[3];  // voxel size  (x, y, z)
    float lo[3];    // point-cloud range minimum
};

struct GeoOut {
    int *q_row;       // (W, nq)      global feature row of each query slot, -1 pad
    int *win1_row;    // (W, cap1)    global row of each win1 voxel, -1 pad
    int *k_row;       // (W, 2K)      global row of each key slot (Q1: pad picks -> sample row 0)
    unsigned char *k_mask;  // (W, 2K) 1 = masked key (fps index 0 beyond slot 0)
    unsigned char *nn_idx;  // (W, cap1, 3) three-NN query slots        (interp only)
    float *nn_w;            // (W, cap1, 3) normalised 1/d weights      (interp only)
    unsigned char *covered; // (N) 1 = row is written by the merge (appears in a win1 list)
    int *fps_idx;     // (W, 2K) raw FPS picks, optional tap (may be null)
    int *counts;      // (W, 4) list lengths odd, even, win1, win2, optional tap (may be null)
    int *rep_row;     // (W, 2K) per scale: rows of the DISTINCT keys (unmasked slots in order, then
                      //         one entry standing for every masked slot), optional (may be null)
    int *meta;        // (W, 4) {#real queries, #win1 voxels, nrep0 | nmask0 << 8, nrep1 | nmask1 << 8}
    int *vox_slot;    // (N) w * cap1 + i of the win1 slot holding the voxel, -1 if none (optional)
    int *odd_row;     // (W, cap[0]) / (W, cap[1]): global rows of the odd / even chessboard lists, -1 pad (optional:
    int *even_row;    //  what mssvt_block_queries needs to serve another cbs_pattern from this geometry)
};

__device__ __forceinline__ unsigned bitrev(unsigned v, int bits) { return bits ? __brev(v) >> (32 - bits) : 0u; }

// Farthest point sampling over one shared-memory list of n slots (cnt real, rest padding at
// offset (0,0,0)), reproducing farthest_point_sampling_kernel<B> (sampling_gpu.cu:100-216)
// including its tie order: among the slots at maximal distance the block reduction keeps the
// one minimising (bit_reverse(k mod B), k).  Offsets are small integers, so the
// fp32 distances of the reference are exact and integer arithmetic gives identical results.
// key = dist << 20 | (B-1-bitrev(k mod B)) << 10 | (1023-k)  ->  one redux.sync per pick.
//
// All padded slots sit at the same point, so they always carry the same distance and only the one with
// the best tie order can ever win: they are folded into ONE candidate (dist d_pad, key tie_pad), and the
// per-pick scan covers the cnt real slots only.  With cnt <= 32 (the usual case) a lane keeps its slot's
// offset, tie key and running distance in registers.
__device__ __forceinline__ unsigned fps_tie(int k, int B, int log2b) {
    return ((unsigned)(B - 1) - bitrev((unsigned)k & (B - 1), log2b)) << 10 | (unsigned)(1023 - k);
}

__device__ __forceinline__ void fps_list(const int *s_off, int cnt, int n, int log2b, int K,
                                         int *s_min, int *s_pick) {
    const int lane = threadIdx.x & 31;
    const int B = 1 << log2b;
    unsigned tie_pad = 0;  // best tie key among the padded slots [cnt, n); 0 = there is none
    for (int k = cnt + lane; k < n; k += 32) tie_pad = max(tie_pad, fps_tie(k, B, log2b));
    tie_pad = __reduce_max_sync(0xffffffffu, tie_pad);
    int d_pad = 0xfff;  // (keys hold 12 bits of distance: squared offsets of a few cells)
    if (lane == 0) s_pick[0] = 0;
    int old = 0, j = 1;
    if (cnt <= 32) {
        const int p = lane < cnt ? s_off[lane] : 0;
        const int px = off_x(p), py = off_y(p), pz = off_z(p);
        const unsigned tie = fps_tie(lane, B, log2b);
        int d = 0xfff;
        for (; j < K; ++j) {
            const int po = __shfl_sync(0xffffffffu, p, old & 31);  // (old >= cnt: a padded slot, offset 0)
            const int ox = old < cnt ? off_x(po) : 0, oy = old < cnt ? off_y(po) : 0, oz = old < cnt ? off_z(po) : 0;
            const int dx = px - ox, dy = py - oy, dz = pz - oz;
            d = min(d, dx * dx + dy * dy + dz * dz);
            d_pad = min(d_pad, ox * ox + oy * oy + oz * oz);
            unsigned best = lane < cnt ? ((unsigned)d << 20) | tie : 0u;
            if (tie_pad) best = max(best, ((unsigned)d_pad << 20) | tie_pad);
            best = __reduce_max_sync(0xffffffffu, best);
            if ((best >> 20) == 0) break;  // every slot coincides with a pick: index 0 from here on
            old = 1023 - (int)(best & 1023u);
            if (lane == 0) s_pick[j] = old;
        }
    } else {
        for (int k = lane; k < cnt; k += 32) s_min[k] = 0xfff;
        __syncwarp();
        for (; j < K; ++j) {
            const int po = old < cnt ? s_off[old] : 0;
            const int ox = off_x(po), oy = off_y(po), oz = off_z(po);
            d_pad = min(d_pad, ox * ox + oy * oy + oz * oz);
            unsigned best = tie_pad ? ((unsigned)d_pad << 20) | tie_pad : 0u;
            for (int k = lane; k < cnt; k += 32) {
                const int p = s_off[k];
                const int dx = off_x(p) - ox, dy = off_y(p) - oy, dz = off_z(p) - oz;
                const int d = min(s_min[k], dx * dx + dy * dy + dz * dz);
                s_min[k] = d;
                best = max(best, ((unsigned)d << 20) | fps_tie(k, B, log2b));
            }
            best = __reduce_max_sync(0xffffffffu, best);
            if ((best >> 20) == 0) break;
            old = 1023 - (int)(best & 1023u);
            if (lane == 0) s_pick[j] = old;
        }
    }
    for (int r = j + lane; r < K; r += 32) s_pick[r] = 0;
    __syncwarp();
}

// three nearest of nq known points (padding sits at the origin, quirk Q4) for one win1 voxel, and the normalised
// 1/d weights: three_nn_kernel_fast (interpolate_gpu.cu:16-59) + mssvt_backbone.py:305-307
// nreal: the first nreal known points are real, the rest is padding.  All padded points coincide, and ties keep the
// earlier index (strict <), so at most the first three of them can ever enter the result: the scan stops there.
__device__ __forceinline__ void three_nn_slot(float ux, float uy, float uz, const float *s_known, int nq, int nreal,
                                              unsigned char *oi, float *ow) {
    const float INF = __int_as_float(0x7f800000);
    float b1 = INF, b2 = INF, b3 = INF;
    int i1 = 0, i2 = 0, i3 = 0;
    const int kmax = min(nq, nreal + 3);
    for (int k = 0; k < kmax; ++k) {
        float dx = __fsub_rn(ux, s_known[3 * k]);
        float dy = __fsub_rn(uy, s_known[3 * k + 1]);
        float dz = __fsub_rn(uz, s_known[3 * k + 2]);
        // nvcc's contraction of the reference expression (SASS of oracle/_ref):
        float d = __fmaf_rn(dz, dz, __fmaf_rn(dx, dx, __fmul_rn(dy, dy)));
        if (d < b1) { b3 = b2; i3 = i2; b2 = b1; i2 = i1; b1 = d; i1 = k; }
        else if (d < b2) { b3 = b2; i3 = i2; b2 = d; i2 = k; }
        else if (d < b3) { b3 = d; i3 = k; }
    }
    // dist = sqrt(d2); w = 1 / clamp(dist, 1e-10); w /= sum(w)  (mssvt_backbone.py:305-307)
    float w1 = __fdiv_rn(1.0f, fmaxf(__fsqrt_rn(b1), 1e-10f));
    float w2 = __fdiv_rn(1.0f, fmaxf(__fsqrt_rn(b2), 1e-10f));
    float w3 = __fdiv_rn(1.0f, fmaxf(__fsqrt_rn(b3), 1e-10f));
    float sum = __fadd_rn(__fadd_rn(w1, w2), w3);
    oi[0] = (unsigned char)i1; oi[1] = (unsigned char)i2; oi[2] = (unsigned char)i3;
    ow[0] = __fdiv_rn(w1, sum); ow[1] = __fdiv_rn(w2, sum); ow[2] = __fdiv_rn(w3, sum);
}

__global__ void __launch_bounds__(GEO_WARPS * 32)
k_block_geometry(GeoParams P, TablePtrs tabs, int win_cap, const int *__restrict__ win_count_total,
                 const int4 *__restrict__ win_list, GridIdx index,
                 const int *__restrict__ v_start, GeoOut out) {
    extern __shared__ int smem[];
    const GatherShape &g = P.g;
    const int total = g.seg[0] + g.seg[1] + g.seg[2] + g.seg[3];
    const int caps = g.cap[0] + g.cap[1] + g.cap[2] + g.cap[3];
    const int K = P.K, cap1 = g.cap[2], cap2 = g.cap[3];
    const int nmax = max(cap1, cap2);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // per-warp: lists (ind + off), fps min-dist (nmax), picks (2K), known points (3 * nq floats)
    const int nq = P.pattern == 0 ? g.cap[1] : P.pattern == 1 ? g.cap[0] : cap1;
    const int per_warp = caps * 2 + nmax + 2 * K + 3 * nq;
    int *s_tab = smem;
    int *s_ind = smem + total * 3 + warp * per_warp;
    int *s_off = s_ind + caps;
    int *s_min = s_off + caps;
    int *s_pick = s_min + nmax;
    float *s_known = (float *)(s_pick + 2 * K);
    load_tables(g, tabs.p, s_tab);
    __syncthreads();
    const int list_at[4] = {0, g.cap[0], g.cap[0] + g.cap[1], g.cap[0] + g.cap[1] + g.cap[2]};
    const int qL = P.pattern == 0 ? 1 : P.pattern == 1 ? 0 : 2;
    const int num_wins = min(win_cap, __ldg(win_count_total));  // never past the caller's allocations

    for (int w = blockIdx.x * GEO_WARPS + warp; w < num_wins; w += gridDim.x * GEO_WARPS) {
        int cnt[4], cx, cy, cz;
        const int4 win = __ldg(win_list + w);
        probe_window(g, win, s_tab, index, s_ind, s_off, list_at, cnt, cx, cy, cz);
        const int row0 = __ldg(v_start + win.x);
        if (out.counts && lane < 4)
            out.counts[4 * w + lane] = lane == 0 ? cnt[0] : lane == 1 ? cnt[1] : lane == 2 ? cnt[2] : cnt[3];
        if (out.meta && lane == 0) {
            out.meta[4 * (size_t)w + 0] = cnt[qL];
            out.meta[4 * (size_t)w + 1] = cnt[2];
        }

        // queries and win1 rows (global rows)
        for (int i = lane; i < nq; i += 32)
            out.q_row[(size_t)w * nq + i] = i < cnt[qL] ? row0 + s_ind[list_at[qL] + i] : -1;
        if (out.odd_row)
            for (int i = lane; i < g.cap[0]; i += 32)
                out.odd_row[(size_t)w * g.cap[0] + i] = i < cnt[0] ? row0 + s_ind[list_at[0] + i] : -1;
        if (out.even_row)
            for (int i = lane; i < g.cap[1]; i += 32)
                out.even_row[(size_t)w * g.cap[1] + i] = i < cnt[1] ? row0 + s_ind[list_at[1] + i] : -1;
        for (int i = lane; i < cap1; i += 32) {
            int r = i < cnt[2] ? row0 + s_ind[list_at[2] + i] : -1;
            out.win1_row[(size_t)w * cap1 + i] = r;
            if (P.interp && r >= 0) out.covered[r] = 1;
            if (out.vox_slot && r >= 0) out.vox_slot[r] = w * cap1 + i;
        }
        if (!P.interp)  // without interpolation the merge writes the query rows instead
            for (int i = lane; i < cnt[qL]; i += 32) out.covered[row0 + s_ind[list_at[qL] + i]] = 1;

        // FPS on the win1 list then on the win2 list; key slot j of scale s is slot s*K + j
#pragma unroll
        for (int s = 0; s < 2; ++s) {
            const int L = 2 + s, n = g.cap[L];
            if (!out.k_row) {
                // Only the SET of distinct keys is asked for (rep_row / meta: what the tensor-core attention reads).  FPS
                // keeps picking while some slot is at a positive distance from the picks, i.e. until every distinct
                // position has been picked: the real voxels (different cells) and, if the list has padded slots and no
                // real voxel sits at offset 0, the one position all padded slots share (a key aliased to the sample's
                // first voxel, quirk Q1).  With at most K distinct positions the set is known without running FPS --
                // always for the 3^3 lists, for nearly all 5^3 lists; the K - #distinct remaining picks are index 0 =
                // the masked key.  Keys are listed in list order instead of pick order (a softmax over a set).
                const int c = cnt[L];
                int ic = -1;   // the real slot at offset 0 (at most one), -1: none
                for (int i0 = 0; i0 < c; i0 += 32) {
                    const int i = i0 + lane;
                    const unsigned m = __ballot_sync(0xffffffffu, i < c && s_off[list_at[L] + i] == 0);
                    if (m) ic = i0 + __ffs(m) - 1;
                }
                // The padded position and a real voxel at offset 0 coincide: they always carry the same distance, the
                // tie order of the reference's block reduction decides which of the two is picked (the other one is at
                // distance 0 from then on).  Slot 0 is the first pick by definition.
                bool alias = false;   // the padded position is a key (aliased to the sample's first voxel)
                int excl = -1;        // real slot that is never picked
                if (c < n) {
                    if (ic < 0) alias = true;
                    else if (ic > 0) {
                        const int B = 1 << P.log2b[s];
                        unsigned tie_pad = 0;
                        for (int k = c + lane; k < n; k += 32) tie_pad = max(tie_pad, fps_tie(k, B, P.log2b[s]));
                        tie_pad = __reduce_max_sync(0xffffffffu, tie_pad);
                        if (tie_pad > fps_tie(ic, B, P.log2b[s])) { alias = true; excl = ic; }
                    }
                }
                const int nreal = c - (excl >= 0 ? 1 : 0), ndist = nreal + (alias ? 1 : 0);
                if (ndist <= K) {
                    int *rr = out.rep_row + (size_t)w * 2 * K + s * K;
                    for (int i = lane; i < c; i += 32)
                        if (i != excl) rr[i - (excl >= 0 && i > excl ? 1 : 0)] = row0 + s_ind[list_at[L] + i];
                    if (lane == 0) {
                        const int nmask = K - ndist;
                        if (alias) rr[nreal] = row0;
                        if (nmask > 0) rr[ndist] = row0 + s_ind[list_at[L]];
                        out.meta[4 * (size_t)w + 2 + s] = (ndist + (nmask > 0 ? 1 : 0)) | (nmask << 8);
                    }
                    __syncwarp();
                    continue;
                }
            }
            fps_list(s_off + list_at[L], cnt[L], n, P.log2b[s], K, s_min, s_pick + s * K);
            for (int j = lane; j < K; j += 32) {
                int f = s_pick[s * K + j];
                int v = f < cnt[L] ? s_ind[list_at[L] + f] : -1;
                // (ind.float() gathered at f, + 0.1).int(): -1 -> 0, anything >= 0 unchanged
                if (out.k_row) {
                    out.k_row[(size_t)w * 2 * K + s * K + j] = row0 + max(v, 0);
                    out.k_mask[(size_t)w * 2 * K + s * K + j] = (j > 0 && f == 0) ? 1 : 0;
                }
                if (out.fps_idx) out.fps_idx[(size_t)w * 2 * K + s * K + j] = f;
            }
            if (out.rep_row) {
                // distinct keys of this scale: all masked slots hold the same key (first voxel of
                // the list, offset zeroed), so one entry with a multiplicity stands for them
                int nrep = 0, nmask = 0, masked_row = -1;
                for (int j0 = 0; j0 < K; j0 += 32) {
                    const int j = j0 + lane;
                    const int f = j < K ? s_pick[s * K + j] : 1;
                    const int v = (j < K && f < cnt[L]) ? s_ind[list_at[L] + f] : -1;
                    const int row = row0 + max(v, 0);
                    const bool masked = j < K && j > 0 && f == 0, live = j < K && !masked;
                    const unsigned mm = __ballot_sync(0xffffffffu, masked), lm = __ballot_sync(0xffffffffu, live);
                    if (live) out.rep_row[(size_t)w * 2 * K + s * K + nrep + __popc(lm & lanemask_lt())] = row;
                    if (mm && masked_row < 0) masked_row = __shfl_sync(0xffffffffu, row, __ffs(mm) - 1);
                    nrep += __popc(lm);
                    nmask += __popc(mm);
                }
                if (nmask > 0) {
                    if (lane == 0) out.rep_row[(size_t)w * 2 * K + s * K + nrep] = masked_row;
                    nrep += 1;
                }
                if (lane == 0) out.meta[4 * (size_t)w + 2 + s] = nrep | (nmask << 8);
            }
            __syncwarp();
        }

        if (P.interp) {
            // known = query slots in world coordinates (padding sits at the origin, Q4)
            for (int k = lane; k < nq; k += 32) {
                float x = 0.f, y = 0.f, z = 0.f;
                if (k < cnt[qL]) {
                    int p = s_off[list_at[qL] + k];
                    x = world_coord(cx + off_x(p), P.cell[0], P.lo[0]);
                    y = world_coord(cy + off_y(p), P.cell[1], P.lo[1]);
                    z = world_coord(cz + off_z(p), P.cell[2], P.lo[2]);
                }
                s_known[3 * k] = x; s_known[3 * k + 1] = y; s_known[3 * k + 2] = z;
            }
            __syncwarp();
            for (int i = lane; i < cap1; i += 32) {
                unsigned char *oi = out.nn_idx + ((size_t)w * cap1 + i) * 3;
                float *ow = out.nn_w + ((size_t)w * cap1 + i) * 3;
                if (i >= cnt[2]) {
                    oi[0] = oi[1] = oi[2] = 0;
                    ow[0] = ow[1] = ow[2] = 0.f;
                    continue;
                }
                int p = s_off[list_at[2] + i];
                three_nn_slot(world_coord(cx + off_x(p), P.cell[0], P.lo[0]), world_coord(cy + off_y(p), P.cell[1], P.lo[1]),
                              world_coord(cz + off_z(p), P.cell[2], P.lo[2]), s_known, nq, cnt[qL], oi, ow);
            }
        }
        __syncwarp();
    }
}

// The pattern-dependent part of the block geometry, from the lists another pattern's mssvt_block_geometry call
// left behind: query rows, #real queries (meta[0]; the key entries of meta are copied), three-NN + weights.
// One warp per window.  src_row: (W, nq) the chessboard list that serves as the query set (odd / even / win1
// rows).  Voxel centres come from the xyz array, which holds the very world_coord() values k_block_geometry
// derives from the cell indices, so indices and weights are bit-identical to a direct call.
__global__ void __launch_bounds__(GEO_WARPS * 32)
k_block_queries(int nq, int cap1, int interp, int win_cap, const int *__restrict__ win_count_total,
                const int *__restrict__ src_row, const int *__restrict__ win1_row, const int4 *__restrict__ meta_in,
                const float *__restrict__ xyz, int *__restrict__ q_row, int4 *__restrict__ meta_out,
                unsigned char *__restrict__ nn_idx, float *__restrict__ nn_w) {
    extern __shared__ int smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float *s_known = (float *)smem + warp * 3 * nq;
    const int num_wins = min(win_cap, __ldg(win_count_total));
    for (int w = blockIdx.x * GEO_WARPS + warp; w < num_wins; w += gridDim.x * GEO_WARPS) {
        int nqr = 0;
        for (int k0 = 0; k0 < nq; k0 += 32) {
            const int k = k0 + lane;
            const int row = k < nq ? __ldg(src_row + (size_t)w * nq + k) : -1;
            if (k < nq) {
                q_row[(size_t)w * nq + k] = row;
                float x = 0.f, y = 0.f, z = 0.f;       // padded queries sit at the world origin (Q4)
                if (row >= 0) { x = __ldg(xyz + 3 * (size_t)row); y = __ldg(xyz + 3 * (size_t)row + 1); z = __ldg(xyz + 3 * (size_t)row + 2); }
                s_known[3 * k] = x; s_known[3 * k + 1] = y; s_known[3 * k + 2] = z;
            }
            nqr += __popc(__ballot_sync(0xffffffffu, row >= 0));
        }
        if (lane == 0) {
            int4 m = __ldg(meta_in + w);
            m.x = nqr;
            meta_out[w] = m;
        }
        __syncwarp();
        if (interp) {
            for (int i = lane; i < cap1; i += 32) {
                unsigned char *oi = nn_idx + ((size_t)w * cap1 + i) * 3;
                float *ow = nn_w + ((size_t)w * cap1 + i) * 3;
                const int row = __ldg(win1_row + (size_t)w * cap1 + i);
                if (row < 0) {
                    oi[0] = oi[1] = oi[2] = 0;
                    ow[0] = ow[1] = ow[2] = 0.f;
                    continue;
                }
                three_nn_slot(__ldg(xyz + 3 * (size_t)row), __ldg(xyz + 3 * (size_t)row + 1), __ldg(xyz + 3 * (size_t)row + 2),
                              s_known, nq, nqr, oi, ow);
            }
        }
        __syncwarp();
    }
}

// one-window gather for the compress block, sync-free: global rows only, -1 padded
__global__ void __launch_bounds__(GEO_WARPS * 32)
k_window_rows(GatherShape g, TablePtrs tabs, int win_cap, const int *__restrict__ win_count_total,
              const int4 *__restrict__ win_list, GridIdx index,
              const int *__restrict__ v_start, int *__restrict__ k_row) {
    extern __shared__ int smem[];
    const int total = g.seg[2], cap = g.cap[2];
    int *s_tab = smem;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    int *s_ind = smem + total * 3 + warp * cap * 2;
    int *s_off = s_ind + cap;
    load_tables(g, tabs.p, s_tab);
    __syncthreads();
    const int list_at[4] = {0, 0, 0, 0};
    const int num_wins = min(win_cap, __ldg(win_count_total));  // never past the caller's allocations
    for (int w = blockIdx.x * GEO_WARPS + warp; w < num_wins; w += gridDim.x * GEO_WARPS) {
        int cnt[4], cx, cy, cz;
        const int4 win = __ldg(win_list + w);
        probe_window(g, win, s_tab, index, s_ind, s_off, list_at, cnt, cx, cy, cz);
        const int row0 = __ldg(v_start + win.x);
        for (int i = lane; i < cap; i += 32) k_row[(size_t)w * cap + i] = i < cnt[2] ? row0 + s_ind[i] : -1;
        __syncwarp();
    }
}

// q_src[q_base[w] + s] = w * nq + s for every real query slot: inverse of the compact query numbering
__global__ void k_query_src(int win_cap, const int *__restrict__ win_count_total, int nq,
                            const int *__restrict__ meta, const int *__restrict__ q_base,
                            int *__restrict__ q_src) {
    const int num_wins = min(win_cap, __ldg(win_count_total));
    const long long total = (long long)num_wins * nq;
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total;
         e += (long long)gridDim.x * blockDim.x) {
        const int w = (int)(e / nq), sl = (int)(e - (long long)w * nq);
        if (sl < __ldg(meta + 4 * (size_t)w)) q_src[__ldg(q_base + w) + sl] = (int)e;
    }
}

}  // namespace mssvt

using namespace mssvt;

static bool shape_ok(const GatherShape &g) {
    if (g.hash_size <= 0 || g.x_ws <= 0 || g.y_ws <= 0 || g.z_ws <= 0) return false;
    for (int c = 0; c < 4; ++c)
        if (g.seg[c] < 0 || g.cap[c] < 0 || g.seg[c] > 4096 || g.cap[c] > 1024) return false;
    return true;
}

// the reference computes its FPS block size on the host as 2^(int)(log(n)/log(2)) capped at
// 1024 (pointnet2_batch/src/cuda_utils.h:10-14); same expression, same libm.
#include <cmath>
static int fps_log2_block(int n) {
    int p = (int)(std::log((double)n) / std::log(2.0));
    if (p > 10) p = 10;
    if (p < 0) p = 0;
    return p;
}

extern "C" {

int mssvt_fps_log2_block(int n) { return n > 0 ? fps_log2_block(n) : 0; }

int mssvt_gather_two_window(int x_max, int y_max, int z_max, int x_ws, int y_ws, int z_ws,
                            int max_odd, int max_even, int max_win1, int max_win2, int num_wins,
                            int hash_size, int num_odd, int num_even, int num_win1, int num_win2,
                            int *ind_odd, int *ind_even, int *ind_win1, int *ind_win2,
                            int *coord_odd, int *coord_even, int *coord_win1, int *coord_win2,
                            const int *q_odd, const int *q_even, const int *q_win1,
                            const int *q_win2, const int *win_indices, const int *table,
                            void *stream) {
    GatherShape g = {x_max, y_max, z_max, x_ws, y_ws, z_ws, hash_size,
                     {num_odd, num_even, num_win1, num_win2},
                     {max_odd, max_even, max_win1, max_win2}};
    if (!shape_ok(g) || num_wins < 0) return MSSVT_ERR_INVALID;
    if (num_wins == 0) return MSSVT_OK;
    if (!win_indices || !table) return MSSVT_ERR_INVALID;
    TablePtrs tabs = {{q_odd, q_even, q_win1, q_win2}};
    GatherOut out = {{ind_odd, ind_even, ind_win1, ind_win2},
                     {coord_odd, coord_even, coord_win1, coord_win2}};
    for (int c = 0; c < 4; ++c) {
        if (g.seg[c] && !tabs.p[c]) return MSSVT_ERR_INVALID;
        if (g.cap[c] && (!out.ind[c] || !out.coord[c])) return MSSVT_ERR_INVALID;
    }
    int total = num_odd + num_even + num_win1 + num_win2;
    int caps = max_odd + max_even + max_win1 + max_win2;
    size_t smem = (size_t)(total * 3 + GEO_WARPS * caps * 2) * sizeof(int);
    if (smem > 200 * 1024) return MSSVT_ERR_INVALID;
    cudaFuncSetAttribute(k_gather_lists, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    int grid = persistent_grid(num_wins, GEO_WARPS, 8);
    ++g_launches;
    k_gather_lists<<<grid, GEO_WARPS * 32, smem, (cudaStream_t)stream>>>(
        g, tabs, num_wins, (const int4 *)win_indices, (const int2 *)table, out);
    return check_launch();
}

int mssvt_gather_one_window(int x_max, int y_max, int z_max, int x_ws, int y_ws, int z_ws,
                            int max_win1, int num_wins, int hash_size, int num_win1, int *ind_win1,
                            int *coord_win1, const int *q_win1, const int *win_indices,
                            const int *table, void *stream) {
    // one table walked into one list: the "win1-rest" segment feeding only the win1 list
    return mssvt_gather_two_window(x_max, y_max, z_max, x_ws, y_ws, z_ws, 0, 0, max_win1, 0,
                                   num_wins, hash_size, 0, 0, num_win1, 0, nullptr, nullptr,
                                   ind_win1, nullptr, nullptr, nullptr, coord_win1, nullptr, nullptr,
                                   nullptr, q_win1, nullptr, win_indices, table, stream);
}

// Fused per-window geometry of a two-window block.  All sizes that depend on the data stay on
// the device (win_count_total), so there is no host synchronisation.
int mssvt_block_geometry(int x_max, int y_max, int z_max, int x_ws, int y_ws, int z_ws,
                         int num_odd, int num_even, int num_win1, int num_win2,
                         int max_win1, int max_win2, int key_num_sample, int cbs_pattern,
                         int use_interp, const float *voxel_size, const float *range_min,
                         const int *q_odd, const int *q_even, const int *q_win1, const int *q_win2,
                         int win_capacity, const int *win_count_total, const int *win_list,
                         const int *grid_cells, const int *grid_vals, const int *v_start,
                         int num_voxels, int *q_row,
                         int *win1_row, int *k_row, unsigned char *k_mask, unsigned char *nn_idx,
                         float *nn_w, unsigned char *covered, int *fps_idx_tap, int *counts_tap,
                         int *rep_row, int *meta, int *vox_slot, int *odd_row, int *even_row, void *stream) {
    GeoParams P;
    P.g = {x_max, y_max, z_max, x_ws, y_ws, z_ws, 1,
           {num_odd, num_even, num_win1, num_win2}, {num_odd, num_even, max_win1, max_win2}};
    if (!shape_ok(P.g) || key_num_sample <= 0 || key_num_sample > 256 || cbs_pattern < 0 ||
        cbs_pattern > 2 || max_win1 <= 0 || max_win2 <= 0 || win_capacity < 0)
        return MSSVT_ERR_INVALID;
    if (!voxel_size || !range_min || !win_count_total || !win_list || !grid_cells || !grid_vals ||
        !v_start || !q_row || !win1_row || !covered)
        return MSSVT_ERR_INVALID;
    // k_row / k_mask are optional together: without them only the distinct-key form (rep_row / meta) is produced
    if ((k_row == nullptr) != (k_mask == nullptr) || (!k_row && !rep_row) || (!k_row && fps_idx_tap)) return MSSVT_ERR_INVALID;
    if (use_interp && (!nn_idx || !nn_w)) return MSSVT_ERR_INVALID;
    int nq = cbs_pattern == 0 ? num_even : cbs_pattern == 1 ? num_odd : max_win1;
    if (nq <= 0 || nq > 255) return MSSVT_ERR_INVALID;
    // integer FPS keys hold the squared distance in 12 bits: 3 * (extent)^2 must stay below 4096
    int ext = x_ws > y_ws ? x_ws : y_ws;
    ext = ext > z_ws ? ext : z_ws;
    // win2 extent is not passed; bound it by the offset range that fits an int8 pack
    if (num_win2 > 0 && max_win2 > 1024) return MSSVT_ERR_INVALID;
    P.K = key_num_sample;
    P.pattern = cbs_pattern;
    P.interp = use_interp ? 1 : 0;
    P.log2b[0] = fps_log2_block(max_win1);
    P.log2b[1] = fps_log2_block(max_win2);
    for (int i = 0; i < 3; ++i) { P.cell[i] = voxel_size[i]; P.lo[i] = range_min[i]; }
    cudaStream_t s = (cudaStream_t)stream;
    cudaError_t e = cudaMemsetAsync(covered, 0, (size_t)num_voxels, s);
    if (e != cudaSuccess) { g_last_cuda_error = (int)e; return MSSVT_ERR_LAUNCH; }
    if (win_capacity == 0) return MSSVT_OK;
    TablePtrs tabs = {{q_odd, q_even, q_win1, q_win2}};
    if ((rep_row == nullptr) != (meta == nullptr)) return MSSVT_ERR_INVALID;
    GeoOut out = {q_row, win1_row, k_row, k_mask, nn_idx, nn_w, covered, fps_idx_tap, counts_tap, rep_row, meta,
                  vox_slot, odd_row, even_row};
    if (vox_slot) {
        e = cudaMemsetAsync(vox_slot, 0xff, (size_t)num_voxels * sizeof(int), s);
        if (e != cudaSuccess) { g_last_cuda_error = (int)e; return MSSVT_ERR_LAUNCH; }
    }
    int total = num_odd + num_even + num_win1 + num_win2;
    int caps = num_odd + num_even + max_win1 + max_win2;
    int nmax = max_win1 > max_win2 ? max_win1 : max_win2;
    size_t per_warp = (size_t)caps * 2 + nmax + 2 * key_num_sample + 3 * nq;
    size_t smem = ((size_t)total * 3 + GEO_WARPS * per_warp) * sizeof(int);
    if (smem > 200 * 1024) return MSSVT_ERR_INVALID;
    cudaFuncSetAttribute(k_block_geometry, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    int grid = persistent_grid(win_capacity, GEO_WARPS, 6);
    ++g_launches;
    const int zw = (z_max + 31) / 32;
    GridIdx index = {(const int2 *)grid_cells, grid_vals, y_max, zw, (long long)x_max * y_max * zw};
    k_block_geometry<<<grid, GEO_WARPS * 32, smem, s>>>(P, tabs, win_capacity, win_count_total,
                                                       (const int4 *)win_list, index, v_start, out);
    (void)ext;
    return check_launch();
}

/* Pattern-dependent part of the block geometry for ANOTHER cbs_pattern, from the outputs of one
 * mssvt_block_geometry call over the same windows (mssvt_backbone.py:220-234, 300-307): src_row = the list that
 * serves as the query set -- odd_row (pattern 1), even_row (pattern 0) or win1_row (pattern 2), nq entries per
 * window; meta_in = that call's meta.  Writes q_row (cap, nq), meta_out (cap, 4) (entry 0 = #real queries of THIS
 * pattern, the rest copied) and, with use_interp, nn_idx / nn_w (cap, max_win1, 3).  Results are bit-identical
 * to a direct mssvt_block_geometry call with that pattern; the chessboard probes and both FPS passes are not
 * repeated. */
int mssvt_block_queries(int nq, int max_win1, int use_interp, int win_capacity, const int *win_count_total,
                        const int *src_row, const int *win1_row, const int *meta_in, const float *xyz, int *q_row,
                        int *meta_out, unsigned char *nn_idx, float *nn_w, void *stream) {
    if (nq <= 0 || nq > 255 || max_win1 <= 0 || win_capacity < 0) return MSSVT_ERR_INVALID;
    if (win_capacity == 0) return MSSVT_OK;
    if (!win_count_total || !src_row || !win1_row || !meta_in || !xyz || !q_row || !meta_out) return MSSVT_ERR_INVALID;
    if (use_interp && (!nn_idx || !nn_w)) return MSSVT_ERR_INVALID;
    const size_t smem = (size_t)GEO_WARPS * 3 * nq * sizeof(float);
    ++g_launches;
    k_block_queries<<<persistent_grid(win_capacity, GEO_WARPS, 8), GEO_WARPS * 32, smem, (cudaStream_t)stream>>>(
        nq, max_win1, use_interp ? 1 : 0, win_capacity, win_count_total, src_row, win1_row, (const int4 *)meta_in, xyz,
        q_row, (int4 *)meta_out, nn_idx, nn_w);
    return check_launch();
}

/* One-window gather of the compress block without host synchronisation: k_row (cap, max_win1)
 * global feature rows in table order, -1 padded. */
int mssvt_window_rows(int x_max, int y_max, int z_max, int x_ws, int y_ws, int z_ws,
                      int num_win1, int max_win1, const int *q_win1, int win_capacity,
                      const int *win_count_total, const int *win_list, const int *grid_cells,
                      const int *grid_vals, const int *v_start, int *k_row, void *stream) {
    GatherShape g = {x_max, y_max, z_max, x_ws, y_ws, z_ws, 1, {0, 0, num_win1, 0},
                     {0, 0, max_win1, 0}};
    if (!shape_ok(g) || max_win1 <= 0 || win_capacity < 0) return MSSVT_ERR_INVALID;
    if (win_capacity == 0) return MSSVT_OK;
    if (!q_win1 || !win_count_total || !win_list || !grid_cells || !grid_vals || !v_start || !k_row)
        return MSSVT_ERR_INVALID;
    TablePtrs tabs = {{nullptr, nullptr, q_win1, nullptr}};
    size_t smem = (size_t)(num_win1 * 3 + GEO_WARPS * max_win1 * 2) * sizeof(int);
    if (smem > 200 * 1024) return MSSVT_ERR_INVALID;
    cudaFuncSetAttribute(k_window_rows, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    const int zw = (z_max + 31) / 32;
    GridIdx index = {(const int2 *)grid_cells, grid_vals, y_max, zw, (long long)x_max * y_max * zw};
    ++g_launches;
    k_window_rows<<<persistent_grid(win_capacity, GEO_WARPS, 8), GEO_WARPS * 32, smem,
                    (cudaStream_t)stream>>>(g, tabs, win_capacity, win_count_total, (const int4 *)win_list,
                                            index, v_start, k_row);
    return check_launch();
}

/* q_src[q_base[w] + s] = w * nq + s for every real query slot (inverse of the compact numbering). */
int mssvt_query_src(int win_capacity, const int *win_count_total, int nq, const int *meta, const int *q_base,
                    int *q_src, void *stream) {
    if (win_capacity < 0 || nq <= 0) return MSSVT_ERR_INVALID;
    if (win_capacity == 0) return MSSVT_OK;
    if (!win_count_total || !meta || !q_base || !q_src) return MSSVT_ERR_INVALID;
    ++g_launches;
    k_query_src<<<persistent_grid((long long)win_capacity * nq, 256, 8), 256, 0, (cudaStream_t)stream>>>(
        win_capacity, win_count_total, nq, meta, q_base, q_src);
    return check_launch();
}

}  // extern "C"
