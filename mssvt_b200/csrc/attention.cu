// attention.cu -- window attention of the MsSVT blocks in exact fp32 (FFMA) arithmetic (sm_100a).
//
//   mssvt_block_attention    <- grouping_operation x7 + pos_proj + MixedScaleAttention +
//                               three-NN interpolation + scatter-merge
//                               mssvt_backbone.py:260-336, mssvt_utils.py:88-157,
//                               group_features_gpu.cu:73-106, group_points_gpu.cu:53-72
//   mssvt_compress_attention <- MixedScaleSparseTransformerCompressBlock.forward:361-383
//
// Nothing padded is ever materialised: a warp owns a window, gathers the layer-normed rows it
// needs straight into shared memory, and writes only the merged rows of its win1 voxels.
// Two exact shortcuts (both leave the mathematical result unchanged):
//   * padded query slots are skipped -- the reference zeroes them after attention;
//   * all masked key slots of one scale carry the same key (first voxel of the list, zero
//     relative offset), so K/V are computed once and enter the softmax with a multiplicity.
// Projections run R keys (or queries) at a time through dense_rows: weights come from shared
// memory once per R rows and every lane carries 2R..4R independent FMA chains, which is what
// makes the kernel throughput- rather than latency-bound (profiles/r01_a: 1.5 ms -> see r01_b).
#include "block_common.cuh"

namespace mssvt {

#define ATT_WARPS 16
#define KEY_R 8      // keys projected per pass (two-window blocks)
#define QRY_R 4      // queries projected per pass
#define CMP_R 4      // keys per pass in the compress block (2-3 real keys per pillar on average)

struct AttnSmem {
    int sd_max, h_max, kv_pitch, in_ld;
    __host__ __device__ AttnSmem(const AttnShape &S) {
        sd_max = 0; h_max = 0;
        for (int g = 0; g < S.G; ++g) {
            sd_max = sd_max > S.sd[g] ? sd_max : S.sd[g];
            h_max = h_max > S.heads[g] ? h_max : S.heads[g];
        }
        kv_pitch = 2 * sd_max + 1;
        in_ld = sd_max;
    }
};

#define QRY_C 4   // queries carried through the key passes together (a window has 1-2 on average)
#define MAX_CPL 2 // channels of a head group per lane (sd <= 64); the kernel is compiled for 1 and 2

__host__ __device__ inline int block_attn_per_warp_floats(const AttnShape &S, const AttnSmem &L) {
    const int nk4 = (S.nk + 3) & ~3;
    const int slots = ((S.nq > KEY_R ? S.nq : KEY_R) + 3) & ~3;
    return S.nq * S.C                                  // s_a   [nq][C]  query inputs -> attention rows
           + QRY_C * S.C                               // s_q   [QRY_C][C] scaled q -> head outputs
           + KEY_R * L.in_ld                           // s_in  [KEY_R][sd_max] key inputs of a pass
           + ((KEY_R * L.kv_pitch + 3) & ~3)           // s_kv  [KEY_R][2 sd_max + 1]
           + QRY_C * L.h_max * KEY_R                   // s_sc  [QRY_C][h_max][KEY_R] scores of a pass
           + 2 * nk4                                   // s_rep, s_mult
           + 4 * slots                                 // s_row [slots], s_rel [slots][3]
           + 3 * QRY_C * L.sd_max;                     // s_run [3][QRY_C][sd_max] online-softmax state
}

// One warp per window.  Keys go through in passes of KEY_R with an online softmax, so the K/V of a
// pass (2 KB) is all a warp keeps; 16 warps fit beside the 36 KB of weights.
template <int CPL>
__global__ void __launch_bounds__(ATT_WARPS * 32)
k_block_attention(AttnShape S, const float *__restrict__ params, int win_cap,
                  const int *__restrict__ win_count_total, const int4 *__restrict__ win_list,
                  const float *__restrict__ xn, const float *__restrict__ xyz,
                  const int *__restrict__ q_row, const int *__restrict__ k_row,
                  const unsigned char *__restrict__ k_mask, const int *__restrict__ win1_row,
                  const unsigned char *__restrict__ nn_idx, const float *__restrict__ nn_w,
                  float *__restrict__ merged) {
    extern __shared__ __align__(16) float smem[];
    const int C = S.C, nq = S.nq, nk = S.nk, hd = S.hd;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const AttnSmem L(S);
    const int kv_pitch = L.kv_pitch, in_ld = L.in_ld, h_max = L.h_max;
    const int nk4 = (nk + 3) & ~3, slots = (max(nq, KEY_R) + 3) & ~3;
    float *s_par = smem;
    float *s_a = smem + S.total_floats + warp * block_attn_per_warp_floats(S, L);
    float *s_q = s_a + nq * C;
    float *s_in = s_q + QRY_C * C;
    float *s_kv = s_in + KEY_R * in_ld;
    float *s_sc = s_kv + ((KEY_R * kv_pitch + 3) & ~3);
    int *s_rep = (int *)(s_sc + QRY_C * h_max * KEY_R);
    int *s_mult = s_rep + nk4;
    int *s_row = s_mult + nk4;
    float *s_rel = (float *)(s_row + slots);
    float *s_run = s_rel + 3 * slots;  // running max / denominator / weighted sum per (query, channel)
    const int run_ld = L.sd_max, run_sz = QRY_C * L.sd_max;
    for (int i = threadIdx.x; i < S.total_floats; i += blockDim.x) s_par[i] = __ldg(params + i);
    __syncthreads();
    const float *s_pos = s_par + S.off_pos_w;
    const int num_wins = min(win_cap, __ldg(win_count_total));

    const int nwarps = blockDim.x >> 5;  // as many warps as fit beside the weights (host decides)
    for (int w = blockIdx.x * nwarps + warp; w < num_wins; w += gridDim.x * nwarps) {
        const int4 win = __ldg(win_list + w);
        const float ctx = world_coord(win.w, S.win_cell[0], S.lo[0]);
        const float cty = world_coord(win.z, S.win_cell[1], S.lo[1]);
        const float ctz = world_coord(win.y, S.win_cell[2], S.lo[2]);
        const int *qr = q_row + (size_t)w * nq;

        // ---- A: lane = query slot: feature row and offset to the window centre
        int nqr = 0;  // real queries are compacted at the front of the list
        for (int s0 = 0; s0 < nq; s0 += 32) {
            const int s = s0 + lane;
            const int row = s < nq ? __ldg(qr + s) : -1;
            if (s < nq) {
                s_row[s] = row;
                if (row >= 0) {
                    s_rel[3 * s + 0] = __fsub_rn(__ldg(xyz + 3 * (size_t)row), ctx);
                    s_rel[3 * s + 1] = __fsub_rn(__ldg(xyz + 3 * (size_t)row + 1), cty);
                    s_rel[3 * s + 2] = __fsub_rn(__ldg(xyz + 3 * (size_t)row + 2), ctz);
                }
            }
            nqr += __popc(__ballot_sync(0xffffffffu, row >= 0));
        }
        __syncwarp();
        // query inputs = layer-normed feature + positional embedding; zero rows for padding
        for (int c = lane; c < C; c += 32) {
            float pw[7];
#pragma unroll
            for (int k = 0; k < 7; ++k) pw[k] = s_pos[k * C + c];
            for (int sl = 0; sl < nq; ++sl) {
                float v = 0.f;
                if (sl < nqr) {
                    float a = pw[6];
                    a = fmaf(pw[0], s_rel[3 * sl], a); a = fmaf(pw[1], s_rel[3 * sl + 1], a);
                    a = fmaf(pw[2], s_rel[3 * sl + 2], a); a = fmaf(pw[3], ctx, a);
                    a = fmaf(pw[4], cty, a); a = fmaf(pw[5], ctz, a);
                    v = __ldg(xn + (size_t)s_row[sl] * C + c) + fmaxf(a, 0.f);
                }
                s_a[sl * C + c] = v;
            }
        }
        __syncwarp();

        for (int q0 = 0; q0 < nqr; q0 += QRY_C) {
            const int nqc = min(QRY_C, nqr - q0);
            // ---- B: q = (Wq x + b) * scale per head group
            for (int g = 0; g < S.G; ++g)
                dense_store<QRY_C>(s_par + S.off_wq[g], s_par + S.off_bq[g], s_a + q0 * C + S.c0[g], C,
                                   S.sd[g], S.sd[g], s_q + S.c0[g], C, nqc, S.scale);
            __syncwarp();

            for (int g = 0; g < S.G; ++g) {
                const int sd = S.sd[g], c0 = S.c0[g], heads = S.heads[g];
                const int *kr = k_row + (size_t)w * S.nk_total + g * nk;
                const unsigned char *km = k_mask + (size_t)w * S.nk_total + g * nk;
                // ---- C1: distinct keys = every unmasked slot + the first masked slot for all masked
                int nrep = 0, first_masked = -1, n_masked = 0;
                for (int j0 = 0; j0 < nk; j0 += 32) {
                    const int j = j0 + lane;
                    const bool masked = j < nk && __ldg(km + j) != 0;
                    const bool live = j < nk && !masked;
                    const unsigned mm = __ballot_sync(0xffffffffu, masked);
                    const unsigned lm = __ballot_sync(0xffffffffu, live);
                    if (mm && first_masked < 0) first_masked = j0 + __ffs(mm) - 1;
                    n_masked += __popc(mm);
                    if (live) {
                        const int t = nrep + __popc(lm & lanemask_lt());
                        s_rep[t] = __ldg(kr + j);
                        s_mult[t] = 1;
                    }
                    nrep += __popc(lm);
                }
                if (first_masked >= 0) {
                    if (lane == 0) { s_rep[nrep] = __ldg(kr + first_masked); s_mult[nrep] = -n_masked; }
                    nrep += 1;
                }
                // positional-embedding weights of this lane's channels stay in registers
                float pw[CPL][7];
#pragma unroll
                for (int jj = 0; jj < CPL; ++jj) {
                    const int c = min(c0 + lane + 32 * jj, C - 1);
#pragma unroll
                    for (int k = 0; k < 7; ++k) pw[jj][k] = s_pos[k * C + c];
                }
                // online-softmax state per (query, channel of the group): running max, denominator, sum
                for (int e = lane; e < run_sz; e += 32) {
                    s_run[e] = -3.0e38f; s_run[run_sz + e] = 0.f; s_run[2 * run_sz + e] = 0.f;
                }
                __syncwarp();

                for (int t0 = 0; t0 < nrep; t0 += KEY_R) {
                    const int valid = min(KEY_R, nrep - t0);
                    // ---- C2: lane = key of the pass: offset to the window centre (0 for the masked key)
                    if (lane < KEY_R) {
                        int row = -1;
                        float rx = 0.f, ry = 0.f, rz = 0.f;
                        if (lane < valid) {
                            row = s_rep[t0 + lane];
                            if (s_mult[t0 + lane] > 0) {
                                rx = __fsub_rn(__ldg(xyz + 3 * (size_t)row), ctx);
                                ry = __fsub_rn(__ldg(xyz + 3 * (size_t)row + 1), cty);
                                rz = __fsub_rn(__ldg(xyz + 3 * (size_t)row + 2), ctz);
                            }
                        }
                        s_row[lane] = row;
                        s_rel[3 * lane] = rx; s_rel[3 * lane + 1] = ry; s_rel[3 * lane + 2] = rz;
                    }
                    __syncwarp();
                    // key inputs: all KEY_R gathers of a lane are independent -> KEY_R loads in flight
#pragma unroll
                    for (int jj = 0; jj < CPL; ++jj) {
                        const int i = lane + 32 * jj;
                        if (i < sd) {
                            float f[KEY_R];
#pragma unroll
                            for (int r = 0; r < KEY_R; ++r) {
                                const int row = s_row[r];
                                f[r] = row >= 0 ? __ldg(xn + (size_t)row * C + c0 + i) : 0.f;
                            }
#pragma unroll
                            for (int r = 0; r < KEY_R; ++r) {
                                float a = pw[jj][6];
                                a = fmaf(pw[jj][0], s_rel[3 * r], a); a = fmaf(pw[jj][1], s_rel[3 * r + 1], a);
                                a = fmaf(pw[jj][2], s_rel[3 * r + 2], a); a = fmaf(pw[jj][3], ctx, a);
                                a = fmaf(pw[jj][4], cty, a); a = fmaf(pw[jj][5], ctz, a);
                                s_in[r * in_ld + i] = f[r] + fmaxf(a, 0.f);
                            }
                        }
                    }
                    __syncwarp();
                    dense_store<KEY_R>(s_par + S.off_wkv[g], s_par + S.off_bkv[g], s_in, in_ld, sd, 2 * sd,
                                       s_kv, kv_pitch, KEY_R);
                    __syncwarp();
                    // ---- D1: scores of the pass, lane = (query, head, key)
                    for (int e = lane; e < nqc * heads * KEY_R; e += 32) {
                        const int t = e % KEY_R, sh = e / KEY_R, h = sh % heads, sq = sh / heads;
                        if (t < valid) {
                            const float *qv = s_q + sq * C + c0 + h * hd, *kk = s_kv + t * kv_pitch + h * hd;
                            float a = 0.f;
                            for (int d = 0; d < hd; ++d) a = fmaf(qv[d], kk[d], a);
                            if (s_mult[t0 + t] < 0) a += -100.0f;  // additive mask of the reference
                            s_sc[(sq * h_max + h) * KEY_R + t] = a;
                        }
                    }
                    __syncwarp();
                    // ---- D2: fold the pass into the running softmax, lane = channel
                    for (int c = lane; c < sd; c += 32) {
                        const int h = c / hd;
#pragma unroll 1
                        for (int sq = 0; sq < nqc; ++sq) {
                            const float *sc = s_sc + (sq * h_max + h) * KEY_R;
                            float *st = s_run + sq * run_ld + c;
                            const float m_old = st[0];
                            float m_new = m_old;
                            for (int t = 0; t < valid; ++t) m_new = fmaxf(m_new, sc[t]);
                            const float keep = exp_neg(m_old - m_new);
                            float l = st[run_sz] * keep, o = st[2 * run_sz] * keep;
                            for (int t = 0; t < valid; ++t) {
                                const int m = s_mult[t0 + t];
                                const float e = exp_neg(sc[t] - m_new) * (float)(m < 0 ? -m : m);
                                l += e;
                                o = fmaf(e, s_kv[t * kv_pitch + sd + c], o);
                            }
                            st[0] = m_new; st[run_sz] = l; st[2 * run_sz] = o;
                        }
                    }
                    __syncwarp();
                }
                // head outputs replace this group's q slice
                for (int c = lane; c < sd; c += 32)
                    for (int sq = 0; sq < nqc; ++sq)
                        s_q[sq * C + c0 + c] = s_run[2 * run_sz + sq * run_ld + c] / s_run[run_sz + sq * run_ld + c];
                __syncwarp();
            }
            // ---- E: output projection per group back into s_a (padded query rows stay zero)
            for (int g = 0; g < S.G; ++g)
                dense_store<QRY_C>(s_par + S.off_wp[g], s_par + S.off_bp[g], s_q + S.c0[g], C, S.sd[g],
                                   S.sd[g], s_a + q0 * C + S.c0[g], C, nqc);
            __syncwarp();
        }

        // ---- F: merge.  interp: every win1 voxel gets the 1/d blend of its 3 nearest queries;
        //         otherwise the query voxels get their own attention rows.
        if (S.interp) {
            const int cap1 = S.cap1;
            for (int i0 = 0; i0 < cap1; i0 += 32) {
                // lane = win1 slot: its row, neighbours and weights, then broadcast slot by slot
                const int i = i0 + lane;
                int row = -1, n0 = 0, n1 = 0, n2 = 0;
                float w0 = 0.f, w1 = 0.f, w2 = 0.f;
                if (i < cap1) {
                    row = __ldg(win1_row + (size_t)w * cap1 + i);
                    if (row >= 0) {
                        const unsigned char *ni = nn_idx + ((size_t)w * cap1 + i) * 3;
                        const float *nw = nn_w + ((size_t)w * cap1 + i) * 3;
                        n0 = ni[0]; n1 = ni[1]; n2 = ni[2];
                        w0 = __ldg(nw); w1 = __ldg(nw + 1); w2 = __ldg(nw + 2);
                    }
                }
                const int cnt = __popc(__ballot_sync(0xffffffffu, row >= 0));  // list is compact
                for (int k = 0; k < cnt; ++k) {
                    const int r = __shfl_sync(0xffffffffu, row, k);
                    const int a0 = __shfl_sync(0xffffffffu, n0, k) * C, a1 = __shfl_sync(0xffffffffu, n1, k) * C;
                    const int a2 = __shfl_sync(0xffffffffu, n2, k) * C;
                    const float b0 = __shfl_sync(0xffffffffu, w0, k), b1 = __shfl_sync(0xffffffffu, w1, k);
                    const float b2 = __shfl_sync(0xffffffffu, w2, k);
                    for (int c = lane; c < C; c += 32)
                        merged[(size_t)r * C + c] = __fadd_rn(__fadd_rn(__fmul_rn(s_a[a0 + c], b0),
                                                                        __fmul_rn(s_a[a1 + c], b1)),
                                                              __fmul_rn(s_a[a2 + c], b2));
                }
                if (cnt < 32) break;
            }
        } else {
            for (int s = 0; s < nqr; ++s) {
                const int row = __ldg(qr + s);
                for (int c = lane; c < C; c += 32) merged[(size_t)row * C + c] = s_a[s * C + c];
            }
        }
        __syncwarp();
    }
}

// ------------------------------------------------------------------------------- compress block

#define CMP_WARPS 16

// One warp per window of a one-window (compress) block: a single query per window = channel-wise
// max over the window's layer-normed rows INCLUDING the zero padding (Q6); keys = the rows plus
// a two-layer positional embedding; padded slots all carry the same key (0 + posemb(-ctr, ctr)),
// computed once and weighted by their count under the -100 mask.  Keys are processed CMP_R at a
// time with an online softmax (running max / denominator / weighted sum per head), so a warp only
// needs CMP_R K/V rows of shared memory and 16 warps fit beside the 85 KB of weights.
__global__ void __launch_bounds__(CMP_WARPS * 32)
k_compress_attention(AttnShape S, const float *__restrict__ params, int win_cap,
                     const int *__restrict__ win_count_total, const int4 *__restrict__ win_list,
                     const float *__restrict__ xn, const float *__restrict__ xyz,
                     const int *__restrict__ k_row_list, float *__restrict__ out) {
    extern __shared__ __align__(16) float smem[];
    const int C = S.C, n1 = S.cap1, nk = S.nk;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const AttnSmem L(S);
    const int kv_pitch = L.kv_pitch;
    float *s_par = smem;
    const int per_warp = 3 * C + 2 * CMP_R * C + ((CMP_R * kv_pitch + 3) & ~3) + L.h_max * CMP_R + CMP_R;
    float *s_qin = smem + S.total_floats + warp * per_warp;  // [C] max-pooled query
    float *s_q = s_qin + C;                                  // [C] projected, scaled q
    float *s_o = s_q + C;                                    // [C] attention output (pre-projection)
    float *s_h = s_o + C;                                    // [CMP_R][C] pos_proj hidden layer
    float *s_key = s_h + CMP_R * C;                          // [CMP_R][C] key inputs of a pass
    float *s_kv = s_key + CMP_R * C;                         // [CMP_R][2 sd_max + 1]
    float *s_sc = s_kv + ((CMP_R * kv_pitch + 3) & ~3);      // [h_max][CMP_R] scores of a pass
    float *s_mul = s_sc + L.h_max * CMP_R;                   // [CMP_R] multiplicity (< 0: pad key)
    for (int i = threadIdx.x; i < S.total_floats; i += blockDim.x) s_par[i] = __ldg(params + i);
    __syncthreads();
    const float *s_pos = s_par + S.off_pos_w;
    const int num_wins = min(win_cap, __ldg(win_count_total));
    const int hd = S.hd;

    const int nwarps = blockDim.x >> 5;
    for (int w = blockIdx.x * nwarps + warp; w < num_wins; w += gridDim.x * nwarps) {
        const int4 win = __ldg(win_list + w);
        const float ctx = world_coord(win.w, S.win_cell[0], S.lo[0]);
        const float cty = world_coord(win.z, S.win_cell[1], S.lo[1]);
        const float ctz = world_coord(win.y, S.win_cell[2], S.lo[2]);
        const int *kr = k_row_list + (size_t)w * n1;
        int cnt = 0;
        for (int s0 = 0; s0 < n1; s0 += 32) {
            const int s = s0 + lane;
            cnt += __popc(__ballot_sync(0xffffffffu, s < n1 && __ldg(kr + s) >= 0));
        }
        // query = channel-wise max over the n1 slots; padded slots contribute zeros
        for (int c = lane; c < C; c += 32) {
            float m = cnt < n1 ? 0.f : -3.0e38f;
            for (int t = 0; t < cnt; ++t) m = fmaxf(m, __ldg(xn + (size_t)__ldg(kr + t) * C + c));
            s_qin[c] = m;
        }
        __syncwarp();
        for (int g = 0; g < S.G; ++g)
            dense_store<1>(s_par + S.off_wq[g], s_par + S.off_bq[g], s_qin + S.c0[g], C, S.sd[g], S.sd[g],
                           s_q + S.c0[g], C, 1, S.scale);
        __syncwarp();

        for (int g = 0; g < S.G; ++g) {
            const int sd = S.sd[g], c0 = S.c0[g];
            // group g sees slots [g nk, (g+1) nk) of the padded list: real slots first, then padding
            const int lo_slot = g * nk, hi_slot = lo_slot + nk;
            const int n_real = max(0, min(hi_slot, cnt) - lo_slot);
            const int n_pad = nk - n_real;
            const int nrep = n_real + (n_pad > 0 ? 1 : 0);
            // online-softmax state of the (up to 4) channels this lane owns in the group
            float run_m[4], run_l[4], run_o[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) { run_m[j] = -3.0e38f; run_l[j] = 0.f; run_o[j] = 0.f; }
            for (int t0 = 0; t0 < nrep; t0 += CMP_R) {
                const int valid = min(CMP_R, nrep - t0);
                // key inputs: feature (0 for the pad key) + two-layer positional embedding
#pragma unroll
                for (int r = 0; r < CMP_R; ++r) {
                    if (r < valid) {
                        const bool pad = t0 + r >= n_real;
                        const int row = pad ? 0 : __ldg(kr + lo_slot + t0 + r);
                        // padded slots: the grouped coordinate is 0, so relative = 0 - centre
                        const float px = pad ? 0.f : __ldg(xyz + 3 * (size_t)row);
                        const float py = pad ? 0.f : __ldg(xyz + 3 * (size_t)row + 1);
                        const float pz = pad ? 0.f : __ldg(xyz + 3 * (size_t)row + 2);
                        const float rx = __fsub_rn(px, ctx), ry = __fsub_rn(py, cty), rz = __fsub_rn(pz, ctz);
                        for (int c = lane; c < C; c += 32) {
                            s_key[r * C + c] = pad ? 0.f : __ldg(xn + (size_t)row * C + c);
                            s_h[r * C + c] = pos_embed(s_pos, C, c, rx, ry, rz, ctx, cty, ctz);
                        }
                        if (lane == 0) s_mul[r] = pad ? -(float)n_pad : 1.0f;
                    }
                }
                __syncwarp();
                if (S.pos_layers == 2) {
                    dense_store<CMP_R>(s_par + S.off_pos2_w, s_par + S.off_pos2_b, s_h, C, C, C, s_key, C, CMP_R,
                                       1.0f, DENSE_RELU | DENSE_ADD_DST);
                } else {
                    for (int e = lane; e < CMP_R * C; e += 32) s_key[e] += s_h[e];
                }
                __syncwarp();
                dense_store<CMP_R>(s_par + S.off_wkv[g], s_par + S.off_bkv[g], s_key + c0, C, sd, 2 * sd, s_kv,
                                   kv_pitch, CMP_R);
                __syncwarp();
                // scores of this pass: lane = (head, key)
                for (int e = lane; e < S.heads[g] * CMP_R; e += 32) {
                    const int h = e / CMP_R, r = e - h * CMP_R;
                    const float *qv = s_q + c0 + h * hd, *kk = s_kv + r * kv_pitch + h * hd;
                    float a = 0.f;
                    for (int d = 0; d < hd; ++d) a = fmaf(qv[d], kk[d], a);
                    if (s_mul[r] < 0.f) a += -100.0f;
                    s_sc[h * CMP_R + r] = a;
                }
                __syncwarp();
                // fold the pass into the running softmax, lane = channel
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const int c = lane + 32 * j;
                    if (c < sd) {
                        const float *sc = s_sc + (c / hd) * CMP_R;
                        float m_new = run_m[j];
                        for (int r = 0; r < valid; ++r) m_new = fmaxf(m_new, sc[r]);
                        const float keep = exp_neg(run_m[j] - m_new);
                        float l = run_l[j] * keep, o = run_o[j] * keep;
                        for (int r = 0; r < valid; ++r) {
                            const float e = exp_neg(sc[r] - m_new) * fabsf(s_mul[r]);
                            l += e;
                            o = fmaf(e, s_kv[r * kv_pitch + sd + c], o);
                        }
                        run_m[j] = m_new; run_l[j] = l; run_o[j] = o;
                    }
                }
                __syncwarp();
            }
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int c = lane + 32 * j;
                if (c < sd) s_o[c0 + c] = run_o[j] / run_l[j];
            }
        }
        __syncwarp();
        for (int g = 0; g < S.G; ++g)
            dense_store<1>(s_par + S.off_wp[g], s_par + S.off_bp[g], s_o + S.c0[g], C, S.sd[g], S.sd[g],
                           out + (size_t)w * C + S.c0[g], C, 1);
        __syncwarp();
    }
}

}  // namespace mssvt

using namespace mssvt;

// warps per CTA: as many as fit in the 227 KB of shared memory beside the weight pack
static int fit_warps(size_t weight_floats, size_t per_warp_floats, int max_warps, size_t *smem_bytes) {
    const size_t budget = 227 * 1024;
    int warps = max_warps;
    while (warps > 0 && (weight_floats + warps * per_warp_floats) * sizeof(float) > budget) --warps;
    *smem_bytes = (weight_floats + (size_t)warps * per_warp_floats) * sizeof(float);
    return warps;
}

static size_t compress_per_warp_floats(const AttnShape &S) {
    const AttnSmem L(S);
    return (size_t)3 * S.C + 2 * CMP_R * S.C + ((CMP_R * L.kv_pitch + 3) & ~3) + L.h_max * CMP_R + CMP_R;
}

static bool attn_shape_ok(const AttnShape &S) {
    if (S.C <= 0 || S.C > 256 || (S.C & 3) || S.G <= 0 || S.G > MAX_GROUPS || S.hd <= 0 || S.nk <= 0)
        return false;
    int c = 0;
    for (int g = 0; g < S.G; ++g) {
        if (S.heads[g] <= 0 || S.heads[g] > 4 || S.sd[g] != S.heads[g] * S.hd || S.c0[g] != c || (S.sd[g] & 3))
            return false;
        c += S.sd[g];
    }
    return c == S.C && S.total_floats > 0 && (S.total_floats & 3) == 0;
}

extern "C" {

// shape: the AttnShape struct as a flat int32/float32 blob built by the host (see
// mssvt_b200/_lib.py: AttnShape mirrors this layout field by field).
int mssvt_block_attention(const void *shape, int shape_bytes, const float *params, int win_capacity,
                          const int *win_count_total, const int *win_list, const float *xn,
                          const float *xyz, const int *q_row, const int *k_row,
                          const unsigned char *k_mask, const int *win1_row,
                          const unsigned char *nn_idx, const float *nn_w, float *merged,
                          void *stream) {
    if (!shape || shape_bytes != (int)sizeof(AttnShape)) return MSSVT_ERR_INVALID;
    AttnShape S = *(const AttnShape *)shape;
    if (!attn_shape_ok(S) || S.nq <= 0 || S.nk * S.G > S.nk_total || win_capacity < 0) return MSSVT_ERR_INVALID;
    if (win_capacity == 0) return MSSVT_OK;
    if (!params || !win_count_total || !win_list || !xn || !xyz || !q_row || !k_row || !k_mask || !merged)
        return MSSVT_ERR_INVALID;
    if (S.interp && (!win1_row || !nn_idx || !nn_w)) return MSSVT_ERR_INVALID;
    for (int g = 0; g < S.G; ++g)
        if (S.sd[g] > 32 * MAX_CPL) return MSSVT_ERR_INVALID;  // a lane owns at most MAX_CPL channels per group
    size_t smem = 0;
    const AttnSmem L(S);
    const int warps = fit_warps(S.total_floats, block_attn_per_warp_floats(S, L), ATT_WARPS, &smem);
    if (warps < 1) return MSSVT_ERR_INVALID;
    int sd_max = 0;
    for (int g = 0; g < S.G; ++g) sd_max = sd_max > S.sd[g] ? sd_max : S.sd[g];
    auto kernel = sd_max <= 32 ? k_block_attention<1> : k_block_attention<2>;
    cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    int grid = persistent_grid(win_capacity, warps, 1, 1);
    ++g_launches;
    kernel<<<grid, warps * 32, smem, (cudaStream_t)stream>>>(
        S, params, win_capacity, win_count_total, (const int4 *)win_list, xn, xyz, q_row, k_row,
        k_mask, win1_row, nn_idx, nn_w, merged);
    return check_launch();
}

int mssvt_compress_attention(const void *shape, int shape_bytes, const float *params,
                             int win_capacity, const int *win_count_total, const int *win_list,
                             const float *xn, const float *xyz, const int *k_row, float *out,
                             void *stream) {
    if (!shape || shape_bytes != (int)sizeof(AttnShape)) return MSSVT_ERR_INVALID;
    AttnShape S = *(const AttnShape *)shape;
    if (!attn_shape_ok(S) || S.cap1 <= 0 || S.nk * S.G > S.cap1 || win_capacity < 0) return MSSVT_ERR_INVALID;
    for (int g = 0; g < S.G; ++g)
        if (S.sd[g] > 128) return MSSVT_ERR_INVALID;  // a lane owns at most 4 channels of a group
    if (win_capacity == 0) return MSSVT_OK;
    if (!params || !win_count_total || !win_list || !xn || !xyz || !k_row || !out) return MSSVT_ERR_INVALID;
    size_t smem = 0;
    const int warps = fit_warps(S.total_floats, compress_per_warp_floats(S), CMP_WARPS, &smem);
    if (warps < 1) return MSSVT_ERR_INVALID;
    cudaFuncSetAttribute(k_compress_attention, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    int grid = persistent_grid(win_capacity, warps, 1, 1);
    ++g_launches;
    k_compress_attention<<<grid, warps * 32, smem, (cudaStream_t)stream>>>(
        S, params, win_capacity, win_count_total, (const int4 *)win_list, xn, xyz, k_row, out);
    return check_launch();
}

int mssvt_sizeof_attn_shape(void) { return (int)sizeof(AttnShape); }

}  // extern "C"
