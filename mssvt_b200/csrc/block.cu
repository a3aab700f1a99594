// block.cu -- feature path of the MsSVT blocks in exact fp32 (FFMA) arithmetic (sm_100a).
//
//   mssvt_layernorm          <- norm1 = nn.LayerNorm(C)            mssvt_backbone.py:37, 210, 352
//   mssvt_block_attention    <- grouping_operation x7 + pos_proj + MixedScaleAttention +
//                               three-NN interpolation + scatter-merge
//                               mssvt_backbone.py:260-336, mssvt_utils.py:88-157,
//                               group_features_gpu.cu:73-106, group_points_gpu.cu:53-72
//   mssvt_compress_attention <- MixedScaleSparseTransformerCompressBlock.forward:361-383
//   mssvt_ffn                <- residual + norm2 + linear1/ReLU/linear2 + residual (+ out_linear)
//                               mssvt_backbone.py:337-343, 384-387
//   mssvt_dense_scatter      <- SparseTensor.dense()               mssvt_utils.py:50-62
//
// Nothing padded is ever materialised: a warp owns a window, gathers the layer-normed rows it
// needs straight into shared memory, and writes only the merged rows of its win1 voxels.
// Two exact shortcuts (both leave the mathematical result unchanged):
//   * padded query slots are skipped -- the reference zeroes them after attention;
//   * all masked key slots of one scale carry the same key (first voxel of the list, zero
//     relative offset), so K/V are computed once and enter the softmax with a multiplicity.
#include "common.cuh"

namespace mssvt {

#define MAX_GROUPS 4

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// ------------------------------------------------------------------------------- LayerNorm

// one warp per row; two-pass mean / biased variance in fp32, y = (x - mean) * rstd * g + b
__global__ void __launch_bounds__(256)
k_layernorm(int n_cap, const int *__restrict__ n_dev, int C, const float *__restrict__ x,
            const float *__restrict__ gamma, const float *__restrict__ beta, float eps,
            float *__restrict__ y) {
    const int n = n_dev ? min(n_cap, __ldg(n_dev)) : n_cap;
    const int lane = threadIdx.x & 31;
    const int warps = (gridDim.x * blockDim.x) >> 5;
    const float inv_c = 1.0f / (float)C;
    for (int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; row < n; row += warps) {
        const float *src = x + (size_t)row * C;
        float v[8];  // C <= 256
        float s = 0.f;
#pragma unroll
        for (int r = 0; r < 8; ++r) {
            int c = lane + 32 * r;
            v[r] = c < C ? __ldg(src + c) : 0.f;
            s += v[r];
        }
        const float mean = warp_sum(s) * inv_c;
        float q = 0.f;
#pragma unroll
        for (int r = 0; r < 8; ++r) {
            int c = lane + 32 * r;
            float d = c < C ? v[r] - mean : 0.f;
            q += d * d;
        }
        const float rstd = rsqrtf(warp_sum(q) * inv_c + eps);
        float *dst = y + (size_t)row * C;
#pragma unroll
        for (int r = 0; r < 8; ++r) {
            int c = lane + 32 * r;
            if (c < C) dst[c] = (v[r] - mean) * rstd * __ldg(gamma + c) + __ldg(beta + c);
        }
    }
}

// ------------------------------------------------------------------------------- weights

// Flat fp32 parameter pack of one block, built once by the host module (offsets in floats).
// All matrices are stored TRANSPOSED ([in][out]) so that lanes, which own outputs, read
// consecutive words.
struct AttnShape {
    int C, G, hd, nq, nk_total, nk, cap1, interp, pos_layers;
    int heads[MAX_GROUPS], sd[MAX_GROUPS], c0[MAX_GROUPS];
    int off_pos_w, off_pos_b;            // [6][C], [C]
    int off_pos2_w, off_pos2_b;          // [C][C], [C]   (one-window blocks only)
    int off_wq[MAX_GROUPS], off_bq[MAX_GROUPS];     // [sd][sd], [sd]
    int off_wkv[MAX_GROUPS], off_bkv[MAX_GROUPS];   // [sd][2sd], [2sd]
    int off_wp[MAX_GROUPS], off_bp[MAX_GROUPS];     // [sd][sd], [sd]
    int total_floats;
    float scale;
    float win_cell[3];  // window size in metres (fp32 of the python double vs * ws)
    float lo[3];
};

#define ATT_WARPS 4

// positional embedding channel c of pos_proj layer 1: ReLU(b[c] + W[c, 0:3] . rel + W[c, 3:6] . ctr)
__device__ __forceinline__ float pos_embed(const float *s_w, int C, int c, float rx, float ry,
                                           float rz, float cx, float cy, float cz) {
    const float *w = s_w;
    float a = w[6 * C + c];  // bias stored right after the [6][C] matrix
    a = fmaf(w[0 * C + c], rx, a);
    a = fmaf(w[1 * C + c], ry, a);
    a = fmaf(w[2 * C + c], rz, a);
    a = fmaf(w[3 * C + c], cx, a);
    a = fmaf(w[4 * C + c], cy, a);
    a = fmaf(w[5 * C + c], cz, a);
    return fmaxf(a, 0.f);
}

// y[o] = b[o] + sum_i Wt[i][o] * x[i] for o = lane, lane+32, ... < n_out  (x in shared memory)
__device__ __forceinline__ void matvec_store(const float *s_wt, const float *s_b, const float *s_x,
                                             int n_in, int n_out, float mul, float *dst) {
    const int lane = threadIdx.x & 31;
    for (int o = lane; o < n_out; o += 32) {
        float a = s_b[o];
        for (int i = 0; i < n_in; ++i) a = fmaf(s_wt[i * n_out + o], s_x[i], a);
        dst[o] = a * mul;
    }
}

// One warp per window.
__global__ void __launch_bounds__(ATT_WARPS * 32)
k_block_attention(AttnShape S, const float *__restrict__ params, int win_cap,
                  const int *__restrict__ win_count_total, const int4 *__restrict__ win_list,
                  const float *__restrict__ xn, const float *__restrict__ xyz,
                  const int *__restrict__ q_row, const int *__restrict__ k_row,
                  const unsigned char *__restrict__ k_mask, const int *__restrict__ win1_row,
                  const unsigned char *__restrict__ nn_idx, const float *__restrict__ nn_w,
                  float *__restrict__ merged) {
    extern __shared__ float smem[];
    const int C = S.C, nq = S.nq, nk = S.nk;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    int sd_max = 0;
    for (int g = 0; g < S.G; ++g) sd_max = max(sd_max, S.sd[g]);
    const int kv_pitch = 2 * sd_max + 1;
    // CTA-shared weights, then per-warp scratch
    float *s_par = smem;
    const int per_warp = 2 * nq * C + sd_max + nk * kv_pitch + nk + 2 * nk;
    float *s_a = smem + S.total_floats + warp * per_warp;  // [nq][C] query inputs -> attn output
    float *s_q = s_a + nq * C;                             // [nq][C] projected q -> head outputs
    float *s_kin = s_q + nq * C;                           // [sd_max]
    float *s_kv = s_kin + sd_max;                          // [nk][2 sd_max + 1]
    float *s_sc = s_kv + nk * kv_pitch;                    // [nk]
    int *s_rep = (int *)(s_sc + nk);                       // [nk] slot of each distinct key
    int *s_mult = s_rep + nk;                              // [nk] multiplicity (masked: count)
    for (int i = threadIdx.x; i < S.total_floats; i += blockDim.x) s_par[i] = __ldg(params + i);
    __syncthreads();
    const float *s_pos = s_par + S.off_pos_w;
    const int num_wins = min(win_cap, __ldg(win_count_total));

    for (int w = blockIdx.x * ATT_WARPS + warp; w < num_wins; w += gridDim.x * ATT_WARPS) {
        const int4 win = __ldg(win_list + w);
        const float ctx = world_coord(win.w, S.win_cell[0], S.lo[0]);
        const float cty = world_coord(win.z, S.win_cell[1], S.lo[1]);
        const float ctz = world_coord(win.y, S.win_cell[2], S.lo[2]);
        const int *qr = q_row + (size_t)w * nq;
        int nqr = 0;  // real queries are compacted at the front of the list
        for (int s0 = 0; s0 < nq; s0 += 32) {
            int s = s0 + lane;
            nqr += __popc(__ballot_sync(0xffffffffu, s < nq && __ldg(qr + s) >= 0));
        }

        // ---- A: query inputs = layer-normed feature + positional embedding
        for (int s = 0; s < nq; ++s) {
            if (s < nqr) {
                const int row = __ldg(qr + s);
                const float rx = __fsub_rn(__ldg(xyz + 3 * (size_t)row), ctx);
                const float ry = __fsub_rn(__ldg(xyz + 3 * (size_t)row + 1), cty);
                const float rz = __fsub_rn(__ldg(xyz + 3 * (size_t)row + 2), ctz);
                for (int c = lane; c < C; c += 32)
                    s_a[s * C + c] = __ldg(xn + (size_t)row * C + c) +
                                     pos_embed(s_pos, C, c, rx, ry, rz, ctx, cty, ctz);
            } else {
                for (int c = lane; c < C; c += 32) s_a[s * C + c] = 0.f;
            }
        }
        __syncwarp();

        if (nqr > 0) {
            // ---- B: q = (Wq x + b) * scale, per head group on its channel slice
            for (int g = 0; g < S.G; ++g)
                for (int s = 0; s < nqr; ++s)
                    matvec_store(s_par + S.off_wq[g], s_par + S.off_bq[g], s_a + s * C + S.c0[g],
                                 S.sd[g], S.sd[g], S.scale, s_q + s * C + S.c0[g]);
            __syncwarp();

            // ---- C/D: per head group: distinct keys -> K,V ; softmax(QK^T - 100 mask) V
            for (int g = 0; g < S.G; ++g) {
                const int sd = S.sd[g], c0 = S.c0[g];
                const int *kr = k_row + (size_t)w * S.nk_total + g * nk;
                const unsigned char *km = k_mask + (size_t)w * S.nk_total + g * nk;
                // distinct keys: every unmasked slot, plus the first masked slot standing for all
                int nrep = 0, first_masked = -1, n_masked = 0;
                for (int j0 = 0; j0 < nk; j0 += 32) {
                    int j = j0 + lane;
                    bool masked = j < nk && __ldg(km + j) != 0;
                    bool live = j < nk && !masked;
                    unsigned mm = __ballot_sync(0xffffffffu, masked);
                    unsigned lm = __ballot_sync(0xffffffffu, live);
                    if (mm && first_masked < 0) first_masked = j0 + __ffs(mm) - 1;
                    n_masked += __popc(mm);
                    if (live) {
                        int t = nrep + __popc(lm & lanemask_lt());
                        s_rep[t] = j;
                        s_mult[t] = 1;
                    }
                    nrep += __popc(lm);
                }
                if (first_masked >= 0) {
                    if (lane == 0) { s_rep[nrep] = first_masked; s_mult[nrep] = -n_masked; }
                    nrep += 1;
                }
                __syncwarp();
                for (int t = 0; t < nrep; ++t) {
                    const int j = s_rep[t];
                    const bool masked = s_mult[t] < 0;
                    const int row = __ldg(kr + j);
                    float rx = 0.f, ry = 0.f, rz = 0.f;  // masked keys: relative offset zeroed
                    if (!masked) {
                        rx = __fsub_rn(__ldg(xyz + 3 * (size_t)row), ctx);
                        ry = __fsub_rn(__ldg(xyz + 3 * (size_t)row + 1), cty);
                        rz = __fsub_rn(__ldg(xyz + 3 * (size_t)row + 2), ctz);
                    }
                    for (int i = lane; i < sd; i += 32)
                        s_kin[i] = __ldg(xn + (size_t)row * C + c0 + i) +
                                   pos_embed(s_pos, C, c0 + i, rx, ry, rz, ctx, cty, ctz);
                    __syncwarp();
                    matvec_store(s_par + S.off_wkv[g], s_par + S.off_bkv[g], s_kin, sd, 2 * sd,
                                 1.0f, s_kv + t * kv_pitch);
                    __syncwarp();
                }
                for (int s = 0; s < nqr; ++s) {
                    for (int h = 0; h < S.heads[g]; ++h) {
                        const float *qv = s_q + s * C + c0 + h * S.hd;
                        // scores over distinct keys (lane = key), additive -100 on masked ones
                        float mx = -3.0e38f;
                        for (int t0 = 0; t0 < nrep; t0 += 32) {
                            int t = t0 + lane;
                            if (t < nrep) {
                                const float *kk = s_kv + t * kv_pitch + h * S.hd;
                                float a = 0.f;
                                for (int d = 0; d < S.hd; ++d) a = fmaf(qv[d], kk[d], a);
                                if (s_mult[t] < 0) a += -100.0f;
                                s_sc[t] = a;
                                mx = fmaxf(mx, a);
                            }
                        }
                        mx = warp_max(mx);
                        __syncwarp();
                        float den = 0.f;
                        for (int t0 = 0; t0 < nrep; t0 += 32) {
                            int t = t0 + lane;
                            if (t < nrep) {
                                int m = s_mult[t];
                                float e = expf(s_sc[t] - mx) * (float)(m < 0 ? -m : m);
                                s_sc[t] = e;
                                den += e;
                            }
                        }
                        den = warp_sum(den);
                        __syncwarp();
                        const float inv = 1.0f / den;
                        __syncwarp();
                        // head output (lane = channel within the head); overwrites this head's q
                        for (int d = lane; d < S.hd; d += 32) {
                            float o = 0.f;
                            for (int t = 0; t < nrep; ++t)
                                o = fmaf(s_sc[t], s_kv[t * kv_pitch + sd + h * S.hd + d], o);
                            s_q[s * C + c0 + h * S.hd + d] = o * inv;
                        }
                        __syncwarp();
                    }
                }
            }
            // ---- E: output projection per group back into s_a (padded query rows stay zero)
            for (int g = 0; g < S.G; ++g)
                for (int s = 0; s < nqr; ++s)
                    matvec_store(s_par + S.off_wp[g], s_par + S.off_bp[g], s_q + s * C + S.c0[g],
                                 S.sd[g], S.sd[g], 1.0f, s_a + s * C + S.c0[g]);
            __syncwarp();
        }

        // ---- F: merge.  interp: every win1 voxel gets the 1/d blend of its 3 nearest queries;
        //         otherwise the query voxels get their own attention rows.
        if (S.interp) {
            const int *wr = win1_row + (size_t)w * S.cap1;
            for (int i = 0; i < S.cap1; ++i) {
                const int row = __ldg(wr + i);
                if (row < 0) break;  // list is compact
                const unsigned char *ni = nn_idx + ((size_t)w * S.cap1 + i) * 3;
                const float *nw = nn_w + ((size_t)w * S.cap1 + i) * 3;
                const int n0 = ni[0], n1 = ni[1], n2 = ni[2];
                const float w0 = __ldg(nw), w1 = __ldg(nw + 1), w2 = __ldg(nw + 2);
                for (int c = lane; c < C; c += 32) {
                    float y = __fadd_rn(__fadd_rn(__fmul_rn(s_a[n0 * C + c], w0),
                                                  __fmul_rn(s_a[n1 * C + c], w1)),
                                        __fmul_rn(s_a[n2 * C + c], w2));
                    merged[(size_t)row * C + c] = y;
                }
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

// One warp per window of a one-window (compress) block: a single query per window = channel-wise
// max over the window's layer-normed rows INCLUDING the zero padding (Q6); keys = the rows plus
// a two-layer positional embedding; padded slots all carry the same key (0 + posemb(-ctr, ctr)),
// computed once and weighted by their count under the -100 mask.
__global__ void __launch_bounds__(ATT_WARPS * 32)
k_compress_attention(AttnShape S, const float *__restrict__ params, int win_cap,
                     const int *__restrict__ win_count_total, const int4 *__restrict__ win_list,
                     const float *__restrict__ xn, const float *__restrict__ xyz,
                     const int *__restrict__ k_row_list, float *__restrict__ out) {
    extern __shared__ float smem[];
    const int C = S.C, n1 = S.cap1, nk = S.nk;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    int sd_max = 0;
    for (int g = 0; g < S.G; ++g) sd_max = max(sd_max, S.sd[g]);
    const int kv_pitch = 2 * sd_max + 1;
    float *s_par = smem;
    const int per_warp = 3 * C + C + (n1 + 1) * C + nk * kv_pitch + nk + nk;
    float *s_qin = smem + S.total_floats + warp * per_warp;  // [C] max-pooled query
    float *s_q = s_qin + C;                                  // [C] projected q -> head outputs
    float *s_o = s_q + C;                                    // [C] attention output
    float *s_h = s_o + C;                                    // [C] hidden of pos_proj layer 1
    float *s_key = s_h + C;                                  // [n1 + 1][C] key inputs (last = pad key)
    float *s_kv = s_key + (n1 + 1) * C;                      // [nk][2 sd_max + 1]
    float *s_sc = s_kv + nk * kv_pitch;                      // [nk]
    int *s_mult = (int *)(s_sc + nk);                        // [nk]
    for (int i = threadIdx.x; i < S.total_floats; i += blockDim.x) s_par[i] = __ldg(params + i);
    __syncthreads();
    const float *s_pos = s_par + S.off_pos_w;
    const int num_wins = min(win_cap, __ldg(win_count_total));

    for (int w = blockIdx.x * ATT_WARPS + warp; w < num_wins; w += gridDim.x * ATT_WARPS) {
        const int4 win = __ldg(win_list + w);
        const float ctx = world_coord(win.w, S.win_cell[0], S.lo[0]);
        const float cty = world_coord(win.z, S.win_cell[1], S.lo[1]);
        const float ctz = world_coord(win.y, S.win_cell[2], S.lo[2]);
        const int *kr = k_row_list + (size_t)w * n1;
        int cnt = 0;
        for (int s0 = 0; s0 < n1; s0 += 32) {
            int s = s0 + lane;
            cnt += __popc(__ballot_sync(0xffffffffu, s < n1 && __ldg(kr + s) >= 0));
        }
        // key inputs for the cnt real slots and, if any slot is padding, the shared pad key
        const int nkeys = cnt < n1 ? cnt + 1 : cnt;
        for (int c = lane; c < C; c += 32) s_qin[c] = cnt < n1 ? 0.f : -3.0e38f;
        __syncwarp();
        for (int t = 0; t < nkeys; ++t) {
            const bool pad = t >= cnt;
            const int row = pad ? 0 : __ldg(kr + t);
            // padded slots: grouped coordinate is 0, so relative = 0 - centre (not masked here)
            const float px = pad ? 0.f : __ldg(xyz + 3 * (size_t)row);
            const float py = pad ? 0.f : __ldg(xyz + 3 * (size_t)row + 1);
            const float pz = pad ? 0.f : __ldg(xyz + 3 * (size_t)row + 2);
            const float rx = __fsub_rn(px, ctx), ry = __fsub_rn(py, cty), rz = __fsub_rn(pz, ctz);
            for (int c = lane; c < C; c += 32) {
                float f = pad ? 0.f : __ldg(xn + (size_t)row * C + c);
                if (!pad) s_qin[c] = fmaxf(s_qin[c], f);
                s_key[t * C + c] = f;
                s_h[c] = pos_embed(s_pos, C, c, rx, ry, rz, ctx, cty, ctz);
            }
            __syncwarp();
            if (S.pos_layers == 2) {
                const float *w2 = s_par + S.off_pos2_w, *b2 = s_par + S.off_pos2_b;
                for (int o = lane; o < C; o += 32) {
                    float a = b2[o];
                    for (int i = 0; i < C; ++i) a = fmaf(w2[i * C + o], s_h[i], a);
                    s_key[t * C + o] += fmaxf(a, 0.f);
                }
            } else {
                for (int c = lane; c < C; c += 32) s_key[t * C + c] += s_h[c];
            }
            __syncwarp();
        }
        // q projection per group
        for (int g = 0; g < S.G; ++g)
            matvec_store(s_par + S.off_wq[g], s_par + S.off_bq[g], s_qin + S.c0[g], S.sd[g], S.sd[g],
                         S.scale, s_q + S.c0[g]);
        __syncwarp();
        for (int g = 0; g < S.G; ++g) {
            const int sd = S.sd[g], c0 = S.c0[g];
            // group g sees slots [g nk, (g+1) nk) of the padded list: real ones, then padding
            const int lo_slot = g * nk, hi_slot = lo_slot + nk;
            const int real_hi = min(hi_slot, cnt);
            int nrep = 0;
            for (int j = lo_slot; j < real_hi; ++j, ++nrep) {
                matvec_store(s_par + S.off_wkv[g], s_par + S.off_bkv[g], s_key + j * C + c0, sd,
                             2 * sd, 1.0f, s_kv + nrep * kv_pitch);
                if (lane == 0) s_mult[nrep] = 1;
            }
            const int n_pad = hi_slot - max(lo_slot, cnt);
            if (n_pad > 0) {
                matvec_store(s_par + S.off_wkv[g], s_par + S.off_bkv[g], s_key + cnt * C + c0, sd,
                             2 * sd, 1.0f, s_kv + nrep * kv_pitch);
                if (lane == 0) s_mult[nrep] = -n_pad;
                nrep += 1;
            }
            __syncwarp();
            for (int h = 0; h < S.heads[g]; ++h) {
                const float *qv = s_q + c0 + h * S.hd;
                float mx = -3.0e38f;
                for (int t0 = 0; t0 < nrep; t0 += 32) {
                    int t = t0 + lane;
                    if (t < nrep) {
                        const float *kk = s_kv + t * kv_pitch + h * S.hd;
                        float a = 0.f;
                        for (int d = 0; d < S.hd; ++d) a = fmaf(qv[d], kk[d], a);
                        if (s_mult[t] < 0) a += -100.0f;
                        s_sc[t] = a;
                        mx = fmaxf(mx, a);
                    }
                }
                mx = warp_max(mx);
                __syncwarp();
                float den = 0.f;
                for (int t0 = 0; t0 < nrep; t0 += 32) {
                    int t = t0 + lane;
                    if (t < nrep) {
                        int m = s_mult[t];
                        float e = expf(s_sc[t] - mx) * (float)(m < 0 ? -m : m);
                        s_sc[t] = e;
                        den += e;
                    }
                }
                den = warp_sum(den);
                const float inv = 1.0f / den;
                __syncwarp();
                for (int d = lane; d < S.hd; d += 32) {
                    float o = 0.f;
                    for (int t = 0; t < nrep; ++t)
                        o = fmaf(s_sc[t], s_kv[t * kv_pitch + sd + h * S.hd + d], o);
                    s_q[c0 + h * S.hd + d] = o * inv;
                }
                __syncwarp();
            }
        }
        for (int g = 0; g < S.G; ++g)
            matvec_store(s_par + S.off_wp[g], s_par + S.off_bp[g], s_q + S.c0[g], S.sd[g], S.sd[g],
                         1.0f, s_o + S.c0[g]);
        __syncwarp();
        for (int c = lane; c < C; c += 32) out[(size_t)w * C + c] = s_o[c];
        __syncwarp();
    }
}

// ------------------------------------------------------------------------------- FFN

struct FfnShape {
    int C, F, C_out;       // in, hidden, out (C_out == 0: no out_linear)
    int mode;              // 0: x_in = merged (compress block); 1: covered ? merged + x : 2 x
    int off_ln_g, off_ln_b, off_w1, off_b1, off_w2, off_b2, off_wo, off_bo;  // transposed mats
    int total_floats;
    float eps;
};

#define FFN_WARPS 8
#define FFN_ROWS 4  // rows per warp per pass

__device__ __forceinline__ float f4_get(const float4 &v, int k) {
    return k == 0 ? v.x : k == 1 ? v.y : k == 2 ? v.z : v.w;
}

// out[r][o] = b[o] + sum_i Wt[i][o] * in[r][i] for R rows at once (in: shared, [R][n_in],
// 16-byte aligned, n_in % 4 == 0).  A lane owns outputs lane, lane+32, lane+64, lane+96 of each
// 128-wide output chunk: per 4 inputs it issues R broadcast LDS.128 + 16 LDS.32 for 16 R FMAs.
template <int R, typename Epi>
__device__ __forceinline__ void dense_rows(const float *wt, const float *b, const float *in,
                                           int n_in, int n_out, Epi epi) {
    const int lane = threadIdx.x & 31;
    for (int ob = 0; ob < n_out; ob += 128) {
        float a[4][R];
        int oc[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            oc[j] = min(ob + lane + 32 * j, n_out - 1);  // clamped: surplus lanes redo the last output
            const float bv = b[oc[j]];
#pragma unroll
            for (int r = 0; r < R; ++r) a[j][r] = bv;
        }
        for (int i = 0; i < n_in; i += 4) {
            float4 v[R];
#pragma unroll
            for (int r = 0; r < R; ++r) v[r] = *(const float4 *)(in + r * n_in + i);
#pragma unroll
            for (int k = 0; k < 4; ++k) {
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const float wv = wt[(i + k) * n_out + oc[j]];
#pragma unroll
                    for (int r = 0; r < R; ++r) a[j][r] = fmaf(wv, f4_get(v[r], k), a[j][r]);
                }
            }
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int o = ob + lane + 32 * j;
            if (o < n_out) {
#pragma unroll
                for (int r = 0; r < R; ++r) epi(r, o, a[j][r]);
            }
        }
    }
}

// y = u + W2 relu(W1 LN(u) + b1) + b2, then optional out_linear; u built from the merge.
// Each warp carries FFN_ROWS rows at once so that every weight word fetched from shared memory
// feeds FFN_ROWS FMAs.
__global__ void __launch_bounds__(FFN_WARPS * 32)
k_ffn(FfnShape S, const float *__restrict__ params, int n_cap, const int *__restrict__ n_dev,
      const float *__restrict__ x, const float *__restrict__ merged,
      const unsigned char *__restrict__ covered, float *__restrict__ y) {
    extern __shared__ __align__(16) float smem[];
    const int C = S.C, F = S.F;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float *s_par = smem;
    const int per_warp = FFN_ROWS * (2 * C + F);
    float *s_u = smem + S.total_floats + warp * per_warp;  // [R][C] residual stream
    float *s_n = s_u + FFN_ROWS * C;                       // [R][C] normed / final
    float *s_h = s_n + FFN_ROWS * C;                       // [R][F] hidden
    for (int i = threadIdx.x; i < S.total_floats; i += blockDim.x) s_par[i] = __ldg(params + i);
    __syncthreads();
    const int n = n_dev ? min(n_cap, __ldg(n_dev)) : n_cap;
    const float inv_c = 1.0f / (float)C;
    const int stride = gridDim.x * FFN_WARPS * FFN_ROWS;
    for (int r0 = (blockIdx.x * FFN_WARPS + warp) * FFN_ROWS; r0 < n; r0 += stride) {
        // residual stream u, then LayerNorm (norm2)
#pragma unroll
        for (int r = 0; r < FFN_ROWS; ++r) {
            const int row = r0 + r;
            const bool live = row < n;
            const bool cov = live && S.mode == 1 && covered[row] != 0;
            float s = 0.f;
            for (int c = lane; c < C; c += 32) {
                float u = 0.f;
                if (live) {
                    if (S.mode == 0) u = __ldg(merged + (size_t)row * C + c);
                    else {
                        const float xv = __ldg(x + (size_t)row * C + c);
                        u = cov ? __ldg(merged + (size_t)row * C + c) + xv : xv + xv;
                    }
                }
                s_u[r * C + c] = u;
                s += u;
            }
            const float mean = warp_sum(s) * inv_c;
            float q = 0.f;
            for (int c = lane; c < C; c += 32) { float d = s_u[r * C + c] - mean; q += d * d; }
            const float rstd = rsqrtf(warp_sum(q) * inv_c + S.eps);
            for (int c = lane; c < C; c += 32)
                s_n[r * C + c] = (s_u[r * C + c] - mean) * rstd * s_par[S.off_ln_g + c] + s_par[S.off_ln_b + c];
        }
        __syncwarp();
        dense_rows<FFN_ROWS>(s_par + S.off_w1, s_par + S.off_b1, s_n, C, F,
                             [&](int r, int o, float a) { s_h[r * F + o] = fmaxf(a, 0.f); });
        __syncwarp();
        dense_rows<FFN_ROWS>(s_par + S.off_w2, s_par + S.off_b2, s_h, F, C,
                             [&](int r, int o, float a) { s_n[r * C + o] = s_u[r * C + o] + a; });
        __syncwarp();
        if (S.C_out == 0) {
#pragma unroll
            for (int r = 0; r < FFN_ROWS; ++r)
                if (r0 + r < n)
                    for (int c = lane; c < C; c += 32) y[(size_t)(r0 + r) * C + c] = s_n[r * C + c];
        } else {
            const int Co = S.C_out;
            dense_rows<FFN_ROWS>(s_par + S.off_wo, s_par + S.off_bo, s_n, C, Co,
                                 [&](int r, int o, float a) {
                                     if (r0 + r < n) y[(size_t)(r0 + r) * Co + o] = a;
                                 });
        }
        __syncwarp();
    }
}

// ------------------------------------------------------------------------------- dense()

// out (B, C, D, H, W) zero-filled, then out[b, :, z, y, x] = features[m, :]
__global__ void k_dense_scatter(int m_cap, const int *__restrict__ m_dev, int C, int D, int H, int Wd,
                                const float *__restrict__ features, const int4 *__restrict__ indices,
                                float *__restrict__ out) {
    const int M = m_dev ? min(m_cap, __ldg(m_dev)) : m_cap;
    const size_t plane = (size_t)D * H * Wd;
    const size_t total = (size_t)M * C;
    for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < total;
         e += (size_t)gridDim.x * blockDim.x) {
        // consecutive threads take consecutive voxels of one channel: rows are sorted along the
        // fastest spatial axes, so the writes of a warp land close together
        int c = (int)(e / M), m = (int)(e - (size_t)c * M);
        int4 id = __ldg(indices + m);
        size_t at = (((size_t)id.x * C + c) * plane) + ((size_t)id.y * H + id.z) * Wd + id.w;
        out[at] = __ldg(features + (size_t)m * C + c);
    }
}

}  // namespace mssvt

using namespace mssvt;

static size_t attn_smem_bytes(const AttnShape &S) {
    int sd_max = 0;
    for (int g = 0; g < S.G; ++g) sd_max = sd_max > S.sd[g] ? sd_max : S.sd[g];
    size_t per_warp = (size_t)2 * S.nq * S.C + sd_max + (size_t)S.nk * (2 * sd_max + 1) + 3 * S.nk;
    return ((size_t)S.total_floats + ATT_WARPS * per_warp) * sizeof(float);
}

static size_t compress_smem_bytes(const AttnShape &S) {
    int sd_max = 0;
    for (int g = 0; g < S.G; ++g) sd_max = sd_max > S.sd[g] ? sd_max : S.sd[g];
    size_t per_warp = (size_t)4 * S.C + (size_t)(S.cap1 + 1) * S.C + (size_t)S.nk * (2 * sd_max + 1) + 2 * S.nk;
    return ((size_t)S.total_floats + ATT_WARPS * per_warp) * sizeof(float);
}

static bool attn_shape_ok(const AttnShape &S) {
    if (S.C <= 0 || S.C > 256 || S.G <= 0 || S.G > MAX_GROUPS || S.hd <= 0 || S.nk <= 0) return false;
    int c = 0;
    for (int g = 0; g < S.G; ++g) {
        if (S.heads[g] <= 0 || S.sd[g] != S.heads[g] * S.hd || S.c0[g] != c) return false;
        c += S.sd[g];
    }
    return c == S.C && S.total_floats > 0;
}

extern "C" {

int mssvt_layernorm(int num_rows, const int *num_rows_dev, int C, const float *x,
                    const float *gamma, const float *beta, float eps, float *y, void *stream) {
    if (num_rows < 0 || C <= 0 || C > 256) return MSSVT_ERR_INVALID;
    if (num_rows == 0) return MSSVT_OK;
    if (!x || !gamma || !beta || !y) return MSSVT_ERR_INVALID;
    ++g_launches;
    k_layernorm<<<persistent_grid(num_rows, 8, 8), 256, 0, (cudaStream_t)stream>>>(
        num_rows, num_rows_dev, C, x, gamma, beta, eps, y);
    return check_launch();
}

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
    size_t smem = attn_smem_bytes(S);
    if (smem > 220 * 1024) return MSSVT_ERR_INVALID;
    cudaFuncSetAttribute(k_block_attention, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    int per_sm = (int)(220 * 1024 / smem);
    per_sm = per_sm > 4 ? 4 : per_sm < 1 ? 1 : per_sm;
    int grid = persistent_grid(win_capacity, ATT_WARPS, per_sm, 1);
    ++g_launches;
    k_block_attention<<<grid, ATT_WARPS * 32, smem, (cudaStream_t)stream>>>(
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
    if (win_capacity == 0) return MSSVT_OK;
    if (!params || !win_count_total || !win_list || !xn || !xyz || !k_row || !out) return MSSVT_ERR_INVALID;
    size_t smem = compress_smem_bytes(S);
    if (smem > 220 * 1024) return MSSVT_ERR_INVALID;
    cudaFuncSetAttribute(k_compress_attention, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    int per_sm = (int)(220 * 1024 / smem);
    per_sm = per_sm > 4 ? 4 : per_sm < 1 ? 1 : per_sm;
    int grid = persistent_grid(win_capacity, ATT_WARPS, per_sm, 1);
    ++g_launches;
    k_compress_attention<<<grid, ATT_WARPS * 32, smem, (cudaStream_t)stream>>>(
        S, params, win_capacity, win_count_total, (const int4 *)win_list, xn, xyz, k_row, out);
    return check_launch();
}

int mssvt_ffn(const void *shape, int shape_bytes, const float *params, int num_rows,
              const int *num_rows_dev, const float *x, const float *merged,
              const unsigned char *covered, float *y, void *stream) {
    if (!shape || shape_bytes != (int)sizeof(FfnShape)) return MSSVT_ERR_INVALID;
    FfnShape S = *(const FfnShape *)shape;
    if (S.C <= 0 || S.F <= 0 || S.C_out < 0 || S.total_floats <= 0 || num_rows < 0) return MSSVT_ERR_INVALID;
    if ((S.C & 3) || (S.F & 3) || (S.total_floats & 3)) return MSSVT_ERR_INVALID;  // 16-byte rows
    if (num_rows == 0) return MSSVT_OK;
    if (!params || !merged || !y || (S.mode == 1 && (!x || !covered))) return MSSVT_ERR_INVALID;
    size_t smem = ((size_t)S.total_floats + (size_t)FFN_WARPS * FFN_ROWS * (2 * S.C + S.F)) * sizeof(float);
    if (smem > 220 * 1024) return MSSVT_ERR_INVALID;
    cudaFuncSetAttribute(k_ffn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    int per_sm = (int)(220 * 1024 / smem);
    per_sm = per_sm > 2 ? 2 : per_sm < 1 ? 1 : per_sm;
    int grid = persistent_grid(num_rows, FFN_WARPS * FFN_ROWS, per_sm, 1);
    ++g_launches;
    k_ffn<<<grid, FFN_WARPS * 32, smem, (cudaStream_t)stream>>>(S, params, num_rows, num_rows_dev, x,
                                                              merged, covered, y);
    return check_launch();
}

int mssvt_dense_scatter(int num_rows, const int *num_rows_dev, int batch_size, int C, int D, int H,
                        int W, const float *features, const int *indices, float *out, void *stream) {
    if (num_rows < 0 || batch_size <= 0 || C <= 0 || D <= 0 || H <= 0 || W <= 0 || !out) return MSSVT_ERR_INVALID;
    cudaStream_t s = (cudaStream_t)stream;
    cudaError_t e = cudaMemsetAsync(out, 0, (size_t)batch_size * C * D * H * W * sizeof(float), s);
    if (e != cudaSuccess) { g_last_cuda_error = (int)e; return MSSVT_ERR_LAUNCH; }
    if (num_rows == 0) return MSSVT_OK;
    if (!features || !indices) return MSSVT_ERR_INVALID;
    ++g_launches;
    k_dense_scatter<<<persistent_grid((long long)num_rows * C, 256, 8), 256, 0, s>>>(
        num_rows, num_rows_dev, C, D, H, W, features, (const int4 *)indices, out);
    return check_launch();
}

int mssvt_sizeof_attn_shape(void) { return (int)sizeof(AttnShape); }
int mssvt_sizeof_ffn_shape(void) { return (int)sizeof(FfnShape); }

}  // extern "C"
