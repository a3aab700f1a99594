// train.cu -- hand-written forward + backward kernels of the TRAINING path on the compact
// (ragged) form of the window attention: the distinct keys of a window and its real queries, no padding.
//
// What the reference differentiates (autograd of mssvt_utils.py:100-157 inside mssvt_backbone.py:260-336):
// per window and head group softmax(q k^T * scale + (-100) * key_mask) v over nk padded key slots, the padded
// slots all holding ONE key (the first voxel of the list, quirks Q1 / Q2) -- so a window's softmax runs over its
// distinct keys, the masked one counted `mult` times, and the gradient of the `mult` identical slots adds up in
// that one key row.  The padded tensors of the autograd form ((W, 64, C) keys: 640 MB per block on a 150 k-voxel
// frame) never exist here: rows are CSR lists per window.
//
//   k_ragged_attn_fwd      thread = (query, head): online softmax over the window's keys, O and the
//                          log-sum-exp of the row (what the backward needs instead of the probabilities)
//   k_ragged_attn_bwd_q    thread = (query, head): delta = dO . O, dQ = scale * sum_k dS_k K_k
//   k_ragged_attn_bwd_kv   thread = (key row, head): dK = scale * sum_q dS Q_q, dV = sum_q P dO_q over the queries
//                          of the key's window -- a key row belongs to ONE window, so there are no atomics and
//                          the result does not depend on the launch order
//   k_interp_merge_fwd/bwd the three-NN blend of the projected query rows back to voxels (mssvt_backbone.py:
//                          318-333): out[v] = sum_j w[v, j] rows[src[v, j]]; uncovered voxels keep x (Q5).  The
//                          backward scatters into the query rows with vector atomics (a query row feeds the
//                          voxels of its window: ~10 adders per row).
// fp32 everywhere: this is the parity-grade training path; the dense projections around these kernels are
// library GEMMs.
#include "common.cuh"

namespace mssvt {

template <int HD>
__device__ __forceinline__ float dot_row(const float *q, const float *__restrict__ row) {
    float a0 = 0.f, a1 = 0.f;
#pragma unroll
    for (int d4 = 0; d4 < HD / 4; ++d4) {
        const float4 k4 = __ldg((const float4 *)row + d4);
        a0 = fmaf(q[4 * d4], k4.x, a0); a1 = fmaf(q[4 * d4 + 1], k4.y, a1);
        a0 = fmaf(q[4 * d4 + 2], k4.z, a0); a1 = fmaf(q[4 * d4 + 3], k4.w, a1);
    }
    return a0 + a1;
}

template <int HD>
__device__ __forceinline__ void load_row(float *dst, const float *__restrict__ row, float mul) {
#pragma unroll
    for (int d4 = 0; d4 < HD / 4; ++d4) {
        const float4 v = __ldg((const float4 *)row + d4);
        dst[4 * d4] = v.x * mul; dst[4 * d4 + 1] = v.y * mul; dst[4 * d4 + 2] = v.z * mul; dst[4 * d4 + 3] = v.w * mul;
    }
}

template <int HD>
__device__ __forceinline__ void store_row(float *row, const float *src, float mul) {
#pragma unroll
    for (int d4 = 0; d4 < HD / 4; ++d4)
        ((float4 *)row)[d4] = make_float4(src[4 * d4] * mul, src[4 * d4 + 1] * mul, src[4 * d4 + 2] * mul, src[4 * d4 + 3] * mul);
}

struct RaggedRows {
    const float *q, *k, *v;
    int ldq, ldk, ldv;
};

template <int HD>
__global__ void __launch_bounds__(128)
k_ragged_attn_fwd(int n_q, int heads, float scale, const int *__restrict__ q_win, const int *__restrict__ key_off,
                  const int *__restrict__ key_mult, RaggedRows R, float *__restrict__ out, int ldo,
                  float *__restrict__ lse) {
    const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= (long long)n_q * heads) return;
    const int qi = (int)(e / heads), h = (int)(e % heads);
    float q[HD], acc[HD];
    load_row<HD>(q, R.q + (size_t)qi * R.ldq + h * HD, scale);
#pragma unroll
    for (int d = 0; d < HD; ++d) acc[d] = 0.f;
    const int w = __ldg(q_win + qi);
    const int k0 = __ldg(key_off + w), k1 = __ldg(key_off + w + 1), mult = __ldg(key_mult + w);
    float m = -INFINITY, l = 0.f;
    for (int k = k0; k < k1; ++k) {
        float s = dot_row<HD>(q, R.k + (size_t)k * R.ldk + h * HD), cnt = 1.f;
        if (k == k1 - 1 && mult > 0) { s += -100.0f; cnt = (float)mult; }   // additive mask of the reference
        const float m_new = fmaxf(m, s);
        const float c = expf(m - m_new), p = expf(s - m_new) * cnt;
        l = fmaf(l, c, p);
        const float *vr = R.v + (size_t)k * R.ldv + h * HD;
#pragma unroll
        for (int d4 = 0; d4 < HD / 4; ++d4) {
            const float4 v4 = __ldg((const float4 *)vr + d4);
            acc[4 * d4] = fmaf(acc[4 * d4], c, p * v4.x); acc[4 * d4 + 1] = fmaf(acc[4 * d4 + 1], c, p * v4.y);
            acc[4 * d4 + 2] = fmaf(acc[4 * d4 + 2], c, p * v4.z); acc[4 * d4 + 3] = fmaf(acc[4 * d4 + 3], c, p * v4.w);
        }
        m = m_new;
    }
    store_row<HD>(out + (size_t)qi * ldo + h * HD, acc, l > 0.f ? 1.0f / l : 0.f);
    lse[e] = l > 0.f ? m + logf(l) : 0.f;
}

template <int HD>
__global__ void __launch_bounds__(128)
k_ragged_attn_bwd_q(int n_q, int heads, float scale, const int *__restrict__ q_win, const int *__restrict__ key_off,
                    const int *__restrict__ key_mult, RaggedRows R, const float *__restrict__ out, int ldo,
                    const float *__restrict__ lse, const float *__restrict__ gout, int ldgo,
                    float *__restrict__ delta, float *__restrict__ gq, int ldgq) {
    const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= (long long)n_q * heads) return;
    const int qi = (int)(e / heads), h = (int)(e % heads);
    float q[HD], go[HD], dq[HD];
    load_row<HD>(q, R.q + (size_t)qi * R.ldq + h * HD, scale);
    load_row<HD>(go, gout + (size_t)qi * ldgo + h * HD, 1.f);
    const float D = dot_row<HD>(go, out + (size_t)qi * ldo + h * HD);
    delta[e] = D;
#pragma unroll
    for (int d = 0; d < HD; ++d) dq[d] = 0.f;
    const int w = __ldg(q_win + qi);
    const int k0 = __ldg(key_off + w), k1 = __ldg(key_off + w + 1), mult = __ldg(key_mult + w);
    const float L = __ldg(lse + e);
    for (int k = k0; k < k1; ++k) {
        const float *kr = R.k + (size_t)k * R.ldk + h * HD;
        float s = dot_row<HD>(q, kr), cnt = 1.f;
        if (k == k1 - 1 && mult > 0) { s += -100.0f; cnt = (float)mult; }
        const float p = expf(s - L) * cnt;
        const float ds = p * (dot_row<HD>(go, R.v + (size_t)k * R.ldv + h * HD) - D);
#pragma unroll
        for (int d4 = 0; d4 < HD / 4; ++d4) {
            const float4 k4 = __ldg((const float4 *)kr + d4);
            dq[4 * d4] = fmaf(ds, k4.x, dq[4 * d4]); dq[4 * d4 + 1] = fmaf(ds, k4.y, dq[4 * d4 + 1]);
            dq[4 * d4 + 2] = fmaf(ds, k4.z, dq[4 * d4 + 2]); dq[4 * d4 + 3] = fmaf(ds, k4.w, dq[4 * d4 + 3]);
        }
    }
    store_row<HD>(gq + (size_t)qi * ldgq + h * HD, dq, scale);
}

template <int HD>
__global__ void __launch_bounds__(128)
k_ragged_attn_bwd_kv(int n_k, int heads, float scale, const int *__restrict__ k_win, const int *__restrict__ q_off,
                     const int *__restrict__ key_off, const int *__restrict__ key_mult, RaggedRows R,
                     const float *__restrict__ lse, const float *__restrict__ delta,
                     const float *__restrict__ gout, int ldgo, float *__restrict__ gk, int ldgk,
                     float *__restrict__ gv, int ldgv) {
    const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= (long long)n_k * heads) return;
    const int k = (int)(e / heads), h = (int)(e % heads);
    float kr[HD], dk[HD], dv[HD];
    load_row<HD>(kr, R.k + (size_t)k * R.ldk + h * HD, scale);      // (scale folded into the key: s = q . (scale k))
    const float *vr = R.v + (size_t)k * R.ldv + h * HD;
#pragma unroll
    for (int d = 0; d < HD; ++d) dk[d] = dv[d] = 0.f;
    const int w = __ldg(k_win + k);
    const int q0 = __ldg(q_off + w), q1 = __ldg(q_off + w + 1), mult = __ldg(key_mult + w);
    const bool masked = mult > 0 && k == __ldg(key_off + w + 1) - 1;
    const float bias = masked ? -100.0f : 0.f, cnt = masked ? (float)mult : 1.f;
    for (int qi = q0; qi < q1; ++qi) {
        const float *qr = R.q + (size_t)qi * R.ldq + h * HD, *gr = gout + (size_t)qi * ldgo + h * HD;
        const float s = dot_row<HD>(kr, qr) + bias;
        const float p = expf(s - __ldg(lse + (size_t)qi * heads + h)) * cnt;
        const float ds = p * (dot_row<HD>(vr, gr) - __ldg(delta + (size_t)qi * heads + h));   // (vr read as the "q" side)
#pragma unroll
        for (int d4 = 0; d4 < HD / 4; ++d4) {
            const float4 q4 = __ldg((const float4 *)qr + d4), g4 = __ldg((const float4 *)gr + d4);
            dk[4 * d4] = fmaf(ds, q4.x, dk[4 * d4]); dk[4 * d4 + 1] = fmaf(ds, q4.y, dk[4 * d4 + 1]);
            dk[4 * d4 + 2] = fmaf(ds, q4.z, dk[4 * d4 + 2]); dk[4 * d4 + 3] = fmaf(ds, q4.w, dk[4 * d4 + 3]);
            dv[4 * d4] = fmaf(p, g4.x, dv[4 * d4]); dv[4 * d4 + 1] = fmaf(p, g4.y, dv[4 * d4 + 1]);
            dv[4 * d4 + 2] = fmaf(p, g4.z, dv[4 * d4 + 2]); dv[4 * d4 + 3] = fmaf(p, g4.w, dv[4 * d4 + 3]);
        }
    }
    store_row<HD>(gk + (size_t)k * ldgk + h * HD, dk, scale);
    store_row<HD>(gv + (size_t)k * ldgv + h * HD, dv, 1.f);
}

// ---- three-NN blend back to voxels: LPR lanes per row, a float4 each
__global__ void __launch_bounds__(256)
k_interp_merge_fwd(int n, int c4n, const int *__restrict__ src, const float *__restrict__ wgt,
                   const float4 *__restrict__ rows, const float4 *__restrict__ x, float4 *__restrict__ out) {
    const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= (long long)n * c4n) return;
    const int v = (int)(e / c4n), c = (int)(e % c4n);
    const int s0 = __ldg(src + 3 * (size_t)v);
    if (s0 == -2) {   // uncovered voxel: keeps its input row (Q5)
        out[e] = __ldg(x + e);
        return;
    }
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int j = 0; j < 3; ++j) {
        const int s = j == 0 ? s0 : __ldg(src + 3 * (size_t)v + j);
        if (s < 0) continue;   // padded query slot: a zero row (mssvt_utils.py:152-153)
        const float w = __ldg(wgt + 3 * (size_t)v + j);
        const float4 r = __ldg(rows + (size_t)s * c4n + c);
        acc.x = fmaf(w, r.x, acc.x); acc.y = fmaf(w, r.y, acc.y); acc.z = fmaf(w, r.z, acc.z); acc.w = fmaf(w, r.w, acc.w);
    }
    out[e] = acc;
}

__global__ void __launch_bounds__(256)
k_interp_merge_bwd(int n, int c4n, const int *__restrict__ src, const float *__restrict__ wgt,
                   const float4 *__restrict__ gout, float4 *__restrict__ grows, float4 *__restrict__ gx) {
    const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= (long long)n * c4n) return;
    const int v = (int)(e / c4n), c = (int)(e % c4n);
    const int s0 = __ldg(src + 3 * (size_t)v);
    const float4 g = __ldg(gout + e);
    if (s0 == -2) {
        gx[e] = g;
        return;
    }
    gx[e] = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int j = 0; j < 3; ++j) {
        const int s = j == 0 ? s0 : __ldg(src + 3 * (size_t)v + j);
        if (s < 0) continue;
        const float w = __ldg(wgt + 3 * (size_t)v + j);
        atomicAdd(grows + (size_t)s * c4n + c, make_float4(w * g.x, w * g.y, w * g.z, w * g.w));   // red.global.add.v4.f32
    }
}


// ---- gather + positional embedding of compact rows (mssvt_backbone.py:288-300 for the rows that exist):
//      out[r, :] = xn[row[r], c0 : c0 + cs] + relu(W[c0 : c0 + cs] . [rel | centre] + b), rel = xyz[row] - centre[win[r]]
//      (zeroed for the masked key, quirk Q1 / Q2).  LPR lanes per row, four channels each.
struct EmbedRows {
    const int *rows, *win;
    const unsigned char *masked;   // may be NULL
    const float *xyz, *centre;     // (N, 3), (W, 3)
    const float *w, *b;            // Conv1d(6 -> C, 1) weight (C, 6) and bias (C)
    int n_rows, c0, ldx;
};

__device__ __forceinline__ void embed_pos(const EmbedRows &E, int r, int row, float pos[6]) {
    const int wi = __ldg(E.win + r);
    const float cx = __ldg(E.centre + 3 * (size_t)wi), cy = __ldg(E.centre + 3 * (size_t)wi + 1), cz = __ldg(E.centre + 3 * (size_t)wi + 2);
    float x = 0.f, y = 0.f, z = 0.f;
    if (row >= 0) { x = __ldg(E.xyz + 3 * (size_t)row); y = __ldg(E.xyz + 3 * (size_t)row + 1); z = __ldg(E.xyz + 3 * (size_t)row + 2); }
    const bool m = E.masked && E.masked[r];
    pos[0] = m ? 0.f : __fsub_rn(x, cx); pos[1] = m ? 0.f : __fsub_rn(y, cy); pos[2] = m ? 0.f : __fsub_rn(z, cz);
    pos[3] = cx; pos[4] = cy; pos[5] = cz;
}

// weights of a lane's four channels: [k][0..5] = W[c + k][:], [k][6] = b[c + k] (28 registers, loaded once per thread)
struct EmbedW {
    float w[4][7];
    __device__ __forceinline__ void load(const EmbedRows &E, int c) {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
#pragma unroll
            for (int j = 0; j < 6; ++j) w[k][j] = __ldg(E.w + 6 * (c + k) + j);
            w[k][6] = __ldg(E.b + c + k);
        }
    }
    __device__ __forceinline__ float pre(int k, const float pos[6]) const {
        float a = w[k][6];
#pragma unroll
        for (int j = 0; j < 6; ++j) a = fmaf(w[k][j], pos[j], a);
        return a;
    }
};

// grid-stride over rows with a FIXED channel group per thread (blockDim.x % LPR == 0), so that the group's weights stay in
// registers
template <int LPR>
__global__ void __launch_bounds__(256)
k_embed_rows_fwd(EmbedRows E, const float *__restrict__ xn, float *__restrict__ out) {
    const int l = threadIdx.x % LPR, c = E.c0 + 4 * l;
    const int rows_per_pass = gridDim.x * (blockDim.x / LPR);
    EmbedW W;
    if (E.w) W.load(E, c);
    for (int r = blockIdx.x * (blockDim.x / LPR) + threadIdx.x / LPR; r < E.n_rows; r += rows_per_pass) {
        const int row = __ldg(E.rows + r);
        float4 v = (xn && row >= 0) ? __ldg((const float4 *)(xn + (size_t)row * E.ldx + c)) : make_float4(0.f, 0.f, 0.f, 0.f);
        if (E.w) {   // (without weights: the gather alone)
            float pos[6];
            embed_pos(E, r, row, pos);
            v.x += fmaxf(W.pre(0, pos), 0.f); v.y += fmaxf(W.pre(1, pos), 0.f);
            v.z += fmaxf(W.pre(2, pos), 0.f); v.w += fmaxf(W.pre(3, pos), 0.f);
        }
        ((float4 *)out)[(size_t)r * LPR + l] = v;
    }
}

// backward: grad_xn rows by vector atomics (a voxel is a key of up to 125 windows), grad_w / grad_b per thread over the
// grid-stride loop, then one shared-memory and one global reduction per CTA
template <int LPR>
__global__ void __launch_bounds__(256)
k_embed_rows_bwd(EmbedRows E, const float *__restrict__ gout, float *__restrict__ gxn, float *__restrict__ gw,
                 float *__restrict__ gb) {
    __shared__ float s_acc[LPR * 4 * 7];
    for (int i = threadIdx.x; i < LPR * 4 * 7; i += blockDim.x) s_acc[i] = 0.f;
    __syncthreads();
    const int l = threadIdx.x % LPR, c = E.c0 + 4 * l;       // (blockDim.x % LPR == 0)
    const int rows_per_pass = gridDim.x * (blockDim.x / LPR);
    EmbedW W;
    if (E.w) W.load(E, c);
    float aw[4][6], ab[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        ab[k] = 0.f;
#pragma unroll
        for (int j = 0; j < 6; ++j) aw[k][j] = 0.f;
    }
    for (int r = blockIdx.x * (blockDim.x / LPR) + threadIdx.x / LPR; r < E.n_rows; r += rows_per_pass) {
        const int row = __ldg(E.rows + r);
        const float4 g = __ldg((const float4 *)gout + (size_t)r * LPR + l);
        if (gxn && row >= 0) atomicAdd((float4 *)(gxn + (size_t)row * E.ldx + c), g);
        if (!E.w) continue;
        float pos[6];
        embed_pos(E, r, row, pos);
        const float gk[4] = {g.x, g.y, g.z, g.w};
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            if (W.pre(k, pos) > 0.f) {
                ab[k] += gk[k];
#pragma unroll
                for (int j = 0; j < 6; ++j) aw[k][j] = fmaf(gk[k], pos[j], aw[k][j]);
            }
        }
    }
    if (!E.w) return;   // (uniform over the grid)
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        atomicAdd(s_acc + (4 * l + k) * 7 + 6, ab[k]);
#pragma unroll
        for (int j = 0; j < 6; ++j) atomicAdd(s_acc + (4 * l + k) * 7 + j, aw[k][j]);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < LPR * 4 * 7; i += blockDim.x) {
        const int ch = E.c0 + i / 7, j = i % 7;
        if (j == 6) atomicAdd(gb + ch, s_acc[i]);
        else atomicAdd(gw + 6 * ch + j, s_acc[i]);
    }
}

// ---- max over the rows of a window (the max-pooled query of the compress block, mssvt_backbone.py:373: the compact key
//      rows of a window are contiguous and include its zero pad row when the window has padded slots, quirk Q6).
//      arg = the row that supplied each channel (first one on ties, like torch.max): the backward routes the gradient
//      there; a row belongs to one window, so the backward writes every row once and needs no atomics.
__global__ void __launch_bounds__(256)
k_segment_max_fwd(int n_win, int c4n, const int *__restrict__ key_off, const float4 *__restrict__ rows,
                  float4 *__restrict__ out, int4 *__restrict__ arg) {
    const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= (long long)n_win * c4n) return;
    const int w = (int)(e / c4n), c = (int)(e % c4n);
    const int k0 = __ldg(key_off + w), k1 = __ldg(key_off + w + 1);
    float4 m = make_float4(0.f, 0.f, 0.f, 0.f);
    int4 a = make_int4(-1, -1, -1, -1);
    for (int k = k0; k < k1; ++k) {
        const float4 v = __ldg(rows + (size_t)k * c4n + c);
        if (k == k0 || v.x > m.x) { m.x = v.x; a.x = k; }
        if (k == k0 || v.y > m.y) { m.y = v.y; a.y = k; }
        if (k == k0 || v.z > m.z) { m.z = v.z; a.z = k; }
        if (k == k0 || v.w > m.w) { m.w = v.w; a.w = k; }
    }
    out[e] = m;
    arg[e] = a;
}

__global__ void __launch_bounds__(256)
k_segment_max_bwd(int n_rows, int c4n, const int *__restrict__ k_win, const int4 *__restrict__ arg,
                  const float4 *__restrict__ gout, float4 *__restrict__ grows) {
    const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= (long long)n_rows * c4n) return;
    const int k = (int)(e / c4n), c = (int)(e % c4n);
    const int w = __ldg(k_win + k);
    const int4 a = __ldg(arg + (size_t)w * c4n + c);
    const float4 g = __ldg(gout + (size_t)w * c4n + c);
    grows[e] = make_float4(a.x == k ? g.x : 0.f, a.y == k ? g.y : 0.f, a.z == k ? g.z : 0.f, a.w == k ? g.w : 0.f);
}

// ---- LayerNorm backward over rows of C = 4 * LPR channels (statistics recomputed from x): LPR lanes per row with a
//      fixed channel group, grad_gamma / grad_beta in registers over the grid-stride loop, reduced once per CTA
template <int LPR>
__global__ void __launch_bounds__(256)
k_layernorm_bwd(int n, const float *__restrict__ x, const float *__restrict__ gamma, float eps,
                const float *__restrict__ gy, float *__restrict__ gx, float *__restrict__ ggamma,
                float *__restrict__ gbeta) {
    constexpr int C = 4 * LPR;
    __shared__ float s_acc[2 * C];
    for (int i = threadIdx.x; i < 2 * C; i += blockDim.x) s_acc[i] = 0.f;
    __syncthreads();
    const int l = threadIdx.x % LPR;
    const int rows_per_pass = gridDim.x * (blockDim.x / LPR);
    const float4 g4 = __ldg((const float4 *)gamma + l);
    float4 ag = make_float4(0.f, 0.f, 0.f, 0.f), abt = ag;
    const int r0 = blockIdx.x * (blockDim.x / LPR) + threadIdx.x / LPR;
    // (every lane of a warp runs the same number of iterations: the shuffles below are warp-wide)
    const int iters = (n + rows_per_pass - 1) / rows_per_pass;
    for (int it = 0; it < iters; ++it) {
        const int r = r0 + it * rows_per_pass;
        const bool live = r < n;
        const float4 v = live ? __ldg((const float4 *)(x + (size_t)r * C) + l) : make_float4(0.f, 0.f, 0.f, 0.f);
        const float4 d = live ? __ldg((const float4 *)(gy + (size_t)r * C) + l) : make_float4(0.f, 0.f, 0.f, 0.f);
        float s = (v.x + v.y) + (v.z + v.w);
#pragma unroll
        for (int o = LPR / 2; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        const float mean = s * (1.0f / C);
        const float dx = v.x - mean, dy = v.y - mean, dz = v.z - mean, dw = v.w - mean;
        float q = (dx * dx + dy * dy) + (dz * dz + dw * dw);
#pragma unroll
        for (int o = LPR / 2; o > 0; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
        const float rstd = rsqrtf(q * (1.0f / C) + eps);
        const float hx = dx * rstd, hy = dy * rstd, hz = dz * rstd, hw = dw * rstd;        // x hat
        const float ex = d.x * g4.x, ey = d.y * g4.y, ez = d.z * g4.z, ew = d.w * g4.w;    // dy * gamma
        float c1 = (ex + ey) + (ez + ew), c2 = (ex * hx + ey * hy) + (ez * hz + ew * hw);
#pragma unroll
        for (int o = LPR / 2; o > 0; o >>= 1) {
            c1 += __shfl_xor_sync(0xffffffffu, c1, o);
            c2 += __shfl_xor_sync(0xffffffffu, c2, o);
        }
        c1 *= 1.0f / C; c2 *= 1.0f / C;
        if (live) {
            *((float4 *)(gx + (size_t)r * C) + l) = make_float4(rstd * (ex - c1 - hx * c2), rstd * (ey - c1 - hy * c2),
                                                                rstd * (ez - c1 - hz * c2), rstd * (ew - c1 - hw * c2));
            ag.x = fmaf(d.x, hx, ag.x); ag.y = fmaf(d.y, hy, ag.y); ag.z = fmaf(d.z, hz, ag.z); ag.w = fmaf(d.w, hw, ag.w);
            abt.x += d.x; abt.y += d.y; abt.z += d.z; abt.w += d.w;
        }
    }
    atomicAdd(s_acc + 4 * l, ag.x); atomicAdd(s_acc + 4 * l + 1, ag.y); atomicAdd(s_acc + 4 * l + 2, ag.z); atomicAdd(s_acc + 4 * l + 3, ag.w);
    atomicAdd(s_acc + C + 4 * l, abt.x); atomicAdd(s_acc + C + 4 * l + 1, abt.y);
    atomicAdd(s_acc + C + 4 * l + 2, abt.z); atomicAdd(s_acc + C + 4 * l + 3, abt.w);
    __syncthreads();
    for (int i = threadIdx.x; i < 2 * C; i += blockDim.x) atomicAdd((i < C ? ggamma : gbeta - C) + i, s_acc[i]);
}

template <int HD>
static int ragged_fwd(int heads, float scale, int nq, const int *q_win, const int *key_off, const int *key_mult,
                      RaggedRows R, float *out, int ldo, float *lse, cudaStream_t s) {
    const long long items = (long long)nq * heads;
    k_ragged_attn_fwd<HD><<<div_up(items, 128), 128, 0, s>>>(nq, heads, scale, q_win, key_off, key_mult, R, out, ldo, lse);
    ++g_launches;
    return check_launch();
}

template <int HD>
static int ragged_bwd(int heads, float scale, int nq, int nk, const int *q_win, const int *k_win, const int *q_off,
                      const int *key_off, const int *key_mult, RaggedRows R, const float *out, int ldo, const float *lse,
                      const float *gout, int ldgo, float *delta, float *gq, int ldgq, float *gk, int ldgk, float *gv,
                      int ldgv, cudaStream_t s) {
    if (nq > 0) {
        k_ragged_attn_bwd_q<HD><<<div_up((long long)nq * heads, 128), 128, 0, s>>>(
            nq, heads, scale, q_win, key_off, key_mult, R, out, ldo, lse, gout, ldgo, delta, gq, ldgq);
        ++g_launches;
    }
    if (nk > 0) {
        k_ragged_attn_bwd_kv<HD><<<div_up((long long)nk * heads, 128), 128, 0, s>>>(
            nk, heads, scale, k_win, q_off, key_off, key_mult, R, lse, delta, gout, ldgo, gk, ldgk, gv, ldgv);
        ++g_launches;
    }
    return check_launch();
}

static bool rows_ok(const void *p, int ld) { return p && ((uintptr_t)p & 15u) == 0 && ld > 0 && ld % 4 == 0; }

}  // namespace mssvt

using namespace mssvt;

extern "C" {

int mssvt_ragged_attention_fwd(int heads, int head_dim, float scale, int num_queries, const int *q_win,
                               const int *key_off, const int *key_mult, const float *q, int ldq, const float *k,
                               int ldk, const float *v, int ldv, float *out, int ldo, float *lse, void *stream) {
    if (heads <= 0 || num_queries < 0) return MSSVT_ERR_INVALID;
    if (num_queries == 0) return MSSVT_OK;
    if (!q_win || !key_off || !key_mult || !lse || !rows_ok(q, ldq) || !rows_ok(k, ldk) || !rows_ok(v, ldv) ||
        !rows_ok(out, ldo))
        return MSSVT_ERR_INVALID;
    const RaggedRows R = {q, k, v, ldq, ldk, ldv};
    cudaStream_t s = (cudaStream_t)stream;
    switch (head_dim) {
    case 8: return ragged_fwd<8>(heads, scale, num_queries, q_win, key_off, key_mult, R, out, ldo, lse, s);
    case 16: return ragged_fwd<16>(heads, scale, num_queries, q_win, key_off, key_mult, R, out, ldo, lse, s);
    case 32: return ragged_fwd<32>(heads, scale, num_queries, q_win, key_off, key_mult, R, out, ldo, lse, s);
    default: return MSSVT_ERR_INVALID;
    }
}

int mssvt_ragged_attention_bwd(int heads, int head_dim, float scale, int num_queries, int num_keys, const int *q_win,
                               const int *k_win, const int *q_off, const int *key_off, const int *key_mult,
                               const float *q, int ldq, const float *k, int ldk, const float *v, int ldv,
                               const float *out, int ldo, const float *lse, const float *grad_out, int ldgo,
                               float *delta, float *grad_q, int ldgq, float *grad_k, int ldgk, float *grad_v,
                               int ldgv, void *stream) {
    if (heads <= 0 || num_queries < 0 || num_keys < 0) return MSSVT_ERR_INVALID;
    if (num_queries == 0 && num_keys == 0) return MSSVT_OK;
    if (!q_win || !k_win || !q_off || !key_off || !key_mult || !lse || !delta || !rows_ok(q, ldq) || !rows_ok(k, ldk) ||
        !rows_ok(v, ldv) || !rows_ok(out, ldo) || !rows_ok(grad_out, ldgo) || !rows_ok(grad_q, ldgq) ||
        !rows_ok(grad_k, ldgk) || !rows_ok(grad_v, ldgv))
        return MSSVT_ERR_INVALID;
    const RaggedRows R = {q, k, v, ldq, ldk, ldv};
    cudaStream_t s = (cudaStream_t)stream;
#define MSSVT_RAGGED_BWD(HD)                                                                                          \
    return ragged_bwd<HD>(heads, scale, num_queries, num_keys, q_win, k_win, q_off, key_off, key_mult, R, out, ldo, \
                          lse, grad_out, ldgo, delta, grad_q, ldgq, grad_k, ldgk, grad_v, ldgv, s)
    switch (head_dim) {
    case 8: MSSVT_RAGGED_BWD(8);
    case 16: MSSVT_RAGGED_BWD(16);
    case 32: MSSVT_RAGGED_BWD(32);
    default: return MSSVT_ERR_INVALID;
    }
#undef MSSVT_RAGGED_BWD
}

int mssvt_interp_merge_fwd(int num_voxels, int C, const int *src, const float *weights, const float *rows,
                           const float *x, float *out, void *stream) {
    if (num_voxels < 0 || C <= 0 || C % 4) return MSSVT_ERR_INVALID;
    if (num_voxels == 0) return MSSVT_OK;
    if (!src || !weights || !x || !out) return MSSVT_ERR_INVALID;   // (rows may be empty: every src < 0 then)
    const long long items = (long long)num_voxels * (C / 4);
    k_interp_merge_fwd<<<div_up(items, 256), 256, 0, (cudaStream_t)stream>>>(
        num_voxels, C / 4, src, weights, (const float4 *)rows, (const float4 *)x, (float4 *)out);
    ++g_launches;
    return check_launch();
}

int mssvt_interp_merge_bwd(int num_voxels, int C, int num_rows, const int *src, const float *weights,
                           const float *grad_out, float *grad_rows, float *grad_x, void *stream) {
    if (num_voxels < 0 || num_rows < 0 || C <= 0 || C % 4) return MSSVT_ERR_INVALID;
    cudaStream_t s = (cudaStream_t)stream;
    if (num_rows > 0) {
        if (!grad_rows) return MSSVT_ERR_INVALID;
        if (cudaMemsetAsync(grad_rows, 0, (size_t)num_rows * C * sizeof(float), s) != cudaSuccess) return check_launch();
    }
    if (num_voxels == 0) return MSSVT_OK;
    if (!src || !weights || !grad_out || !grad_x) return MSSVT_ERR_INVALID;
    const long long items = (long long)num_voxels * (C / 4);
    k_interp_merge_bwd<<<div_up(items, 256), 256, 0, s>>>(num_voxels, C / 4, src, weights, (const float4 *)grad_out,
                                                          (float4 *)grad_rows, (float4 *)grad_x);
    ++g_launches;
    return check_launch();
}

int mssvt_embed_rows_fwd(int num_rows, int c0, int cs, int C, const int *rows, const int *win,
                         const unsigned char *masked, const float *xn, const float *xyz, const float *centre,
                         const float *pos_w, const float *pos_b, float *out, void *stream) {
    if (num_rows < 0 || c0 < 0 || (c0 & 3) || (C & 3) || c0 + cs > C || (cs != 32 && cs != 64)) return MSSVT_ERR_INVALID;
    if (num_rows == 0) return MSSVT_OK;
    if (!rows || !win || !out || (!xn && !pos_w) || (pos_w && (!xyz || !centre || !pos_b))) return MSSVT_ERR_INVALID;
    const EmbedRows E = {rows, win, masked, xyz, centre, pos_w, pos_b, num_rows, c0, C};
    const int grid = persistent_grid((long long)num_rows * (cs / 4), 256, 8, 2);
    if (cs == 32) k_embed_rows_fwd<8><<<grid, 256, 0, (cudaStream_t)stream>>>(E, xn, out);
    else k_embed_rows_fwd<16><<<grid, 256, 0, (cudaStream_t)stream>>>(E, xn, out);
    ++g_launches;
    return check_launch();
}

int mssvt_embed_rows_bwd(int num_rows, int c0, int cs, int C, const int *rows, const int *win,
                         const unsigned char *masked, const float *xyz, const float *centre, const float *pos_w,
                         const float *pos_b, const float *grad_out, float *grad_xn, float *grad_w, float *grad_b,
                         void *stream) {
    if (num_rows < 0 || c0 < 0 || (c0 & 3) || (C & 3) || c0 + cs > C || (cs != 32 && cs != 64)) return MSSVT_ERR_INVALID;
    if (num_rows == 0) return MSSVT_OK;
    if (!rows || !win || !grad_out || (!grad_xn && !pos_w) || (pos_w && (!xyz || !centre || !pos_b || !grad_w || !grad_b)))
        return MSSVT_ERR_INVALID;
    const EmbedRows E = {rows, win, masked, xyz, centre, pos_w, pos_b, num_rows, c0, C};
    const int grid = persistent_grid((long long)num_rows * (cs / 4), 256, 8, 1);
    if (cs == 32) k_embed_rows_bwd<8><<<grid, 256, 0, (cudaStream_t)stream>>>(E, grad_out, grad_xn, grad_w, grad_b);
    else k_embed_rows_bwd<16><<<grid, 256, 0, (cudaStream_t)stream>>>(E, grad_out, grad_xn, grad_w, grad_b);
    ++g_launches;
    return check_launch();
}

int mssvt_layernorm_bwd(int num_rows, int C, const float *x, const float *gamma, float eps, const float *grad_y,
                        float *grad_x, float *grad_gamma, float *grad_beta, void *stream) {
    if (num_rows < 0 || (C != 64 && C != 128)) return MSSVT_ERR_INVALID;
    if (!grad_gamma || !grad_beta) return MSSVT_ERR_INVALID;
    cudaStream_t s = (cudaStream_t)stream;
    if (cudaMemsetAsync(grad_gamma, 0, C * sizeof(float), s) != cudaSuccess ||
        cudaMemsetAsync(grad_beta, 0, C * sizeof(float), s) != cudaSuccess)
        return check_launch();
    if (num_rows == 0) return MSSVT_OK;
    if (!x || !gamma || !grad_y || !grad_x) return MSSVT_ERR_INVALID;
    const int grid = persistent_grid((long long)num_rows * (C / 4), 256, 8, 1);
    if (C == 64) k_layernorm_bwd<16><<<grid, 256, 0, s>>>(num_rows, x, gamma, eps, grad_y, grad_x, grad_gamma, grad_beta);
    else k_layernorm_bwd<32><<<grid, 256, 0, s>>>(num_rows, x, gamma, eps, grad_y, grad_x, grad_gamma, grad_beta);
    ++g_launches;
    return check_launch();
}

int mssvt_segment_max_fwd(int num_windows, int C, const int *key_off, const float *rows, float *out, int *arg,
                          void *stream) {
    if (num_windows < 0 || C <= 0 || (C & 3)) return MSSVT_ERR_INVALID;
    if (num_windows == 0) return MSSVT_OK;
    if (!key_off || !rows || !out || !arg) return MSSVT_ERR_INVALID;
    k_segment_max_fwd<<<div_up((long long)num_windows * (C / 4), 256), 256, 0, (cudaStream_t)stream>>>(
        num_windows, C / 4, key_off, (const float4 *)rows, (float4 *)out, (int4 *)arg);
    ++g_launches;
    return check_launch();
}

int mssvt_segment_max_bwd(int num_rows, int C, const int *k_win, const int *arg, const float *grad_out,
                          float *grad_rows, void *stream) {
    if (num_rows < 0 || C <= 0 || (C & 3)) return MSSVT_ERR_INVALID;
    if (num_rows == 0) return MSSVT_OK;
    if (!k_win || !arg || !grad_out || !grad_rows) return MSSVT_ERR_INVALID;
    k_segment_max_bwd<<<div_up((long long)num_rows * (C / 4), 256), 256, 0, (cudaStream_t)stream>>>(
        num_rows, C / 4, k_win, (const int4 *)arg, (const float4 *)grad_out, (float4 *)grad_rows);
    ++g_launches;
    return check_launch();
}

}  // extern "C"
