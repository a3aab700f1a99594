// block_common.cuh -- descriptors and warp-level building blocks shared by block.cu / attention.cu
#pragma once
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

// exp(x) for softmax arguments (x <= 0 after the running max is subtracted): one FMUL + ex2.approx.
// Relative error <= 2 ulp of ex2 + |x| * 2^-24 from the scaling, far inside the 1e-4 feature budget;
// the full-range expf expands to ~25 instructions per call site and bloated the window kernels.
__device__ __forceinline__ float exp_neg(float x) { return exp2f(x * 1.4426950408889634f); }

// Flat fp32 parameter pack of one block, built once by the host module (offsets in floats).
// All matrices are stored TRANSPOSED ([in][out]) so that lanes, which own outputs, read
// consecutive words.  Mirrored field by field in mssvt_b200/_lib.py.
struct AttnShape {
    int C, G, hd, nq, nk_total, nk, cap1, interp, pos_layers;
    int heads[MAX_GROUPS], sd[MAX_GROUPS], c0[MAX_GROUPS];
    int off_pos_w, off_pos_b;            // [6][C], [C] (bias directly behind the matrix)
    int off_pos2_w, off_pos2_b;          // [C][C], [C]   (one-window blocks only)
    int off_wq[MAX_GROUPS], off_bq[MAX_GROUPS];     // [sd][sd], [sd]
    int off_wkv[MAX_GROUPS], off_bkv[MAX_GROUPS];   // [sd][2sd], [2sd]
    int off_wp[MAX_GROUPS], off_bp[MAX_GROUPS];     // [sd][sd], [sd]
    int total_floats;
    float scale;
    float win_cell[3];  // window size in metres (fp32 of the python double vs * ws)
    float lo[3];
};

struct FfnShape {
    int C, F, C_out;       // in, hidden, out (C_out == 0: no out_linear)
    int mode;              // 0: u = merged (compress block); 1: u = covered ? merged + x : 2 x
    int off_ln_g, off_ln_b, off_w1, off_b1, off_w2, off_b2, off_wo, off_bo;  // transposed mats
    int total_floats;
    float eps;
};

__device__ __forceinline__ float f4_get(const float4 &v, int k) {
    return k == 0 ? v.x : k == 1 ? v.y : k == 2 ? v.z : v.w;
}

// out[r][o] = b[o] + sum_i Wt[i][o] * in[r * ld_in + i] for R rows at once.
//   wt, b, in: shared memory; `in` rows 16-byte aligned, n_in % 4 == 0.
//   A lane owns outputs lane, lane+32, ... (OPL of them) of each 32*OPL-wide chunk of outputs: per
//   4 inputs it issues R broadcast LDS.128 for the rows + 4*OPL LDS.32 for the weights and 4*OPL*R
//   FMAs, so the weight traffic is amortised over R rows and every lane carries OPL*R independent
//   FMA chains.  The loop over inputs is deliberately NOT unrolled: the body (~100 instructions)
//   stays in the instruction cache; a fully unrolled version made the window kernels
//   instruction-fetch bound (profiles/r01_b: stall_no_instruction 4.3 per issue, 152 KB of SASS).
template <int R, int OPL, typename Epi>
__device__ __forceinline__ void dense_rows(const float *wt, const float *b, const float *in, int ld_in,
                                           int n_in, int n_out, Epi epi) {
    const int lane = threadIdx.x & 31;
    for (int ob = 0; ob < n_out; ob += 32 * OPL) {
        float a[OPL][R];
        int oc[OPL];
#pragma unroll
        for (int j = 0; j < OPL; ++j) {
            oc[j] = min(ob + lane + 32 * j, n_out - 1);  // surplus lanes recompute the last output
            const float bv = b[oc[j]];
#pragma unroll
            for (int r = 0; r < R; ++r) a[j][r] = bv;
        }
        const float *wrow = wt;
#pragma unroll 1
        for (int i = 0; i < n_in; i += 4, wrow += 4 * n_out) {
            float4 v[R];
#pragma unroll
            for (int r = 0; r < R; ++r) v[r] = *(const float4 *)(in + r * ld_in + i);
#pragma unroll
            for (int k = 0; k < 4; ++k) {
#pragma unroll
                for (int j = 0; j < OPL; ++j) {
                    const float wv = wrow[k * n_out + oc[j]];
#pragma unroll
                    for (int r = 0; r < R; ++r) a[j][r] = fmaf(wv, f4_get(v[r], k), a[j][r]);
                }
            }
        }
#pragma unroll
        for (int j = 0; j < OPL; ++j) {
            const int o = ob + lane + 32 * j;
            if (o < n_out) {
#pragma unroll
                for (int r = 0; r < R; ++r) epi(r, o, a[j][r]);
            }
        }
    }
}

// Epilogue flags of dense_store: a = b[o] + W x for row r, output o
//   v = RELU ? max(a, 0) : a * mul;  v += ADD_DST ? dst[r][o] : 0;  v += ADD_SRC ? src2[r][o] : 0
enum DenseMode { DENSE_SET = 0, DENSE_RELU = 1, DENSE_ADD_DST = 2, DENSE_ADD_SRC = 4 };

// Out-of-line so that each (R, OPL) variant exists once per kernel: the window kernels call it from
// several places and an inlined copy per call site blew the instruction cache (see dense_rows).
template <int R, int OPL>
__device__ __noinline__ void dense_store_impl(const float *wt, const float *b, const float *in, int ld_in,
                                              int n_in, int n_out, float *dst, int ld_out, int valid,
                                              float mul, int mode, const float *src2) {
    const bool relu = mode & DENSE_RELU, add_dst = mode & DENSE_ADD_DST, add_src = mode & DENSE_ADD_SRC;
    dense_rows<R, OPL>(wt, b, in, ld_in, n_in, n_out, [&](int r, int o, float a) {
        if (r < valid) {
            float *d = dst + r * ld_out + o;
            float v = relu ? fmaxf(a, 0.f) : a * mul;
            if (add_dst) v += *d;
            if (add_src) v += src2[r * ld_out + o];
            *d = v;
        }
    });
}

// outputs-per-lane chosen from the runtime width: <= 32 -> 1, <= 64 -> 2, else 4 per 128-chunk
template <int R>
__device__ __forceinline__ void dense_store(const float *wt, const float *b, const float *in, int ld_in,
                                            int n_in, int n_out, float *dst, int ld_out, int valid,
                                            float mul = 1.0f, int mode = DENSE_SET,
                                            const float *src2 = nullptr) {
    if (n_out <= 32) dense_store_impl<R, 1>(wt, b, in, ld_in, n_in, n_out, dst, ld_out, valid, mul, mode, src2);
    else if (n_out <= 64) dense_store_impl<R, 2>(wt, b, in, ld_in, n_in, n_out, dst, ld_out, valid, mul, mode, src2);
    else dense_store_impl<R, 4>(wt, b, in, ld_in, n_in, n_out, dst, ld_out, valid, mul, mode, src2);
}

// positional embedding channel c of pos_proj layer 1: ReLU(b[c] + W[c, 0:3] . rel + W[c, 3:6] . ctr)
// s_w: [6][C] transposed weight, bias stored right behind it
__device__ __forceinline__ float pos_embed(const float *s_w, int C, int c, float rx, float ry,
                                           float rz, float cx, float cy, float cz) {
    float a = s_w[6 * C + c];
    a = fmaf(s_w[0 * C + c], rx, a);
    a = fmaf(s_w[1 * C + c], ry, a);
    a = fmaf(s_w[2 * C + c], rz, a);
    a = fmaf(s_w[3 * C + c], cx, a);
    a = fmaf(s_w[4 * C + c], cy, a);
    a = fmaf(s_w[5 * C + c], cz, a);
    return fmaxf(a, 0.f);
}

}  // namespace mssvt
