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
#include "block_common.cuh"

namespace mssvt {

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

// C = 64: 16 lanes per row (one float4 each), 8 rows per warp and pass with all loads issued before the
// first reduction: 2 KB in flight per warp instead of 256 bytes (the one-row version ran at 2.1 TB/s)
__global__ void __launch_bounds__(256)
k_layernorm64(int n_cap, const int *__restrict__ n_dev, const float *__restrict__ x,
              const float *__restrict__ gamma, const float *__restrict__ beta, float eps, float *__restrict__ y) {
    pdl_launch_dependents();
    pdl_wait();
    const int n = n_dev ? min(n_cap, __ldg(n_dev)) : n_cap;
    const int lane = threadIdx.x & 31, sub = lane >> 4, c4 = lane & 15;
    const int warps = (gridDim.x * blockDim.x) >> 5;
    const float4 g = __ldg((const float4 *)gamma + c4), b = __ldg((const float4 *)beta + c4);
    for (int base = ((blockIdx.x * blockDim.x + threadIdx.x) >> 5) * 8; base < n; base += warps * 8) {
        float4 v[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int row = base + 2 * i + sub;
            v[i] = row < n ? __ldg((const float4 *)(x + (size_t)row * 64) + c4) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int row = base + 2 * i + sub;
            float s = (v[i].x + v[i].y) + (v[i].z + v[i].w);
#pragma unroll
            for (int o = 8; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
            const float mean = s * (1.0f / 64.0f);
            const float dx = v[i].x - mean, dy = v[i].y - mean, dz = v[i].z - mean, dw = v[i].w - mean;
            float q = (dx * dx + dy * dy) + (dz * dz + dw * dw);
#pragma unroll
            for (int o = 8; o > 0; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
            const float rstd = rsqrtf(q * (1.0f / 64.0f) + eps);
            if (row < n)
                *((float4 *)(y + (size_t)row * 64) + c4) =
                    make_float4(dx * rstd * g.x + b.x, dy * rstd * g.y + b.y, dz * rstd * g.z + b.z, dw * rstd * g.w + b.w);
        }
    }
}

// ------------------------------------------------------------------------------- FFN

#define FFN_WARPS 8
#define FFN_ROWS 4  // rows per warp per pass

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
        const int live_rows = min(FFN_ROWS, n - r0);
        dense_store<FFN_ROWS>(s_par + S.off_w1, s_par + S.off_b1, s_n, C, C, F, s_h, F, FFN_ROWS, 1.0f, DENSE_RELU);
        __syncwarp();
        if (S.C_out == 0) {
            // y = u + W2 h + b2 straight to global memory
            dense_store<FFN_ROWS>(s_par + S.off_w2, s_par + S.off_b2, s_h, F, F, C, y + (size_t)r0 * C, C,
                                  live_rows, 1.0f, DENSE_ADD_SRC, s_u);
        } else {
            dense_store<FFN_ROWS>(s_par + S.off_w2, s_par + S.off_b2, s_h, F, F, C, s_n, C, FFN_ROWS, 1.0f,
                                  DENSE_ADD_SRC, s_u);
            __syncwarp();
            dense_store<FFN_ROWS>(s_par + S.off_wo, s_par + S.off_bo, s_n, C, C, S.C_out,
                                  y + (size_t)r0 * S.C_out, S.C_out, live_rows);
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

extern "C" {

int mssvt_layernorm(int num_rows, const int *num_rows_dev, int C, const float *x,
                    const float *gamma, const float *beta, float eps, float *y, void *stream) {
    if (num_rows < 0 || C <= 0 || C > 256) return MSSVT_ERR_INVALID;
    if (num_rows == 0) return MSSVT_OK;
    if (!x || !gamma || !beta || !y) return MSSVT_ERR_INVALID;
    ++g_launches;
    if (C == 64)
        launch_pdl(k_layernorm64, dim3(persistent_grid(num_rows, 64, 8)), dim3(256), 0, (cudaStream_t)stream, 
            num_rows, num_rows_dev, x, gamma, beta, eps, y);
    else
        k_layernorm<<<persistent_grid(num_rows, 8, 8), 256, 0, (cudaStream_t)stream>>>(
            num_rows, num_rows_dev, C, x, gamma, beta, eps, y);
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

int mssvt_sizeof_ffn_shape(void) { return (int)sizeof(FfnShape); }

}  // extern "C"
