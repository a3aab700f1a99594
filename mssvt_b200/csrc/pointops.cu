// pointops.cu -- the generic (op-level) gather / sampling operators of the hot path (sm_100a).
//
//   mssvt_group_features(_grad) <- group_features(_grad)_kernel_stack
//                                  pcdet/ops/mssvt/src/group_features_gpu.cu:73-129 / 15-70
//   mssvt_fps                   <- farthest_point_sampling_kernel  pointnet2_batch/src/sampling_gpu.cu:100-260
//   mssvt_gather_points         <- gather_points_kernel_fast       pointnet2_batch/src/sampling_gpu.cu:15-51
//   mssvt_three_nn              <- three_nn_kernel_fast            pointnet2_batch/src/interpolate_gpu.cu:16-81
//   mssvt_group_points(_grad)   <- group_points(_grad)_kernel_fast pointnet2_batch/src/group_points_gpu.cu:53-92 / 14-50
//
// The fused backbone path does not call these (it never materialises the padded tensors); they
// exist so that code written against the reference's operator API keeps working, and they are
// what the K/V gather micro-benchmark (BASELINE config 3) measures.
#include "common.cuh"

namespace mssvt {

// ------------------------------------------------------------------------------- group_features

#define GF_THREADS 256
#define GF_STILE 32  // samples staged per pass: each output run out[m, c, s0:s0+32] is 128 B

__device__ __forceinline__ int stacked_row_start(int m, int B, const int *__restrict__ idx_batch_cnt,
                                                 const int *__restrict__ features_batch_cnt) {
    // sample of idx row m (group_features_gpu.cu:91-99), then the first feature row of that sample
    int b = 0, upto = __ldg(idx_batch_cnt);
    for (int k = 1; k < B; ++k) {
        if (m < upto) break;
        upto += __ldg(idx_batch_cnt + k);
        b = k;
    }
    int start = 0;
    for (int k = 0; k < b; ++k) start += __ldg(features_batch_cnt + k);
    return start;
}

// One CTA per idx row m.  Feature rows are read whole (coalesced, 16-byte vectors when C % 4 == 0)
// into a padded shared tile, then written channel-major: consecutive lanes write consecutive s.
__global__ void __launch_bounds__(GF_THREADS)
k_group_features(int B, int M, int C, int ns, const float *__restrict__ features,
                 const int *__restrict__ features_batch_cnt, const int *__restrict__ idx,
                 const int *__restrict__ idx_batch_cnt, float *__restrict__ out) {
    extern __shared__ float tile[];  // GF_STILE x (C + 1)
    __shared__ int s_rows[GF_STILE];
    __shared__ int s_start;
    const int pitch = C + 1;
    for (int m = blockIdx.x; m < M; m += gridDim.x) {
        if (threadIdx.x == 0) s_start = stacked_row_start(m, B, idx_batch_cnt, features_batch_cnt);
        __syncthreads();
        for (int s0 = 0; s0 < ns; s0 += GF_STILE) {
            const int cur = min(GF_STILE, ns - s0);
            if (threadIdx.x < cur) {
                int v = __ldg(idx + (size_t)m * ns + s0 + threadIdx.x);
                s_rows[threadIdx.x] = v < 0 ? -1 : s_start + v;
            }
            __syncthreads();
            if ((C & 3) == 0) {
                const int c4 = C >> 2;
                for (int e = threadIdx.x; e < cur * c4; e += blockDim.x) {
                    int s = e / c4, q = e - s * c4;
                    int row = s_rows[s];
                    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (row >= 0) v = __ldg((const float4 *)(features + (size_t)row * C) + q);
                    float *t = tile + s * pitch + 4 * q;
                    t[0] = v.x; t[1] = v.y; t[2] = v.z; t[3] = v.w;
                }
            } else {
                for (int e = threadIdx.x; e < cur * C; e += blockDim.x) {
                    int s = e / C, c = e - s * C;
                    int row = s_rows[s];
                    tile[s * pitch + c] = row >= 0 ? __ldg(features + (size_t)row * C + c) : 0.f;
                }
            }
            __syncthreads();
            float *dst = out + (size_t)m * C * ns + s0;
            for (int e = threadIdx.x; e < C * GF_STILE; e += blockDim.x) {
                int c = e / GF_STILE, s = e - c * GF_STILE;
                if (s < cur) dst[(size_t)c * ns + s] = tile[s * pitch + c];
            }
            __syncthreads();
        }
    }
}

__global__ void __launch_bounds__(GF_THREADS)
k_group_features_grad(int B, int M, int C, int ns, const float *__restrict__ grad_out,
                      const int *__restrict__ idx, const int *__restrict__ idx_batch_cnt,
                      const int *__restrict__ features_batch_cnt, float *__restrict__ grad_features) {
    extern __shared__ float tile[];  // GF_STILE x (C + 1)
    __shared__ int s_rows[GF_STILE];
    __shared__ int s_start;
    const int pitch = C + 1;
    for (int m = blockIdx.x; m < M; m += gridDim.x) {
        if (threadIdx.x == 0) s_start = stacked_row_start(m, B, idx_batch_cnt, features_batch_cnt);
        __syncthreads();
        for (int s0 = 0; s0 < ns; s0 += GF_STILE) {
            const int cur = min(GF_STILE, ns - s0);
            if (threadIdx.x < cur) {
                int v = __ldg(idx + (size_t)m * ns + s0 + threadIdx.x);
                s_rows[threadIdx.x] = v < 0 ? -1 : s_start + v;
            }
            const float *src = grad_out + (size_t)m * C * ns + s0;
            for (int e = threadIdx.x; e < C * GF_STILE; e += GF_THREADS) {
                int c = e / GF_STILE, s = e - c * GF_STILE;
                if (s < cur) tile[s * pitch + c] = __ldg(src + (size_t)c * ns + s);
            }
            __syncthreads();
            for (int e = threadIdx.x; e < cur * C; e += GF_THREADS) {
                int s = e / C, c = e - s * C;
                int row = s_rows[s];
                if (row >= 0) atomicAdd(grad_features + (size_t)row * C + c, tile[s * pitch + c]);
            }
            __syncthreads();
        }
    }
}

// ------------------------------------------------------------------------------- FPS (float)

// 64-bit selection key, maximised: [ fp32 bits of min-dist | B-1-bitrev(k mod B) | 2^21-1-k ].
// It reproduces the reference block reduction's tie order for any block size B.
__device__ __forceinline__ unsigned fps_tie(int k, int log2b) {
    unsigned B1 = (1u << log2b) - 1u;
    unsigned rev = log2b ? (__brev((unsigned)k & B1) >> (32 - log2b)) : 0u;
    return ((B1 - rev) << 21) | (unsigned)(0x1fffff - k);
}

// the reference's contraction of (x2-x1)^2 + (y2-y1)^2 + (z2-z1)^2 under nvcc (checked in SASS)
__device__ __forceinline__ float fps_dist(float x, float y, float z, float x1, float y1, float z1) {
    float dx = __fsub_rn(x, x1), dy = __fsub_rn(y, y1), dz = __fsub_rn(z, z1);
    return __fmaf_rn(dz, dz, __fmaf_rn(dx, dx, __fmul_rn(dy, dy)));
}

#define FPS_WARPS 4
#define FPS_WARP_MAXN 256

// rows with n <= 256: one warp per row, points and running min-distance in shared memory
__global__ void __launch_bounds__(FPS_WARPS * 32)
k_fps_warp(int b, int n, int m, int log2b, const float *__restrict__ dataset, int *__restrict__ idxs) {
    extern __shared__ float sm[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float *pts = sm + (size_t)warp * n * 4;
    float *tmin = pts + 3 * n;
    for (int row = blockIdx.x * FPS_WARPS + warp; row < b; row += gridDim.x * FPS_WARPS) {
        const float *src = dataset + (size_t)row * n * 3;
        for (int i = lane; i < 3 * n; i += 32) pts[i] = __ldg(src + i);
        for (int i = lane; i < n; i += 32) tmin[i] = 1e10f;
        int *out = idxs + (size_t)row * m;
        if (lane == 0) out[0] = 0;
        __syncwarp();
        int old = 0;
        for (int j = 1; j < m; ++j) {
            float x1 = pts[3 * old], y1 = pts[3 * old + 1], z1 = pts[3 * old + 2];
            unsigned hi = 0, lo = 0;
            for (int k = lane; k < n; k += 32) {
                float d = fminf(fps_dist(pts[3 * k], pts[3 * k + 1], pts[3 * k + 2], x1, y1, z1), tmin[k]);
                tmin[k] = d;
                unsigned h = __float_as_uint(d), l = fps_tie(k, log2b);
                if (h > hi || (h == hi && l > lo)) { hi = h; lo = l; }
            }
            unsigned top = __reduce_max_sync(0xffffffffu, hi);
            lo = __reduce_max_sync(0xffffffffu, hi == top ? lo : 0u);
            old = 0x1fffff - (int)(lo & 0x1fffffu);
            if (lane == 0) out[j] = old;
        }
        __syncwarp();
    }
}

#define FPS_CTA 512

// larger rows: one CTA per row; running min-distance lives in `temp` (global, caller-provided,
// as in the reference API) unless the row fits shared memory
__global__ void __launch_bounds__(FPS_CTA)
k_fps_cta(int n, int m, int log2b, int in_smem, const float *__restrict__ dataset,
          float *__restrict__ temp, int *__restrict__ idxs) {
    extern __shared__ float sm[];
    __shared__ unsigned s_hi[FPS_CTA / 32], s_lo[FPS_CTA / 32];
    __shared__ int s_old;
    const int row = blockIdx.x, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const float *pts = dataset + (size_t)row * n * 3;
    float *tmin = in_smem ? sm : temp + (size_t)row * n;
    for (int i = threadIdx.x; i < n; i += FPS_CTA) tmin[i] = 1e10f;
    int *out = idxs + (size_t)row * m;
    if (threadIdx.x == 0) { out[0] = 0; s_old = 0; }
    __syncthreads();
    for (int j = 1; j < m; ++j) {
        int old = s_old;
        float x1 = __ldg(pts + 3 * old), y1 = __ldg(pts + 3 * old + 1), z1 = __ldg(pts + 3 * old + 2);
        unsigned hi = 0, lo = 0;
        for (int k = threadIdx.x; k < n; k += FPS_CTA) {
            float d = fminf(fps_dist(__ldg(pts + 3 * k), __ldg(pts + 3 * k + 1), __ldg(pts + 3 * k + 2), x1, y1, z1), tmin[k]);
            tmin[k] = d;
            unsigned h = __float_as_uint(d), l = fps_tie(k, log2b);
            if (h > hi || (h == hi && l > lo)) { hi = h; lo = l; }
        }
        unsigned top = __reduce_max_sync(0xffffffffu, hi);
        lo = __reduce_max_sync(0xffffffffu, hi == top ? lo : 0u);
        if (lane == 0) { s_hi[warp] = top; s_lo[warp] = lo; }
        __syncthreads();
        if (warp == 0) {
            unsigned h = lane < FPS_CTA / 32 ? s_hi[lane] : 0u, l = lane < FPS_CTA / 32 ? s_lo[lane] : 0u;
            unsigned t2 = __reduce_max_sync(0xffffffffu, h);
            l = __reduce_max_sync(0xffffffffu, h == t2 ? l : 0u);
            if (lane == 0) {
                int pick = 0x1fffff - (int)(l & 0x1fffffu);
                s_old = pick;
                out[j] = pick;
            }
        }
        __syncthreads();
    }
}

// ------------------------------------------------------------------------------- small gathers

__global__ void k_gather_points(int b, int c, int n, int m, const float *__restrict__ points,
                                const int *__restrict__ idx, float *__restrict__ out) {
    size_t total = (size_t)b * c * m;
    for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < total;
         e += (size_t)gridDim.x * blockDim.x) {
        int j = (int)(e % m);
        size_t bc = e / m;
        int bi = (int)(bc / c);
        out[e] = __ldg(points + bc * n + __ldg(idx + (size_t)bi * m + j));
    }
}

__global__ void k_group_points(int b, int c, int n, int npoints, int nsample,
                               const float *__restrict__ points, const int *__restrict__ idx,
                               float *__restrict__ out) {
    const size_t per = (size_t)npoints * nsample, total = (size_t)b * c * per;
    for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < total;
         e += (size_t)gridDim.x * blockDim.x) {
        size_t bc = e / per, r = e - bc * per;
        int bi = (int)(bc / c);
        out[e] = __ldg(points + bc * n + __ldg(idx + (size_t)bi * per + r));
    }
}

__global__ void k_group_points_grad(int b, int c, int n, int npoints, int nsample,
                                    const float *__restrict__ grad_out, const int *__restrict__ idx,
                                    float *__restrict__ grad_points) {
    const size_t per = (size_t)npoints * nsample, total = (size_t)b * c * per;
    for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < total;
         e += (size_t)gridDim.x * blockDim.x) {
        size_t bc = e / per, r = e - bc * per;
        int bi = (int)(bc / c);
        atomicAdd(grad_points + bc * n + __ldg(idx + (size_t)bi * per + r), __ldg(grad_out + e));
    }
}

// three nearest known points of every unknown point; distances evaluated with the reference's
// FMA contraction, compared with strict < in index order (lowest index wins ties)
__global__ void k_three_nn(int b, int n, int m, const float *__restrict__ unknown,
                           const float *__restrict__ known, float *__restrict__ dist2,
                           int *__restrict__ idx) {
    int bi = blockIdx.y;
    int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (bi >= b || p >= n) return;
    const float *u = unknown + ((size_t)bi * n + p) * 3;
    const float *kn = known + (size_t)bi * m * 3;
    float ux = __ldg(u), uy = __ldg(u + 1), uz = __ldg(u + 2);
    const float INF = __int_as_float(0x7f800000);
    float b1 = INF, b2 = INF, b3 = INF;  // the reference starts from double 1e40 (> FLT_MAX)
    int i1 = 0, i2 = 0, i3 = 0;
    for (int k = 0; k < m; ++k) {
        float dx = __fsub_rn(ux, __ldg(kn + 3 * k)), dy = __fsub_rn(uy, __ldg(kn + 3 * k + 1));
        float dz = __fsub_rn(uz, __ldg(kn + 3 * k + 2));
        float d = __fmaf_rn(dz, dz, __fmaf_rn(dx, dx, __fmul_rn(dy, dy)));
        if (d < b1) { b3 = b2; i3 = i2; b2 = b1; i2 = i1; b1 = d; i1 = k; }
        else if (d < b2) { b3 = b2; i3 = i2; b2 = d; i2 = k; }
        else if (d < b3) { b3 = d; i3 = k; }
    }
    float *od = dist2 + ((size_t)bi * n + p) * 3;
    int *oi = idx + ((size_t)bi * n + p) * 3;
    od[0] = b1; od[1] = b2; od[2] = b3;
    oi[0] = i1; oi[1] = i2; oi[2] = i3;
}

}  // namespace mssvt

using namespace mssvt;

extern "C" {

int mssvt_fps_log2_block(int n);

int mssvt_group_features(int B, int M, int C, int nsample, const float *features,
                         const int *features_batch_cnt, const int *idx, const int *idx_batch_cnt,
                         float *out, void *stream) {
    if (B <= 0 || M < 0 || C <= 0 || nsample <= 0 || C > 4096) return MSSVT_ERR_INVALID;
    if (M == 0) return MSSVT_OK;
    if (!features || !features_batch_cnt || !idx || !idx_batch_cnt || !out) return MSSVT_ERR_INVALID;
    size_t smem = (size_t)GF_STILE * (C + 1) * sizeof(float);
    cudaFuncSetAttribute(k_group_features, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    // a pass of one CTA is a chain idx -> rows -> transposed stores with a barrier between the links; narrow rows
    // (C <= 32: 4 KB per pass) take smaller CTAs, twice as many of them per SM
    const int threads = C <= 32 ? 128 : GF_THREADS, cap = C <= 32 ? 16 : 8;
    int per_sm = (int)(200 * 1024 / (smem + 1024));
    per_sm = per_sm > cap ? cap : per_sm < 1 ? 1 : per_sm;
    ++g_launches;
    k_group_features<<<persistent_grid(M, 1, per_sm), threads, smem, (cudaStream_t)stream>>>(
        B, M, C, nsample, features, features_batch_cnt, idx, idx_batch_cnt, out);
    return check_launch();
}

int mssvt_group_features_grad(int B, int M, int C, int N, int nsample, const float *grad_out,
                              const int *idx, const int *idx_batch_cnt,
                              const int *features_batch_cnt, float *grad_features, void *stream) {
    if (B <= 0 || M < 0 || C <= 0 || nsample <= 0 || N < 0 || C > 4096) return MSSVT_ERR_INVALID;
    cudaStream_t s = (cudaStream_t)stream;
    if (N) {
        if (!grad_features) return MSSVT_ERR_INVALID;
        cudaError_t e = cudaMemsetAsync(grad_features, 0, (size_t)N * C * sizeof(float), s);
        if (e != cudaSuccess) { g_last_cuda_error = (int)e; return MSSVT_ERR_LAUNCH; }
    }
    if (M == 0) return MSSVT_OK;
    if (!grad_out || !idx || !idx_batch_cnt || !features_batch_cnt) return MSSVT_ERR_INVALID;
    size_t smem = (size_t)GF_STILE * (C + 1) * sizeof(float);
    cudaFuncSetAttribute(k_group_features_grad, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    int per_sm = (int)(200 * 1024 / (smem + 1024));
    per_sm = per_sm > 8 ? 8 : per_sm < 1 ? 1 : per_sm;
    ++g_launches;
    k_group_features_grad<<<persistent_grid(M, 1, per_sm), GF_THREADS, smem, s>>>(
        B, M, C, nsample, grad_out, idx, idx_batch_cnt, features_batch_cnt, grad_features);
    return check_launch();
}

// temp: (b, n) fp32 scratch, only touched when a row does not fit shared memory (n > ~50k);
// may be null otherwise.  The reference's API passes it always (pointnet2_utils.py:26).
int mssvt_fps(int b, int n, int m, const float *dataset, float *temp, int *idxs, void *stream) {
    if (b < 0 || n <= 0 || m < 0 || n >= (1 << 21)) return MSSVT_ERR_INVALID;
    if (b == 0 || m == 0) return MSSVT_OK;
    if (!dataset || !idxs) return MSSVT_ERR_INVALID;
    int log2b = mssvt_fps_log2_block(n);
    cudaStream_t s = (cudaStream_t)stream;
    if (n <= FPS_WARP_MAXN) {
        size_t smem = (size_t)FPS_WARPS * n * 4 * sizeof(float);
        ++g_launches;
        k_fps_warp<<<persistent_grid(b, FPS_WARPS, 8), FPS_WARPS * 32, smem, s>>>(b, n, m, log2b,
                                                                                  dataset, idxs);
    } else {
        size_t smem = (size_t)n * sizeof(float);
        int in_smem = smem <= 200 * 1024;
        if (!in_smem && !temp) return MSSVT_ERR_WORKSPACE;
        if (in_smem)
            cudaFuncSetAttribute(k_fps_cta, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        ++g_launches;
        k_fps_cta<<<b, FPS_CTA, in_smem ? smem : 0, s>>>(n, m, log2b, in_smem, dataset, temp, idxs);
    }
    return check_launch();
}

int mssvt_gather_points(int b, int c, int n, int m, const float *points, const int *idx,
                        float *out, void *stream) {
    if (b < 0 || c < 0 || n < 0 || m < 0) return MSSVT_ERR_INVALID;
    long long total = (long long)b * c * m;
    if (total == 0) return MSSVT_OK;
    if (!points || !idx || !out) return MSSVT_ERR_INVALID;
    ++g_launches;
    k_gather_points<<<persistent_grid(total, 256, 8), 256, 0, (cudaStream_t)stream>>>(b, c, n, m,
                                                                                      points, idx, out);
    return check_launch();
}

int mssvt_three_nn(int b, int n, int m, const float *unknown, const float *known, float *dist2,
                   int *idx, void *stream) {
    if (b < 0 || n < 0 || m < 0 || b > 65535 * 1024) return MSSVT_ERR_INVALID;
    if (b == 0 || n == 0) return MSSVT_OK;
    if (!unknown || (m && !known) || !dist2 || !idx) return MSSVT_ERR_INVALID;
    // rows are independent; fold large b into gridDim.y chunks of 65535
    for (int b0 = 0; b0 < b; b0 += 65535) {
        int bb = b - b0 < 65535 ? b - b0 : 65535;
        dim3 grid(div_up(n, 128), bb);
        ++g_launches;
        k_three_nn<<<grid, 128, 0, (cudaStream_t)stream>>>(
            bb, n, m, unknown + (size_t)b0 * n * 3, known + (size_t)b0 * m * 3,
            dist2 + (size_t)b0 * n * 3, idx + (size_t)b0 * n * 3);
    }
    return check_launch();
}

int mssvt_group_points(int b, int c, int n, int npoints, int nsample, const float *points,
                       const int *idx, float *out, void *stream) {
    if (b < 0 || c < 0 || n < 0 || npoints < 0 || nsample < 0) return MSSVT_ERR_INVALID;
    long long total = (long long)b * c * npoints * nsample;
    if (total == 0) return MSSVT_OK;
    if (!points || !idx || !out) return MSSVT_ERR_INVALID;
    ++g_launches;
    k_group_points<<<persistent_grid(total, 256, 8), 256, 0, (cudaStream_t)stream>>>(
        b, c, n, npoints, nsample, points, idx, out);
    return check_launch();
}

int mssvt_group_points_grad(int b, int c, int n, int npoints, int nsample, const float *grad_out,
                            const int *idx, float *grad_points, void *stream) {
    if (b < 0 || c < 0 || n < 0 || npoints < 0 || nsample < 0) return MSSVT_ERR_INVALID;
    cudaStream_t s = (cudaStream_t)stream;
    if ((long long)b * c * n) {
        if (!grad_points) return MSSVT_ERR_INVALID;
        cudaError_t e = cudaMemsetAsync(grad_points, 0, (size_t)b * c * n * sizeof(float), s);
        if (e != cudaSuccess) { g_last_cuda_error = (int)e; return MSSVT_ERR_LAUNCH; }
    }
    long long total = (long long)b * c * npoints * nsample;
    if (total == 0) return MSSVT_OK;
    if (!grad_out || !idx) return MSSVT_ERR_INVALID;
    ++g_launches;
    k_group_points_grad<<<persistent_grid(total, 256, 8), 256, 0, s>>>(b, c, n, npoints, nsample,
                                                                      grad_out, idx, grad_points);
    return check_launch();
}

}  // extern "C"
