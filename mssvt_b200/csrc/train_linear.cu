// train_linear.cu -- the row-wise linear maps of the TRAINING path (q / kv / output projections over compact window
// rows, FFN layers: nn.Linear with 32 / 64 / 128 channels over 10^5 .. 10^6 rows) forward and backward.
//
// These are tall-skinny products: 21 flop per byte at most, so they are bound by HBM as soon as the multiply runs on a
// tensor pipe; what they need is one pass over the rows with the small weight matrix resident on chip.  Library GEMMs
// treat the weight gradient (a reduction over up to 10^6 rows into a 64 x 32 matrix) as a general split-K problem:
// 160 - 440 us per call in fp32, and milliseconds under bf16 autocast, against 40 us of HBM time.
//
//   k_linear_rows    y = x W^T (+ b) (ReLU): CTA = 64-row tile x all outputs, W in shared memory for the life of the
//                    CTA, the next tile prefetched into registers while the current one is multiplied.  Also the input
//                    gradient (dx = dy W: the same kernel on the transposed weight).
//   k_linear_wgrad   dW = dy^T x and db = sum dy: rows in tiles of 32 through shared memory, the N x K result split over
//                    the 8 warps of a CTA and kept in registers across the CTA's tiles; per-CTA partial results are
//                    written out and summed by k_linear_wgrad_reduce in a fixed order (no atomics: deterministic).
// Arithmetic: warp-level mma.sync m16n8k8 TF32 with fp32 accumulation; SPLIT = every operand as hi + lo TF32 and three
// MMAs per product (the 3xTF32 scheme of the inference kernels): fp32-grade
// results (the parity-grade default).
// tcgen05 would not move these kernels: the tensor pipe is idle most of the time either way.
#include "common.cuh"

namespace mssvt {

// Operand split for the 3xTF32 scheme with integer operations instead of conversion instructions (cvt.rna.tf32 runs on the
// quarter-rate conversion pipe): hi = v rounded to TF32 (ties away from zero: one ADD + one AND on the bits), lo = v - hi
// (exact in fp32), rounded the same way -- the accuracy of the cvt.rna form: |v - hi - lo| <= 2^-22 |v|.
// -DMSSVT_TF32_SPLIT_TRUNC: both halves truncated instead (two operations per operand, 2^-20): measured, not used -- logits
// of keys far from their window (the aliased key of quirk Q1) are sums of large products, where the coarser halves show.
__device__ __forceinline__ void tf32_split(float v, uint32_t &hi, uint32_t &lo) {
#ifdef MSSVT_TF32_SPLIT_TRUNC
    hi = __float_as_uint(v) & 0xffffe000u;
    lo = __float_as_uint(v - __uint_as_float(hi));
#else
    hi = (__float_as_uint(v) + 0x1000u) & 0xffffe000u;
    lo = (__float_as_uint(v - __uint_as_float(hi)) + 0x1000u) & 0xffffe000u;
#endif
}

// plain TF32 operand, rounded to nearest (ties away from zero) with two integer operations
__device__ __forceinline__ uint32_t tf32_round(float v) { return (__float_as_uint(v) + 0x1000u) & 0xffffe000u; }

// D (16 x 8) += A (16 x 8, row) B (8 x 8, col).  Lane = 4 g + t: a0 (g, t) a1 (g + 8, t) a2 (g, t + 4) a3 (g + 8, t + 4);
// b0 (k = t, n = g) b1 (k = t + 4, n = g); c0 (g, 2t) c1 (g, 2t + 1) c2 (g + 8, 2t) c3 (g + 8, 2t + 1)
__device__ __forceinline__ void mma_tf32(float *c, const uint32_t *a, const uint32_t *b) {
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}

#define LIN_ROWS 64       // rows per tile of k_linear_rows (4 warps x 16)
#define LIN_THREADS 128
#define WG_ROWS 32        // rows per tile of k_linear_wgrad
#define WG_THREADS 256
#define WG_CTAS (2 * MSSVT_NUM_SMS)

template <int K, int N, bool SPLIT>
__global__ void __launch_bounds__(LIN_THREADS)
k_linear_rows(int R, const float *__restrict__ X, int ldx, const float *__restrict__ W, const float *__restrict__ bias,
              int relu, float *__restrict__ Y, int ldy) {
    constexpr int LDK = K + 4;                   // row stride of both tiles: fragment loads hit 32 different banks
    constexpr int PV = LIN_ROWS * K / 4 / LIN_THREADS;   // float4 per thread of a row tile
    extern __shared__ __align__(16) float lin_smem[];
    float *sW = lin_smem, *sX = lin_smem + N * LDK;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
    for (int i = tid; i < N * K / 4; i += LIN_THREADS) {
        const int n = i / (K / 4), k4 = i % (K / 4);
        *(float4 *)(sW + n * LDK + 4 * k4) = __ldg((const float4 *)(W + (size_t)n * K) + k4);
    }
    const int tiles = (R + LIN_ROWS - 1) / LIN_ROWS;
    float4 pre[PV];
    auto fetch = [&](int tile) {
#pragma unroll
        for (int j = 0; j < PV; ++j) {
            const int i = tid + j * LIN_THREADS, r = tile * LIN_ROWS + i / (K / 4), k4 = i % (K / 4);
            pre[j] = r < R ? __ldg((const float4 *)(X + (size_t)r * ldx) + k4) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
    };
    int tile = blockIdx.x;
    if (tile < tiles) fetch(tile);
    for (; tile < tiles; tile += gridDim.x) {
        __syncthreads();                          // the previous tile has been consumed (first pass: nothing to wait for)
#pragma unroll
        for (int j = 0; j < PV; ++j) {
            const int i = tid + j * LIN_THREADS;
            *(float4 *)(sX + (i / (K / 4)) * LDK + 4 * (i % (K / 4))) = pre[j];
        }
        __syncthreads();                          // (also: the weights are in place)
        if (tile + (int)gridDim.x < tiles) fetch(tile + gridDim.x);
        float acc[N / 8][4];
#pragma unroll
        for (int nb = 0; nb < N / 8; ++nb) acc[nb][0] = acc[nb][1] = acc[nb][2] = acc[nb][3] = 0.f;
        const float *xa = sX + (warp * 16 + g) * LDK + t;
#pragma unroll 2
        for (int ks = 0; ks < K / 8; ++ks) {
            const float af[4] = {xa[ks * 8], xa[8 * LDK + ks * 8], xa[ks * 8 + 4], xa[8 * LDK + ks * 8 + 4]};
            uint32_t ahi[4], alo[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                if (SPLIT) tf32_split(af[i], ahi[i], alo[i]);
                else ahi[i] = tf32_round(af[i]);
            }
#pragma unroll
            for (int nb = 0; nb < N / 8; ++nb) {
                const float *wb = sW + (nb * 8 + g) * LDK + ks * 8 + t;
                uint32_t bhi[2], blo[2];
                if (SPLIT) {
                    tf32_split(wb[0], bhi[0], blo[0]);
                    tf32_split(wb[4], bhi[1], blo[1]);
                    mma_tf32(acc[nb], alo, bhi);
                    mma_tf32(acc[nb], ahi, blo);
                } else {
                    bhi[0] = tf32_round(wb[0]);
                    bhi[1] = tf32_round(wb[4]);
                }
                mma_tf32(acc[nb], ahi, bhi);
            }
        }
        const int r_lo = tile * LIN_ROWS + warp * 16 + g, r_hi = r_lo + 8;
#pragma unroll
        for (int nb = 0; nb < N / 8; ++nb) {
            const int col = nb * 8 + 2 * t;
            const float b0 = bias ? __ldg(bias + col) : 0.f, b1 = bias ? __ldg(bias + col + 1) : 0.f;
            float v0 = acc[nb][0] + b0, v1 = acc[nb][1] + b1, v2 = acc[nb][2] + b0, v3 = acc[nb][3] + b1;
            if (relu) { v0 = fmaxf(v0, 0.f); v1 = fmaxf(v1, 0.f); v2 = fmaxf(v2, 0.f); v3 = fmaxf(v3, 0.f); }
            if (r_lo < R) *(float2 *)(Y + (size_t)r_lo * ldy + col) = make_float2(v0, v1);
            if (r_hi < R) *(float2 *)(Y + (size_t)r_hi * ldy + col) = make_float2(v2, v3);
        }
    }
}

template <int K, int N, bool SPLIT>
__global__ void __launch_bounds__(WG_THREADS, 2)
k_linear_wgrad(int R, const float *__restrict__ dY, int ldy, const float *__restrict__ X, int ldx,
               float *__restrict__ part_w, float *__restrict__ part_b) {
    constexpr int LDN = N + 8, LDK = K + 8;       // fragment loads: bank = 8 t + g
    constexpr int WN = N / 16 < 4 ? N / 16 : 4, WK = 8 / WN;   // warps along n / along k
    constexpr int A = N / (16 * WN), B = K / (8 * WK);         // MMA tiles per warp
    constexpr int YV = WG_ROWS * N / 4 / WG_THREADS, XV = WG_ROWS * K / 4 / WG_THREADS;
    __shared__ __align__(16) float sY[WG_ROWS * LDN];
    __shared__ __align__(16) float sX[WG_ROWS * LDK];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
    const int n_base = (warp % WN) * 16 * A, k_base = (warp / WN) * 8 * B;
    float acc[A][B][4];
#pragma unroll
    for (int a = 0; a < A; ++a)
#pragma unroll
        for (int b = 0; b < B; ++b) acc[a][b][0] = acc[a][b][1] = acc[a][b][2] = acc[a][b][3] = 0.f;
    float bsum = 0.f;
    float4 py[YV], px[XV];
    const int tiles = (R + WG_ROWS - 1) / WG_ROWS;
    auto fetch = [&](int tile) {
#pragma unroll
        for (int j = 0; j < YV; ++j) {
            const int i = tid + j * WG_THREADS, r = tile * WG_ROWS + i / (N / 4);
            py[j] = r < R ? __ldg((const float4 *)(dY + (size_t)r * ldy) + i % (N / 4)) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
#pragma unroll
        for (int j = 0; j < XV; ++j) {
            const int i = tid + j * WG_THREADS, r = tile * WG_ROWS + i / (K / 4);
            px[j] = r < R ? __ldg((const float4 *)(X + (size_t)r * ldx) + i % (K / 4)) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
    };
    int tile = blockIdx.x;
    if (tile < tiles) fetch(tile);
    for (; tile < tiles; tile += gridDim.x) {
        __syncthreads();
#pragma unroll
        for (int j = 0; j < YV; ++j) {
            const int i = tid + j * WG_THREADS;
            *(float4 *)(sY + (i / (N / 4)) * LDN + 4 * (i % (N / 4))) = py[j];
        }
#pragma unroll
        for (int j = 0; j < XV; ++j) {
            const int i = tid + j * WG_THREADS;
            *(float4 *)(sX + (i / (K / 4)) * LDK + 4 * (i % (K / 4))) = px[j];
        }
        __syncthreads();
        if (tile + (int)gridDim.x < tiles) fetch(tile + gridDim.x);
        if (tid < N) {
#pragma unroll 8
            for (int r = 0; r < WG_ROWS; ++r) bsum += sY[r * LDN + tid];
        }
#pragma unroll
        for (int rs = 0; rs < WG_ROWS / 8; ++rs) {
            uint32_t bhi[B][2], blo[B][2];
#pragma unroll
            for (int b = 0; b < B; ++b) {
                const float *xb = sX + (rs * 8 + t) * LDK + k_base + 8 * b + g;
                if (SPLIT) {
                    tf32_split(xb[0], bhi[b][0], blo[b][0]);
                    tf32_split(xb[4 * LDK], bhi[b][1], blo[b][1]);
                } else {
                    bhi[b][0] = tf32_round(xb[0]);
                    bhi[b][1] = tf32_round(xb[4 * LDK]);
                }
            }
#pragma unroll
            for (int a = 0; a < A; ++a) {
                const float *ya = sY + (rs * 8 + t) * LDN + n_base + 16 * a + g;
                const float af[4] = {ya[0], ya[8], ya[4 * LDN], ya[4 * LDN + 8]};
                uint32_t ahi[4], alo[4];
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    if (SPLIT) tf32_split(af[i], ahi[i], alo[i]);
                    else ahi[i] = tf32_round(af[i]);
                }
#pragma unroll
                for (int b = 0; b < B; ++b) {
                    if (SPLIT) {
                        mma_tf32(acc[a][b], alo, bhi[b]);
                        mma_tf32(acc[a][b], ahi, blo[b]);
                    }
                    mma_tf32(acc[a][b], ahi, bhi[b]);
                }
            }
        }
    }
    // this CTA's partial sums: (N, K) + (N)
    float *pw = part_w + (size_t)blockIdx.x * N * K;
#pragma unroll
    for (int a = 0; a < A; ++a)
#pragma unroll
        for (int b = 0; b < B; ++b) {
            const int n = n_base + 16 * a + g, k = k_base + 8 * b + 2 * t;
            *(float2 *)(pw + (size_t)n * K + k) = make_float2(acc[a][b][0], acc[a][b][1]);
            *(float2 *)(pw + (size_t)(n + 8) * K + k) = make_float2(acc[a][b][2], acc[a][b][3]);
        }
    if (tid < N) part_b[(size_t)blockIdx.x * N + tid] = bsum;
}

// sums of the per-CTA partial results in a fixed order: 32 outputs x 8 slices of the partial list per CTA (a slice
// adds its partials in order, the 8 slice sums are added in order)
__global__ void __launch_bounds__(256)
k_linear_wgrad_reduce(int ctas, int nk, int n, const float *__restrict__ part_w, const float *__restrict__ part_b,
                      float *__restrict__ gw, float *__restrict__ gb) {
    __shared__ float s_part[8][32];
    const int lane = threadIdx.x & 31, slice = threadIdx.x >> 5;
    const int i = blockIdx.x * 32 + lane;
    const int per = (ctas + 7) / 8, c0 = slice * per, c1 = min(ctas, c0 + per);
    float s = 0.f;
    if (i < nk) {
#pragma unroll 4
        for (int c = c0; c < c1; ++c) s += __ldg(part_w + (size_t)c * nk + i);
    } else if (i < nk + n) {
#pragma unroll 4
        for (int c = c0; c < c1; ++c) s += __ldg(part_b + (size_t)c * n + (i - nk));
    }
    s_part[slice][lane] = s;
    __syncthreads();
    if (slice == 0 && i < nk + n) {
        float t = 0.f;
#pragma unroll
        for (int k = 0; k < 8; ++k) t += s_part[k][lane];
        if (i < nk) gw[i] = t;
        else if (gb) gb[i - nk] = t;
    }
}

template <int K, int N, bool SPLIT>
static int linear_rows(int R, const float *x, int ldx, const float *w, const float *bias, int relu, float *y, int ldy,
                       cudaStream_t s) {
    const size_t smem = (size_t)(N + LIN_ROWS) * (K + 4) * sizeof(float);
    cudaFuncSetAttribute(k_linear_rows<K, N, SPLIT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    const int grid = persistent_grid(R, LIN_ROWS, 4, 1);
    k_linear_rows<K, N, SPLIT><<<grid, LIN_THREADS, smem, s>>>(R, x, ldx, w, bias, relu, y, ldy);
    ++g_launches;
    return check_launch();
}

template <int K, int N, bool SPLIT>
static int linear_wgrad(int R, const float *gy, int ldgy, const float *x, int ldx, float *ws, float *gw, float *gb,
                        cudaStream_t s) {
    const int ctas = R > 0 ? persistent_grid(R, WG_ROWS, 2, 1) : 0;
    float *part_w = ws, *part_b = ws + (size_t)WG_CTAS * N * K;
    if (ctas > 0) {
        k_linear_wgrad<K, N, SPLIT><<<ctas, WG_THREADS, 0, s>>>(R, gy, ldgy, x, ldx, part_w, part_b);
        ++g_launches;
    }
    k_linear_wgrad_reduce<<<div_up(N * K + N, 32), 256, 0, s>>>(ctas, N * K, N, part_w, part_b, gw, gb);
    ++g_launches;
    return check_launch();
}

static bool lin_dim(int d) { return d == 32 || d == 64 || d == 128; }
static bool lin_rows_ok(const void *p, int ld) { return p && ((uintptr_t)p & 15u) == 0 && ld > 0 && ld % 4 == 0; }

}  // namespace mssvt

using namespace mssvt;

#define MSSVT_LIN_DISPATCH(FN, ...)                                                                   \
    do {                                                                                              \
        const int code = (K == 32 ? 0 : K == 64 ? 1 : 2) * 3 + (N == 32 ? 0 : N == 64 ? 1 : 2);        \
        if (terms == 3) switch (code) {                                                               \
            case 0: return FN<32, 32, true>(__VA_ARGS__); case 1: return FN<32, 64, true>(__VA_ARGS__);   \
            case 2: return FN<32, 128, true>(__VA_ARGS__); case 3: return FN<64, 32, true>(__VA_ARGS__);  \
            case 4: return FN<64, 64, true>(__VA_ARGS__); case 5: return FN<64, 128, true>(__VA_ARGS__);  \
            case 6: return FN<128, 32, true>(__VA_ARGS__); case 7: return FN<128, 64, true>(__VA_ARGS__); \
            default: return FN<128, 128, true>(__VA_ARGS__);                                          \
        }                                                                                             \
        switch (code) {                                                                               \
            case 0: return FN<32, 32, false>(__VA_ARGS__); case 1: return FN<32, 64, false>(__VA_ARGS__);   \
            case 2: return FN<32, 128, false>(__VA_ARGS__); case 3: return FN<64, 32, false>(__VA_ARGS__);  \
            case 4: return FN<64, 64, false>(__VA_ARGS__); case 5: return FN<64, 128, false>(__VA_ARGS__);  \
            case 6: return FN<128, 32, false>(__VA_ARGS__); case 7: return FN<128, 64, false>(__VA_ARGS__); \
            default: return FN<128, 128, false>(__VA_ARGS__);                                         \
        }                                                                                             \
    } while (0)

extern "C" {

int mssvt_linear_rows_fwd(int num_rows, int K, int N, int terms, const float *x, int ldx, const float *w,
                          const float *bias, int relu, float *y, int ldy, void *stream) {
    if (num_rows < 0 || !lin_dim(K) || !lin_dim(N) || (terms != 1 && terms != 3)) return MSSVT_ERR_INVALID;
    if (num_rows == 0) return MSSVT_OK;
    if (!w || !lin_rows_ok(x, ldx) || !lin_rows_ok(y, ldy) || ((uintptr_t)w & 15u)) return MSSVT_ERR_INVALID;
    MSSVT_LIN_DISPATCH(linear_rows, num_rows, x, ldx, w, bias, relu, y, ldy, (cudaStream_t)stream);
}

long long mssvt_linear_rows_wgrad_workspace_floats(int K, int N) { return (long long)WG_CTAS * ((long long)N * K + N); }

int mssvt_linear_rows_wgrad(int num_rows, int K, int N, int terms, const float *grad_y, int ldgy, const float *x,
                            int ldx, float *workspace, float *grad_w, float *grad_b, void *stream) {
    if (num_rows < 0 || !lin_dim(K) || !lin_dim(N) || (terms != 1 && terms != 3)) return MSSVT_ERR_INVALID;
    if (!workspace || !grad_w) return MSSVT_ERR_INVALID;
    if (num_rows > 0 && (!lin_rows_ok(grad_y, ldgy) || !lin_rows_ok(x, ldx))) return MSSVT_ERR_INVALID;
    MSSVT_LIN_DISPATCH(linear_wgrad, num_rows, grad_y, ldgy, x, ldx, workspace, grad_w, grad_b, (cudaStream_t)stream);
}

}  // extern "C"
