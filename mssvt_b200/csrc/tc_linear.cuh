// tc_linear.cuh -- out[r] = (W in[r] + b) * mul for 64-channel rows on the tcgen05 tensor cores.
// The small projections of the attention blocks (queries, outputs) are row-wise 64 x 64 linear maps over
// 10^4..10^5 rows; with one thread per (row, 8 outputs) on the FP32 pipe they are bound by the shared-memory
// reads of the weights.  Here a CTA of 256 threads owns a tile of 128 rows (threads t and t + 128 share
// row t: 32 channels each, both reach TMEM lane t); a PRODUCER functor builds the input row (gather,
// positional embedding, max-pool ...), the thread stores it TF32-rounded as the A operand, one thread
// issues 8 tcgen05.mma (M = 128, N = 64, K = 8) against the packed weight (mssvt_pack_operand_tf32; two
// head groups = one block-diagonal 64 x 64 matrix) and every thread reads its 32 outputs back from TMEM.
// 49 KB of shared memory and 64 TMEM columns per CTA: four CTAs per SM.
#pragma once
#include "tc_common.cuh"

namespace mssvt {

#define TCL_ROWS 128
#define TCL_THREADS 256
#define TCL_C 64

struct TclParams {
    const float *w_packed;  // [64][64] packed
    const float *bias;      // [64] (bias_hi == nullptr) or [32] + bias_hi [32]
    const float *bias_hi;
    float mul;
};

static inline size_t tcl_smem_bytes(int terms) {
    return (size_t)(terms == 3 ? 2 : 1) * (TCL_ROWS * TCL_C * 4 + TCL_C * TCL_C * 4) + TCL_C * 4 + 2048 + 8 + 16 + 128;
}

// Producer: struct with
//   __device__ void init(float *s_extra)            cooperative, before the first barrier (s_extra: 2 KB)
//   __device__ int rows() const                     number of rows (device-side count)
//   struct Ctx;                                      per-row context carried from src() to finish()
//   __device__ const float4 *src(int row, int half, Ctx &) const     128-byte segment holding the row's 32 inputs
//   __device__ void finish(const Ctx &, int half, const float *s_extra, float *in /*[32]*/) const   in-place fix-up

// rows copied from a dense (rows, 64) array; row count = min(cap, *count), optionally mapped through
// an index array (count_map[n], e.g. a prefix sum)
struct TclCopyRows {
    const float *rows_in;
    const int *count, *count_map;
    int cap;
    __device__ void init(float *) const {}
    __device__ int rows() const {
        const int n = min(cap, __ldg(count));
        return count_map ? __ldg(count_map + n) : n;
    }
    struct Ctx {};
    __device__ const float4 *src(int row, int half, Ctx &) const {
        return (const float4 *)(rows_in + (size_t)row * TCL_C + half * 32);
    }
    __device__ void finish(const Ctx &, int, const float *, float *) const {}
};
template <class Producer, int TERMS>
__global__ void __launch_bounds__(TCL_THREADS, TERMS == 3 ? 2 : 4)
k_tc_linear(TclParams P, Producer prod, float *__restrict__ out) {
    constexpr int NT = TERMS == 3 ? 2 : 1;     // operand tiles: hi [, lo]
    extern __shared__ __align__(128) char smem_raw[];
    pdl_launch_dependents();
    const int tid = threadIdx.x, warp = tid >> 5;
    const int r = tid & (TCL_ROWS - 1), half = tid >> 7;
    char *sA = smem_raw;                           // NT x [16][128][16 B]
    char *sW = sA + NT * TCL_ROWS * TCL_C * 4;     // NT x [16][64][16 B]
    float *sB = (float *)(sW + NT * TCL_C * TCL_C * 4); // [64]
    float *sExtra = sB + TCL_C;                    // producer scratch (2 KB)
    uint64_t *sBar = (uint64_t *)(sExtra + 512);
    uint32_t *sTmem = (uint32_t *)(sBar + 1);

    stage_packed(P.w_packed, NT * TCL_C * TCL_C, sW);
    if (tid < TCL_C) sB[tid] = P.bias_hi && tid >= 32 ? __ldg(P.bias_hi + tid - 32) : __ldg(P.bias + tid);
    prod.init(sExtra);
    const uint32_t bar = smem_u32(sBar);
    if (tid == 0) {
        mbar_init(bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) tmem_alloc(smem_u32(sTmem), 64);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_d = *sTmem;
    const uint32_t lane_off = (uint32_t)((warp & 3) * 32) << 16;
    const uint32_t idesc = umma_idesc_tf32(TCL_ROWS, TCL_C);
    const uint32_t sA_u = smem_u32(sA), sW_u = smem_u32(sW);
    const uint32_t a_lbo = TCL_ROWS * 16, w_lbo = TCL_C * 16;
    const uint32_t my_row_off = (uint32_t)(r >> 3) * 128u + (uint32_t)(r & 7) * 16u;
    char *stg = sA + warp * 4096;  // warp-private staging (inside the A tile, see the barriers below)
    pdl_wait();  // (everything above touched static parameters only)
    const int n = prod.rows();
    const int tiles = (n + TCL_ROWS - 1) / TCL_ROWS;
    uint32_t phase = 0;
    for (int tile = blockIdx.x; tile < tiles; tile += gridDim.x, phase ^= 1u) {
        const int row = tile * TCL_ROWS + r;
        const bool live = row < n;
        float in[32];
        {
            typename Producer::Ctx ctx;
            const float4 *my_src = live ? prod.src(row, half, ctx) : nullptr;
            float4 v[8];
            warp_rows_load(stg, my_src, v);
#pragma unroll
            for (int c = 0; c < 8; ++c) { in[4 * c] = v[c].x; in[4 * c + 1] = v[c].y; in[4 * c + 2] = v[c].z; in[4 * c + 3] = v[c].w; }
            if (live) prod.finish(ctx, half, sExtra, in);
        }
        __syncthreads();  // the staging areas alias the A tile: every warp is done reading before A is written
#pragma unroll
        for (int c = 0; c < 8; ++c) {
            float4 hi, lo;
            split_tf32(make_float4(in[4 * c], in[4 * c + 1], in[4 * c + 2], in[4 * c + 3]), hi, lo);
            *(float4 *)(sA + (uint32_t)(half * 8 + c) * a_lbo + my_row_off) = hi;
            if (TERMS == 3) *(float4 *)(sA + TCL_ROWS * TCL_C * 4 + (uint32_t)(half * 8 + c) * a_lbo + my_row_off) = lo;
        }
        stage_packed_wait();
        fence_async_smem();
        __syncthreads();
        if (issuer_elected()) {
            tc_fence_after();
#pragma unroll
            for (int k = 0; k < TCL_C / 8; ++k)
                umma_step<TERMS>(tmem_d, sA_u + (uint32_t)k * 2u * a_lbo, a_lbo, TCL_ROWS * TCL_C * 4,
                                 sW_u + (uint32_t)k * 2u * w_lbo, w_lbo, TCL_C * TCL_C * 4, idesc, k == 0);
            umma_commit(bar);
        }
        mbar_wait(bar, phase);
        tc_fence_after();
        float d[32];
        tmem_ld32(tmem_d + lane_off + (uint32_t)(half * 32), d);
        {
            const float *bb = sB + half * 32;
            float4 o[8];
#pragma unroll
            for (int q = 0; q < 8; ++q)
                o[q] = make_float4((d[4 * q] + bb[4 * q]) * P.mul, (d[4 * q + 1] + bb[4 * q + 1]) * P.mul,
                                   (d[4 * q + 2] + bb[4 * q + 2]) * P.mul, (d[4 * q + 3] + bb[4 * q + 3]) * P.mul);
            // (the MMA has consumed the A tile: its memory serves as the staging area again)
            warp_rows_store(stg, live ? (float4 *)(out + (size_t)row * TCL_C + half * 32) : nullptr, o);
        }
        tc_fence_before();
        __syncthreads();  // TMEM and the A tile are free for the next tile
    }
    if (warp == 0) tmem_dealloc(tmem_d, 64);
}

template <class Producer, int TERMS>
static inline void tcl_launch_t(const TclParams &P, const Producer &prod, int row_capacity, float *out, cudaStream_t s) {
    const size_t smem = tcl_smem_bytes(TERMS);
    int tiles = (row_capacity + TCL_ROWS - 1) / TCL_ROWS;
    int grid = MSSVT_NUM_SMS * (TERMS == 3 ? 2 : 4);
    if (grid > tiles) grid = tiles;
    if (grid < 1) return;
    cudaFuncSetAttribute(k_tc_linear<Producer, TERMS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    ++g_launches;
    launch_pdl(k_tc_linear<Producer, TERMS>, dim3(grid), dim3(TCL_THREADS), smem, s, P, prod, out);
}

// terms = 1: TF32 operands; terms = 3: split operands (w_packed = [hi | lo])
template <class Producer>
static inline void tcl_launch(const TclParams &P, const Producer &prod, int row_capacity, float *out, cudaStream_t s,
                              int terms = 1) {
    if (terms == 3) tcl_launch_t<Producer, 3>(P, prod, row_capacity, out, s);
    else tcl_launch_t<Producer, 1>(P, prod, row_capacity, out, s);
}

}  // namespace mssvt
