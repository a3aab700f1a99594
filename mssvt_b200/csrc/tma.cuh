// tma.cuh -- tensor-map (TMA) tile movement for row-major (rows, 64) fp32 feature matrices, sm_100a.
//
// The feature matrices of the path (x, y, the next block's LayerNorm rows) are dense row-major arrays, and the
// kernels own their rows in groups of 32 consecutive rows per warp.  A warp's 32 x 128-byte half rows are one TMA
// box: cp.async.bulk.tensor moves it between global memory and a 4 KB shared-memory region without a register or
// an LSU data-pipe pass (the LDG -> STS -> LDS staging it replaces crosses that pipe three times), completion on an
// mbarrier (loads) or a bulk group (stores).  The box lands with the 128-byte swizzle: 16-byte chunk c of row r sits
// at chunk c ^ (r & 7) of the row's 128 bytes -- the same conflict-free layout the register staging used, so every
// lane reads / writes its own row with 16-byte accesses that hit eight different bank groups per quarter warp.
// SASS: UTMALDG / UTMASTG.  The tensor map is built on the host per launch (cuTensorMapEncodeTiled, fetched through
// cudaGetDriverEntryPoint: the library links no libcuda) and passed as a __grid_constant__ kernel parameter.
#pragma once
#include <cuda.h>
#include "tc_common.cuh"

namespace mssvt {

#define TMA_BOX_ROWS 32
#define TMA_BOX_COLS 32                        // floats: 128 bytes, the span of the 128-byte swizzle
#define TMA_BOX_BYTES (TMA_BOX_ROWS * TMA_BOX_COLS * 4)

// byte offset of 16-byte chunk `c` of row `r` inside a swizzled box (region 1024-byte aligned)
__device__ __forceinline__ uint32_t tma_swz(int r, int c) { return (uint32_t)(r * 128 + ((c ^ (r & 7)) << 4)); }

// box (col0 .. col0 + 31, row0 .. row0 + 31) -> dst; completes `TMA_BOX_BYTES` on the mbarrier (rows / columns outside
// the tensor arrive as zeros and count as bytes)
__device__ __forceinline__ void tma_load_box(char *dst, const CUtensorMap *map, int col0, int row0, uint32_t mbar) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(
            smem_u32(dst)),
        "l"(map), "r"(col0), "r"(row0), "r"(mbar)
        : "memory");
}
// src -> box (col0 .., row0 ..); rows / columns outside the tensor are dropped.  Part of the issuing thread's current
// bulk group.
__device__ __forceinline__ void tma_store_box(const CUtensorMap *map, int col0, int row0, const char *src) {
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.tile.bulk_group [%0, {%1, %2}], [%3];" ::"l"(map), "r"(col0),
                 "r"(row0), "r"(smem_u32(src))
                 : "memory");
}
__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
// every bulk store of this thread has finished READING shared memory (the region may be overwritten)
__device__ __forceinline__ void tma_store_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void tma_prefetch_map(const CUtensorMap *map) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(map) : "memory");
}

// ---- host side
typedef CUresult (*TmaEncodeFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static inline TmaEncodeFn tma_encode_fn() {
    static TmaEncodeFn fn = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = (TmaEncodeFn)p;
    }
    return fn;
}

// tensor map of a row-major (rows, cols) fp32 matrix (cols a multiple of 32), boxes of 32 rows x 32 floats, 128-byte
// swizzle.  false: the driver entry point is missing or the arguments were rejected.
static inline bool tma_rows_map(CUtensorMap *map, const float *base, long long rows, int cols) {
    TmaEncodeFn enc = tma_encode_fn();
    if (!enc || !base || rows <= 0 || cols <= 0 || (cols % TMA_BOX_COLS) != 0 || ((uintptr_t)base & 15)) return false;
    const cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    const cuuint64_t strides[1] = {(cuuint64_t)cols * 4};
    const cuuint32_t box[2] = {TMA_BOX_COLS, TMA_BOX_ROWS};
    const cuuint32_t estr[2] = {1, 1};
    return enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void *)base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
               CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

}  // namespace mssvt
