// common.cuh -- shared device helpers for the mssvt_b200 kernels (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#define MSSVT_OK 0
#define MSSVT_ERR_INVALID (-1)   // bad argument (null pointer, size out of the supported range)
#define MSSVT_ERR_LAUNCH (-2)    // kernel launch / CUDA runtime failure (see mssvt_last_cuda_error)
#define MSSVT_ERR_WORKSPACE (-3) // workspace too small

#define MSSVT_EMPTY (-1)         // EMPTY_KEY of the reference table (ms_cuda_utils.h:9)
#define MSSVT_NUM_SMS 148

namespace mssvt {

extern int g_last_cuda_error;
extern long long g_launches;  // kernels launched by this library since load

static inline int check_launch() {
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        g_last_cuda_error = (int)e;
        return MSSVT_ERR_LAUNCH;
    }
    return MSSVT_OK;
}

static inline int div_up(long long a, long long b) { return (int)((a + b - 1) / b); }

// grid for a grid-stride kernel: enough CTAs to cover `items`, capped at `waves` full waves of
// the 148 SMs times the resident CTAs per SM.
// ---- programmatic dependent launch (PDL).  A kernel launched with launch_pdl() may become resident while
// its predecessor in the stream is still draining: its CTAs run their prologue (stage packed weights,
// barriers, TMEM allocation -- nothing the predecessor writes) and then block in pdl_wait() until the
// predecessor has completed and its writes are visible.  Every kernel of such a chain calls
// pdl_launch_dependents() first thing, which lets the successor start as soon as all CTAs of this grid
// are running.  This removes the launch gap and the tail idle time between the ~45 kernels of a frame.
// RULES: (1) in a kernel launched with launch_pdl(), every read of data that is not a static parameter
// (weights, biases) comes after pdl_wait().  (2) The L1 of an SM is not invalidated between the early start
// of the dependent's CTAs and their pdl_wait(), so a dependent must not read -- through L1-cached loads
// (__ldg / const __restrict__) -- addresses that the predecessor both reads and writes (its own loads may have
// left stale lines in that L1).  The feature-kernel chain satisfies this (every kernel only writes its
// outputs); the hash / window / scan kernels, which read and update the same tables, are launched normally.
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

template <typename... KArgs, typename... Args>
static inline void launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t s,
                              Args &&...args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

static inline int persistent_grid(long long items, int per_cta, int ctas_per_sm, int waves = 4) {
    long long need = (items + per_cta - 1) / per_cta;
    long long cap = (long long)MSSVT_NUM_SMS * ctas_per_sm * waves;
    if (need < 1) need = 1;
    return (int)(need < cap ? need : cap);
}

// ---- reference hash table: (H, 2) int32 [key, value], h(k) = k % H, linear probing
// (ms_sparse_attention_gpu.cu:18-64).  One 8-byte load per probe.
__device__ __forceinline__ int table_find(const int2 *__restrict__ tab, int hash_size, int key) {
    int slot = (int)((unsigned)key % (unsigned)hash_size);
    for (int probes = 0; probes < hash_size; ++probes) {
        int2 kv = __ldg(tab + slot);
        if (kv.x == key) return kv.y;
        if (kv.x == MSSVT_EMPTY) return MSSVT_EMPTY;
        slot = slot + 1 == hash_size ? 0 : slot + 1;
    }
    return MSSVT_EMPTY;
}

// ---- voxel lookup structures.  Both expose find(b, x, y, z) -> per-sample voxel index or -1.
//
// HashIdx: the reference's table (drop-in contract of the op-level API).
struct HashIdx {
    const int2 *table;
    int hash_size, y_max, z_max;
    __device__ __forceinline__ int find(int b, int x, int y, int z) const {
        return table_find(table + (size_t)b * hash_size, hash_size, x * y_max * z_max + y * z_max + z);
    }
};

// GridIdx: occupancy bitmap + rank, used by the fused path.  One int2 {bits, base} per 32 cells
// of a z column; a hit costs one more load from `vals` at base + popcount(bits below z).  At most
// two dependent loads, no probing loop, no clustering, and the 125 probes of a window touch only
// 25 neighbouring words (the z neighbours of an offset share a word).  For the S0 grid the
// structure is 1.75 MB + 4 N bytes: it lives in L2.
struct GridIdx {
    const int2 *cells;  // (B, x_max * y_max * zw)
    const int *vals;    // (N) per-sample voxel index, grouped by word, ordered by z inside a word
    int y_max, zw;      // zw = words per column = ceil(z_max / 32)
    long long words_per_sample;
    __device__ __forceinline__ int find(int b, int x, int y, int z) const {
        int2 c = __ldg(cells + (size_t)b * words_per_sample + ((size_t)x * y_max + y) * zw + (z >> 5));
        unsigned bit = 1u << (z & 31);
        if (!((unsigned)c.x & bit)) return MSSVT_EMPTY;
        return __ldg(vals + c.y + __popc((unsigned)c.x & (bit - 1u)));
    }
};

__device__ __forceinline__ unsigned lanemask_lt() {
    unsigned m;
    asm("mov.u32 %0, %%lanemask_lt;" : "=r"(m));
    return m;
}

// voxel-centre world coordinate exactly as with_coords (mssvt_backbone.py:133-137) evaluates it
// in PyTorch: three separately rounded fp32 ops, never contracted into an FMA.
__device__ __forceinline__ float world_coord(int idx, float cell, float lo) {
    return __fadd_rn(__fmul_rn(__fadd_rn((float)idx, 0.5f), cell), lo);
}

}  // namespace mssvt
