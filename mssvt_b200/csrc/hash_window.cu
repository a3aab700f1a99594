// hash_window.cu -- voxel hash build, hash lookup, window partition (sm_100a).
//
// Replaces, behind the C-ABI of include/mssvt_b200.h:
//   build_mapping_with_hash_kernel  pcdet/ops/mssvt/src/ms_sparse_attention_gpu.cu:66-115
//   window_with_hash_kernel         pcdet/ops/mssvt/src/ms_sparse_attention_gpu.cu:117-191
// and the CPU fill + H2D copies / per-sample Python loops around them in
// pcdet/ops/mssvt/mssvt_ops.py:7-60.
//
// Table contract is the reference's: (B, H, 2) int32 [key, value], empty = -1, h(k) = k % H,
// linear probing, so tables are interchangeable with the reference kernels.  Differences in HOW:
//  * the -1 fill happens on the device with 16-byte stores (no 8*B*H-byte H2D copy);
//  * a slot is claimed with ONE 64-bit CAS carrying key and value together, probes are single
//    8-byte loads;
//  * windows are numbered deterministically by first occurrence in voxel order (atomicMin of the
//    voxel index + a two-level scan) instead of atomicAdd arrival order.
#include "common.cuh"

namespace mssvt {

int g_last_cuda_error = 0;
long long g_launches = 0;

// ------------------------------------------------------------------------------- fill

__global__ void k_fill_i32x4(int4 *__restrict__ p, size_t n16, int *__restrict__ tail, int ntail,
                             int value) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    size_t stride = (size_t)gridDim.x * blockDim.x;
    int4 v = make_int4(value, value, value, value);
    for (; i < n16; i += stride) p[i] = v;
    if (blockIdx.x == 0 && threadIdx.x < ntail) tail[threadIdx.x] = value;
}

// fill `count` int32 with `value`; p must be 16-byte aligned (torch allocations are).
int fill_i32(int *p, size_t count, int value, cudaStream_t s) {
    if (count == 0) return MSSVT_OK;
    if (value == 0 || value == -1) {   // a byte pattern: a memset node instead of a kernel
        if (cudaMemsetAsync(p, value == 0 ? 0 : 0xff, count * sizeof(int), s) != cudaSuccess) return MSSVT_ERR_LAUNCH;
        return MSSVT_OK;
    }
    size_t n16 = count / 4;
    int ntail = (int)(count % 4);
    int grid = persistent_grid((long long)(n16 ? n16 : 1), 256, 8);
    ++g_launches;
    k_fill_i32x4<<<grid, 256, 0, s>>>((int4 *)p, n16, p + n16 * 4, ntail, value);
    return check_launch();
}

// ------------------------------------------------------------------------------- hash build

template <int MAXB>
__device__ __forceinline__ int sample_start(const int *__restrict__ v_bs_cnt, int b) {
    int s = 0;
    for (int i = 0; i < b; ++i) s += __ldg(v_bs_cnt + i);
    return s;
}

__global__ void k_hash_insert(int x_max, int y_max, int z_max, int num_voxels, int hash_size,
                              const int4 *__restrict__ v_indices, const int *__restrict__ v_bs_cnt,
                              unsigned long long *__restrict__ table) {
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= num_voxels) return;
    int4 c = __ldg(v_indices + t);  // [b, z, y, x]
    int b = c.x, z = c.y, y = c.z, x = c.w;
    if (x >= x_max || x < 0 || y < 0 || y >= y_max || z < 0 || z >= z_max) return;
    int local = t - sample_start<0>(v_bs_cnt, b);
    int key = x * y_max * z_max + y * z_max + z;
    unsigned long long *tab = table + (size_t)b * hash_size;
    unsigned long long packed = (unsigned long long)(unsigned)key | ((unsigned long long)(unsigned)local << 32);
    int slot = (int)((unsigned)key % (unsigned)hash_size);
    for (int probes = 0; probes < hash_size; ++probes) {
        // walk occupied slots with plain loads (keys never change once written); only an empty
        // slot costs an atomic.  `k % H` clusters badly on voxel keys, so chains can be long.
        unsigned long long prev = *((volatile unsigned long long *)(tab + slot));
        if (prev == ~0ull) prev = atomicCAS(tab + slot, ~0ull, packed);
        if (prev == ~0ull) return;
        if ((int)(unsigned)(prev & 0xffffffffull) == key) {
            // duplicate coordinate: the reference lets the last writer win; sequential order
            // (the oracle) means the highest voxel index, which atomicMax reproduces.
            atomicMax((int *)(tab + slot) + 1, local);
            return;
        }
        slot = slot + 1 == hash_size ? 0 : slot + 1;
    }
}

__global__ void k_hash_lookup(int hash_size, int n, const int *__restrict__ batch_ids,
                              const int *__restrict__ keys, const int2 *__restrict__ table,
                              int *__restrict__ values) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    values[i] = table_find(table + (size_t)batch_ids[i] * hash_size, hash_size, keys[i]);
}

// world coordinates of every voxel, (N, 3) fp32 [x, y, z] (with_coords, mssvt_backbone.py:132-137)
__global__ void k_world_coords(int n, const int4 *__restrict__ v_indices, float vx, float vy,
                               float vz, float lx, float ly, float lz, float *__restrict__ xyz) {
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    int4 c = __ldg(v_indices + t);
    xyz[3 * t + 0] = world_coord(c.w, vx, lx);
    xyz[3 * t + 1] = world_coord(c.z, vy, ly);
    xyz[3 * t + 2] = world_coord(c.y, vz, lz);
}

// per-sample voxel counts -> v_bs_cnt (B) and exclusive prefix v_start (B + 1); replaces the
// per-sample `.sum().item()` loops (mssvt_utils.py:35-38, mssvt_backbone.py:124-130).
__global__ void k_count_samples(int n, int batch_size, const int4 *__restrict__ v_indices,
                                int *__restrict__ counts) {
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    int b = t < n ? __ldg(v_indices + t).x : -1;
    // samples are contiguous, so a warp usually holds one or two batch ids
    unsigned active = __activemask();
    if (b >= 0 && b < batch_size) {
        unsigned peers = __match_any_sync(__activemask(), b);
        if ((peers & lanemask_lt()) == 0) atomicAdd(counts + b, __popc(peers));
    }
    (void)active;
}

__global__ void k_prefix_small(int batch_size, const int *__restrict__ counts, int *__restrict__ start) {
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        int s = 0;
        for (int b = 0; b < batch_size; ++b) { start[b] = s; s += counts[b]; }
        start[batch_size] = s;
    }
}

// ------------------------------------------------------------------------------- strided scan
// dst[i] = sum_{j < i} src[j * stride], dst[n] = total, for n = min(n_cap, *n_dev) read on the device
// (two-level: 1024-element block sums, then prefix of the sums + in-block scan).

__global__ void __launch_bounds__(1024)
k_scan_block_sums(int n_cap, const int *__restrict__ n_dev, const int *__restrict__ src, int stride,
                  int *__restrict__ block_sums) {
    __shared__ int s_warp[32];
    const int n = n_dev ? min(n_cap, __ldg(n_dev)) : n_cap;
    const int i = blockIdx.x * 1024 + threadIdx.x;
    int v = i < n ? __ldg(src + (size_t)i * stride) : 0;
    v = __reduce_add_sync(0xffffffffu, v);
    if ((threadIdx.x & 31) == 0) s_warp[threadIdx.x >> 5] = v;
    __syncthreads();
    if (threadIdx.x < 32) {
        int t = s_warp[threadIdx.x];
        t = __reduce_add_sync(0xffffffffu, t);
        if (threadIdx.x == 0) block_sums[blockIdx.x] = t;
    }
}

__global__ void __launch_bounds__(1024)
k_scan_emit(int n_cap, const int *__restrict__ n_dev, const int *__restrict__ src, int stride,
            const int *__restrict__ block_sums, int *__restrict__ dst) {
    __shared__ int s_warp[32];
    __shared__ int s_base;
    const int n = n_dev ? min(n_cap, __ldg(n_dev)) : n_cap;
    if (blockIdx.x * 1024 > n) return;  // (the block holding index n still writes the total)
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int part = 0;
    if (block_sums) {
        for (int b = threadIdx.x; b < (int)blockIdx.x; b += 1024) part += block_sums[b];
    } else {
        // single pass: the block sums everything in front of it itself (b independent loads per thread; the lists this
        // scans are 10^4..10^5 long, so the last block reads ~100 elements per thread out of L2 -- cheaper than a second
        // kernel in the serial coordinate prefix of a frame)
        const int limit = (int)blockIdx.x * 1024;
        int j = threadIdx.x, p1 = 0, p2 = 0, p3 = 0, p4 = 0, p5 = 0, p6 = 0, p7 = 0;
        for (; j + 7 * 1024 < limit; j += 8 * 1024) {   // eight independent loads in flight per thread
            part += __ldg(src + (size_t)j * stride);              p1 += __ldg(src + (size_t)(j + 1024) * stride);
            p2 += __ldg(src + (size_t)(j + 2048) * stride);       p3 += __ldg(src + (size_t)(j + 3072) * stride);
            p4 += __ldg(src + (size_t)(j + 4096) * stride);       p5 += __ldg(src + (size_t)(j + 5120) * stride);
            p6 += __ldg(src + (size_t)(j + 6144) * stride);       p7 += __ldg(src + (size_t)(j + 7168) * stride);
        }
        for (; j < limit; j += 1024) part += __ldg(src + (size_t)j * stride);
        part += p1 + p2 + p3 + p4 + p5 + p6 + p7;
    }
    part = __reduce_add_sync(0xffffffffu, part);
    if (lane == 0) s_warp[warp] = part;
    __syncthreads();
    if (threadIdx.x == 0) {
        int t = 0;
        for (int w = 0; w < 32; ++w) t += s_warp[w];
        s_base = t;
    }
    __syncthreads();
    const int base = s_base;
    __syncthreads();
    const int i = blockIdx.x * 1024 + threadIdx.x;
    const int v = i < n ? __ldg(src + (size_t)i * stride) : 0;
    int incl = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        int t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += t;
    }
    if (lane == 31) s_warp[warp] = incl;
    __syncthreads();
    int before = 0;
    for (int w = 0; w < warp; ++w) before += s_warp[w];
    if (i <= n) dst[i] = base + before + incl - v;  // i == n: the total
}

// ------------------------------------------------------------------------------- grid index
// Occupancy bitmap + rank (GridIdx in common.cuh), built in four small passes:
//   bits:  atomicOr one bit per voxel            scan: per-1024-word block popcount sums
//   base:  exclusive prefix of the popcounts     vals: vals[base + rank] = per-sample voxel index

__global__ void k_grid_bits(int x_max, int y_max, int z_max, int zw, long long words_per_sample,
                            int num_voxels, const int4 *__restrict__ v_indices, int2 *__restrict__ cells) {
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= num_voxels) return;
    int4 c = __ldg(v_indices + t);
    int z = c.y, y = c.z, x = c.w;
    if (x >= x_max || x < 0 || y < 0 || y >= y_max || z < 0 || z >= z_max) return;
    atomicOr(&cells[(size_t)c.x * words_per_sample + ((size_t)x * y_max + y) * zw + (z >> 5)].x, 1 << (z & 31));
}

#define GRID_BLOCK 1024
__global__ void __launch_bounds__(GRID_BLOCK)
k_grid_block_sums(long long total_words, const int2 *__restrict__ cells, int *__restrict__ block_sums) {
    __shared__ int s_warp[GRID_BLOCK / 32];
    long long i = (long long)blockIdx.x * GRID_BLOCK + threadIdx.x;
    int c = i < total_words ? __popc((unsigned)cells[i].x) : 0;
    c = __reduce_add_sync(0xffffffffu, c);
    if ((threadIdx.x & 31) == 0) s_warp[threadIdx.x >> 5] = c;
    __syncthreads();
    if (threadIdx.x < 32) {
        int v = s_warp[threadIdx.x];
        v = __reduce_add_sync(0xffffffffu, v);
        if (threadIdx.x == 0) block_sums[blockIdx.x] = v;
    }
}

__global__ void __launch_bounds__(GRID_BLOCK)
k_grid_base(long long total_words, int2 *__restrict__ cells, const int *__restrict__ block_sums) {
    __shared__ int s_warp[GRID_BLOCK / 32];
    __shared__ int s_base;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int part = 0;
    for (int i = threadIdx.x; i < (int)blockIdx.x; i += GRID_BLOCK) part += block_sums[i];
    part = __reduce_add_sync(0xffffffffu, part);
    if (lane == 0) s_warp[warp] = part;
    __syncthreads();
    if (threadIdx.x == 0) {
        int s = 0;
        for (int i = 0; i < GRID_BLOCK / 32; ++i) s += s_warp[i];
        s_base = s;
    }
    __syncthreads();
    const int base = s_base;
    __syncthreads();
    long long i = (long long)blockIdx.x * GRID_BLOCK + threadIdx.x;
    int c = i < total_words ? __popc((unsigned)cells[i].x) : 0;
    // inclusive warp scan, then add the sums of the warps before
    int incl = c;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        int n = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += n;
    }
    if (lane == 31) s_warp[warp] = incl;
    __syncthreads();
    int before = 0;
    for (int w = 0; w < warp; ++w) before += s_warp[w];
    if (i < total_words) cells[i].y = base + before + incl - c;
}

__global__ void k_grid_vals(int x_max, int y_max, int z_max, int zw, long long words_per_sample,
                            int num_voxels, const int4 *__restrict__ v_indices,
                            const int *__restrict__ v_start, const int2 *__restrict__ cells,
                            int *__restrict__ vals) {
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= num_voxels) return;
    int4 c = __ldg(v_indices + t);
    int z = c.y, y = c.z, x = c.w;
    if (x >= x_max || x < 0 || y < 0 || y >= y_max || z < 0 || z >= z_max) return;
    int2 cell = cells[(size_t)c.x * words_per_sample + ((size_t)x * y_max + y) * zw + (z >> 5)];
    unsigned bit = 1u << (z & 31);
    vals[cell.y + __popc((unsigned)cell.x & (bit - 1u))] = t - __ldg(v_start + c.x);
}

// ------------------------------------------------------------------------------- window partition

#define WIN_BLOCK 1024

// pass A: claim the window's slot, remember the lowest voxel index that maps to it
__global__ void k_win_insert(int x_wgs, int y_wgs, int z_wgs, int x_ws, int y_ws, int z_ws,
                             int num_voxels, int hash_size, const int4 *__restrict__ v_indices,
                             int *__restrict__ table, int *__restrict__ slot_of) {
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= num_voxels) return;
    int4 c = __ldg(v_indices + t);
    int wz = c.y / z_ws, wy = c.z / y_ws, wx = c.w / x_ws;
    int found = -1;
    if (!(wx < 0 || wx >= x_wgs || wy < 0 || wy >= y_wgs || wz < 0 || wz >= z_wgs)) {
        int *tab = table + (size_t)c.x * hash_size * 2;
        int key = wx * y_wgs * z_wgs + wy * z_wgs + wz;
        int slot = (int)((unsigned)key % (unsigned)hash_size);
        for (int probes = 0; probes < hash_size; ++probes) {
            int prev = atomicCAS(tab + 2 * slot, MSSVT_EMPTY, key);
            if (prev == MSSVT_EMPTY || prev == key) {
                atomicMin((unsigned *)(tab + 2 * slot + 1), (unsigned)t);
                found = slot;
                break;
            }
            slot = slot + 1 == hash_size ? 0 : slot + 1;
        }
    }
    slot_of[t] = found;
}

// pass B: a voxel opens a window iff it is the lowest index in the slot; count per block and
// per sample.  slot_of[t] is set to -1 for every other voxel.
__global__ void __launch_bounds__(WIN_BLOCK)
k_win_count(int num_voxels, int hash_size, int batch_size, const int4 *__restrict__ v_indices,
            const int *__restrict__ table, int *__restrict__ slot_of, int *__restrict__ block_counts,
            int *__restrict__ win_count) {
    __shared__ int s_total;
    if (threadIdx.x == 0) s_total = 0;
    __syncthreads();
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    int opens = 0, b = -1;
    if (t < num_voxels) {
        int slot = slot_of[t];
        if (slot >= 0) {
            b = __ldg(v_indices + t).x;
            opens = table[((size_t)b * hash_size + slot) * 2 + 1] == t;
            if (!opens) slot_of[t] = -1;
        }
    }
    unsigned ball = __ballot_sync(0xffffffffu, opens);
    if ((threadIdx.x & 31) == 0 && ball) atomicAdd(&s_total, __popc(ball));
    if (opens) {
        unsigned peers = __match_any_sync(__activemask(), b);
        if ((peers & lanemask_lt()) == 0) atomicAdd(win_count + b, __popc(peers));
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        block_counts[blockIdx.x] = s_total;
        if (s_total) atomicAdd(win_count + batch_size, s_total);
    }
}

// pass C: rank = (#openers in earlier blocks) + (#openers earlier in this block); row `rank` of
// the concatenated list gets [b, wz, wy, wx]; the table value becomes the per-sample row id.
__global__ void __launch_bounds__(WIN_BLOCK)
k_win_emit(int x_ws, int y_ws, int z_ws, int num_voxels, int hash_size, int batch_size,
           int max_wins, int list_capacity, const int4 *__restrict__ v_indices,
           int *__restrict__ table, const int *__restrict__ slot_of,
           const int *__restrict__ block_counts, int *win_count,
           int4 *__restrict__ win_list, int *__restrict__ overflow) {
    __shared__ int s_warp[WIN_BLOCK / 32];
    __shared__ int s_base;
    int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    // prefix over earlier blocks
    int part = 0;
    for (int i = threadIdx.x; i < blockIdx.x; i += blockDim.x) part += block_counts[i];
    part = __reduce_add_sync(0xffffffffu, part);
    if (lane == 0) s_warp[warp] = part;
    __syncthreads();
    if (threadIdx.x == 0) {
        int s = 0;
        for (int i = 0; i < WIN_BLOCK / 32; ++i) s += s_warp[i];
        s_base = s;
    }
    __syncthreads();
    int base = s_base;
    __syncthreads();

    // win_count[B] (all openers, written by k_win_count and read by nobody in this kernel) becomes the number of
    // rows actually in the list, so that every consumer's loop bound is the kept count
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        int kept = 0;
        for (int i = 0; i < batch_size; ++i) kept += min(win_count[i], max_wins);
        win_count[batch_size] = min(kept, list_capacity);
    }
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    int slot = t < num_voxels ? slot_of[t] : -1;
    int opens = slot >= 0;
    unsigned ball = __ballot_sync(0xffffffffu, opens);
    if (lane == 0) s_warp[warp] = __popc(ball);
    __syncthreads();
    int before = 0;
    for (int i = 0; i < warp; ++i) before += s_warp[i];
    if (!opens) return;
    int rank = base + before + __popc(ball & lanemask_lt());
    int4 c = __ldg(v_indices + t);
    // A sample keeps its first max_wins windows; the kept windows of all samples are COMPACTED (no holes
    // between samples when one overflows), and the list holds list_capacity rows at most.
    int first_row = 0, first_kept = 0;
    for (int i = 0; i < c.x; ++i) {
        int n = win_count[i];
        first_row += n;
        first_kept += min(n, max_wins);
    }
    int local = rank - first_row, row = first_kept + local;
    if (local >= max_wins || row >= list_capacity) {
        atomicAdd(overflow, 1);
        return;
    }
    win_list[row] = make_int4(c.x, c.y / z_ws, c.z / y_ws, c.w / x_ws);
    table[((size_t)c.x * hash_size + slot) * 2 + 1] = local;
}

// ------------------------------------------------------------------------------- window partition, dense form
//
// The fused path needs the window LIST only (voxels are found through the grid index, not through the window
// hash), so it numbers the windows through a dense array over the window grid: first[b][cell] = lowest voxel
// index that falls into the window (atomicMin), a voxel opens its window iff it is that voxel, openers are
// ranked in voxel order (same first-occurrence numbering as the hash form, bit-identical lists).  No probing,
// no (B, H, 2) table to fill: 1 MB instead of 3.2 MB of initialisation for the S0 grid.

__device__ __forceinline__ long long wd_cell(int4 c, int x_wgs, int y_wgs, int z_wgs, int x_ws, int y_ws, int z_ws) {
    const int wz = c.y / z_ws, wy = c.z / y_ws, wx = c.w / x_ws;
    if (c.y < 0 || c.z < 0 || c.w < 0 || wx >= x_wgs || wy >= y_wgs || wz >= z_wgs) return -1;
    return ((long long)c.x * x_wgs + wx) * y_wgs * z_wgs + (long long)wy * z_wgs + wz;
}

__global__ void k_wd_first(int x_wgs, int y_wgs, int z_wgs, int x_ws, int y_ws, int z_ws, int num_voxels,
                           const int4 *__restrict__ v_indices, int *__restrict__ first) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= num_voxels) return;
    const long long cell = wd_cell(__ldg(v_indices + t), x_wgs, y_wgs, z_wgs, x_ws, y_ws, z_ws);
    if (cell >= 0) atomicMin(first + cell, t);
}

// openers per block and per sample (opens[t] kept as a flag byte for the emit pass)
__global__ void __launch_bounds__(WIN_BLOCK)
k_wd_count(int x_wgs, int y_wgs, int z_wgs, int x_ws, int y_ws, int z_ws, int num_voxels, int batch_size,
           const int4 *__restrict__ v_indices, const int *__restrict__ first, unsigned char *__restrict__ opens_flag,
           int *__restrict__ block_counts, int *__restrict__ win_count) {
    __shared__ int s_total;
    if (threadIdx.x == 0) s_total = 0;
    __syncthreads();
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    int opens = 0, b = -1;
    if (t < num_voxels) {
        const int4 c = __ldg(v_indices + t);
        const long long cell = wd_cell(c, x_wgs, y_wgs, z_wgs, x_ws, y_ws, z_ws);
        b = c.x;
        opens = cell >= 0 && first[cell] == t;
        opens_flag[t] = (unsigned char)opens;
    }
    const unsigned ball = __ballot_sync(0xffffffffu, opens);
    if ((threadIdx.x & 31) == 0 && ball) atomicAdd(&s_total, __popc(ball));
    if (opens) {
        const unsigned peers = __match_any_sync(__activemask(), b);
        if ((peers & lanemask_lt()) == 0) atomicAdd(win_count + b, __popc(peers));
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        block_counts[blockIdx.x] = s_total;
        if (s_total) atomicAdd(win_count + batch_size, s_total);
    }
}

__global__ void __launch_bounds__(WIN_BLOCK)
k_wd_emit(int x_ws, int y_ws, int z_ws, int num_voxels, int batch_size, int max_wins, int list_capacity,
          const int4 *__restrict__ v_indices, const unsigned char *__restrict__ opens_flag,
          const int *__restrict__ block_counts, int *win_count, int4 *__restrict__ win_list,
          int *__restrict__ overflow) {
    __shared__ int s_warp[WIN_BLOCK / 32];
    __shared__ int s_base;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int part = 0;
    for (int i = threadIdx.x; i < blockIdx.x; i += blockDim.x) part += block_counts[i];
    part = __reduce_add_sync(0xffffffffu, part);
    if (lane == 0) s_warp[warp] = part;
    __syncthreads();
    if (threadIdx.x == 0) {
        int sum = 0;
        for (int i = 0; i < WIN_BLOCK / 32; ++i) sum += s_warp[i];
        s_base = sum;
    }
    __syncthreads();
    const int base = s_base;
    __syncthreads();
    if (blockIdx.x == 0 && threadIdx.x == 0) {   // (see k_win_emit: the list length every consumer uses)
        int kept = 0;
        for (int i = 0; i < batch_size; ++i) kept += min(win_count[i], max_wins);
        win_count[batch_size] = min(kept, list_capacity);
    }
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    const int opens = t < num_voxels ? opens_flag[t] : 0;
    const unsigned ball = __ballot_sync(0xffffffffu, opens);
    if (lane == 0) s_warp[warp] = __popc(ball);
    __syncthreads();
    int before = 0;
    for (int i = 0; i < warp; ++i) before += s_warp[i];
    if (!opens) return;
    const int rank = base + before + __popc(ball & lanemask_lt());
    const int4 c = __ldg(v_indices + t);
    int first_row = 0, first_kept = 0;
    for (int i = 0; i < c.x; ++i) {
        const int n = win_count[i];
        first_row += n;
        first_kept += min(n, max_wins);
    }
    const int local = rank - first_row, row = first_kept + local;
    if (local >= max_wins || row >= list_capacity) {
        atomicAdd(overflow, 1);
        return;
    }
    win_list[row] = make_int4(c.x, c.y / z_ws, c.z / y_ws, c.w / x_ws);
}

}  // namespace mssvt

using namespace mssvt;

extern "C" {

long long mssvt_window_list_workspace_bytes(int x_wgs, int y_wgs, int z_wgs, int batch_size, int num_voxels) {
    const long long cells = (long long)batch_size * x_wgs * y_wgs * z_wgs;
    const long long blocks = (num_voxels + WIN_BLOCK - 1) / WIN_BLOCK;
    return (cells + blocks + 8) * 4 + ((long long)num_voxels + 15) / 16 * 16;
}

/* The window LIST of mssvt_window_partition without the window hash table (what the fused path needs: it
 * finds voxels through the grid index).  Same rows in the same first-occurrence order, same win_count layout
 * (per-sample counts, [B] rows in the list, [B + 1] windows dropped).  workspace:
 * mssvt_window_list_workspace_bytes(...) bytes. */
int mssvt_window_list(int x_wgs, int y_wgs, int z_wgs, int x_ws, int y_ws, int z_ws, int num_voxels, int max_wins,
                      int batch_size, int list_capacity, const int *v_indices, int *win_list, int *win_count,
                      void *workspace, long long workspace_bytes, void *stream) {
    if (!win_count || batch_size <= 0 || num_voxels < 0 || x_ws <= 0 || y_ws <= 0 || z_ws <= 0 || x_wgs <= 0 ||
        y_wgs <= 0 || z_wgs <= 0)
        return MSSVT_ERR_INVALID;
    if (workspace_bytes < mssvt_window_list_workspace_bytes(x_wgs, y_wgs, z_wgs, batch_size, num_voxels))
        return MSSVT_ERR_WORKSPACE;
    cudaStream_t s = (cudaStream_t)stream;
    int rc = fill_i32(win_count, batch_size + 2, 0, s);
    if (rc || num_voxels == 0) return rc;
    if (!v_indices || !win_list || !workspace) return MSSVT_ERR_INVALID;
    const long long cells = (long long)batch_size * x_wgs * y_wgs * z_wgs;
    const int blocks = div_up(num_voxels, WIN_BLOCK);
    int *first = (int *)workspace;
    int *block_counts = first + cells;
    unsigned char *opens_flag = (unsigned char *)(block_counts + blocks + 8);
    cudaError_t e = cudaMemsetAsync(first, 0x7f, (size_t)cells * 4, s);   // 0x7f7f7f7f: above every voxel index
    if (e != cudaSuccess) { g_last_cuda_error = (int)e; return MSSVT_ERR_LAUNCH; }
    ++g_launches;
    k_wd_first<<<div_up(num_voxels, 256), 256, 0, s>>>(x_wgs, y_wgs, z_wgs, x_ws, y_ws, z_ws, num_voxels,
                                                       (const int4 *)v_indices, first);
    ++g_launches;
    k_wd_count<<<blocks, WIN_BLOCK, 0, s>>>(x_wgs, y_wgs, z_wgs, x_ws, y_ws, z_ws, num_voxels, batch_size,
                                            (const int4 *)v_indices, first, opens_flag, block_counts, win_count);
    ++g_launches;
    k_wd_emit<<<blocks, WIN_BLOCK, 0, s>>>(x_ws, y_ws, z_ws, num_voxels, batch_size, max_wins, list_capacity,
                                           (const int4 *)v_indices, opens_flag, block_counts, win_count,
                                           (int4 *)win_list, win_count + batch_size + 1);
    return check_launch();
}

int mssvt_last_cuda_error(void) { return g_last_cuda_error; }
long long mssvt_launch_count(void) { return g_launches; }

const char *mssvt_version(void) { return "mssvt_b200 0.1 (sm_100a)"; }

int mssvt_fill_i32(int *p, long long count, int value, void *stream) {
    if (!p || count < 0) return MSSVT_ERR_INVALID;
    return fill_i32(p, (size_t)count, value, (cudaStream_t)stream);
}

int mssvt_build_hash_table(int x_max, int y_max, int z_max, int num_voxels, int hash_size,
                           int batch_size, const int *v_indices, const int *v_bs_cnt, int *table,
                           void *stream) {
    if (!table || hash_size <= 0 || batch_size <= 0 || num_voxels < 0) return MSSVT_ERR_INVALID;
    if (num_voxels && (!v_indices || !v_bs_cnt)) return MSSVT_ERR_INVALID;
    cudaStream_t s = (cudaStream_t)stream;
    int rc = fill_i32(table, (size_t)batch_size * hash_size * 2, MSSVT_EMPTY, s);
    if (rc || num_voxels == 0) return rc;
    ++g_launches;
    k_hash_insert<<<div_up(num_voxels, 256), 256, 0, s>>>(
        x_max, y_max, z_max, num_voxels, hash_size, (const int4 *)v_indices, v_bs_cnt,
        (unsigned long long *)table);
    return check_launch();
}

int mssvt_hash_lookup(int hash_size, int num_queries, const int *batch_ids, const int *keys,
                      const int *table, int *values, void *stream) {
    if (num_queries < 0 || hash_size <= 0) return MSSVT_ERR_INVALID;
    if (num_queries == 0) return MSSVT_OK;
    if (!batch_ids || !keys || !table || !values) return MSSVT_ERR_INVALID;
    ++g_launches;
    k_hash_lookup<<<div_up(num_queries, 256), 256, 0, (cudaStream_t)stream>>>(
        hash_size, num_queries, batch_ids, keys, (const int2 *)table, values);
    return check_launch();
}

int mssvt_voxel_world_coords(int num_voxels, const int *v_indices, const float *voxel_size,
                             const float *range_min, float *xyz, void *stream) {
    if (num_voxels < 0 || !voxel_size || !range_min) return MSSVT_ERR_INVALID;
    if (num_voxels == 0) return MSSVT_OK;
    if (!v_indices || !xyz) return MSSVT_ERR_INVALID;
    ++g_launches;
    k_world_coords<<<div_up(num_voxels, 256), 256, 0, (cudaStream_t)stream>>>(
        num_voxels, (const int4 *)v_indices, voxel_size[0], voxel_size[1], voxel_size[2],
        range_min[0], range_min[1], range_min[2], xyz);
    return check_launch();
}

int mssvt_count_samples(int num_rows, int batch_size, const int *indices, int *counts, int *start,
                        void *stream) {
    if (num_rows < 0 || batch_size <= 0 || !counts || !start) return MSSVT_ERR_INVALID;
    cudaStream_t s = (cudaStream_t)stream;
    int rc = fill_i32(counts, batch_size, 0, s);
    if (rc) return rc;
    if (num_rows) {
        if (!indices) return MSSVT_ERR_INVALID;
        ++g_launches;
        k_count_samples<<<div_up(num_rows, 256), 256, 0, s>>>(num_rows, batch_size,
                                                              (const int4 *)indices, counts);
    }
    ++g_launches;
    k_prefix_small<<<1, 32, 0, s>>>(batch_size, counts, start);
    return check_launch();
}

/* Grid index of the fused path (GridIdx): cells (B * x*y*ceil(z/32), 2) int32, vals (N) int32.
 * workspace: ceil(words / 1024) + 1 ints. */
long long mssvt_grid_index_words(int x_max, int y_max, int z_max, int batch_size) {
    return (long long)batch_size * x_max * y_max * ((z_max + 31) / 32);
}

int mssvt_grid_index_build(int x_max, int y_max, int z_max, int num_voxels, int batch_size,
                           const int *v_indices, const int *v_start, int *cells, int *vals,
                           int *workspace, void *stream) {
    if (x_max <= 0 || y_max <= 0 || z_max <= 0 || batch_size <= 0 || num_voxels < 0 || !cells)
        return MSSVT_ERR_INVALID;
    cudaStream_t s = (cudaStream_t)stream;
    const int zw = (z_max + 31) / 32;
    const long long wps = (long long)x_max * y_max * zw, total = wps * batch_size;
    cudaError_t e = cudaMemsetAsync(cells, 0, (size_t)total * 8, s);
    if (e != cudaSuccess) { g_last_cuda_error = (int)e; return MSSVT_ERR_LAUNCH; }
    if (num_voxels == 0) return MSSVT_OK;
    if (!v_indices || !v_start || !vals || !workspace) return MSSVT_ERR_INVALID;
    const int blocks = div_up(total, GRID_BLOCK);
    ++g_launches;
    k_grid_bits<<<div_up(num_voxels, 256), 256, 0, s>>>(x_max, y_max, z_max, zw, wps, num_voxels,
                                                        (const int4 *)v_indices, (int2 *)cells);
    ++g_launches;
    k_grid_block_sums<<<blocks, GRID_BLOCK, 0, s>>>(total, (const int2 *)cells, workspace);
    ++g_launches;
    k_grid_base<<<blocks, GRID_BLOCK, 0, s>>>(total, (int2 *)cells, workspace);
    ++g_launches;
    k_grid_vals<<<div_up(num_voxels, 256), 256, 0, s>>>(x_max, y_max, z_max, zw, wps, num_voxels,
                                                        (const int4 *)v_indices, v_start,
                                                        (const int2 *)cells, vals);
    return check_launch();
}

/* dst[i] = sum of src[j * stride] for j < i, i in [0, n]; n = min(n_cap, *n_dev) (n_dev may be NULL).
 * dst holds n_cap + 1 ints, workspace ceil((n_cap + 1) / 1024) + 1 ints. */
int mssvt_exclusive_scan(int n_cap, const int *n_dev, const int *src, int stride, int *dst, int *workspace,
                         void *stream) {
    if (n_cap < 0 || stride <= 0 || !dst || !workspace || (n_cap && !src)) return MSSVT_ERR_INVALID;
    cudaStream_t s = (cudaStream_t)stream;
    const int blocks = div_up((long long)n_cap + 1, 1024);
    if (n_cap <= 262144) {   // one pass: every block sums what precedes it itself (see k_scan_emit)
        ++g_launches;
        k_scan_emit<<<blocks, 1024, 0, s>>>(n_cap, n_dev, src, stride, nullptr, dst);
        return check_launch();
    }
    ++g_launches;
    k_scan_block_sums<<<blocks, 1024, 0, s>>>(n_cap, n_dev, src, stride, workspace);
    ++g_launches;
    k_scan_emit<<<blocks, 1024, 0, s>>>(n_cap, n_dev, src, stride, workspace, dst);
    return check_launch();
}

long long mssvt_window_partition_workspace_bytes(int num_voxels) {
    long long blocks = (num_voxels + WIN_BLOCK - 1) / WIN_BLOCK;
    return ((long long)num_voxels + blocks + 8) * 4;
}

// win_count: (batch_size + 2) int32 -> per-sample counts (before clamping), [B] rows in the list (kept windows,
// compacted), [B+1] windows that did not fit (max_wins per sample / list_capacity)
int mssvt_window_partition(int x_wgs, int y_wgs, int z_wgs, int x_ws, int y_ws, int z_ws,
                           int num_voxels, int max_wins, int hash_size, int batch_size,
                           int list_capacity, const int *v_indices, int *win_list, int *table,
                           int *win_count, void *workspace, long long workspace_bytes,
                           void *stream) {
    if (!table || !win_count || hash_size <= 0 || batch_size <= 0 || num_voxels < 0 ||
        x_ws <= 0 || y_ws <= 0 || z_ws <= 0)
        return MSSVT_ERR_INVALID;
    if (workspace_bytes < mssvt_window_partition_workspace_bytes(num_voxels)) return MSSVT_ERR_WORKSPACE;
    cudaStream_t s = (cudaStream_t)stream;
    int rc = fill_i32(table, (size_t)batch_size * hash_size * 2, MSSVT_EMPTY, s);
    if (rc) return rc;
    rc = fill_i32(win_count, batch_size + 2, 0, s);
    if (rc || num_voxels == 0) return rc;
    if (!v_indices || !win_list || !workspace) return MSSVT_ERR_INVALID;
    int *slot_of = (int *)workspace;
    int *block_counts = slot_of + num_voxels;
    int blocks = div_up(num_voxels, WIN_BLOCK);
    ++g_launches;
    k_win_insert<<<div_up(num_voxels, 256), 256, 0, s>>>(x_wgs, y_wgs, z_wgs, x_ws, y_ws, z_ws,
                                                         num_voxels, hash_size,
                                                         (const int4 *)v_indices, table, slot_of);
    ++g_launches;
    k_win_count<<<blocks, WIN_BLOCK, 0, s>>>(num_voxels, hash_size, batch_size,
                                             (const int4 *)v_indices, table, slot_of, block_counts,
                                             win_count);
    ++g_launches;
    k_win_emit<<<blocks, WIN_BLOCK, 0, s>>>(x_ws, y_ws, z_ws, num_voxels, hash_size, batch_size,
                                            max_wins, list_capacity, (const int4 *)v_indices, table,
                                            slot_of, block_counts, win_count, (int4 *)win_list,
                                            win_count + batch_size + 1);
    return check_launch();
}

}  // extern "C"
