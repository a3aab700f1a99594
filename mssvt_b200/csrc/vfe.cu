// vfe.cu -- dynamic voxelisation + point-feature encoder: the producer of the backbone's inputs
// (pcdet/models/backbones_3d/vfe/dynamic_vfe.py:71-130, eval mode).
//
//   reference                                           here
//   torch.unique(merge_coords, sorted, inverse) :93     occupancy bitmap over the voxel grid in key order
//                                                       (b, x, y, z) + popcount scan: the rank of a set bit IS
//                                                       the row of its voxel in the sorted unique list; no sort
//   torch_scatter.scatter_mean(xyz) :98                 atomicAdd of (x, y, z, 1) per point
//   PFN: Linear + BatchNorm1d + ReLU :124-130           BatchNorm (eval) folded into the linear layer by the
//                                                       caller; one warp per point, lane = output channel
//   torch_scatter.scatter_max :108, :128                atomicMax on the int view (values are >= 0 after ReLU)
#include "tc_common.cuh"

namespace mssvt {

struct VfeGrid {
    int gx, gy, gz, zw, batch;       // zw = words per z column
    float vs[3], lo[3];
};

// floor((p - lo) / vs) exactly as the reference computes it in fp32 (dynamic_vfe.py:85); false if outside
__device__ __forceinline__ bool vfe_cell(const VfeGrid &G, const float *pt, int &b, int &x, int &y, int &z) {
    b = (int)pt[0];
    x = (int)floorf(__fdiv_rn(__fsub_rn(pt[1], G.lo[0]), G.vs[0]));
    y = (int)floorf(__fdiv_rn(__fsub_rn(pt[2], G.lo[1]), G.vs[1]));
    z = (int)floorf(__fdiv_rn(__fsub_rn(pt[3], G.lo[2]), G.vs[2]));
    return x >= 0 && x < G.gx && y >= 0 && y < G.gy && z >= 0 && z < G.gz && b >= 0 && b < G.batch;
}
__device__ __forceinline__ long long vfe_word(const VfeGrid &G, int b, int x, int y, int z) {
    return (((long long)b * G.gx + x) * G.gy + y) * G.zw + (z >> 5);
}

__global__ void k_vfe_bits(VfeGrid G, int n, const float *__restrict__ points, int stride, unsigned *__restrict__ bits) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n) return;
    int b, x, y, z;
    if (vfe_cell(G, points + (size_t)p * stride, b, x, y, z)) atomicOr(bits + vfe_word(G, b, x, y, z), 1u << (z & 31));
}

__global__ void k_vfe_popc(long long words, const unsigned *__restrict__ bits, int *__restrict__ pc) {
    const long long w = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (w < words) pc[w] = __popc(bits[w]);
}

// point -> voxel row; per-voxel coordinate sums for the cluster centre
__global__ void k_vfe_assign(VfeGrid G, int n, const float *__restrict__ points, int stride,
                             const unsigned *__restrict__ bits, const int *__restrict__ base,
                             int *__restrict__ point_voxel, float *__restrict__ xyz_sum) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n) return;
    const float *pt = points + (size_t)p * stride;
    int b, x, y, z, v = -1;
    if (vfe_cell(G, pt, b, x, y, z)) {
        const long long w = vfe_word(G, b, x, y, z);
        v = base[w] + __popc(bits[w] & ((1u << (z & 31)) - 1u));
        if (xyz_sum) {
            atomicAdd(xyz_sum + 4 * (size_t)v, pt[1]); atomicAdd(xyz_sum + 4 * (size_t)v + 1, pt[2]);
            atomicAdd(xyz_sum + 4 * (size_t)v + 2, pt[3]); atomicAdd(xyz_sum + 4 * (size_t)v + 3, 1.0f);
        }
    }
    point_voxel[p] = v;
}

// voxel_coords[row] = [b, z, y, x] for every set bit, rows in key order (dynamic_vfe.py:111-116)
__global__ void k_vfe_coords(VfeGrid G, long long words, const unsigned *__restrict__ bits,
                             const int *__restrict__ base, int4 *__restrict__ voxel_coords) {
    const long long w = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (w >= words) return;
    unsigned m = bits[w];
    if (!m) return;
    const int zwi = (int)(w % G.zw);
    long long r = w / G.zw;
    const int y = (int)(r % G.gy); r /= G.gy;
    const int x = (int)(r % G.gx), b = (int)(r / G.gx);
    int row = base[w];
    while (m) {
        const int bit = __ffs(m) - 1;
        m &= m - 1;
        voxel_coords[row++] = make_int4(b, zwi * 32 + bit, y, x);
    }
}

// ---- point-feature network -----------------------------------------------------------------------------
#define VFE_WARPS 8
#define VFE_MAX_IN 20     // raw point features + 3 + 3 + 1
#define VFE_OPL 4         // outputs per lane: layer widths up to 128

struct VfeFeat {
    int n, stride, nfeat, cluster, centre, dist, in0, c0, c1;   // c1 = 0: one layer
    float vs[3], lo[3], off[3];
    const float *w0, *b0, *w1, *b1;                             // BatchNorm already folded in
};

// One warp per point, lane = output channel (o = lane + 32 k): the first layer's weights live in registers,
// a point's output row is written / max-reduced as one contiguous segment.
// LAYER 0: out = max over the voxel's points of relu(W0 x + b0);
// LAYER 1: x1 = [relu(W0 x + b0), vmax0[voxel]], out = max of relu(W1 x1 + b1)   (W1 rows in shared memory)
template <int LAYER>
__global__ void __launch_bounds__(VFE_WARPS * 32)
k_vfe_pfn(VfeFeat F, const float *__restrict__ points, const int *__restrict__ point_voxel,
          const float *__restrict__ xyz_sum, const float *__restrict__ vmax0, float *__restrict__ out) {
    extern __shared__ float sm[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int in1 = 2 * F.c0, pitch1 = in1 + 1;
    float *sW1 = sm;                                              // [c1][in1 + 1]   (LAYER 1)
    float *sY = sW1 + (LAYER ? F.c1 * pitch1 : 0) + warp * in1;   // [warps][in1]    (LAYER 1)
    if (LAYER) {
        for (int i = threadIdx.x; i < F.c1 * in1; i += blockDim.x) sW1[(i / in1) * pitch1 + i % in1] = __ldg(F.w1 + i);
        __syncthreads();
    }
    float w[VFE_OPL][VFE_MAX_IN], bias[VFE_OPL];
#pragma unroll
    for (int k = 0; k < VFE_OPL; ++k) {
        const int o = lane + 32 * k;
        bias[k] = o < F.c0 ? __ldg(F.b0 + o) : 0.f;
#pragma unroll
        for (int i = 0; i < VFE_MAX_IN; ++i) w[k][i] = (o < F.c0 && i < F.in0) ? __ldg(F.w0 + o * F.in0 + i) : 0.f;
    }
    const int warps = gridDim.x * VFE_WARPS;
    for (int p = blockIdx.x * VFE_WARPS + warp; p < F.n; p += warps) {
        const int v = __ldg(point_voxel + p);
        if (v < 0) continue;
        // input features of the point (dynamic_vfe.py:95-107); every lane builds the same small vector
        const float *pt = points + (size_t)p * F.stride;
        float x[VFE_MAX_IN];
#pragma unroll
        for (int i = 0; i < VFE_MAX_IN; ++i) x[i] = 0.f;
        int at = 0;
        for (int i = 0; i < F.nfeat; ++i) x[at++] = __ldg(pt + 1 + i);
        const float px = x[0], py = x[1], pz = x[2];
        float cnt = 0.f;
        if (xyz_sum) {
            const float4 s = __ldg((const float4 *)(xyz_sum + 4 * (size_t)v));
            cnt = s.w;
            if (F.cluster) {
                const float c = fmaxf(s.w, 1.0f);
                x[at++] = __fsub_rn(px, __fdiv_rn(s.x, c)); x[at++] = __fsub_rn(py, __fdiv_rn(s.y, c));
                x[at++] = __fsub_rn(pz, __fdiv_rn(s.z, c));
            }
        }
        if (F.centre) {
            const float q[3] = {px, py, pz};
#pragma unroll
            for (int d = 0; d < 3; ++d) {
                const float c = floorf(__fdiv_rn(__fsub_rn(q[d], F.lo[d]), F.vs[d]));
                x[at++] = __fsub_rn(q[d], __fadd_rn(__fmul_rn(c, F.vs[d]), F.off[d]));
            }
        }
        if (F.dist) x[at++] = __fsqrt_rn(px * px + py * py + pz * pz);
        const bool alone = cnt == 1.0f;   // a voxel with a single point needs no atomic
        float y[VFE_OPL];
#pragma unroll
        for (int k = 0; k < VFE_OPL; ++k) {
            float a = bias[k];
#pragma unroll
            for (int i = 0; i < VFE_MAX_IN; ++i) a = fmaf(w[k][i], x[i], a);
            y[k] = fmaxf(a, 0.f);
        }
        if (!LAYER) {
#pragma unroll
            for (int k = 0; k < VFE_OPL; ++k) {
                const int o = lane + 32 * k;
                if (o < F.c0) {
                    if (alone) out[(size_t)v * F.c0 + o] = y[k];
                    else atomicMax((int *)out + (size_t)v * F.c0 + o, __float_as_int(y[k]));   // y >= 0: int order = float order
                }
            }
        } else {
            __syncwarp();
#pragma unroll
            for (int k = 0; k < VFE_OPL; ++k) {
                const int o = lane + 32 * k;
                if (o < F.c0) { sY[o] = y[k]; sY[F.c0 + o] = __ldg(vmax0 + (size_t)v * F.c0 + o); }
            }
            __syncwarp();
            for (int o = lane; o < F.c1; o += 32) {
                float a = __ldg(F.b1 + o);
                const float *wr = sW1 + o * pitch1;
                for (int i = 0; i < in1; ++i) a = fmaf(wr[i], sY[i], a);
                a = fmaxf(a, 0.f);
                if (alone) out[(size_t)v * F.c1 + o] = a;
                else atomicMax((int *)out + (size_t)v * F.c1 + o, __float_as_int(a));
            }
        }
    }
}

// First layer, thread = point: the point's features are built once, its outputs are computed 32 at a time
// against weights in shared memory (rows padded to a multiple of 4 inputs: 16-byte broadcast loads), and a
// chunk of 32 outputs = 128 bytes of the voxel's row leaves through the warp-cooperative row store
// (tc_common.cuh) when the voxel holds a single point, through atomicMax otherwise.
__global__ void __launch_bounds__(VFE_WARPS * 32)
k_vfe_pfn0(VfeFeat F, const float *__restrict__ points, const int *__restrict__ point_voxel,
           const float *__restrict__ xyz_sum, float *__restrict__ out) {
    extern __shared__ __align__(16) float sm[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int pitch = (F.in0 + 3) & ~3;                       // <= VFE_MAX_IN
    float *sW = sm;                                           // [c0][pitch], zero padded
    float *sB = sW + F.c0 * pitch;                            // [c0]
    char *stg = (char *)(sB + ((F.c0 + 3) & ~3)) + warp * 4096;
    for (int i = threadIdx.x; i < F.c0 * pitch; i += blockDim.x) {
        const int o = i / pitch, k = i - o * pitch;
        sW[i] = k < F.in0 ? __ldg(F.w0 + o * F.in0 + k) : 0.f;
    }
    for (int i = threadIdx.x; i < F.c0; i += blockDim.x) sB[i] = __ldg(F.b0 + i);
    __syncthreads();
    const int stride_pts = gridDim.x * blockDim.x;
    for (int p0 = blockIdx.x * blockDim.x + warp * 32; p0 < F.n; p0 += stride_pts) {   // (warp-uniform loop)
        const int p = p0 + lane;
        const int v = p < F.n ? __ldg(point_voxel + p) : -1;
        float x[VFE_MAX_IN];
#pragma unroll
        for (int i = 0; i < VFE_MAX_IN; ++i) x[i] = 0.f;
        bool alone = false;
        if (v >= 0) {   // input features of the point (dynamic_vfe.py:95-107)
            const float *pt = points + (size_t)p * F.stride;
            const float px = __ldg(pt + 1), py = __ldg(pt + 2), pz = __ldg(pt + 3);
            int at = 0;
#pragma unroll
            for (int i = 0; i < VFE_MAX_IN - 7; ++i)
                if (i < F.nfeat) x[at++] = __ldg(pt + 1 + i);
            if (xyz_sum) {
                const float4 s = __ldg((const float4 *)(xyz_sum + 4 * (size_t)v));
                alone = s.w == 1.0f;
                if (F.cluster) {
                    const float c = fmaxf(s.w, 1.0f);
                    x[at++] = __fsub_rn(px, __fdiv_rn(s.x, c)); x[at++] = __fsub_rn(py, __fdiv_rn(s.y, c));
                    x[at++] = __fsub_rn(pz, __fdiv_rn(s.z, c));
                }
            }
            if (F.centre) {
                const float q[3] = {px, py, pz};
#pragma unroll
                for (int d = 0; d < 3; ++d) {
                    const float c = floorf(__fdiv_rn(__fsub_rn(q[d], F.lo[d]), F.vs[d]));
                    x[at++] = __fsub_rn(q[d], __fadd_rn(__fmul_rn(c, F.vs[d]), F.off[d]));
                }
            }
            if (F.dist) x[at++] = __fsqrt_rn(px * px + py * py + pz * pz);
        }
        for (int c = 0; c < F.c0; c += 32) {   // 32 outputs = one 128-byte segment of the voxel's row
            float4 y[8];
#pragma unroll
            for (int j = 0; j < 32; ++j) {
                float a = 0.f;
                if (c + j < F.c0) {
                    a = sB[c + j];
                    const float4 *wr = (const float4 *)(sW + (c + j) * pitch);
#pragma unroll
                    for (int i4 = 0; i4 < VFE_MAX_IN / 4; ++i4)
                        if (4 * i4 < pitch) {
                            const float4 wv = wr[i4];
                            a = fmaf(wv.x, x[4 * i4], a); a = fmaf(wv.y, x[4 * i4 + 1], a);
                            a = fmaf(wv.z, x[4 * i4 + 2], a); a = fmaf(wv.w, x[4 * i4 + 3], a);
                        }
                    a = fmaxf(a, 0.f);
                }
                ((float *)y)[j] = a;
            }
            const bool full = c + 32 <= F.c0;
            float *row = v >= 0 ? out + (size_t)v * F.c0 + c : nullptr;
            warp_rows_store(stg, (full && alone) ? (float4 *)row : nullptr, y);
            if (row && !(full && alone)) {
#pragma unroll
                for (int j = 0; j < 32; ++j)
                    if (c + j < F.c0) atomicMax((int *)row + j, __float_as_int(((float *)y)[j]));   // y >= 0
            }
        }
    }
}

}  // namespace mssvt

using namespace mssvt;

extern "C" {

int mssvt_exclusive_scan(int n_cap, const int *n_dev, const int *src, int stride, int *dst, int *workspace,
                         void *stream);

/* words of the occupancy bitmap of mssvt_vfe_voxelize: batch * gx * gy * ceil(gz / 32) */
long long mssvt_vfe_bitmap_words(int batch_size, int gx, int gy, int gz) {
    return (long long)batch_size * gx * gy * ((gz + 31) / 32);
}

/* Dynamic voxelisation (dynamic_vfe.py:85-93, 111-116).  points (P, stride) fp32 rows [batch, x, y, z, ...].
 * bitmap (words) / counts (words) / base (words + 1) / scan_workspace ((words + 1) / 1024 + 2) int scratch.
 * Outputs: point_voxel (P) voxel row of every point or -1 (outside the range); voxel_coords (P, 4) [b, z, y, x]
 * rows [0, num_voxels) in ascending (b, x, y, z) key order -- the order of torch.unique in the reference;
 * num_voxels = base[words] stays on the device; xyz_sum (P, 4) optional: per-voxel (sum x, sum y, sum z, count). */
int mssvt_vfe_voxelize(int num_points, const float *points, int point_stride, int batch_size, int gx, int gy,
                       int gz, const float *voxel_size, const float *range_min, int *bitmap, int *counts,
                       int *base, int *scan_workspace, int *point_voxel, int *voxel_coords, float *xyz_sum,
                       void *stream) {
    if (num_points < 0 || point_stride < 4 || batch_size <= 0 || gx <= 0 || gy <= 0 || gz <= 0) return MSSVT_ERR_INVALID;
    if (!voxel_size || !range_min || !bitmap || !counts || !base || !scan_workspace) return MSSVT_ERR_INVALID;
    if (num_points > 0 && (!points || !point_voxel || !voxel_coords)) return MSSVT_ERR_INVALID;
    const long long words = mssvt_vfe_bitmap_words(batch_size, gx, gy, gz);
    if (words > 0x7fffffffLL) return MSSVT_ERR_INVALID;
    cudaStream_t s = (cudaStream_t)stream;
    VfeGrid G = {gx, gy, gz, (gz + 31) / 32, batch_size, {voxel_size[0], voxel_size[1], voxel_size[2]},
                 {range_min[0], range_min[1], range_min[2]}};
    if (cudaMemsetAsync(bitmap, 0, (size_t)words * 4, s) != cudaSuccess) return MSSVT_ERR_LAUNCH;
    if (xyz_sum && cudaMemsetAsync(xyz_sum, 0, (size_t)num_points * 16, s) != cudaSuccess) return MSSVT_ERR_LAUNCH;
    const int nb = (num_points + 255) / 256, wb = (int)((words + 255) / 256);
    if (num_points > 0) { ++g_launches; k_vfe_bits<<<nb, 256, 0, s>>>(G, num_points, points, point_stride, (unsigned *)bitmap); }
    ++g_launches;
    k_vfe_popc<<<wb, 256, 0, s>>>(words, (const unsigned *)bitmap, counts);
    int rc = mssvt_exclusive_scan((int)words, nullptr, counts, 1, base, scan_workspace, stream);
    if (rc != MSSVT_OK) return rc;
    if (num_points > 0) {
        ++g_launches;
        k_vfe_assign<<<nb, 256, 0, s>>>(G, num_points, points, point_stride, (const unsigned *)bitmap, base, point_voxel,
                                        xyz_sum);
        ++g_launches;
        k_vfe_coords<<<wb, 256, 0, s>>>(G, words, (const unsigned *)bitmap, base, (int4 *)voxel_coords);
    }
    return check_launch();
}

/* Point-feature network + per-voxel max (dynamic_vfe.py:95-108, 124-130), eval mode: the caller folds each
 * BatchNorm1d into its Linear (w' = w * g / sqrt(var + eps), b' = (b - mean) * g / sqrt(var + eps) + beta).
 * One or two layers: w0 (c0, in0), b0 (c0); w1 (c1, 2 * c0), b1 (c1) or c1 = 0.  in0 = num_point_features
 * + 3 (cluster centre) + 3 (voxel centre) + 1 (distance) as enabled.  centre_offset = voxel_size / 2 + range_min.
 * scratch: voxel_capacity * c0 floats (two layers only).  out (voxel_capacity, c_last), rows >= num_voxels zero. */
int mssvt_vfe_features(int num_points, const float *points, int point_stride, int num_point_features,
                       int with_cluster_center, int with_voxel_center, int with_distance, const float *voxel_size,
                       const float *range_min, const float *centre_offset, const int *point_voxel,
                       const float *xyz_sum, int voxel_capacity, const float *w0, const float *b0, int c0,
                       const float *w1, const float *b1, int c1, float *scratch, float *out, void *stream) {
    const int in0 = num_point_features + (with_cluster_center ? 3 : 0) + (with_voxel_center ? 3 : 0) + (with_distance ? 1 : 0);
    if (num_points < 0 || num_point_features < 3 || point_stride < 1 + num_point_features || in0 > VFE_MAX_IN || c0 <= 0 ||
        c1 < 0 || voxel_capacity < 0)
        return MSSVT_ERR_INVALID;
    if (!voxel_size || !range_min || !centre_offset || !w0 || !b0 || !out || (c1 && (!w1 || !b1 || !scratch)) ||
        (with_cluster_center && !xyz_sum))
        return MSSVT_ERR_INVALID;
    cudaStream_t s = (cudaStream_t)stream;
    const int c_last = c1 ? c1 : c0;
    if (cudaMemsetAsync(out, 0, (size_t)voxel_capacity * c_last * 4, s) != cudaSuccess) return MSSVT_ERR_LAUNCH;
    if (num_points == 0 || voxel_capacity == 0) return MSSVT_OK;
    if (!points || !point_voxel) return MSSVT_ERR_INVALID;
    VfeFeat F = {num_points, point_stride, num_point_features, with_cluster_center ? 1 : 0, with_voxel_center ? 1 : 0,
                 with_distance ? 1 : 0, in0, c0, c1,
                 {voxel_size[0], voxel_size[1], voxel_size[2]}, {range_min[0], range_min[1], range_min[2]},
                 {centre_offset[0], centre_offset[1], centre_offset[2]}, w0, b0, w1, b1};
    if (c0 > 32 * VFE_OPL || c1 > 32 * VFE_OPL) return MSSVT_ERR_INVALID;
    const int grid = persistent_grid(num_points, VFE_WARPS, 8);
    const int grid0 = persistent_grid(num_points, VFE_WARPS * 32, 8);
    const size_t sm0 = (size_t)(c0 * ((in0 + 3) & ~3) + ((c0 + 3) & ~3)) * 4 + VFE_WARPS * 4096;
    if (sm0 > 200 * 1024 || (c0 & 3)) return MSSVT_ERR_INVALID;
    cudaFuncSetAttribute(k_vfe_pfn0, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm0);
    if (!c1) {
        ++g_launches;
        k_vfe_pfn0<<<grid0, VFE_WARPS * 32, sm0, s>>>(F, points, point_voxel, xyz_sum, out);
        return check_launch();
    }
    // two layers: pass 1 = per-voxel max of layer 0 (scratch), pass 2 recomputes layer 0 and applies layer 1
    if (cudaMemsetAsync(scratch, 0, (size_t)voxel_capacity * c0 * 4, s) != cudaSuccess) return MSSVT_ERR_LAUNCH;
    VfeFeat F0 = F;
    F0.c1 = 0;
    ++g_launches;
    k_vfe_pfn0<<<grid0, VFE_WARPS * 32, sm0, s>>>(F0, points, point_voxel, xyz_sum, scratch);
    const size_t sm1 = (size_t)(c1 * (2 * c0 + 1) + VFE_WARPS * 2 * c0) * 4;
    if (sm1 > 200 * 1024) return MSSVT_ERR_INVALID;
    cudaFuncSetAttribute(k_vfe_pfn<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm1);
    ++g_launches;
    k_vfe_pfn<1><<<grid, VFE_WARPS * 32, sm1, s>>>(F, points, point_voxel, xyz_sum, scratch, out);
    return check_launch();
}

}  // extern "C"
