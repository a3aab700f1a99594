// vfe.cu -- dynamic voxelisation + point-feature encoder: the producer of the backbone's inputs
// (pcdet/models/backbones_3d/vfe/dynamic_vfe.py:71-130, eval mode).
//
//   reference                                           here
//   torch.unique(merge_coords, sorted, inverse) :93     occupancy bitmap over the voxel grid in key order
//                                                       (b, x, y, z) + popcount scan: the rank of a set bit IS
//                                                       the row of its voxel in the sorted unique list; no sort
//   torch_scatter.scatter_mean(xyz) :98                 atomicAdd of (x, y, z, 1) per point
//   PFN: Linear + BatchNorm1d + ReLU :124-130           BatchNorm (eval) folded into the linear layer by the
//                                                       caller; weights in shared memory, 32 points per CTA pass
//   torch_scatter.scatter_max :108, :128                atomicMax on the int view (values are >= 0 after ReLU)
#include "common.cuh"

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
#define VFE_PTS 32        // points per CTA pass
#define VFE_THREADS 128   // thread = (point, quarter of the outputs)
#define VFE_MAX_IN 20     // raw point features + 3 + 3 + 1

struct VfeFeat {
    int n, stride, nfeat, cluster, centre, dist, in0, c0, c1;   // c1 = 0: one layer
    float vs[3], lo[3], off[3];
    const float *w0, *b0, *w1, *b1;                             // BatchNorm already folded in
};

// LAYER 0: out = max over the voxel's points of relu(W0 x + b0);  LAYER 1: x1 = [relu(W0 x + b0), vmax0[voxel]],
// out = max of relu(W1 x1 + b1)
template <int LAYER>
__global__ void __launch_bounds__(VFE_THREADS)
k_vfe_pfn(VfeFeat F, const float *__restrict__ points, const int *__restrict__ point_voxel,
          const float *__restrict__ xyz_sum, const float *__restrict__ vmax0, float *__restrict__ out) {
    extern __shared__ float sm[];
    const int in1 = 2 * F.c0;
    float *sW0 = sm;                                     // [c0][in0]
    float *sB0 = sW0 + F.c0 * F.in0;                     // [c0]
    float *sW1 = sB0 + F.c0;                             // [c1][in1]        (LAYER 1)
    float *sB1 = sW1 + (LAYER ? F.c1 * in1 : 0);         // [c1]
    float *sX = sB1 + (LAYER ? F.c1 : 0);                // [PTS][in0 + 1]
    float *sY = sX + VFE_PTS * (F.in0 + 1);              // [PTS][in1 + 1]   (LAYER 1)
    for (int i = threadIdx.x; i < F.c0 * F.in0; i += VFE_THREADS) sW0[i] = __ldg(F.w0 + i);
    for (int i = threadIdx.x; i < F.c0; i += VFE_THREADS) sB0[i] = __ldg(F.b0 + i);
    if (LAYER) {
        for (int i = threadIdx.x; i < F.c1 * in1; i += VFE_THREADS) sW1[i] = __ldg(F.w1 + i);
        for (int i = threadIdx.x; i < F.c1; i += VFE_THREADS) sB1[i] = __ldg(F.b1 + i);
    }
    const int pl = threadIdx.x & (VFE_PTS - 1), part = threadIdx.x / VFE_PTS;   // 4 parts
    for (int p0 = blockIdx.x * VFE_PTS; p0 < F.n; p0 += gridDim.x * VFE_PTS) {
        __syncthreads();
        const int p = p0 + pl;
        const int v = p < F.n ? __ldg(point_voxel + p) : -1;
        if (part == 0 && v >= 0) {   // input features of the point (dynamic_vfe.py:95-107)
            const float *pt = points + (size_t)p * F.stride;
            float *x = sX + pl * (F.in0 + 1);
            int at = 0;
            for (int i = 0; i < F.nfeat; ++i) x[at++] = pt[1 + i];
            if (F.cluster) {
                const float4 s = *(const float4 *)(xyz_sum + 4 * (size_t)v);
                const float cnt = fmaxf(s.w, 1.0f);
                x[at++] = __fsub_rn(pt[1], __fdiv_rn(s.x, cnt));
                x[at++] = __fsub_rn(pt[2], __fdiv_rn(s.y, cnt));
                x[at++] = __fsub_rn(pt[3], __fdiv_rn(s.z, cnt));
            }
            if (F.centre) {
#pragma unroll
                for (int d = 0; d < 3; ++d) {
                    const float c = floorf(__fdiv_rn(__fsub_rn(pt[1 + d], F.lo[d]), F.vs[d]));
                    x[at++] = __fsub_rn(pt[1 + d], __fadd_rn(__fmul_rn(c, F.vs[d]), F.off[d]));
                }
            }
            if (F.dist) x[at++] = __fsqrt_rn(pt[1] * pt[1] + pt[2] * pt[2] + pt[3] * pt[3]);
        }
        __syncthreads();
        if (v >= 0) {   // layer 0: this thread's quarter of the c0 outputs
            const float *x = sX + pl * (F.in0 + 1);
            for (int o = part; o < F.c0; o += VFE_THREADS / VFE_PTS) {
                float a = sB0[o];
                for (int i = 0; i < F.in0; ++i) a = fmaf(sW0[o * F.in0 + i], x[i], a);
                a = fmaxf(a, 0.f);
                if (LAYER) {
                    sY[pl * (in1 + 1) + o] = a;
                    sY[pl * (in1 + 1) + F.c0 + o] = __ldg(vmax0 + (size_t)v * F.c0 + o);
                } else {
                    atomicMax((int *)out + (size_t)v * F.c0 + o, __float_as_int(a));   // a >= 0: int order = float order
                }
            }
        }
        if (LAYER) {
            __syncthreads();
            if (v >= 0) {
                const float *y = sY + pl * (in1 + 1);
                for (int o = part; o < F.c1; o += VFE_THREADS / VFE_PTS) {
                    float a = sB1[o];
                    for (int i = 0; i < in1; ++i) a = fmaf(sW1[o * in1 + i], y[i], a);
                    atomicMax((int *)out + (size_t)v * F.c1 + o, __float_as_int(fmaxf(a, 0.f)));
                }
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
    const int grid = persistent_grid(num_points, VFE_PTS, 8);
    const size_t sm0 = (size_t)(c0 * in0 + c0 + VFE_PTS * (in0 + 1)) * 4;
    if (!c1) {
        if (sm0 > 200 * 1024) return MSSVT_ERR_INVALID;
        cudaFuncSetAttribute(k_vfe_pfn<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm0);
        ++g_launches;
        k_vfe_pfn<0><<<grid, VFE_THREADS, sm0, s>>>(F, points, point_voxel, xyz_sum, nullptr, out);
        return check_launch();
    }
    // two layers: pass 1 = per-voxel max of layer 0 (scratch), pass 2 recomputes layer 0 and applies layer 1
    if (cudaMemsetAsync(scratch, 0, (size_t)voxel_capacity * c0 * 4, s) != cudaSuccess) return MSSVT_ERR_LAUNCH;
    VfeFeat F0 = F;
    F0.c1 = 0;
    if (sm0 > 200 * 1024) return MSSVT_ERR_INVALID;
    cudaFuncSetAttribute(k_vfe_pfn<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm0);
    ++g_launches;
    k_vfe_pfn<0><<<grid, VFE_THREADS, sm0, s>>>(F0, points, point_voxel, xyz_sum, nullptr, scratch);
    const size_t sm1 = sm0 + (size_t)(c1 * 2 * c0 + c1 + VFE_PTS * (2 * c0 + 1)) * 4;
    if (sm1 > 200 * 1024) return MSSVT_ERR_INVALID;
    cudaFuncSetAttribute(k_vfe_pfn<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm1);
    ++g_launches;
    k_vfe_pfn<1><<<grid, VFE_THREADS, sm1, s>>>(F, points, point_voxel, xyz_sum, scratch, out);
    return check_launch();
}

}  // extern "C"
