// train_lists.cu -- the compact (CSR) window lists of the training path, built on the device from the maps of
// mssvt_block_geometry / mssvt_window_rows (what mssvt_b200/mssvt_backbone.py::_window_lists needs): per-window key counts,
// the row lists of the real queries and of every head group's distinct keys, the three-NN map of every voxel in compact
// query ids.  The list LENGTHS come from exclusive scans over the counts (mssvt_exclusive_scan); the host reads them once
// to size the lists, then the fill kernels write them.
#include "common.cuh"

namespace mssvt {

// cnt (cap, 4) = {#real queries, #distinct keys of group 0, of group 1, 0} with the key counts zeroed for windows without a
// query (nobody reads their keys); mult (2, cap) = multiplicity of the masked key of each group (0: none)
__global__ void __launch_bounds__(256)
k_lists_count(int cap, const int *__restrict__ total, const int4 *__restrict__ meta, int4 *__restrict__ cnt,
              int *__restrict__ mult) {
    const int w = blockIdx.x * blockDim.x + threadIdx.x;
    if (w >= cap) return;
    int4 c = make_int4(0, 0, 0, 0);
    int m0 = 0, m1 = 0;
    if (w < min(cap, __ldg(total))) {
        const int4 m = __ldg(meta + w);
        c.x = m.x;
        if (m.x > 0) { c.y = m.z & 0xff; c.z = m.w & 0xff; }
        m0 = m.z >> 8; m1 = m.w >> 8;
    }
    cnt[w] = c;
    mult[w] = m0;
    mult[cap + w] = m1;
}

struct ListsOut {
    int *q_rows, *q_win;          // (#queries)
    int *k_rows[2], *k_win[2];    // (#keys of the group)
    unsigned char *k_masked[2];
};

// one warp per window: its real queries and its distinct keys go to their places in the compact lists
__global__ void __launch_bounds__(256)
k_lists_fill(int cap, const int *__restrict__ total, int nq, int K, const int4 *__restrict__ cnt,
             const int *__restrict__ mult, const int *__restrict__ q_off, const int *__restrict__ key_off0,
             const int *__restrict__ key_off1, const int *__restrict__ q_row, const int *__restrict__ rep_row,
             ListsOut out) {
    const int lane = threadIdx.x & 31;
    const int num = min(cap, __ldg(total));
    for (int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; w < num; w += (gridDim.x * blockDim.x) >> 5) {
        const int4 c = __ldg(cnt + w);
        const int q0 = __ldg(q_off + w);
        for (int i = lane; i < c.x; i += 32) {
            out.q_rows[q0 + i] = __ldg(q_row + (size_t)w * nq + i);
            out.q_win[q0 + i] = w;
        }
#pragma unroll
        for (int s = 0; s < 2; ++s) {
            const int n = s ? c.z : c.y, k0 = __ldg((s ? key_off1 : key_off0) + w), m = __ldg(mult + s * cap + w);
            for (int j = lane; j < n; j += 32) {
                out.k_rows[s][k0 + j] = __ldg(rep_row + (size_t)w * 2 * K + s * K + j);
                out.k_win[s][k0 + j] = w;
                out.k_masked[s][k0 + j] = (m > 0 && j == n - 1) ? 1 : 0;
            }
        }
    }
}

// three-NN map of every voxel in compact query ids: -2 = voxel outside every window (keeps x, quirk Q5), -1 = padded query slot
__global__ void __launch_bounds__(256)
k_lists_merge_map(int n_vox, int cap1, const int *__restrict__ vox_slot, const int4 *__restrict__ meta,
                  const int *__restrict__ q_off, const unsigned char *__restrict__ nn_idx,
                  const float *__restrict__ nn_w, int *__restrict__ src, float *__restrict__ wgt) {
    const int v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= n_vox) return;
    const int slot = __ldg(vox_slot + v);
    if (slot < 0) {
        src[3 * v] = src[3 * v + 1] = src[3 * v + 2] = -2;
        wgt[3 * v] = wgt[3 * v + 1] = wgt[3 * v + 2] = 0.f;
        return;
    }
    const int w = slot / cap1;
    const int nqr = __ldg(meta + w).x, q0 = __ldg(q_off + w);
#pragma unroll
    for (int j = 0; j < 3; ++j) {
        const int i = nn_idx[3 * (size_t)slot + j];
        src[3 * v + j] = i < nqr ? q0 + i : -1;
        wgt[3 * v + j] = __ldg(nn_w + 3 * (size_t)slot + j);
    }
}

// compress block: keys of a window = its voxels (k_row, -1 padded) + ONE pad key when slots are left (quirk Q6)
__global__ void __launch_bounds__(256)
k_clists_count(int cap, const int *__restrict__ total, int n1, const int *__restrict__ k_row, int *__restrict__ cnt,
               int *__restrict__ mult) {
    const int w = blockIdx.x * blockDim.x + threadIdx.x;
    if (w >= cap) return;
    int c = 0, m = 0;
    if (w < min(cap, __ldg(total))) {
        int real = 0;
        for (int i = 0; i < n1; ++i) real += __ldg(k_row + (size_t)w * n1 + i) >= 0 ? 1 : 0;
        m = n1 - real;
        c = real + (m > 0 ? 1 : 0);
    }
    cnt[w] = c;
    mult[w] = m;
}

__global__ void __launch_bounds__(256)
k_clists_fill(int cap, const int *__restrict__ total, int n1, const int *__restrict__ cnt, const int *__restrict__ mult,
              const int *__restrict__ key_off, const int *__restrict__ k_row, int *__restrict__ rows,
              int *__restrict__ k_win) {
    const int lane = threadIdx.x & 31;
    const int num = min(cap, __ldg(total));
    for (int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; w < num; w += (gridDim.x * blockDim.x) >> 5) {
        const int n = __ldg(cnt + w), k0 = __ldg(key_off + w), real = n - (__ldg(mult + w) > 0 ? 1 : 0);
        for (int j = lane; j < n; j += 32) {
            rows[k0 + j] = j < real ? __ldg(k_row + (size_t)w * n1 + j) : -1;      // the pad key: no voxel
            k_win[k0 + j] = w;
        }
    }
}

}  // namespace mssvt

using namespace mssvt;

extern "C" {

int mssvt_ragged_lists_count(int win_capacity, const int *win_count_total, const int *meta, int *counts, int *mult,
                             void *stream) {
    if (win_capacity < 0) return MSSVT_ERR_INVALID;
    if (win_capacity == 0) return MSSVT_OK;
    if (!win_count_total || !meta || !counts || !mult) return MSSVT_ERR_INVALID;
    k_lists_count<<<div_up(win_capacity, 256), 256, 0, (cudaStream_t)stream>>>(win_capacity, win_count_total,
                                                                             (const int4 *)meta, (int4 *)counts, mult);
    ++g_launches;
    return check_launch();
}

int mssvt_ragged_lists_fill(int win_capacity, const int *win_count_total, int nq, int K, const int *counts,
                            const int *mult, const int *q_off, const int *key_off0, const int *key_off1,
                            const int *q_row, const int *rep_row, int *q_rows, int *q_win, int *k_rows0, int *k_win0,
                            unsigned char *k_masked0, int *k_rows1, int *k_win1, unsigned char *k_masked1,
                            void *stream) {
    if (win_capacity < 0 || nq <= 0 || K <= 0) return MSSVT_ERR_INVALID;
    if (win_capacity == 0) return MSSVT_OK;
    if (!win_count_total || !counts || !mult || !q_off || !key_off0 || !key_off1 || !q_row || !rep_row)
        return MSSVT_ERR_INVALID;
    ListsOut out = {q_rows, q_win, {k_rows0, k_rows1}, {k_win0, k_win1}, {k_masked0, k_masked1}};
    k_lists_fill<<<persistent_grid(win_capacity, 8, 8), 256, 0, (cudaStream_t)stream>>>(
        win_capacity, win_count_total, nq, K, (const int4 *)counts, mult, q_off, key_off0, key_off1, q_row, rep_row, out);
    ++g_launches;
    return check_launch();
}

int mssvt_ragged_merge_map(int num_voxels, int max_win1, const int *vox_slot, const int *meta, const int *q_off,
                           const unsigned char *nn_idx, const float *nn_w, int *src, float *weights, void *stream) {
    if (num_voxels < 0 || max_win1 <= 0) return MSSVT_ERR_INVALID;
    if (num_voxels == 0) return MSSVT_OK;
    if (!vox_slot || !meta || !q_off || !nn_idx || !nn_w || !src || !weights) return MSSVT_ERR_INVALID;
    k_lists_merge_map<<<div_up(num_voxels, 256), 256, 0, (cudaStream_t)stream>>>(
        num_voxels, max_win1, vox_slot, (const int4 *)meta, q_off, nn_idx, nn_w, src, weights);
    ++g_launches;
    return check_launch();
}

int mssvt_compress_lists_count(int win_capacity, const int *win_count_total, int max_win1, const int *k_row,
                               int *counts, int *mult, void *stream) {
    if (win_capacity < 0 || max_win1 <= 0) return MSSVT_ERR_INVALID;
    if (win_capacity == 0) return MSSVT_OK;
    if (!win_count_total || !k_row || !counts || !mult) return MSSVT_ERR_INVALID;
    k_clists_count<<<div_up(win_capacity, 256), 256, 0, (cudaStream_t)stream>>>(win_capacity, win_count_total, max_win1,
                                                                              k_row, counts, mult);
    ++g_launches;
    return check_launch();
}

int mssvt_compress_lists_fill(int win_capacity, const int *win_count_total, int max_win1, const int *counts,
                              const int *mult, const int *key_off, const int *k_row, int *rows, int *k_win,
                              void *stream) {
    if (win_capacity < 0 || max_win1 <= 0) return MSSVT_ERR_INVALID;
    if (win_capacity == 0) return MSSVT_OK;
    if (!win_count_total || !counts || !mult || !key_off || !k_row) return MSSVT_ERR_INVALID;
    k_clists_fill<<<persistent_grid(win_capacity, 8, 8), 256, 0, (cudaStream_t)stream>>>(
        win_capacity, win_count_total, max_win1, counts, mult, key_off, k_row, rows, k_win);
    ++g_launches;
    return check_launch();
}

}  // extern "C"
