/*
 * mssvt_oracle.c -- CPU restatement of the native kernels on the MsSVT backbone hot path.
 *
 * THIS IS TEST INFRASTRUCTURE, NOT PRODUCT.  Only tests/, __graft_entry__.smoke() and the
 * cpu_baseline / --impl reference legs of bench.py may load it.  The product path
 * (mssvt_b200/) never links, imports or calls anything in oracle/.
 *
 * Every function restates one CUDA kernel of the reference with the kernel's threads executed
 * sequentially in thread-index order (one legal interleaving; it fixes the hash slot layout and
 * the window numbering).  Citations are reference paths relative to
 * /root/reference/pcdet/ops/.
 *
 * Pinning status: the reference ships no tests or golden vectors.  These functions are pinned
 *   (1) on the GPU box against the reference's own kernels compiled unmodified for sm_100a
 *       (oracle/_ref/libmssvt_ref.so, tests/test_gpu_ref_kernels.py), and
 *   (2) here, as stand-ins for the CUDA extension underneath the reference's unmodified Python
 *       layer (oracle/pin_against_reference.py -> tests/golden/).
 *
 * Build: make -C oracle   (gcc -O2 -ffp-contract=off -fopenmp -shared -fPIC)
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define EMPTY_KEY (-1) /* mssvt/src/ms_cuda_utils.h:9 */

/* ------------------------------------------------------------------ hash table primitives */

/* mssvt/src/ms_sparse_attention_gpu.cu:18-41 -- h(k) = k % H, linear probing, probe cap H.
 * A key already present gets its value overwritten (last writer in thread order wins). */
static void table_insert(int key, int value, int hash_size, int *table) {
    int slot = key % hash_size;
    for (int probes = 0; probes < hash_size; ++probes) {
        int seen = table[2 * slot];
        if (seen == EMPTY_KEY) table[2 * slot] = key;
        if (seen == EMPTY_KEY || seen == key) {
            table[2 * slot + 1] = value;
            return;
        }
        slot = (slot + 1) % hash_size;
    }
}

/* mssvt/src/ms_sparse_attention_gpu.cu:43-64 */
static int table_find(int key, int hash_size, const int *table) {
    int slot = key % hash_size;
    for (int probes = 0; probes < hash_size; ++probes) {
        int seen = table[2 * slot];
        if (seen == key) return table[2 * slot + 1];
        if (seen == EMPTY_KEY) return EMPTY_KEY;
        slot = (slot + 1) % hash_size;
    }
    return EMPTY_KEY;
}

/* mssvt/src/ms_sparse_attention_gpu.cu:66-97.  table is (B, H, 2), pre-filled with -1 by the
 * caller exactly as mssvt_ops.py:16-17 does. */
void orc_build_hash_table(int x_max, int y_max, int z_max, int num_voxels, int hash_size,
                          const int *v_indices, const int *v_bs_cnt, int *table) {
    for (int t = 0; t < num_voxels; ++t) {
        int b = v_indices[4 * t + 0], z = v_indices[4 * t + 1];
        int y = v_indices[4 * t + 2], x = v_indices[4 * t + 3];
        int before = 0;
        for (int s = b - 1; s >= 0; --s) before += v_bs_cnt[s];
        int local = t - before;
        if (x >= x_max || x < 0 || y < 0 || y >= y_max || z < 0 || z >= z_max) continue;
        int key = x * y_max * z_max + y * z_max + z;
        table_insert(key, local, hash_size, table + (size_t)b * hash_size * 2);
    }
}

/* Plain lookups (no reference kernel; used by tests to compare tables by content). */
void orc_hash_lookup(int hash_size, int num_queries, const int *batch_ids, const int *keys,
                     const int *table, int *values) {
    for (int i = 0; i < num_queries; ++i)
        values[i] = table_find(keys[i], hash_size, table + (size_t)batch_ids[i] * hash_size * 2);
}

/* mssvt/src/ms_sparse_attention_gpu.cu:117-168.  w_indices is (B, max_wins, 3) pre-filled -1,
 * vcount (B) pre-filled 0, table pre-filled -1 (mssvt_ops.py:36-41).  Sequential execution =>
 * windows numbered by first occurrence in voxel order.  Unlike the reference this refuses to
 * write past max_wins: it returns the number of windows that did not fit (0 = ok). */
int orc_window_partition(int x_wgs, int y_wgs, int z_wgs, int x_ws, int y_ws, int z_ws,
                         int num_voxels, int max_wins, int hash_size, const int *v_indices,
                         int *w_indices, int *table, int *vcount) {
    int overflow = 0;
    for (int t = 0; t < num_voxels; ++t) {
        int b = v_indices[4 * t + 0];
        int wz = v_indices[4 * t + 1] / z_ws;
        int wy = v_indices[4 * t + 2] / y_ws;
        int wx = v_indices[4 * t + 3] / x_ws;
        if (wx < 0 || wx >= x_wgs || wy < 0 || wy >= y_wgs || wz < 0 || wz >= z_wgs) continue;
        int *tab = table + (size_t)b * hash_size * 2;
        int *rows = w_indices + (size_t)b * max_wins * 3;
        int key = wx * y_wgs * z_wgs + wy * z_wgs + wz;
        int slot = key % hash_size;
        for (int probes = 0; probes < hash_size; ++probes) {
            int seen = tab[2 * slot];
            if (seen == EMPTY_KEY) {
                int row = vcount[b]++;
                if (row >= max_wins) { overflow++; break; }
                tab[2 * slot] = key;
                rows[3 * row + 0] = wz;
                rows[3 * row + 1] = wy;
                rows[3 * row + 2] = wx;
                tab[2 * slot + 1] = row;
                break;
            }
            if (seen == key) break;
            slot = (slot + 1) % hash_size;
        }
    }
    return overflow;
}

/* ------------------------------------------------------------------ chessboard gather */

typedef struct {
    int *ind;   /* (W, cap)    */
    int *coord; /* (W, cap, 3) */
    int cap, count;
} list_t;

static void list_push(list_t *l, size_t w, int v, int ox, int oy, int oz) {
    if (l->count >= l->cap) return;
    size_t at = w * l->cap + l->count;
    l->ind[at] = v;
    l->coord[3 * at + 0] = ox;
    l->coord[3 * at + 1] = oy;
    l->coord[3 * at + 2] = oz;
    l->count++;
}

/* mssvt/src/ms_sparse_attention_gpu.cu:193-350.  Offset tables are walked in the order
 * odd, even, win1-rest, win2-rest; a hit is appended to every list that "owns" the table:
 *   odd  -> {odd, win1, win2}   even -> {even, win1, win2}
 *   win1 -> {win1, win2}        win2 -> {win2}
 * each list capped; the reference's early `return`s only fire when every remaining list is
 * full, so they do not change the result.  Outputs pre-filled (-1 / 0) by the caller. */
void orc_gather_two_window(int x_max, int y_max, int z_max, int x_ws, int y_ws, int z_ws,
                           int max_odd, int max_even, int max_win1, int max_win2, int num_wins,
                           int hash_size, int num_odd, int num_even, int num_win1, int num_win2,
                           int *ind_odd, int *ind_even, int *ind_win1, int *ind_win2,
                           int *coord_odd, int *coord_even, int *coord_win1, int *coord_win2,
                           const int *q_odd, const int *q_even, const int *q_win1,
                           const int *q_win2, const int *win_indices, const int *table) {
    const int *tables[4] = {q_odd, q_even, q_win1, q_win2};
    const int sizes[4] = {num_odd, num_even, num_win1, num_win2};
#pragma omp parallel for schedule(static)
    for (int w = 0; w < num_wins; ++w) {
        int b = win_indices[4 * w + 0];
        int cz = win_indices[4 * w + 1] * z_ws + z_ws / 2;
        int cy = win_indices[4 * w + 2] * y_ws + y_ws / 2;
        int cx = win_indices[4 * w + 3] * x_ws + x_ws / 2;
        const int *tab = table + (size_t)b * hash_size * 2;
        list_t odd = {ind_odd, coord_odd, max_odd, 0}, even = {ind_even, coord_even, max_even, 0};
        list_t win1 = {ind_win1, coord_win1, max_win1, 0}, win2 = {ind_win2, coord_win2, max_win2, 0};
        for (int which = 0; which < 4; ++which) {
            for (int q = 0; q < sizes[which]; ++q) {
                int ox = tables[which][3 * q + 0], oy = tables[which][3 * q + 1];
                int oz = tables[which][3 * q + 2];
                int sx = cx + ox, sy = cy + oy, sz = cz + oz;
                if (sx >= x_max || sx < 0 || sy >= y_max || sy < 0 || sz >= z_max || sz < 0) continue;
                int v = table_find(sx * y_max * z_max + sy * z_max + sz, hash_size, tab);
                if (v == EMPTY_KEY) continue;
                if (which == 0) list_push(&odd, w, v, ox, oy, oz);
                if (which == 1) list_push(&even, w, v, ox, oy, oz);
                if (which <= 2) list_push(&win1, w, v, ox, oy, oz);
                list_push(&win2, w, v, ox, oy, oz);
            }
        }
    }
}

/* mssvt/src/ms_sparse_attention_gpu.cu:383-433 */
void orc_gather_one_window(int x_max, int y_max, int z_max, int x_ws, int y_ws, int z_ws,
                           int max_win1, int num_wins, int hash_size, int num_win1, int *ind_win1,
                           int *coord_win1, const int *q_win1, const int *win_indices,
                           const int *table) {
#pragma omp parallel for schedule(static)
    for (int w = 0; w < num_wins; ++w) {
        int b = win_indices[4 * w + 0];
        int cz = win_indices[4 * w + 1] * z_ws + z_ws / 2;
        int cy = win_indices[4 * w + 2] * y_ws + y_ws / 2;
        int cx = win_indices[4 * w + 3] * x_ws + x_ws / 2;
        const int *tab = table + (size_t)b * hash_size * 2;
        list_t win1 = {ind_win1, coord_win1, max_win1, 0};
        for (int q = 0; q < num_win1; ++q) {
            int ox = q_win1[3 * q + 0], oy = q_win1[3 * q + 1], oz = q_win1[3 * q + 2];
            int sx = cx + ox, sy = cy + oy, sz = cz + oz;
            if (sx >= x_max || sx < 0 || sy >= y_max || sy < 0 || sz >= z_max || sz < 0) continue;
            int v = table_find(sx * y_max * z_max + sy * z_max + sz, hash_size, tab);
            if (v != EMPTY_KEY) list_push(&win1, w, v, ox, oy, oz);
        }
    }
}

/* ------------------------------------------------------------------ stacked feature grouping */

/* rows of idx belong to samples by idx_batch_cnt; group_features_gpu.cu:91-99 */
static int sample_of_row(int row, int B, const int *idx_batch_cnt) {
    int b = 0, upto = idx_batch_cnt[0];
    for (int k = 1; k < B; ++k) {
        if (row < upto) break;
        upto += idx_batch_cnt[k];
        b = k;
    }
    return b;
}

/* mssvt/src/group_features_gpu.cu:73-106.  out (M, C, ns) pre-zeroed by the caller. */
void orc_group_features(int B, int M, int C, int nsample, const float *features,
                        const int *features_batch_cnt, const int *idx, const int *idx_batch_cnt,
                        float *out) {
#pragma omp parallel for schedule(static)
    for (int m = 0; m < M; ++m) {
        int b = sample_of_row(m, B, idx_batch_cnt);
        size_t start = 0;
        for (int k = 0; k < b; ++k) start += features_batch_cnt[k];
        for (int s = 0; s < nsample; ++s) {
            int v = idx[(size_t)m * nsample + s];
            if (v < 0) continue;
            const float *src = features + (start + v) * C;
            for (int c = 0; c < C; ++c) out[((size_t)m * C + c) * nsample + s] = src[c];
        }
    }
}

/* mssvt/src/group_features_gpu.cu:15-47.  grad_features (N, C) pre-zeroed.  The reference
 * accumulates with float atomicAdd in arbitrary order; here the order is (m, c, s) ascending. */
void orc_group_features_grad(int B, int M, int C, int N, int nsample, const float *grad_out,
                             const int *idx, const int *idx_batch_cnt,
                             const int *features_batch_cnt, float *grad_features) {
    (void)N;
    for (int m = 0; m < M; ++m) {
        int b = sample_of_row(m, B, idx_batch_cnt);
        size_t start = 0;
        for (int k = 0; k < b; ++k) start += features_batch_cnt[k];
        for (int c = 0; c < C; ++c)
            for (int s = 0; s < nsample; ++s) {
                int v = idx[(size_t)m * nsample + s];
                if (v < 0) continue;
                grad_features[(start + v) * C + c] += grad_out[((size_t)m * C + c) * nsample + s];
            }
    }
}

/* ------------------------------------------------------------------ pointnet2_batch ops */

/* pointnet2/pointnet2_batch/src/cuda_utils.h:10-14 */
static int fps_block_size(int n) {
    int p = (int)(log((double)n) / log(2.0));
    int t = 1 << p;
    if (t > 1024) t = 1024;
    return t < 1 ? 1 : t;
}

/* pointnet2/pointnet2_batch/src/sampling_gpu.cu:100-216, literal simulation of one thread
 * block per row: per-thread strided scan keeping the FIRST strictly larger candidate, then the
 * shared-memory tree reduction where a tie keeps the lower slot (:93-98).  temp (b, n) is
 * pre-filled with 1e10 by the caller (pointnet2_utils.py:26). */
void orc_fps(int b, int n, int m, const float *dataset, float *temp, int *idxs) {
    if (m <= 0) return;
    int T = fps_block_size(n);
#pragma omp parallel
    {
        float *dist = (float *)malloc(sizeof(float) * T);
        int *disti = (int *)malloc(sizeof(int) * T);
#pragma omp for schedule(static)
        for (int row = 0; row < b; ++row) {
            const float *pts = dataset + (size_t)row * n * 3;
            float *tmp = temp + (size_t)row * n;
            int *out = idxs + (size_t)row * m;
            int old = 0;
            out[0] = 0;
            for (int j = 1; j < m; ++j) {
                float x1 = pts[old * 3 + 0], y1 = pts[old * 3 + 1], z1 = pts[old * 3 + 2];
                for (int tid = 0; tid < T; ++tid) {
                    int besti = 0;
                    float best = -1.0f;
                    for (int k = tid; k < n; k += T) {
                        float dx = pts[k * 3 + 0] - x1, dy = pts[k * 3 + 1] - y1;
                        float dz = pts[k * 3 + 2] - z1;
                        /* nvcc 12.9 -> sm_100a contracts the source expression to
                         * FMUL dy,dy; FFMA dx,dx,.; FFMA dz,dz,. (cuobjdump -sass of
                         * oracle/_ref, same shape as three_nn).  Exact anyway for the integer
                         * offsets the backbone feeds in. */
                        float d = fmaf(dz, dz, fmaf(dx, dx, dy * dy));
                        float d2 = d < tmp[k] ? d : tmp[k];
                        tmp[k] = d2;
                        if (d2 > best) { best = d2; besti = k; }
                    }
                    dist[tid] = best;
                    disti[tid] = besti;
                }
                for (int half = T / 2; half >= 1; half /= 2)
                    for (int tid = 0; tid < half; ++tid)
                        if (dist[tid + half] > dist[tid]) {
                            dist[tid] = dist[tid + half];
                            disti[tid] = disti[tid + half];
                        }
                old = disti[0];
                out[j] = old;
            }
        }
        free(dist);
        free(disti);
    }
}

/* pointnet2/pointnet2_batch/src/sampling_gpu.cu:15-31: out[b,c,j] = points[b,c,idx[b,j]] */
void orc_gather_points(int b, int c, int n, int m, const float *points, const int *idx,
                       float *out) {
    for (int bi = 0; bi < b; ++bi)
        for (int ci = 0; ci < c; ++ci)
            for (int j = 0; j < m; ++j)
                out[((size_t)bi * c + ci) * m + j] =
                    points[((size_t)bi * c + ci) * n + idx[(size_t)bi * m + j]];
}

/* pointnet2/pointnet2_batch/src/interpolate_gpu.cu:16-59.  The float distance is evaluated the
 * way nvcc 12.9 contracts `(ux-x)*(ux-x) + (uy-y)*(uy-y) + (uz-z)*(uz-z)` for sm_100a:
 * t = dy*dy; t = fma(dx,dx,t); d = fma(dz,dz,t) (re-checked by
 * oracle/build_ref.sh which dumps the SASS).  Bests are doubles initialised to 1e40 and
 * compared with strict <, so the lowest index wins ties. */
void orc_three_nn(int b, int n, int m, const float *unknown, const float *known, float *dist2,
                  int *idx) {
#pragma omp parallel for schedule(static)
    for (int row = 0; row < b; ++row) {
        const float *kn = known + (size_t)row * m * 3;
        for (int p = 0; p < n; ++p) {
            const float *u = unknown + ((size_t)row * n + p) * 3;
            double b1 = 1e40, b2 = 1e40, b3 = 1e40;
            int i1 = 0, i2 = 0, i3 = 0;
            for (int k = 0; k < m; ++k) {
                float dx = u[0] - kn[3 * k + 0], dy = u[1] - kn[3 * k + 1];
                float dz = u[2] - kn[3 * k + 2];
                float t = dy * dy;
                t = fmaf(dx, dx, t);
                float d = fmaf(dz, dz, t);
                if (d < b1) { b3 = b2; i3 = i2; b2 = b1; i2 = i1; b1 = d; i1 = k; }
                else if (d < b2) { b3 = b2; i3 = i2; b2 = d; i2 = k; }
                else if (d < b3) { b3 = d; i3 = k; }
            }
            float *d2 = dist2 + ((size_t)row * n + p) * 3;
            int *o = idx + ((size_t)row * n + p) * 3;
            d2[0] = (float)b1; d2[1] = (float)b2; d2[2] = (float)b3;
            o[0] = i1; o[1] = i2; o[2] = i3;
        }
    }
}

/* pointnet2/pointnet2_batch/src/group_points_gpu.cu:53-72:
 * out[b,c,p,s] = points[b,c,idx[b,p,s]] */
void orc_group_points(int b, int c, int n, int npoints, int nsample, const float *points,
                      const int *idx, float *out) {
#pragma omp parallel for schedule(static)
    for (int bi = 0; bi < b; ++bi)
        for (int ci = 0; ci < c; ++ci) {
            const float *src = points + ((size_t)bi * c + ci) * n;
            float *dst = out + ((size_t)bi * c + ci) * npoints * nsample;
            const int *ii = idx + (size_t)bi * npoints * nsample;
            for (int e = 0; e < npoints * nsample; ++e) dst[e] = src[ii[e]];
        }
}

/* pointnet2/pointnet2_batch/src/group_points_gpu.cu:14-31 (atomicAdd scatter; order here is
 * ascending) */
void orc_group_points_grad(int b, int c, int n, int npoints, int nsample, const float *grad_out,
                           const int *idx, float *grad_points) {
    for (int bi = 0; bi < b; ++bi)
        for (int ci = 0; ci < c; ++ci) {
            float *dst = grad_points + ((size_t)bi * c + ci) * n;
            const float *src = grad_out + ((size_t)bi * c + ci) * npoints * nsample;
            const int *ii = idx + (size_t)bi * npoints * nsample;
            for (int e = 0; e < npoints * nsample; ++e) dst[ii[e]] += src[e];
        }
}

void orc_set_num_threads(int n) {
#ifdef _OPENMP
    extern void omp_set_num_threads(int);
    if (n > 0) omp_set_num_threads(n);
#else
    (void)n;
#endif
}

int orc_num_threads(void) {
#ifdef _OPENMP
    extern int omp_get_max_threads(void);
    return omp_get_max_threads();
#else
    return 1;
#endif
}
