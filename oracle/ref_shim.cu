// ref_shim.cu -- extern "C" doors onto the reference's own kernel launchers (test infrastructure).
//
// oracle/Makefile (target `ref`) compiles the reference's .cu files UNMODIFIED, from where they
// lie under /root/reference, together with this file into oracle/_ref/libmssvt_ref.so.  The
// launchers have C++ linkage and take raw device pointers; this file only re-declares their
// prototypes (pcdet/ops/mssvt/src/ms_sparse_attention_gpu.h:14-74, group_features_gpu.h:21-29,
// pcdet/ops/pointnet2/pointnet2_batch/src/{sampling,interpolate,group_points}_gpu.h) and forwards
// to them so tests on the GPU box can call the real reference kernels through ctypes.
// All launches go to the legacy default stream, as in the reference.
#include <cuda_runtime_api.h>

void build_mapping_with_hash_kernel_launcher(int, int, int, int, int, const int *, const int *, int *);
void window_with_hash_kernel_launcher(int, int, int, int, int, int, int, int, int, const int *,
                                      int *, int *, int *);
void gather_two_window_voxels_with_hash_kernel_launcher(
    int, int, int, int, int, int, int, int, int, int, int, int, int, int, int, int, int *, int *,
    int *, int *, int *, int *, int *, int *, const int *, const int *, const int *, const int *,
    const int *, const int *);
void gather_one_window_voxels_with_hash_kernel_launcher(int, int, int, int, int, int, int, int,
                                                        int, int, int *, int *, const int *,
                                                        const int *, const int *);
void group_features_kernel_launcher_stack(int, int, int, int, const float *, const int *,
                                          const int *, const int *, float *);
void group_features_grad_kernel_launcher_stack(int, int, int, int, int, const float *,
                                               const int *, const int *, const int *, float *);
void farthest_point_sampling_kernel_launcher(int, int, int, const float *, float *, int *);
void gather_points_kernel_launcher_fast(int, int, int, int, const float *, const int *, float *);
void three_nn_kernel_launcher_fast(int, int, int, const float *, const float *, float *, int *);
void group_points_kernel_launcher_fast(int, int, int, int, int, const float *, const int *,
                                       float *);

static int done() {
    cudaError_t e = cudaDeviceSynchronize();
    if (e == cudaSuccess) e = cudaGetLastError();
    return (int)e;
}

extern "C" {

int ref_build_hash_table(int x_max, int y_max, int z_max, int num_voxels, int hash_size,
                         const int *v_indices, const int *v_bs_cnt, int *table) {
    build_mapping_with_hash_kernel_launcher(x_max, y_max, z_max, num_voxels, hash_size, v_indices,
                                            v_bs_cnt, table);
    return done();
}

int ref_window_partition(int x_wgs, int y_wgs, int z_wgs, int x_ws, int y_ws, int z_ws,
                         int num_voxels, int max_wins, int hash_size, const int *v_indices,
                         int *w_indices, int *table, int *vcount) {
    window_with_hash_kernel_launcher(x_wgs, y_wgs, z_wgs, x_ws, y_ws, z_ws, num_voxels, max_wins,
                                     hash_size, v_indices, w_indices, table, vcount);
    return done();
}

int ref_gather_two_window(int x_max, int y_max, int z_max, int x_ws, int y_ws, int z_ws,
                          int max_odd, int max_even, int max_win1, int max_win2, int num_wins,
                          int hash_size, int num_odd, int num_even, int num_win1, int num_win2,
                          int *ind_odd, int *ind_even, int *ind_win1, int *ind_win2,
                          int *coord_odd, int *coord_even, int *coord_win1, int *coord_win2,
                          const int *q_odd, const int *q_even, const int *q_win1,
                          const int *q_win2, const int *win_indices, const int *table) {
    gather_two_window_voxels_with_hash_kernel_launcher(
        x_max, y_max, z_max, x_ws, y_ws, z_ws, max_odd, max_even, max_win1, max_win2, num_wins,
        hash_size, num_odd, num_even, num_win1, num_win2, ind_odd, ind_even, ind_win1, ind_win2,
        coord_odd, coord_even, coord_win1, coord_win2, q_odd, q_even, q_win1, q_win2, win_indices,
        table);
    return done();
}

int ref_gather_one_window(int x_max, int y_max, int z_max, int x_ws, int y_ws, int z_ws,
                          int max_win1, int num_wins, int hash_size, int num_win1, int *ind_win1,
                          int *coord_win1, const int *q_win1, const int *win_indices,
                          const int *table) {
    gather_one_window_voxels_with_hash_kernel_launcher(x_max, y_max, z_max, x_ws, y_ws, z_ws,
                                                       max_win1, num_wins, hash_size, num_win1,
                                                       ind_win1, coord_win1, q_win1, win_indices,
                                                       table);
    return done();
}

int ref_group_features(int B, int M, int C, int nsample, const float *features,
                       const int *features_batch_cnt, const int *idx, const int *idx_batch_cnt,
                       float *out) {
    group_features_kernel_launcher_stack(B, M, C, nsample, features, features_batch_cnt, idx,
                                         idx_batch_cnt, out);
    return done();
}

int ref_group_features_grad(int B, int M, int C, int N, int nsample, const float *grad_out,
                            const int *idx, const int *idx_batch_cnt,
                            const int *features_batch_cnt, float *grad_features) {
    group_features_grad_kernel_launcher_stack(B, M, C, N, nsample, grad_out, idx, idx_batch_cnt,
                                              features_batch_cnt, grad_features);
    return done();
}

int ref_fps(int b, int n, int m, const float *dataset, float *temp, int *idxs) {
    farthest_point_sampling_kernel_launcher(b, n, m, dataset, temp, idxs);
    return done();
}

int ref_gather_points(int b, int c, int n, int m, const float *points, const int *idx,
                      float *out) {
    gather_points_kernel_launcher_fast(b, c, n, m, points, idx, out);
    return done();
}

int ref_three_nn(int b, int n, int m, const float *unknown, const float *known, float *dist2,
                 int *idx) {
    three_nn_kernel_launcher_fast(b, n, m, unknown, known, dist2, idx);
    return done();
}

int ref_group_points(int b, int c, int n, int npoints, int nsample, const float *points,
                     const int *idx, float *out) {
    group_points_kernel_launcher_fast(b, c, n, npoints, nsample, points, idx, out);
    return done();
}

}  // extern "C"
