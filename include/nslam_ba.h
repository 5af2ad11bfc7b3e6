/* C ABI of the dense bundle-adjustment path.
 *
 * Replaces, in the reference:
 *   reduced_camera_matrix_cuda  src/droid_kernels.cu:1681-1768  (droid_backends.reduced_camera_matrix, src/droid.cpp:167-196)
 *   solve_depth_cuda            src/droid_kernels.cu:1772-1825  (droid_backends.solve_depth,           src/droid.cpp:198-218)
 *   solve_poses_cuda            src/droid_kernels.cu:1836-1847  (droid_backends.solve_poses,           src/droid.cpp:220-228)
 *   ba_cuda                     src/droid_kernels.cu:1441-1568  (droid_backends.ba,                    src/droid.cpp:133-165)
 *   the gtsam dense solve + retract of RaftVisualFrontend.ba    slam/visual_frontends/visual_frontend.py:1097-1162
 *   the covariance extraction                                    slam/visual_frontends/visual_frontend.py:1164-1230
 *
 * All pointers are DEVICE pointers unless stated; all kernels run on `stream`; no call
 * synchronises the host.  Return value: 0 on success, otherwise a cudaError_t value.
 */
#ifndef NSLAM_BA_H_
#define NSLAM_BA_H_

#ifdef __cplusplus
extern "C" {
#endif

/* Host-built description of the factor graph for one BA window (int32 arrays on the device).
 * Built by nerf_slam_b200/ba_graph.py; mirrors what the reference derives on the CPU in
 * reduced_camera_matrix_cuda / schur_block / accum_cuda (src/droid_kernels.cu:1065-1115,1349-1438). */
typedef struct nslam_ba_graph {
  int E;              /* number of edges (measurements)                                   */
  int P;              /* kf1 - kf0: poses in the window                                    */
  int K;              /* depth maps touched: |unique(arange(kf0,kf1) U ii)|                */
  int kf0;
  int NR;             /* total Schur rows (sum over k of R_k)                              */
  int NPAIR;          /* total Schur 6x6 blocks (sum over k of R_k^2)                      */
  int RMAX;           /* max_k R_k                                                         */
  int NHC;            /* number of H contributions (CSR payload length)                    */
  int NVC;            /* number of v contributions                                         */
  const int* ii;        /* [E] source frame of edge e                                      */
  const int* jj;        /* [E] target frame of edge e                                      */
  const int* kx;        /* [K] sorted frame ids of the depth maps                          */
  const int* src_ptr;   /* [K+1] CSR over edges grouped by source depth map                */
  const int* src_edges; /* [E]                                                             */
  const int* row_ptr;   /* [K+1] CSR over Schur rows per depth map                         */
  const int* row_pose;  /* [NR] pose index in [0,P)                                        */
  const int* row_erow;  /* [NR] row of the E tensor: pose p -> p, edge e -> P+e            */
  const int* pair_off;  /* [K+1] offset (in 6x6 blocks) of depth map k's R_k*R_k blocks    */
  const int* hc_ptr;    /* [P*P+1] CSR: contributions to dense block (a,b)                 */
  const int* hc_idx;    /* [NHC] >=0: Hs block index (which*E+e); <0: -(Schur block id)-1  */
  const int* vc_ptr;    /* [P+1]                                                           */
  const int* vc_idx;    /* [NVC] >=0: vs index (which*E+e); <0: -(Schur row id)-1          */
} nslam_ba_graph;

typedef struct nslam_ba_buffers {
  /* inputs */
  const float* poses;       /* [N,7] cam0_T_world                                          */
  const float* disps_sens;  /* [N,ht,wd] or NULL                                            */
  const float* intrinsics;  /* [4]                                                          */
  const float* extrinsics;  /* [7] cam_T_body                                               */
  const float* targets;     /* [E,2,ht,wd]                                                  */
  const float* weights;     /* [E,2,ht,wd]                                                  */
  const float* eta;         /* [K,ht,wd] damping                                            */
  float* disps;             /* [N,ht,wd]  (updated in place by nslam_ba_depth)              */
  /* outputs with the reference's layouts */
  float* H;                 /* [6P,6P]                                                      */
  float* v;                 /* [6P]                                                         */
  float* Q;                 /* [K,HW]                                                       */
  float* Emat;              /* [P+E,6,HW]                                                   */
  float* w;                 /* [K,HW]                                                       */
  float* Hs;                /* [4,E,6,6]                                                    */
  float* vs;                /* [2,E,6]                                                      */
  /* scratch */
  float* edge_aux;          /* [E,80]: rel pose (7) + stereo flag + Mi(36) + Mj(36)         */
  float* part;              /* [E,T,27] per-tile partial sums of G (21) and g (6)           */
  float* spart;             /* [NPAIR*36 + NR*6, T] per-tile Schur partials                 */
  float* sblk;              /* [NPAIR,36] + [NR,6] reduced Schur blocks                     */
  int ht, wd;
  int T;                    /* pixel tiles of 256: ceil(ht*wd/256)                          */
} nslam_ba_buffers;

/* A7-A11: linearise, accumulate, Schur-complement, assemble dense reduced camera matrix. */
int nslam_ba_reduced_camera_matrix(const nslam_ba_graph* g, const nslam_ba_buffers* b, void* stream);

/* HOST: edge list -> the tables of nslam_ba_graph, packed into one int32 buffer (each table padded to 4 ints) ready for a
 * single upload.  Replaces what reduced_camera_matrix_cuda / accum_cuda / schur_block compute on the CPU on every call
 * (src/droid_kernels.cu:1697-1706, 1065-1115, 1349-1399).  ii, jj: HOST int64 [E]; out: HOST int32 [capacity];
 * meta: HOST int [36] = {E,P,K,kf0,NR,NPAIR,RMAX,NHC,NVC,total, offsets[13], lengths[13]} in the table order
 * ii,jj,kx,src_ptr,src_edges,row_ptr,row_pose,row_erow,pair_off,hc_ptr,hc_idx,vc_ptr,vc_idx.
 * returns 0 ok, 1 capacity too small (meta[9] = ints needed), 2 invalid window. */
int nslam_ba_graph_build(const long long* ii, const long long* jj, int E, int kf0, int kf1, int* out, int capacity,
                         int* meta);

/* HOST (A18): order-sensitive choice of new edges from the pairwise flow distances — add_proximity_factors
 * (visual_frontend.py:712-775).  d: HOST fp32 [(t-kf0)*(t-kf1)], row-major over (i in [kf0,t), j in [kf1,t)), modified in
 * place; ii1/jj1 [n1]: existing edges (active + bad + inactive); es: HOST int64 [cap][2] out, *n_out edges (may exceed cap:
 * then 1 is returned and the call is to be repeated with cap >= *n_out). */
int nslam_proximity_edges(float* d, int kf0, int kf1, int t, const long long* ii1, const long long* jj1, int n1, int rad,
                          int nms, float thresh, int max_factors, int stereo, long long* es, int cap, int* n_out);

/* A12: dense solve (fp64 Cholesky in one CTA).
 *   Hin [n,n] fp32, vin [n] fp32 (n = 6P);  dx [P,6] fp32 out.
 *   prior_pose_idx >= 0: adds the 1e-4-sigma PriorFactorPose3 on that pose (visual_frontend.py:1234-1252)
 *       using prior_err[6] (device, Logmap(prior^-1 x0) in [omega,t] order) and prior_info (=1/sigma^2).
 *   lm, ep: Levenberg damping diag += ep + lm*diag (SparseBlock::solve, src/droid_kernels.cu:1320-1324); pass 0,0 for none.
 *   work: fp64 scratch, 2*n*n + 2n doubles (also receives L, lower triangular, row major, for nslam_ba_cov).
 *   Linv: optional fp32 [n,n] out = L^-1 (for the covariance step), NULL to skip.
 *   status (device int): 0 ok, 1 factorisation failed (dx is zeroed like the reference). */
int nslam_ba_solve(const float* Hin, const float* vin, int P, int prior_pose_idx,
                   const float* prior_err, float prior_info, float lm, float ep, double* work,
                   float* dx, float* Linv, int* status, void* stream);

/* A12: gtsam Pose3 retract of world_T_body (right perturbation, [omega,t], full Expmap) and
 * refresh cam0_T_world = cam0_T_body * world_T_body^-1 (visual_frontend.py:1145-1158). */
int nslam_ba_retract(float* world_T_body, float* cam_T_world, const float* cam_T_body,
                     const float* dx, int kf0, int P, void* stream);

/* pose_retr_kernel (src/droid_kernels.cu:1014-1048): left retraction exp(xi)*T, xi=[tau,phi]. */
int nslam_pose_retr(float* poses, const float* dx, int kf0, int kf1, void* stream);

/* Log(prior^-1 * x) in [omega,t] order for the prior factor; poses as [t,q]. */
int nslam_pose_prior_error(const float* world_T_body_k, const float* prior_pose, float* err6,
                           void* stream);

/* A13: depth back-substitution dz = Q (w - E^T dx), disps[kx] += dz (+ clamp to >= clamp_min if > 0). */
int nslam_ba_depth(const nslam_ba_graph* g, const nslam_ba_buffers* b, const float* dx,
                   float clamp_min, void* stream);

/* A14: per-pixel inverse-depth marginal variance  z_cov = Q + sum_j ((Q x^T L^-1)_j)^2
 * and depth variance z_cov / idepth^4 (visual_frontend.py:1196-1230).
 *   Linv [6P,6P] fp32 from nslam_ba_solve;  z_cov, depth_cov: [K,HW] out. */
int nslam_ba_cov(const nslam_ba_graph* g, const nslam_ba_buffers* b, const float* Linv,
                 float* Mscratch /* [6P,6P] */, float* z_cov, float* depth_cov, void* stream);

/* A14, reference-exact: as nslam_ba_cov, but reproducing what the reference's covariance block really computes for
 * the depth maps of OPTIMISED frames — its assignment `Ej[range(P), kf0-min:kf1-min] = Ei[range(P)]`
 * (visual_frontend.py:1214) broadcasts Ei[q] into every pose row of column q (see csrc/ba_cov_ref.cu).
 *   Mscratch: [6P*6P + 36] floats. */
int nslam_ba_cov_reference(const nslam_ba_graph* g, const nslam_ba_buffers* b, const float* Linv,
                           float* Mscratch /* [6P,6P] + 36 */, float* z_cov, float* depth_cov, void* stream);

/* block-diagonal 6x6 blocks of (L L^T)^-1 = Linv^T Linv: sigma_g [P,6,6] */
int nslam_ba_pose_cov(const float* Linv, int P, float* sigma_g, void* stream);

/* `iters` full Gauss-Newton iterations (A7-A13) in one host call; see csrc/ba.cu.
 * prior_pose [7] DEVICE (t,q) or NULL when prior_pose_idx < 0; prior_err [6] DEVICE scratch. */
int nslam_ba_gn_iterations(const nslam_ba_graph* g, const nslam_ba_buffers* b, int iters,
                           float* world_T_body, float* cam_T_world, const float* cam_T_body,
                           int prior_pose_idx, const float* prior_pose, float prior_info,
                           double* work, float* dx, float* Linv, float* prior_err, int* status,
                           float clamp_min, void* stream);

/* A14 as the live path uses it: covariances written in place into the keyframe arenas (idepths_cov, depths_cov
 * [buffer,HW]: rows kx[k]; pose_cov [buffer,6,6]: rows kf0..kf0+P-1; visual_frontend.py:1192-1194,1228-1230).
 * mode 1 = reference-exact (as nslam_ba_cov_reference), 0 = intended formula.  guard: DEVICE int or NULL; non-zero
 * (failed factorisation) -> nothing is written.  Mscratch: [6P*6P + 36] floats. */
int nslam_ba_cov_arena(const nslam_ba_graph* g, const nslam_ba_buffers* b, const float* Linv, float* Mscratch,
                       int mode, const int* guard, float* idepths_cov, float* depths_cov, float* pose_cov,
                       void* stream);

/* The live front end's BA step in one host call (RaftVisualFrontend.ba, visual_frontend.py:1071-1232):
 * nslam_ba_gn_iterations + nslam_ba_cov_arena (cov_mode < 0: no covariances).  status: DEVICE int[2] —
 * [0] = 1 when the LAST factorisation failed, [1] += 1 per failed factorisation (cumulative counter owned by the
 * caller).  A failed iteration leaves poses / depths / covariances untouched (the reference's gtsam solve raises).
 * lm, ep: Levenberg damping diag += ep + lm*diag of the pose system (0, 0 on the live path). */
int nslam_ba_frontend_update(const nslam_ba_graph* g, const nslam_ba_buffers* b, int iters,
                             float* world_T_body, float* cam_T_world, const float* cam_T_body,
                             int prior_pose_idx, const float* prior_pose, float prior_info,
                             float lm, float ep,
                             double* work, float* dx, float* Linv, float* prior_err, int* status,
                             float clamp_min, int cov_mode, float* Mscratch, float* idepths_cov,
                             float* depths_cov, float* pose_cov, void* stream);

/* pixels per CTA tile of the BA kernels: nslam_ba_buffers.T = ceil(ht*wd / tile) */
int nslam_ba_tile_pixels(void);

#ifdef __cplusplus
}
#endif
#endif
