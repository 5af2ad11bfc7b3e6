/* C ABI of libnslam_sm100a.so — hand-written sm_100a kernels behind the reference's
 * `droid_backends` operator surface (reference src/droid.cpp:347-363) and the Python-side hot
 * loops around it.  Every entry point takes raw DEVICE pointers, sizes and a cudaStream_t
 * (as void*), launches on that stream, never synchronises the host and returns 0 or a
 * cudaError_t value.  dtype codes: 0 = fp16, 1 = fp32.
 *
 * Each function cites the reference interface it replaces (paths relative to /root/reference).
 * The BA entry points live in nslam_ba.h, the conv/update-operator ones in nslam_nn.h, the
 * NeRF ones in nslam_ngp.h.
 */
#ifndef NSLAM_H_
#define NSLAM_H_

#include "nslam_ba.h"

#ifdef __cplusplus
extern "C" {
#endif

/* droid_backends.corr_index_forward (src/droid.cpp:280-288; kernel src/correlation_kernels.cu:19-70)
 * volume [n,h1,w1,h2,w2], coords [n,2,h1,w1] fp32, out [n,2r+1,2r+1,h1,w1] (dtype of volume). */
int nslam_corr_index_forward(const void* volume, int dtype, const float* coords, void* out, int n,
                             int h1, int w1, int h2, int w2, int radius, void* stream);

/* CorrBlock.__call__ (networks/modules/corr.py:40-50) fused over the pyramid:
 * volumes: HOST array of num_levels device pointers, h2s/w2s: HOST int arrays;
 * coords [n,2,h1,w1] in level-0 pixels (level l samples coords/2^l); out [n,L*(2r+1)^2,h1,w1].
 * slots: optional DEVICE int32 [n]: edge n reads volume slots[n] (correlation arena), NULL = n.
 * nhwc_stride > 0: out is channels-last [n,h1,w1,nhwc_stride] (tail channels zero-filled);
 * coords_nhwc != 0: coords are [n,h1,w1,2]. */
int nslam_corr_lookup_pyramid(const void* const* volumes, const int* h2s, const int* w2s,
                              int num_levels, int dtype, const float* coords, void* out, int n,
                              int h1, int w1, int radius, const int* slots, int nhwc_stride,
                              int coords_nhwc, void* stream);

/* CorrBlock.__init__ + CorrBlock.corr (networks/modules/corr.py:23-38,63-72): all-pairs
 * correlation (f1/4).(f2/4) and the 3 avg-pooled levels, fp16, in one tcgen05 kernel.
 * fmaps [NF,H,W,C=128] fp16 channels-last; ii/jj [E] int32 DEVICE frame indices;
 * out_l [E,H,W,H>>l,W>>l] fp16. */
int nslam_corr_volume_build(const void* fmaps, int NF, int H, int W, int C, const int* ii,
                            const int* jj, int E, void* out0, void* out1, void* out2, void* out3,
                            void* stream);
/* same contract, plain SIMT kernels (test cross-check only) */
int nslam_corr_volume_build_simt(const void* fmaps, int NF, int H, int W, int C, const int* ii,
                                 const int* jj, int E, void* out0, void* out1, void* out2,
                                 void* out3, void* stream);

/* The same operation with row-pair tiles (the default for the shapes it covers; bit-identical output,
 * tests/test_gpu_parity.py) — level 0 leaves the SM as 320 contiguous bytes per source pixel instead of
 * 32-byte pieces (csrc/corr_volume_rows.cu).  C = 128, H even, W in {64, 80}; cudaErrorNotSupported otherwise. */
int nslam_corr_volume_build_rows(const void* fmaps, int NF, int H, int W, int C, const int* ii, const int* jj,
                                 int E, void* out0, void* out1, void* out2, void* out3, void* stream);
/* the same kernel writing edge e into slot slots[e] of pyramid ARENAS out0..3 = [n_slots,H,W,H>>l,W>>l]: all new edges of
 * a keyframe in one launch (CorrPool.build); cudaErrorNotSupported for shapes the row-pair kernel does not cover. */
int nslam_corr_volume_build_slots(const void* fmaps, int NF, int H, int W, int C, const int* ii, const int* jj,
                                  const int* slots, int E, int n_slots, void* out0, void* out1, void* out2, void* out3,
                                  void* stream);

/* droid_backends.altcorr_forward (src/droid.cpp:303-313; kernel src/altcorr_kernel.cu:27-149)
 * fmap1 [B,H1,W1,C], fmap2 [B,H2,W2,C], coords [B,N,H1,W1,2] fp32, corr [B,N,(2r+1)^2,H1,W1]. */
int nslam_altcorr_forward(const void* fmap1, const void* fmap2, int dtype, const float* coords,
                          void* corr, int B, int N, int H1, int W1, int H2, int W2, int C,
                          int radius, void* stream);

/* pops.projective_transform, jacobian=False (networks/geom/projective_ops.py:98-145) as used by
 * RaftVisualFrontend.reproject (slam/visual_frontends/visual_frontend.py:909-918).
 * poses [N,7], disps [N,ht,wd], intrinsics [N,4] (intr_stride=4) or [4] (intr_stride=0),
 * ii/jj [E] int64; coords [E,ht,wd,2], valid [E,ht,wd,1] (may be NULL). */
int nslam_reproject(const float* poses, const float* disps, const float* intrinsics,
                    int intr_stride, const long long* ii, const long long* jj, int num_edges,
                    int ht, int wd, float* coords, float* valid, void* stream);

/* droid_backends.frame_distance (src/droid.cpp:230-246; kernel src/droid_kernels.cu:630-769) */
int nslam_frame_distance(const float* poses, const float* disps, const float* intrinsics,
                         const long long* ii, const long long* jj, int num, int ht, int wd,
                         float beta, float* dist, void* stream);

/* droid_backends.projmap (src/droid.cpp:249-264): coords [n,ht,wd,3], valid [n,ht,wd,1] */
int nslam_projmap(const float* poses, const float* disps, const float* intrinsics,
                  const long long* ii, const long long* jj, int num, int ht, int wd,
                  float* coords, float* valid, void* stream);

/* droid_backends.iproj (src/droid.cpp:267-276): points [N,ht,wd,3] */
int nslam_iproj(const float* poses, const float* disps, const float* intrinsics, int num, int ht,
                int wd, float* points, void* stream);

/* droid_backends.depth_filter (src/droid.cpp:330-344): counter [n,ht,wd] */
int nslam_depth_filter(const float* poses, const float* disps, const float* intrinsics,
                       const long long* inds, const float* thresh, int num_inds, int num_frames,
                       int ht, int wd, float* counter, void* stream);

/* cvx_upsample (utils/flow_viz.py:166-183): data [K,ht,wd] fp32, mask [K,576,ht,wd]
 * (or [K,ht,wd,576] when mask_nhwc != 0), out [K,8ht,8wd] */
int nslam_cvx_upsample(const float* data, const void* mask, int mask_dtype, float* out, int K,
                       int ht, int wd, float pw, int mask_nhwc, void* stream);
/* two planes through one softmax (inverse depths and their covariances share the mask,
 * visual_frontend.py:444-446); data2/out2 may be NULL.  index (int64 [K], or NULL = identity): mask plane
 * k is applied to row index[k] of the data and out planes (x[kx] = cvx_upsample(y[kx], mask) gather/scatter). */
int nslam_cvx_upsample2(const float* data, const float* data2, const void* mask, int mask_dtype,
                        float* out, float* out2, const long long* index, int K, int ht, int wd, float pw,
                        int mask_nhwc, void* stream);

/* §8(f3) Sigma / TSDF fusion of the SLAM packet into a DENSE voxel grid — the per-voxel update of
 * TsdfFusion.custom_volume_integrate (reference fusion/tsdf_fusion.py:185-302; Open3D VoxelBlockGrid there).
 * tsdf, weight [nz,ny,nx] fp32, color [nz,ny,nx,3] fp32 (DEVICE, updated in place); origin3 = world position of voxel
 * (0,0,0) (HOST); idepth_up / depth_cov_up [H,W] fp32, rgb u8 [3,H,W] (DEVICE; depth_cov_up NULL = uniform weights,
 * the "tsdf" mode; given = weights 1/sqrt(variance), the "sigma" mode); intr4 = fx, fy, cx, cy at full resolution (HOST);
 * cam_T_world_tq = the packet's pose [t, q_xyzw] (DEVICE, 7 floats: no host copy of the pose). */
int nslam_tsdf_integrate(float* tsdf, float* weight, float* color, int nx, int ny, int nz, const float* origin3_host,
                         float voxel_size, const float* idepth_up, const float* depth_cov_up, const unsigned char* rgb_chw,
                         int H, int W, const float* intr4_host, const float* cam_T_world_tq, float max_depth,
                         float sdf_trunc, float max_weight, float max_depth_sigma, void* stream);

#ifdef __cplusplus
}
#endif
#endif
