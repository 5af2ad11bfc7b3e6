// Projective geometry kernels: A6 reproject, A16 frame_distance, and the exported-but-unused
// projmap / iproj / depth_filter of droid_backends (reference src/droid_kernels.cu:539-967).
#include "common.cuh"

namespace nslam {

// ------------------------------------------------------------------------------------------
// A6  reproject: coords1 = pi(G_j G_i^-1 pi^-1(p, d_i)), valid mask.
// Follows networks/geom/projective_ops.py:98-145 (jacobian=False branch):
//   MIN_DEPTH = 0.2 (python side, NOT the 0.25 of the BA kernels), Z < 0.1 -> Z = 1,
//   stereo edges (ii == jj) use the fixed [-0.1,0,0, 0,0,0,1] extrinsics,
//   per-frame intrinsics: inverse projection with K[ii], projection with K[jj].
// One launch for all edges, one thread per pixel (the reference runs ~30 small torch kernels).
// grid: (ceil(hw/256), E)
__global__ void __launch_bounds__(256)
reproject_kernel(const float* __restrict__ poses, const float* __restrict__ disps,
                 const float* __restrict__ intrinsics, int intr_stride,
                 const long long* __restrict__ ii, const long long* __restrict__ jj,
                 float* __restrict__ coords, float* __restrict__ valid, int ht, int wd) {
  const int e = blockIdx.y;
  const int hw = ht * wd;
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  __shared__ float tq[7];
  __shared__ float Ki[4], Kj[4];
  const int i = (int)ii[e], j = (int)jj[e];
  if (threadIdx.x == 0) {
    if (i == j) {
      tq[0] = -0.1f; tq[1] = 0.f; tq[2] = 0.f; tq[3] = 0.f; tq[4] = 0.f; tq[5] = 0.f; tq[6] = 1.f;
    } else {
      se3_rel(poses + 7 * i, poses + 7 * i + 3, poses + 7 * j, poses + 7 * j + 3, tq, tq + 3);
    }
  }
  if (threadIdx.x < 4) {
    Ki[threadIdx.x] = intrinsics[(size_t)i * intr_stride + threadIdx.x];
    Kj[threadIdx.x] = intrinsics[(size_t)j * intr_stride + threadIdx.x];
  }
  __syncthreads();
  if (k >= hw) return;
  const float u = (float)(k % wd), v = (float)(k / wd);
  float X0[4] = {(u - Ki[2]) / Ki[0], (v - Ki[3]) / Ki[1], 1.0f, disps[(size_t)i * hw + k]};
  float X1[4];
  se3_act4(tq, tq + 3, X0, X1);
  float Z = X1[2];
  const float Zs = (Z < 0.1f) ? 1.0f : Z;
  const float d = 1.0f / Zs;
  float2 c;
  c.x = Kj[0] * (X1[0] * d) + Kj[2];
  c.y = Kj[1] * (X1[1] * d) + Kj[3];
  reinterpret_cast<float2*>(coords)[(size_t)e * hw + k] = c;
  if (valid) valid[(size_t)e * hw + k] = (Z > 0.2f) ? 1.0f : 0.0f;
}

// ------------------------------------------------------------------------------------------
// A16 frame_distance.  One CTA per (i,j) pair, 256 threads.
// The summation tree is the one of the reference (thread-strided partial sums, then strides
// 128,64,32,16,...,1) and the per-pixel expression order is kept, so the fp32 distances order
// identically under argsort — this is what makes the selected edge indices reproducible
// (SURVEY.md §9.20).  Note the reference loop runs a single direction (`n<1`,
// src/droid_kernels.cu:682); bidirectionality is done by the python caller.
__global__ void __launch_bounds__(256)
frame_distance_kernel(const float* __restrict__ poses, const float* __restrict__ disps,
                      const float* __restrict__ intr, const long long* __restrict__ ii,
                      const long long* __restrict__ jj, float* __restrict__ dist, int ht, int wd,
                      float beta) {
  const int tid = threadIdx.x;
  const int ix = (int)ii[blockIdx.x], jx = (int)jj[blockIdx.x];
  const float fx = intr[0], fy = intr[1], cx = intr[2], cy = intr[3];
  float tij[3], qij[4];
  se3_rel(poses + 7 * ix, poses + 7 * ix + 3, poses + 7 * jx, poses + 7 * jx + 3, tij, qij);

  float accum = 0.f, valid = 0.f, total = 0.f;
  const int hw = ht * wd;
  const float* __restrict__ dsp = disps + (size_t)ix * hw;
  // Per-pixel arithmetic with EXPLICIT rounding points.  They reproduce, operation for operation,
  // what nvcc 12.9 emits for the reference kernel (checked against the SASS of oracle/_ref):
  // a*b - c*d  -> fma(a, b, -(c*d));  the three translation products d*t are plain multiplies
  // shared by the full-motion and the translation-only projections (NOT fused into the adds);
  // fx*(x/z)+cx -> fma;  du^2 + dv^2 -> fma(du, du, dv*dv);  accum += w*d -> fma(d, w, accum).
  const float q0 = qij[0], q1 = qij[1], q2 = qij[2], q3 = qij[3];
  const float omb = 1 - beta;
  for (int k = tid; k < hw; k += 256) {
    const float u = (float)(k % wd), v = (float)(k / wd);
    const float X = __fdiv_rn(__fsub_rn(u, cx), fx), Y = __fdiv_rn(__fsub_rn(v, cy), fy);
    const float dsk = dsp[k];
    float uv0 = __fmaf_rn(q2, -Y, q1);                 uv0 = __fadd_rn(uv0, uv0);
    float uv1 = __fmaf_rn(q2, X, -q0);                 uv1 = __fadd_rn(uv1, uv1);
    float uv2 = __fmaf_rn(q0, Y, -__fmul_rn(q1, X));   uv2 = __fadd_rn(uv2, uv2);
    const float r0 = __fadd_rn(__fmaf_rn(q3, uv0, X), __fmaf_rn(q1, uv2, -__fmul_rn(q2, uv1)));
    const float r1 = __fadd_rn(__fmaf_rn(q3, uv1, Y), __fmaf_rn(q2, uv0, -__fmul_rn(q0, uv2)));
    const float r2 = __fadd_rn(__fmaf_rn(q3, uv2, 1.0f), __fmaf_rn(q0, uv1, -__fmul_rn(q1, uv0)));
    const float p0 = __fmul_rn(tij[0], dsk), p1 = __fmul_rn(tij[1], dsk), p2 = __fmul_rn(tij[2], dsk);
    float x = __fadd_rn(r0, p0), y = __fadd_rn(r1, p1), z = __fadd_rn(r2, p2);
    float du = __fsub_rn(__fmaf_rn(fx, __fdiv_rn(x, z), cx), u);
    float dv = __fsub_rn(__fmaf_rn(fy, __fdiv_rn(y, z), cy), v);
    float d = __fsqrt_rn(__fmaf_rn(du, du, __fmul_rn(dv, dv)));
    total = __fadd_rn(total, beta);
    if (z > NSLAM_MIN_DEPTH) {
      accum = __fmaf_rn(d, beta, accum);
      valid = __fadd_rn(valid, beta);
    }
    // translation-only flow
    x = __fadd_rn(p0, X); y = __fadd_rn(p1, Y); z = __fadd_rn(p2, 1.0f);
    du = __fsub_rn(__fmaf_rn(fx, __fdiv_rn(x, z), cx), u);
    dv = __fsub_rn(__fmaf_rn(fy, __fdiv_rn(y, z), cy), v);
    d = __fsqrt_rn(__fmaf_rn(du, du, __fmul_rn(dv, dv)));
    total = __fadd_rn(total, omb);
    if (z > NSLAM_MIN_DEPTH) {
      accum = __fmaf_rn(d, omb, accum);
      valid = __fadd_rn(valid, omb);
    }
  }
  // tree: v[t] += v[t+128]; v[t] += v[t+64]; v[t] += v[t+32]; then 16..1 inside warp 0
  __shared__ float s[3][256];
  s[0][tid] = accum; s[1][tid] = total; s[2][tid] = valid;
  __syncthreads();
  if (tid < 128) {
#pragma unroll
    for (int q = 0; q < 3; q++) s[q][tid] += s[q][tid + 128];
  }
  __syncthreads();
  if (tid < 64) {
#pragma unroll
    for (int q = 0; q < 3; q++) s[q][tid] += s[q][tid + 64];
  }
  __syncthreads();
  if (tid < 32) {
    float r[3];
#pragma unroll
    for (int q = 0; q < 3; q++) {
      float x = s[q][tid] + s[q][tid + 32];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) x += __shfl_down_sync(0xffffffffu, x, o);
      r[q] = x;
    }
    if (tid == 0)
      dist[blockIdx.x] = (r[2] / (r[1] + 1e-8) < 0.75) ? 1000.0f : r[0] / r[2];
  }
}

// ------------------------------------------------------------------------------------------
// projmap (src/droid_kernels.cu:539-628): coords[n][ht][wd][3] (3rd channel untouched = 0),
// valid[n][ht][wd][1]
__global__ void __launch_bounds__(256)
projmap_kernel(const float* __restrict__ poses, const float* __restrict__ disps,
               const float* __restrict__ intr, const long long* __restrict__ ii,
               const long long* __restrict__ jj, float* __restrict__ coords,
               float* __restrict__ valid, int ht, int wd) {
  const int e = blockIdx.y;
  const int hw = ht * wd;
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  __shared__ float tq[7];
  const int ix = (int)ii[e], jx = (int)jj[e];
  if (threadIdx.x == 0)
    se3_rel(poses + 7 * ix, poses + 7 * ix + 3, poses + 7 * jx, poses + 7 * jx + 3, tq, tq + 3);
  __syncthreads();
  if (k >= hw) return;
  const float fx = intr[0], fy = intr[1], cx = intr[2], cy = intr[3];
  const float u = (float)(k % wd), v = (float)(k / wd);
  float Xi[4] = {(u - cx) / fx, (v - cy) / fy, 1.0f, disps[(size_t)ix * hw + k]};
  float Xj[4];
  se3_act4(tq, tq + 3, Xi, Xj);
  float c0 = u, c1 = v;
  if (Xj[2] > 0.01f) {
    c0 = fx * (Xj[0] / Xj[2]) + cx;
    c1 = fy * (Xj[1] / Xj[2]) + cy;
  }
  float* c = coords + ((size_t)e * hw + k) * 3;
  c[0] = c0; c[1] = c1; c[2] = 0.f;
  valid[(size_t)e * hw + k] = (Xj[2] > NSLAM_MIN_DEPTH) ? 1.0f : 0.0f;
}

// iproj (src/droid_kernels.cu:896-967): points[n][ht][wd][3] = (G_n * X)[:3] / (G_n * X)[3]
__global__ void __launch_bounds__(256)
iproj_kernel(const float* __restrict__ poses, const float* __restrict__ disps,
             const float* __restrict__ intr, float* __restrict__ points, int ht, int wd) {
  const int n = blockIdx.y;
  const int hw = ht * wd;
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= hw) return;
  const float fx = intr[0], fy = intr[1], cx = intr[2], cy = intr[3];
  const float u = (float)(k % wd), v = (float)(k / wd);
  float Xi[4] = {(u - cx) / fx, (v - cy) / fy, 1.0f, disps[(size_t)n * hw + k]};
  float Xj[4];
  se3_act4(poses + 7 * n, poses + 7 * n + 3, Xi, Xj);
  float* p = points + ((size_t)n * hw + k) * 3;
  p[0] = Xj[0] / Xj[3]; p[1] = Xj[1] / Xj[3]; p[2] = Xj[2] / Xj[3];
}

// depth_filter (src/droid_kernels.cu:773-892): for each selected frame and its 6 temporal
// neighbours (i-1,i-2,i-3,i+3,i+4,i+5), count neighbours whose depth agrees within thresh.
// One thread per (frame, pixel) loops the 6 neighbours => no atomics, deterministic.
__global__ void __launch_bounds__(256)
depth_filter_kernel(const float* __restrict__ poses, const float* __restrict__ disps,
                    const float* __restrict__ intr, const long long* __restrict__ inds,
                    const float* __restrict__ thresh, float* __restrict__ counter, int num,
                    int ht, int wd) {
  const int b = blockIdx.y;
  const int hw = ht * wd;
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  const int ix = (int)inds[b];
  __shared__ float tq[6][7];
  __shared__ int jxs[6];
  if (threadIdx.x < 6) {
    const int nb = threadIdx.x;
    const int jx = (nb < 3) ? ix - nb - 1 : ix + nb;
    jxs[nb] = jx;
    if (jx >= 0 && jx < num)
      se3_rel(poses + 7 * ix, poses + 7 * ix + 3, poses + 7 * jx, poses + 7 * jx + 3, tq[nb],
              tq[nb] + 3);
  }
  __syncthreads();
  if (k >= hw) return;
  const float fx = intr[0], fy = intr[1], cx = intr[2], cy = intr[3];
  const float t = thresh[b];
  const float ui = (float)(k % wd), vi = (float)(k / wd);
  float Xi[4] = {(ui - cx) / fx, (vi - cy) / fy, 1.0f, disps[(size_t)ix * hw + k]};
  float cnt = 0.f;
  for (int nb = 0; nb < 6; nb++) {
    const int jx = jxs[nb];
    if (jx < 0 || jx >= num) continue;
    float Xj[4];
    se3_act4(tq[nb], tq[nb] + 3, Xi, Xj);
    const float uj = fx * (Xj[0] / Xj[2]) + cx;
    const float vj = fy * (Xj[1] / Xj[2]) + cy;
    const float dj = Xj[3] / Xj[2];
    const int u0 = (int)floorf(uj), v0 = (int)floorf(vj);
    if (u0 >= 0 && v0 >= 0 && u0 < wd - 1 && v0 < ht - 1) {
      const float* dp = disps + (size_t)jx * hw;
      const float d00 = dp[v0 * wd + u0], d01 = dp[v0 * wd + u0 + 1];
      const float d10 = dp[(v0 + 1) * wd + u0], d11 = dp[(v0 + 1) * wd + u0 + 1];
      const float inv = 1.0f / dj;
      if (fabsf(inv - 1.0f / d00) < t) cnt += 1.0f;
      else if (fabsf(inv - 1.0f / d01) < t) cnt += 1.0f;
      else if (fabsf(inv - 1.0f / d10) < t) cnt += 1.0f;
      else if (fabsf(inv - 1.0f / d11) < t) cnt += 1.0f;
    }
  }
  counter[(size_t)b * hw + k] = cnt;
}

}  // namespace nslam

extern "C" {

int nslam_reproject(const float* poses, const float* disps, const float* intrinsics,
                    int intr_stride, const long long* ii, const long long* jj, int num_edges,
                    int ht, int wd, float* coords, float* valid, void* stream) {
  if (num_edges == 0) return 0;
  dim3 grid((ht * wd + 255) / 256, num_edges);
  nslam::reproject_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(
      poses, disps, intrinsics, intr_stride, ii, jj, coords, valid, ht, wd);
  NSLAM_CHECK_LAUNCH();
  return 0;
}

int nslam_frame_distance(const float* poses, const float* disps, const float* intrinsics,
                         const long long* ii, const long long* jj, int num, int ht, int wd,
                         float beta, float* dist, void* stream) {
  if (num == 0) return 0;
  nslam::frame_distance_kernel<<<num, 256, 0, (cudaStream_t)stream>>>(poses, disps, intrinsics,
                                                                      ii, jj, dist, ht, wd, beta);
  NSLAM_CHECK_LAUNCH();
  return 0;
}

int nslam_projmap(const float* poses, const float* disps, const float* intrinsics,
                  const long long* ii, const long long* jj, int num, int ht, int wd,
                  float* coords, float* valid, void* stream) {
  if (num == 0) return 0;
  dim3 grid((ht * wd + 255) / 256, num);
  nslam::projmap_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(poses, disps, intrinsics, ii, jj,
                                                                coords, valid, ht, wd);
  NSLAM_CHECK_LAUNCH();
  return 0;
}

int nslam_iproj(const float* poses, const float* disps, const float* intrinsics, int num, int ht,
                int wd, float* points, void* stream) {
  if (num == 0) return 0;
  dim3 grid((ht * wd + 255) / 256, num);
  nslam::iproj_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(poses, disps, intrinsics, points,
                                                              ht, wd);
  NSLAM_CHECK_LAUNCH();
  return 0;
}

int nslam_depth_filter(const float* poses, const float* disps, const float* intrinsics,
                       const long long* inds, const float* thresh, int num_inds, int num_frames,
                       int ht, int wd, float* counter, void* stream) {
  if (num_inds == 0) return 0;
  dim3 grid((ht * wd + 255) / 256, num_inds);
  nslam::depth_filter_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(
      poses, disps, intrinsics, inds, thresh, counter, num_frames, ht, wd);
  NSLAM_CHECK_LAUNCH();
  return 0;
}

}  // extern "C"
