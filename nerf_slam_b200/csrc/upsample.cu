// A17 — 8x convex upsampling with border-masked softmax over the 3x3 neighbourhood.
//
// Replaces cvx_upsample (reference utils/flow_viz.py:166-183), called on inverse depths and on
// depth covariances at slam/visual_frontends/visual_frontend.py:444-446.
//   data [K,ht,wd] fp32, mask [K,576,ht,wd] (fp16 under autocast, or fp32),  out [K,8ht,8wd] fp32
//   mask channel = n9*64 + sy*8 + sx, n9 = 3*(dy+1) + (dx+1); neighbours that fall outside the
//   image get -inf BEFORE the softmax (the reference writes -inf into the mask tensor in place;
//   we only mask functionally and never mutate the caller's tensor);
//   for an fp16 mask the softmax weights are rounded to fp16 like torch.softmax(half) does.
//
// Reference: F.unfold + a 7-D broadcast product.  Here one CTA handles one coarse row segment of
// 32 pixels: the 576x32 mask slab is read once, coalesced along x, into shared memory and the
// 8x256 output patch is written with coalesced rows.  Algorithmic bytes per depth map at 640x480:
// 576*4800*2 (mask fp16) + 4800*4 + 307200*4 = 6.8 MB.
#include "common.cuh"

namespace nslam {

template <typename T> __device__ __forceinline__ float to_f(T v);
template <> __device__ __forceinline__ float to_f<float>(float v) { return v; }
template <> __device__ __forceinline__ float to_f<__half>(__half v) { return __half2float(v); }

constexpr int UX = 32;  // coarse pixels per CTA

// grid: (ceil(wd/UX), ht, K)  block: 256
template <typename T>
__global__ void __launch_bounds__(256)
cvx_upsample_kernel(const float* __restrict__ data, const T* __restrict__ mask,
                    float* __restrict__ out, int ht, int wd, float pw, int mask_nhwc) {
  extern __shared__ unsigned char smraw[];
  T* ms = reinterpret_cast<T*>(smraw);  // [576][UX]
  __shared__ float nb[3][UX + 2];
  const int k = blockIdx.z, y = blockIdx.y, x0 = blockIdx.x * UX;
  const int hw = ht * wd;
  if (mask_nhwc) {
    // channels-last mask [K,ht,wd,576] (what the tensor-core update operator produces)
    const T* mk = mask + ((size_t)k * hw + (size_t)y * wd + x0) * 576;
    for (int id = threadIdx.x; id < 576 * UX; id += 256) {
      const int xx = id / 576, c = id % 576;
      ms[c * UX + xx] = (x0 + xx < wd) ? mk[(size_t)xx * 576 + c] : T(0);
    }
  } else {
    const T* mk = mask + (size_t)k * 576 * hw + (size_t)y * wd + x0;
    for (int id = threadIdx.x; id < 576 * UX; id += 256) {
      const int c = id / UX, xx = id % UX;
      ms[id] = (x0 + xx < wd) ? mk[(size_t)c * hw + xx] : T(0);
    }
  }
  for (int id = threadIdx.x; id < 3 * (UX + 2); id += 256) {
    const int r = id / (UX + 2), xx = id % (UX + 2);
    const int yy = y + r - 1, xg = x0 + xx - 1;
    nb[r][xx] = (yy >= 0 && yy < ht && xg >= 0 && xg < wd) ? data[(size_t)k * hw + yy * wd + xg] : 0.f;
  }
  __syncthreads();
  // 8 x (8*UX) outputs; thread -> (sy, X)
  const int W8 = 8 * wd;
  for (int id = threadIdx.x; id < 8 * 8 * UX; id += 256) {
    const int sy = id / (8 * UX), X = id % (8 * UX);
    const int xx = X / 8, sx = X % 8;
    const int x = x0 + xx;
    if (x >= wd) continue;
    float m[9], mx = -INFINITY;
#pragma unroll
    for (int n = 0; n < 9; n++) {
      const int dy = n / 3 - 1, dx = n % 3 - 1;
      const bool inb = (y + dy >= 0) && (y + dy < ht) && (x + dx >= 0) && (x + dx < wd);
      m[n] = inb ? to_f<T>(ms[(n * 64 + sy * 8 + sx) * UX + xx]) : -INFINITY;
      mx = fmaxf(mx, m[n]);
    }
    float e[9], sum = 0.f;
#pragma unroll
    for (int n = 0; n < 9; n++) { e[n] = expf(m[n] - mx); sum += e[n]; }
    float acc = 0.f;
#pragma unroll
    for (int n = 0; n < 9; n++) {
      float wgt = e[n] / sum;
      if (sizeof(T) == 2) wgt = __half2float(__float2half_rn(wgt));
      if (pw != 1.0f) wgt = powf(wgt, pw);
      acc += wgt * nb[n / 3][xx + n % 3];
    }
    out[((size_t)k * 8 * ht + (size_t)y * 8 + sy) * W8 + (size_t)x0 * 8 + X] = acc;
  }
}

}  // namespace nslam

extern "C" int nslam_cvx_upsample(const float* data, const void* mask, int mask_dtype, float* out,
                                  int K, int ht, int wd, float pw, int mask_nhwc, void* stream) {
  using namespace nslam;
  if (K == 0) return 0;
  dim3 grid((wd + UX - 1) / UX, ht, K);
  cudaStream_t st = (cudaStream_t)stream;
  if (mask_dtype == 0) {
    const size_t smem = 576 * UX * sizeof(__half);
    cvx_upsample_kernel<__half><<<grid, 256, smem, st>>>(data, (const __half*)mask, out, ht, wd, pw, mask_nhwc);
  } else if (mask_dtype == 1) {
    const size_t smem = 576 * UX * sizeof(float);
    static bool configured = false;
    if (!configured) {
      cudaError_t e = cudaFuncSetAttribute(cvx_upsample_kernel<float>,
                                           cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      if (e != cudaSuccess) return (int)e;
      configured = true;
    }
    cvx_upsample_kernel<float><<<grid, 256, smem, st>>>(data, (const float*)mask, out, ht, wd, pw, mask_nhwc);
  } else {
    return (int)cudaErrorInvalidValue;
  }
  NSLAM_CHECK_LAUNCH();
  return 0;
}
