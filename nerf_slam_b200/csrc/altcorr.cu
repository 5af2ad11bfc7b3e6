// A4 — on-the-fly ("alt") correlation: no stored volume, dot products are recomputed per lookup.
//
// Replaces altcorr_forward_kernel (reference src/altcorr_kernel.cu:27-149, host :290-319) and is
// driven by AltCorrBlock.corr_fn (networks/modules/corr.py:107-126).
//   fmap1 [B,H1,W1,C], fmap2 [B,H2,W2,C] channels-last, coords [B,N,H1,W1,2] (x,y)
//   corr  [B,N,(2r+1)^2,H1,W1], channel = iy + (2r+1)*ix   (x-offset major)
//
// The reference uses 32-thread CTAs, stages 32-channel slabs through shared memory with a
// __syncthreads per tap and accumulates into global memory with 4 RMWs per tap.
// Here: one warp per source pixel; f1[pix] lives in registers (C/32 values per lane... in fact
// broadcast from smem), every lane owns (2r+2)^2/32 taps and streams the corresponding f2 rows
// with 128-bit loads (the feature maps are L2 resident: 1.2-2.4 MB per frame), the (2r+2)^2
// dot products are exchanged through shared memory and each output channel is written once.
// Algorithmic bytes: SURVEY.md §8(d) A4 (9.5 MB/edge unique footprint at 640x480, fp32).
#include "common.cuh"

namespace nslam {

template <typename T> struct Ld4;
template <> struct Ld4<float> {
  static __device__ __forceinline__ void ld(const float* p, float* o) {
    float4 v = __ldg(reinterpret_cast<const float4*>(p));
    o[0] = v.x; o[1] = v.y; o[2] = v.z; o[3] = v.w;
  }
};
template <> struct Ld4<__half> {
  static __device__ __forceinline__ void ld(const __half* p, float* o) {
    uint2 v = __ldg(reinterpret_cast<const uint2*>(p));
    float2 a = __half22float2(*reinterpret_cast<__half2*>(&v.x));
    float2 b = __half22float2(*reinterpret_cast<__half2*>(&v.y));
    o[0] = a.x; o[1] = a.y; o[2] = b.x; o[3] = b.y;
  }
};
template <typename T> __device__ __forceinline__ T from_float(float v);
template <> __device__ __forceinline__ float from_float<float>(float v) { return v; }
template <> __device__ __forceinline__ __half from_float<__half>(float v) { return __float2half_rn(v); }

constexpr int AC_WARPS = 4;

// grid: (ceil(H1*W1 / AC_WARPS), N, B)   block: 32*AC_WARPS.  Requires C % 4 == 0, C <= 512.
template <typename T, int R>
__global__ void __launch_bounds__(32 * AC_WARPS)
altcorr_forward_kernel(const T* __restrict__ fmap1, const T* __restrict__ fmap2,
                       const float* __restrict__ coords, T* __restrict__ corr, int N, int H1,
                       int W1, int H2, int W2, int C) {
  constexpr int RD = 2 * R + 1, NT = RD + 1, NTAP = NT * NT;
  constexpr int TPL = (NTAP + 31) / 32;  // taps per lane
  extern __shared__ float sm[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float* f1s = sm + warp * C;                                  // [C]
  float* dots = sm + AC_WARPS * C + warp * (NTAP + 1);          // [NTAP]
  const int hw = H1 * W1;
  const int pix = blockIdx.x * AC_WARPS + warp;
  const int n = blockIdx.y, b = blockIdx.z;
  if (pix >= hw) return;  // warp-uniform

  const T* f1 = fmap1 + ((size_t)b * hw + pix) * C;
  for (int c = lane * 4; c < C; c += 128) {
    float v[4];
    Ld4<T>::ld(f1 + c, v);
    f1s[c] = v[0]; f1s[c + 1] = v[1]; f1s[c + 2] = v[2]; f1s[c + 3] = v[3];
  }
  const float* cp = coords + (((size_t)b * N + n) * hw + pix) * 2;
  const float x0 = cp[0], y0 = cp[1];
  const float fx0 = floorf(x0), fy0 = floorf(y0);
  const float dx = x0 - fx0, dy = y0 - fy0;
  const int xb = (int)fx0 - R, yb = (int)fy0 - R;
  __syncwarp();

  const T* f2b = fmap2 + (size_t)b * H2 * W2 * C;
#pragma unroll
  for (int q = 0; q < TPL; q++) {
    const int t = lane + 32 * q;
    if (t < NTAP) {
      const int iy = t / NT, ix = t % NT;
      const int h2 = yb + iy, w2 = xb + ix;
      float s = 0.f;
      if (h2 >= 0 && h2 < H2 && w2 >= 0 && w2 < W2) {
        const T* f2 = f2b + ((size_t)h2 * W2 + w2) * C;
        float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
        for (int c = 0; c < C; c += 4) {
          float v[4];
          Ld4<T>::ld(f2 + c, v);
          s0 += f1s[c] * v[0]; s1 += f1s[c + 1] * v[1];
          s2 += f1s[c + 2] * v[2]; s3 += f1s[c + 3] * v[3];
        }
        s = (s0 + s1) + (s2 + s3);
      }
      dots[t] = s;
    }
  }
  __syncwarp();
  T* o = corr + (((size_t)b * N + n) * (RD * RD)) * hw + pix;
  const float w00 = (1.f - dy) * (1.f - dx), w01 = (1.f - dy) * dx;
  const float w10 = dy * (1.f - dx), w11 = dy * dx;
  for (int c = lane; c < RD * RD; c += 32) {
    const int ox = c / RD, oy = c % RD;  // channel = oy + RD*ox
    const float v = dots[oy * NT + ox] * w00 + dots[oy * NT + ox + 1] * w01 +
                    dots[(oy + 1) * NT + ox] * w10 + dots[(oy + 1) * NT + ox + 1] * w11;
    o[(size_t)c * hw] = from_float<T>(v);
  }
}

template <typename T>
static int launch_altcorr(const void* f1, const void* f2, const float* coords, void* corr, int B,
                          int N, int H1, int W1, int H2, int W2, int C, int radius,
                          cudaStream_t st) {
  if (B == 0 || N == 0) return 0;
  if (C % 4 != 0 || C > 512) return (int)cudaErrorInvalidValue;
  dim3 grid((H1 * W1 + AC_WARPS - 1) / AC_WARPS, N, B), block(32 * AC_WARPS);
  const int ntap = (2 * radius + 2) * (2 * radius + 2);
  const size_t smem = (size_t)(AC_WARPS * C + AC_WARPS * (ntap + 1)) * sizeof(float);
#define NSLAM_AC(RR)                                                                         \
  altcorr_forward_kernel<T, RR><<<grid, block, smem, st>>>((const T*)f1, (const T*)f2, coords, \
                                                           (T*)corr, N, H1, W1, H2, W2, C)
  switch (radius) {
    case 1: NSLAM_AC(1); break;
    case 2: NSLAM_AC(2); break;
    case 3: NSLAM_AC(3); break;
    case 4: NSLAM_AC(4); break;
    default: return (int)cudaErrorInvalidValue;
  }
#undef NSLAM_AC
  NSLAM_CHECK_LAUNCH();
  return 0;
}

}  // namespace nslam

extern "C" int nslam_altcorr_forward(const void* fmap1, const void* fmap2, int dtype,
                                     const float* coords, void* corr, int B, int N, int H1,
                                     int W1, int H2, int W2, int C, int radius, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype == 0)
    return nslam::launch_altcorr<__half>(fmap1, fmap2, coords, corr, B, N, H1, W1, H2, W2, C,
                                         radius, st);
  if (dtype == 1)
    return nslam::launch_altcorr<float>(fmap1, fmap2, coords, corr, B, N, H1, W1, H2, W2, C,
                                        radius, st);
  return (int)cudaErrorInvalidValue;
}
