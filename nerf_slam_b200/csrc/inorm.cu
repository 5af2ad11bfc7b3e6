// A1 — instance normalisation of the feature encoder, fused with ReLU and the residual add.
//
// Replaces F.instance_norm + ReLU (+ residual add + ReLU) behind BasicEncoder / ResidualBlock
// (reference networks/modules/extractor.py:6-60,118-198; norm_fn='instance' for fnet).  The
// library path reshapes channels-last tensors to [1, B*C, H, W] (two layout copies), runs two
// batch-norm kernels and separate clamp / add kernels: ~60 launches and 0.86 ms per 640x480 frame.
// Here: activations stay NHWC fp16; one pass accumulates per-(image, channel) sum / sum-of-squares
// in fp32, one pass applies (x - mean) * rstd -> ReLU (-> + residual [itself normalised for the
// strided blocks] -> ReLU).  HBM-bound: 2 bytes/element read per pass, 2 written.
#include "common.cuh"

namespace nslam {

constexpr int IN_PIX = 512;   // pixels per CTA in the statistics pass

// x [B, HW, C] fp16 ; stats [B, C, 2] fp32 (sum, sumsq), zero on entry.  block 256
__global__ void __launch_bounds__(256)
inorm_stats_kernel(const __half* __restrict__ x, float* __restrict__ stats, int HW, int C) {
  __shared__ float acc[128 * 2];
  const int b = blockIdx.y, p0 = blockIdx.x * IN_PIX;
  const int CG = C / 8, PL = 256 / CG;
  const int cg = threadIdx.x % CG, pl = threadIdx.x / CG;
  for (int i = threadIdx.x; i < 2 * C; i += 256) acc[i] = 0.f;
  __syncthreads();
  float s[8] = {0, 0, 0, 0, 0, 0, 0, 0}, q[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  const int pend = min(p0 + IN_PIX, HW);
  for (int p = p0 + pl; p < pend; p += PL) {
    const uint4 raw = *reinterpret_cast<const uint4*>(x + ((size_t)b * HW + p) * C + cg * 8);
    const __half* h = reinterpret_cast<const __half*>(&raw);
#pragma unroll
    for (int j = 0; j < 8; j++) { const float v = __half2float(h[j]); s[j] += v; q[j] = fmaf(v, v, q[j]); }
  }
  // lanes that own the same channel group (lane % CG equal) are reduced by shuffles first, so that each
  // warp issues one shared-memory atomic per channel instead of one per thread
  const int lane = threadIdx.x & 31;
#pragma unroll
  for (int j = 0; j < 8; j++) {
    for (int o = 16; o >= CG; o >>= 1) {
      s[j] += __shfl_xor_sync(0xffffffffu, s[j], o);
      q[j] += __shfl_xor_sync(0xffffffffu, q[j], o);
    }
  }
  if (lane < CG || CG >= 32) {
#pragma unroll
    for (int j = 0; j < 8; j++) { atomicAdd(&acc[(cg * 8 + j) * 2], s[j]); atomicAdd(&acc[(cg * 8 + j) * 2 + 1], q[j]); }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 2 * C; i += 256) atomicAdd(stats + (size_t)b * C * 2 + i, acc[i]);
}

// y = relu?((x - mean) * rstd) ; if res: y = relu(res' + y), res' = res or its own instance norm.
// one thread = 8 channels of one pixel
__global__ void __launch_bounds__(256)
inorm_apply_kernel(const __half* __restrict__ x, const float* __restrict__ stats, const __half* __restrict__ res,
                   const float* __restrict__ res_stats, __half* __restrict__ out, int HW, int C, float eps,
                   int relu) {
  const int b = blockIdx.y;
  const int CG = C / 8;
  const size_t id = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (id >= (size_t)HW * CG) return;
  const int cg = (int)(id % CG);
  const size_t off = ((size_t)b * HW) * C + id * 8;
  const float inv_n = 1.f / (float)HW;
  const uint4 raw = *reinterpret_cast<const uint4*>(x + off);
  const __half* h = reinterpret_cast<const __half*>(&raw);
  uint4 rr = make_uint4(0, 0, 0, 0);
  if (res) rr = *reinterpret_cast<const uint4*>(res + off);
  const __half* hr = reinterpret_cast<const __half*>(&rr);
  __half o[8];
#pragma unroll
  for (int j = 0; j < 8; j++) {
    const int c = cg * 8 + j;
    float y = __half2float(h[j]);
    if (stats) {                                                  // NULL: no normalisation (context encoder)
      const float2 st = *reinterpret_cast<const float2*>(stats + ((size_t)b * C + c) * 2);
      const float mean = st.x * inv_n;
      const float var = fmaxf(st.y * inv_n - mean * mean, 0.f);
      // the library path rounds the normalised tensor to fp16 before the ReLU / add
      y = __half2float(__float2half_rn((y - mean) * rsqrtf(var + eps)));
    }
    if (relu) y = fmaxf(y, 0.f);
    if (res) {
      float r = __half2float(hr[j]);
      if (res_stats) {
        const float2 sr = *reinterpret_cast<const float2*>(res_stats + ((size_t)b * C + c) * 2);
        const float mr = sr.x * inv_n;
        const float vr = fmaxf(sr.y * inv_n - mr * mr, 0.f);
        r = __half2float(__float2half_rn((r - mr) * rsqrtf(vr + eps)));
      }
      y = fmaxf(__half2float(__float2half_rn(r + y)), 0.f);
    }
    o[j] = __float2half_rn(y);
  }
  *reinterpret_cast<uint4*>(out + off) = *reinterpret_cast<const uint4*>(o);
}

}  // namespace nslam

extern "C" {

/* x [B,HW,C] fp16 NHWC (C multiple of 8, <= 128); stats [B,C,2] fp32 is zeroed here unless the caller
 * passes zero_first = 0 (it then provides an already-zeroed buffer, e.g. a slice of one arena per forward) */
int nslam_inorm_stats(const void* x, float* stats, int B, int HW, int C, int zero_first, void* stream) {
  if (C % 8 != 0 || C > 128 || 256 % (C / 8) != 0) return (int)cudaErrorInvalidValue;
  cudaStream_t st = (cudaStream_t)stream;
  if (zero_first) cudaMemsetAsync(stats, 0, (size_t)B * C * 2 * sizeof(float), st);
  dim3 grid((HW + nslam::IN_PIX - 1) / nslam::IN_PIX, B);
  nslam::inorm_stats_kernel<<<grid, 256, 0, st>>>((const __half*)x, stats, HW, C);
  NSLAM_CHECK_LAUNCH();
  return 0;
}

/* out = relu?(IN(x)) ; with res: out = relu(res' + relu?(IN(x))), res' = IN(res) when res_stats != NULL.
 * stats == NULL: IN(x) := x (plain residual add + ReLU of the un-normalised context encoder).  out may alias x. */
int nslam_inorm_apply(const void* x, const float* stats, const void* res, const float* res_stats, void* out,
                      int B, int HW, int C, float eps, int relu, void* stream) {
  if (C % 8 != 0) return (int)cudaErrorInvalidValue;
  const size_t per = (size_t)HW * (C / 8);
  dim3 grid((unsigned)((per + 255) / 256), B);
  nslam::inorm_apply_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>((const __half*)x, stats, (const __half*)res, res_stats,
                                                                    (__half*)out, HW, C, eps, relu);
  NSLAM_CHECK_LAUNCH();
  return 0;
}

}  // extern "C"
