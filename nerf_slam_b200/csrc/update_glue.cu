// A19 — small fused kernels around the update operator (what PyTorch would run as ~25 elementwise /
// indexing launches per update(): reference slam/visual_frontends/visual_frontend.py:371-470 and
// networks/droid_net.py:59-75,118-150).
//
//   nslam_motion_im2col   motion = clamp([coords1 - coords0 | target - coords1], +-64) and its 7x7
//                         im2col (49 taps x 4 channels, zero padded, fp16) in one pass: the 7x7
//                         flow-encoder conv then runs as a 1x1 tensor-core GEMM (K = 196 -> 200).
//   nslam_flow_heads_post delta/weight heads -> new flow target = coords1 + delta, confidence =
//                         sigmoid(.), written both in the frontend's [E,ht,wd,2] state and in the
//                         BA's planar [E,2,ht,wd] input buffers.
//   nslam_segment_mean    GraphAgg's scatter_mean over edges that share the source keyframe
//                         (droid_net.py:67), fp32 accumulation in edge order, fp16 out.
//   nslam_eta_damping     eta = 0.01 softplus(.) -> damping[ux]; BA damping = 0.2 damping[kx] + EP.
// All HBM-bound, one pass over their tensors.
#include "common.cuh"

namespace nslam {

constexpr int MI_C = 200;    // im2col row: 196 real values + 4 zeros (TMA rows must be 16-byte multiples)

// CTA = 32 consecutive pixels: the 32 x 200 fp16 im2col rows are assembled in shared memory (one thread per
// (pixel, tap): 4 motion channels of the tap's source pixel) and written out with coalesced 16-byte stores
constexpr int MI_PIX = 32;
__global__ void __launch_bounds__(256)
motion_im2col_kernel(const float* __restrict__ coords1, const float* __restrict__ coords0,
                     const float* __restrict__ target, __half* __restrict__ out,
                     int E, int ht, int wd) {
  __shared__ __align__(16) __half tile[MI_PIX * MI_C];
  const size_t npix = (size_t)E * ht * wd;
  const size_t p0 = (size_t)blockIdx.x * MI_PIX;
  // work item = (pixel, ky): the 7 taps of one window row read 7 horizontally adjacent source pixels
  for (int idx = threadIdx.x; idx < MI_PIX * 8; idx += 256) {
    const int px = idx >> 3, ky = idx & 7;
    const size_t pix = p0 + px;
    if (ky == 7) {                                              // columns 196..199: zero tail
      *reinterpret_cast<uint2*>(tile + px * MI_C + 196) = make_uint2(0u, 0u);
      continue;
    }
    const bool pok = pix < npix;
    const unsigned p32 = (unsigned)pix;                         // E * ht * wd < 2^31: 32-bit divisions
    const int x = (int)(p32 % (unsigned)wd), y = (int)((p32 / (unsigned)wd) % (unsigned)ht);
    const size_t e = p32 / ((unsigned)wd * (unsigned)ht);
    const int yy = y + ky - 3;
    const bool rowok = pok && yy >= 0 && yy < ht;
    const size_t rowq = (e * ht + (rowok ? yy : 0)) * wd;
    const size_t row0 = (size_t)(rowok ? yy : 0) * wd;
#pragma unroll
    for (int kx = 0; kx < 7; kx++) {
      const int xx = x + kx - 3;
      uint2 v = make_uint2(0u, 0u);
      if (rowok && xx >= 0 && xx < wd) {
        const float2 c1 = reinterpret_cast<const float2*>(coords1)[rowq + xx];
        const float2 c0 = reinterpret_cast<const float2*>(coords0)[row0 + xx];
        const float2 tg = target ? reinterpret_cast<const float2*>(target)[rowq + xx] : c1;
        const float m0 = fminf(fmaxf(c1.x - c0.x, -64.f), 64.f), m1 = fminf(fmaxf(c1.y - c0.y, -64.f), 64.f);
        const float m2 = fminf(fmaxf(tg.x - c1.x, -64.f), 64.f), m3 = fminf(fmaxf(tg.y - c1.y, -64.f), 64.f);
        const __half2 a = __floats2half2_rn(m0, m1), b = __floats2half2_rn(m2, m3);
        v.x = *reinterpret_cast<const uint32_t*>(&a); v.y = *reinterpret_cast<const uint32_t*>(&b);
      }
      *reinterpret_cast<uint2*>(tile + px * MI_C + (ky * 7 + kx) * 4) = v;
    }
  }
  __syncthreads();
  constexpr int V = MI_C / 8;                  // 25 x 16 bytes per pixel
  for (int idx = threadIdx.x; idx < MI_PIX * V; idx += 256) {
    const int px = idx / V, j = idx % V;
    const size_t pix = p0 + px;
    if (pix < npix) reinterpret_cast<uint4*>(out + pix * MI_C)[j] = reinterpret_cast<const uint4*>(tile + px * MI_C)[j];
  }
}

// h2 [E,ht,wd,16] fp16: cols 0,1 = delta, cols 2,3 = weight logits
__global__ void flow_heads_post_kernel(const __half* __restrict__ h2, const float* __restrict__ coords1,
                                       float* __restrict__ flow, float* __restrict__ conf,
                                       float* __restrict__ ba_target, float* __restrict__ ba_weight,
                                       int E, int hw) {
  const size_t id = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (id >= (size_t)E * hw) return;
  const uint2 raw = *reinterpret_cast<const uint2*>(h2 + id * 16);
  const __half2 d = *reinterpret_cast<const __half2*>(&raw.x), w = *reinterpret_cast<const __half2*>(&raw.y);
  const float2 c1 = reinterpret_cast<const float2*>(coords1)[id];
  const float2 f = make_float2(c1.x + __low2float(d), c1.y + __high2float(d));
  const float2 s = make_float2(1.f / (1.f + __expf(-__low2float(w))), 1.f / (1.f + __expf(-__high2float(w))));
  reinterpret_cast<float2*>(flow)[id] = f;
  reinterpret_cast<float2*>(conf)[id] = s;
  if (ba_target) {
    const size_t e = id / hw, p = id % hw;
    ba_target[(e * 2 + 0) * hw + p] = f.x; ba_target[(e * 2 + 1) * hw + p] = f.y;
    ba_weight[(e * 2 + 0) * hw + p] = s.x; ba_weight[(e * 2 + 1) * hw + p] = s.y;
  }
}

// out[k, p, :] = mean_{e in seg k} a[e, p, :]   (128 channels; one thread = 8 channels of one pixel)
__global__ void segment_mean_kernel(const __half* __restrict__ a, const int* __restrict__ seg_ptr,
                                    const int* __restrict__ seg_edges, __half* __restrict__ out, int K, int hw) {
  const size_t id = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (id >= (size_t)K * hw * 16) return;
  const int c8 = (int)(id % 16);
  const size_t p = (id / 16) % hw;
  const int k = (int)(id / ((size_t)16 * hw));
  const int s0 = seg_ptr[k], s1 = seg_ptr[k + 1];
  float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  for (int s = s0; s < s1; s++) {
    const uint4 raw = *reinterpret_cast<const uint4*>(a + ((size_t)seg_edges[s] * hw + p) * 128 + c8 * 8);
    const __half* h = reinterpret_cast<const __half*>(&raw);
#pragma unroll
    for (int j = 0; j < 8; j++) acc[j] += __half2float(h[j]);
  }
  const float inv = 1.f / (float)max(s1 - s0, 1);
  __half2 o[4];
#pragma unroll
  for (int j = 0; j < 4; j++) o[j] = __floats2half2_rn(acc[2 * j] * inv, acc[2 * j + 1] * inv);
  *reinterpret_cast<uint4*>(out + ((size_t)k * hw + p) * 128 + c8 * 8) = *reinterpret_cast<const uint4*>(o);
}

// e16 [K,hw,16] fp16 (col 0 = eta logit) -> damping[ux[k]] = 0.01 softplus ; then (second launch)
// ba_damp[j] = 0.2 damping[kx[j]] + EP
__global__ void eta_kernel(const __half* __restrict__ e16, const long long* __restrict__ ux,
                           float* __restrict__ damping, int K, int hw) {
  const size_t id = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (id >= (size_t)K * hw) return;
  const float x = __half2float(e16[id * 16]);
  const float sp = x > 20.f ? x : log1pf(expf(x));           // F.softplus (threshold 20)
  damping[(size_t)ux[id / hw] * hw + id % hw] = 0.01f * sp;
}
__global__ void ba_damp_kernel(const float* __restrict__ damping, const long long* __restrict__ kx,
                               float* __restrict__ out, int K, int hw, float ep) {
  const size_t id = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (id >= (size_t)K * hw) return;
  out[id] = 0.2f * damping[(size_t)kx[id / hw] * hw + id % hw] + ep;
}

}  // namespace nslam

extern "C" {

int nslam_motion_im2col(const float* coords1, const float* coords0, const float* target, void* out,
                        int E, int ht, int wd, void* stream) {
  const size_t npix = (size_t)E * ht * wd;
  if (npix == 0) return 0;
  nslam::motion_im2col_kernel<<<(unsigned)((npix + nslam::MI_PIX - 1) / nslam::MI_PIX), 256, 0, (cudaStream_t)stream>>>(
      coords1, coords0, target, (__half*)out, E, ht, wd);
  NSLAM_CHECK_LAUNCH();
  return 0;
}

int nslam_flow_heads_post(const void* h2, const float* coords1, float* flow, float* conf, float* ba_target,
                          float* ba_weight, int E, int hw, void* stream) {
  const size_t total = (size_t)E * hw;
  if (total == 0) return 0;
  nslam::flow_heads_post_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
      (const __half*)h2, coords1, flow, conf, ba_target, ba_weight, E, hw);
  NSLAM_CHECK_LAUNCH();
  return 0;
}

int nslam_segment_mean(const void* a, const int* seg_ptr, const int* seg_edges, void* out, int K, int hw,
                       void* stream) {
  const size_t total = (size_t)K * hw * 16;
  if (total == 0) return 0;
  nslam::segment_mean_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
      (const __half*)a, seg_ptr, seg_edges, (__half*)out, K, hw);
  NSLAM_CHECK_LAUNCH();
  return 0;
}

int nslam_eta_damping(const void* e16, const long long* ux, float* damping, int K, const long long* kx_ba,
                      float* ba_damp, int Kba, int hw, float ep, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  if (K > 0) {
    nslam::eta_kernel<<<(unsigned)(((size_t)K * hw + 255) / 256), 256, 0, st>>>((const __half*)e16, ux, damping, K, hw);
    NSLAM_CHECK_LAUNCH();
  }
  if (Kba > 0 && ba_damp) {
    nslam::ba_damp_kernel<<<(unsigned)(((size_t)Kba * hw + 255) / 256), 256, 0, st>>>(damping, kx_ba, ba_damp, Kba, hw, ep);
    NSLAM_CHECK_LAUNCH();
  }
  return 0;
}

}  // extern "C"

// ------------------------------------------------------------------------------------------
// A1: im2col of the encoders' first layer (7x7, stride 2, pad 3, 3 input channels -> K = 147, padded to
// 152 so that rows are 16-byte multiples): the layer then runs as a 1x1 tensor-core GEMM.
//   x [B,3,H,W] fp32 (normalised image, NCHW)  ->  out [B,H/2,W/2,152] fp16, K index = (ky*7 + kx)*3 + c
namespace nslam {
constexpr int I7_K = 147, I7_KP = 152, I7_PIX = 32;

__global__ void __launch_bounds__(256)
im2col7_s2_kernel(const float* __restrict__ x, __half* __restrict__ out, int B, int H, int W) {
  __shared__ __align__(16) __half tile[I7_PIX * I7_KP];
  const int Ho = H / 2, Wo = W / 2;
  const size_t npix = (size_t)B * Ho * Wo;
  const size_t p0 = (size_t)blockIdx.x * I7_PIX;
  // work item = (pixel, ky, channel): 7 horizontally adjacent input values -> 7 smem stores (stride 3 halfs)
  for (int idx = threadIdx.x; idx < I7_PIX * 21; idx += 256) {
    const int px = idx / 21, r = idx % 21, ky = r / 3, c = r % 3;
    const size_t pid = p0 + px;
    __half* dst = tile + px * I7_KP + ky * 21 + c;
    if (pid < npix) {
      const unsigned p32 = (unsigned)pid;                        // B * Ho * Wo < 2^31: 32-bit divisions
      const int ox = (int)(p32 % (unsigned)Wo), oy = (int)((p32 / (unsigned)Wo) % (unsigned)Ho), b = (int)(p32 / ((unsigned)Wo * (unsigned)Ho));
      const int iy = 2 * oy + ky - 3, ix0 = 2 * ox - 3;
      const float* src = x + (((size_t)b * 3 + c) * H + (iy >= 0 && iy < H ? iy : 0)) * W;
      const bool rowok = iy >= 0 && iy < H;
#pragma unroll
      for (int kx = 0; kx < 7; kx++) {
        const int ix = ix0 + kx;
        dst[kx * 3] = __float2half_rn((rowok && ix >= 0 && ix < W) ? src[ix] : 0.f);
      }
    } else {
#pragma unroll
      for (int kx = 0; kx < 7; kx++) dst[kx * 3] = __float2half_rn(0.f);
    }
  }
  for (int idx = threadIdx.x; idx < I7_PIX * (I7_KP - I7_K); idx += 256)        // zero the K padding
    tile[(idx / (I7_KP - I7_K)) * I7_KP + I7_K + idx % (I7_KP - I7_K)] = __float2half_rn(0.f);
  __syncthreads();
  constexpr int V = I7_KP / 8;                 // 19 x 16 bytes per pixel
  for (int idx = threadIdx.x; idx < I7_PIX * V; idx += 256) {
    const int px = idx / V, j = idx % V;
    const size_t pid = p0 + px;
    if (pid < npix) reinterpret_cast<uint4*>(out + pid * I7_KP)[j] = reinterpret_cast<const uint4*>(tile + px * I7_KP)[j];
  }
}
}  // namespace nslam

extern "C" int nslam_im2col7_s2(const float* x, void* out, int B, int H, int W, void* stream) {
  if (H % 2 || W % 2) return (int)cudaErrorInvalidValue;
  const size_t npix = (size_t)B * (H / 2) * (W / 2);
  if (npix == 0) return 0;
  nslam::im2col7_s2_kernel<<<(unsigned)((npix + nslam::I7_PIX - 1) / nslam::I7_PIX), 256, 0, (cudaStream_t)stream>>>(
      x, (__half*)out, B, H, W);
  NSLAM_CHECK_LAUNCH();
  return 0;
}
