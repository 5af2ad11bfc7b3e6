// Path B — tiny-MLP forward on tcgen05 tensor cores, fused with the hash-grid encoding (sm_100a).
//
// Same function as ngp::forward_kernel (csrc/ngp_train.cu): coords -> (rgb, sigma), but the five
// dense layers run as tcgen05.mma on a 128-sample tile:
//   thread = sample (128 threads / CTA, persistent CTAs):
//     hash encode (128 independent fp16 gathers) -> row of the A tile in shared memory (fp16,
//     K-major, 128B-swizzled) -> one thread issues M128 x N{64,16} x K16 MMAs against the layer's
//     pre-packed weight image -> accumulator in TMEM -> tcgen05.ld of the thread's own row ->
//     ReLU / exp / SH / sigmoid in registers -> next layer's A row ... activations never leave the SM.
// Weights: fp16 copies of the fp32 master, packed by ngp_pack_mlp_kernel after every Adam step into
// UMMA-ready images B_l[n][k] = W_l[k][n] (K padded to 64).
// FLOPs per sample: 2 * 10 240; roofline: tensor pipe (north_star target for the tiny MLP), in
// practice bounded by the gather + epilogue.
#include "ngp_common.cuh"
#include "tc.cuh"
#include "../../include/nslam_ngp.h"

namespace ngp {

// packed weight images (bytes): W1 [64][128B], W2 [16][128B], W3 [64][128B], W4 [64][128B], W5 [16][128B]
constexpr int PW1 = 0, PW2 = 8192, PW3 = 10240, PW4 = 18432, PW5 = 26624, PW_TOTAL = 28672;

__global__ void pack_mlp_kernel(const float* __restrict__ mlp, unsigned char* __restrict__ packed) {
  // one thread per (layer, n, k) element of the padded [N][64] images
  const int id = blockIdx.x * blockDim.x + threadIdx.x;
  const int offs[5] = {PW1, PW2, PW3, PW4, PW5};
  const int Ns[5] = {64, 16, 64, 64, 16};
  const int Ks[5] = {32, 64, 32, 64, 64};
  const int woff[5] = {W1_OFF, W2_OFF, W3_OFF, W4_OFF, W5_OFF};
  int base = 0;
  for (int l = 0; l < 5; l++) {
    const int cnt = Ns[l] * 64;
    if (id < base + cnt) {
      const int e = id - base, n = e / 64, k = e % 64;
      const float v = (k < Ks[l]) ? mlp[woff[l] + k * Ns[l] + n] : 0.f;
      // row n, 16-byte chunk (k/8) stored at chunk position (k/8) ^ (n & 7)
      __half* dst = reinterpret_cast<__half*>(packed + offs[l] + n * 128 + (((k >> 3) ^ (n & 7)) << 4)) + (k & 7);
      *dst = __float2half_rn(v);
      return;
    }
    base += cnt;
  }
}

struct FwdSmem {
  static constexpr int W = 0;                 // PW_TOTAL (1024-aligned images)
  static constexpr int A = 28672;             // [128][128 B]
  static constexpr int BAR = A + 16384;
  static constexpr int TOTAL = BAR + 64;
};

__device__ __forceinline__ void store_row_chunk(unsigned char* arow, int row, int chunk, const float* v8) {
  __half2 h2[4];
#pragma unroll
  for (int j = 0; j < 4; j++) h2[j] = __floats2half2_rn(v8[2 * j], v8[2 * j + 1]);
  *reinterpret_cast<uint4*>(arow + ((chunk ^ (row & 7)) << 4)) = *reinterpret_cast<const uint4*>(h2);
}

// one layer: A tile (already written + fenced + synced) x packed weights -> TMEM cols [0,N)
template <int N>
__device__ __forceinline__ void issue_layer(uint32_t a_addr, uint32_t b_addr, uint32_t tmem, int ksteps, uint64_t* bar) {
  constexpr uint32_t idesc = tc::umma_idesc_f16(128, N, 0);
  tc::tc_fence_after();
  for (int k = 0; k < ksteps; k++)
    tc::umma_f16(tmem, tc::umma_desc_sw128(a_addr + k * 32), tc::umma_desc_sw128(b_addr + k * 32), idesc, k ? 1u : 0u);
  tc::umma_commit(bar);
}

__global__ void __launch_bounds__(128)
forward_tc_kernel(const float* __restrict__ coords, const int* __restrict__ counters, int n_fixed,
                  const __half2* __restrict__ grid, LevelInfo lv, const unsigned char* __restrict__ packed,
                  float* __restrict__ rgbsigma) {
  extern __shared__ unsigned char smem_raw[];
  unsigned char* sm = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t* bar = reinterpret_cast<uint64_t*>(sm + FwdSmem::BAR);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar + 1);
  const int tid = threadIdx.x, warp = tid >> 5;
  const int n = n_fixed >= 0 ? n_fixed : counters[0];
  const int ntiles = (n + 127) / 128;

  for (int i = tid; i < PW_TOTAL / 16; i += 128)
    reinterpret_cast<uint4*>(sm + FwdSmem::W)[i] = reinterpret_cast<const uint4*>(packed)[i];
  if (tid == 0) { tc::mbar_init(bar, 1); tc::fence_barrier_init(); }
  if (warp == 0) tc::tmem_alloc<64>(tmem_slot);
  tc::fence_proxy_async();
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16);
  const uint32_t a_addr = tc::smem_u32(sm + FwdSmem::A);
  const uint32_t w_addr = tc::smem_u32(sm + FwdSmem::W);
  unsigned char* arow = sm + FwdSmem::A + tid * 128;
  uint32_t phase = 0;

  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const int s = tile * 128 + tid;
    const bool act = s < n;
    float c7[7] = {0.5f, 0.5f, 0.5f, 0.f, 0.f, 0.f, 1.f};
    if (act) {
#pragma unroll
      for (int k = 0; k < 7; k++) c7[k] = coords[(size_t)s * 7 + k];
    }
    // ---- encode -> A row (cols 0..31), zero cols 32..63
    {
      float enc[ENC_DIM];
      hash_encode(c7, grid, lv, enc);
      const float z8[8] = {0, 0, 0, 0, 0, 0, 0, 0};
#pragma unroll
      for (int c = 0; c < 4; c++) store_row_chunk(arow, tid, c, enc + 8 * c);
#pragma unroll
      for (int c = 4; c < 8; c++) store_row_chunk(arow, tid, c, z8);
    }
    tc::fence_proxy_async();
    tc::tc_fence_before();
    __syncthreads();
    // ---- layer 1: 32 -> 64, ReLU
    if (tid == 0) issue_layer<64>(a_addr, w_addr + PW1, tmem, 2, bar);
    tc::mbar_wait(bar, phase & 1); phase++;
    tc::tc_fence_after();
    {
      uint32_t r[32];
      float v[8];
#pragma unroll
      for (int h = 0; h < 2; h++) {
        tc::tmem_ld_32x32(taddr + h * 32, r);
        tc::tmem_ld_wait();
#pragma unroll
        for (int c = 0; c < 4; c++) {
#pragma unroll
          for (int j = 0; j < 8; j++) v[j] = fmaxf(__uint_as_float(r[c * 8 + j]), 0.f);
          store_row_chunk(arow, tid, h * 4 + c, v);
        }
      }
    }
    tc::fence_proxy_async();
    tc::tc_fence_before();
    __syncthreads();
    // ---- layer 2: 64 -> 16 (density head)
    if (tid == 0) issue_layer<16>(a_addr, w_addr + PW2, tmem, 4, bar);
    tc::mbar_wait(bar, phase & 1); phase++;
    tc::tc_fence_after();
    float sigma;
    {
      uint32_t r[32];
      tc::tmem_ld_32x32(taddr, r);     // 16 valid columns
      tc::tmem_ld_wait();
      float o[16], sh[16];
#pragma unroll
      for (int j = 0; j < 16; j++) o[j] = __uint_as_float(r[j]);
      sigma = __expf(o[0]);
      sh4(c7 + 4, sh);
      const float z8[8] = {0, 0, 0, 0, 0, 0, 0, 0};
      store_row_chunk(arow, tid, 0, o); store_row_chunk(arow, tid, 1, o + 8);
      store_row_chunk(arow, tid, 2, sh); store_row_chunk(arow, tid, 3, sh + 8);
#pragma unroll
      for (int c = 4; c < 8; c++) store_row_chunk(arow, tid, c, z8);
    }
    tc::fence_proxy_async();
    tc::tc_fence_before();
    __syncthreads();
    // ---- layers 3, 4: -> 64, ReLU
#pragma unroll 1
    for (int l = 0; l < 2; l++) {
      if (tid == 0) issue_layer<64>(a_addr, w_addr + (l == 0 ? PW3 : PW4), tmem, l == 0 ? 2 : 4, bar);
      tc::mbar_wait(bar, phase & 1); phase++;
      tc::tc_fence_after();
      uint32_t r[32];
      float v[8];
#pragma unroll
      for (int h = 0; h < 2; h++) {
        tc::tmem_ld_32x32(taddr + h * 32, r);
        tc::tmem_ld_wait();
#pragma unroll
        for (int c = 0; c < 4; c++) {
#pragma unroll
          for (int j = 0; j < 8; j++) v[j] = fmaxf(__uint_as_float(r[c * 8 + j]), 0.f);
          store_row_chunk(arow, tid, h * 4 + c, v);
        }
      }
      tc::fence_proxy_async();
      tc::tc_fence_before();
      __syncthreads();
    }
    // ---- layer 5: 64 -> 3 (16), sigmoid
    if (tid == 0) issue_layer<16>(a_addr, w_addr + PW5, tmem, 4, bar);
    tc::mbar_wait(bar, phase & 1); phase++;
    tc::tc_fence_after();
    {
      uint32_t r[32];
      tc::tmem_ld_32x32(taddr, r);
      tc::tmem_ld_wait();
      if (act) {
        float4 out;
        out.x = 1.f / (1.f + __expf(-__uint_as_float(r[0])));
        out.y = 1.f / (1.f + __expf(-__uint_as_float(r[1])));
        out.z = 1.f / (1.f + __expf(-__uint_as_float(r[2])));
        out.w = sigma;
        reinterpret_cast<float4*>(rgbsigma)[s] = out;
      }
    }
    tc::tc_fence_before();
    __syncthreads();      // TMEM + A tile reusable by the next tile
  }
  tc::tc_fence_before();
  __syncthreads();
  if (warp == 0) tc::tmem_dealloc<64>(tmem);
}

}  // namespace ngp

extern "C" {

int nslam_ngp_pack_mlp(const float* mlp, void* packed, void* stream) {
  const int total = (64 + 16 + 64 + 64 + 16) * 64;
  ngp::pack_mlp_kernel<<<(total + 255) / 256, 256, 0, (cudaStream_t)stream>>>(mlp, (unsigned char*)packed);
  cudaError_t e = cudaGetLastError();
  return (int)e;
}

/* tensor-core variant of nslam_ngp_forward: coords [n,7] (n < 0: read counters[0]) -> rgbsigma [n,4] */
int nslam_ngp_forward_tc(const nslam_ngp_model* m, const void* packed, const float* coords, const int* counters,
                         int n, int max_samples, float* rgbsigma, int num_sms, void* stream) {
  using namespace ngp;
  LevelInfo lv;
  for (int l = 0; l < N_LEVELS; l++) {
    lv.scale[l] = m->scale[l]; lv.res[l] = m->res[l]; lv.size[l] = m->size[l];
    lv.offset[l] = m->offset[l]; lv.dense[l] = m->dense[l];
  }
  const int smem = FwdSmem::TOTAL + 1024;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(forward_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return (int)e;
    configured = true;
  }
  const int cap = n >= 0 ? n : max_samples;
  if (cap == 0) return 0;
  int grid = (cap + 127) / 128;
  if (grid > 4 * num_sms) grid = 4 * num_sms;
  forward_tc_kernel<<<grid, 128, smem, (cudaStream_t)stream>>>(coords, counters, n, (const __half2*)m->grid_half, lv,
                                                              (const unsigned char*)packed, rgbsigma);
  cudaError_t e = cudaGetLastError();
  return (int)e;
}

}  // extern "C"
