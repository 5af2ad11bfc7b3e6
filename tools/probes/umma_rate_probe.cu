// Probe: issue rate of back-to-back tcgen05.mma (M128 x N x K16, fp16, both operands in shared memory, SWIZZLE_128B
// K-major) with NO loads in flight — separates the tensor pipe / operand-fetch cost of one MMA from everything around it
// (TMA, L2, epilogue).  Variants: N in {16, 64, 128, 256}; A descriptor aligned (start row 0, SBO 1024) or row-shifted
// inside a halo box (start row s, SBO 2304), as csrc/conv_halo.cu uses it.
//
// Each CTA issues `reps` MMAs that cycle through 4 K-steps (+32 B) of `stages` operand buffers, like the convolution
// main loop, commits once and waits.  out[cta] = cycles between the first issue and the arrival of the commit.
//   driver: tools/probes/run_umma_rate_probe.py
#include <cuda_fp16.h>
#include "tc.cuh"

namespace {

constexpr int A_BYTES = 48 * 1024;      // one A buffer: up to 38 + 15*18 + 8 rows of 128 B
constexpr int B_BYTES = 32 * 1024;      // one B buffer: 256 rows of 128 B
constexpr int STAGES = 2;

template <int N>
__global__ void __launch_bounds__(128, 1) rate_kernel(int reps, int shift_rows, int sbo_bytes, long long* out) {
  extern __shared__ unsigned char raw[];
  unsigned char* sm = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < STAGES * (A_BYTES + B_BYTES) / 4; i += 128) reinterpret_cast<uint32_t*>(sm)[i] = 0x3c003c00u;  // 1.0h
  if (tid == 0) { tc::mbar_init(&bar, 1); tc::fence_barrier_init(); }
  if (warp == 0) tc::tmem_alloc<512>(&slot);
  tc::fence_proxy_async();
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem = slot;
  long long t0 = 0;
  if (warp == 0) {
    const uint32_t lead = tc::elect_one() ? 1u : 0u;
    constexpr uint32_t idesc = tc::umma_idesc_f16(128, N, 0);
    const uint32_t a0 = tc::smem_u32(sm) + shift_rows * 128, b0 = tc::smem_u32(sm) + STAGES * A_BYTES;
    uint64_t ad[STAGES], bd[STAGES];
    for (int s = 0; s < STAGES; s++) {
      ad[s] = tc::umma_desc_sw128_sbo(a0 + s * A_BYTES, sbo_bytes);
      bd[s] = tc::umma_desc_sw128(b0 + s * B_BYTES);
    }
    t0 = clock64();
    for (int r = 0; r < reps; r += 4 * STAGES) {
#pragma unroll
      for (int s = 0; s < STAGES; s++) {
#pragma unroll
        for (int k = 0; k < 4; k++)
          tc::umma_f16_lead(tmem + (uint32_t)((r / (4 * STAGES)) & 1) * (N <= 256 ? 256 : 0), ad[s] + 2 * k, bd[s] + 2 * k, idesc,
                            (r | s | k) ? 1u : 0u, lead);
      }
    }
    tc::umma_commit_lead(&bar, lead);
  }
  tc::mbar_wait(&bar, 0);
  const long long t1 = clock64();
  if (tid == 0) out[blockIdx.x] = t1 - t0;
  tc::tc_fence_before();
  __syncthreads();
  if (warp == 0) tc::tmem_dealloc<512>(tmem);
}

template <int N>
int run(int ctas, int reps, int shift_rows, int sbo_bytes, long long* d) {
  const int smem = STAGES * (A_BYTES + B_BYTES) + 1024;
  cudaError_t e = cudaFuncSetAttribute(rate_kernel<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  if (e != cudaSuccess) return (int)e;
  rate_kernel<N><<<ctas, 128, smem>>>(reps, shift_rows, sbo_bytes, d);
  return (int)cudaDeviceSynchronize();
}

// L2 -> shared-memory fill rate of one SM while every SM does the same: `reps` bulk copies of 16 KB into a 4-stage ring.
// shared_src = 1: every CTA streams the SAME `span` bytes (the weight set of a layer); 0: its own `span` bytes.
__global__ void __launch_bounds__(32, 1) fill_kernel(const unsigned char* src, int reps, int span, int shared_src, long long* out) {
  extern __shared__ unsigned char raw[];
  unsigned char* sm = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t bar[4];
  if (threadIdx.x == 0) { for (int i = 0; i < 4; i++) tc::mbar_init(&bar[i], 1); tc::fence_barrier_init(); }
  __syncwarp();
  const uint32_t lead = tc::elect_one() ? 1u : 0u;
  const unsigned char* base = src + (shared_src ? 0 : (size_t)blockIdx.x * span);
  const long long t0 = clock64();
  for (int r = 0; r < reps; r++) {
    const int st = r & 3;
    if (r >= 4) tc::mbar_wait(&bar[st], ((r >> 2) - 1) & 1);
    tc::mbar_arrive_expect_tx_lead(&bar[st], 16384, lead);
    tc::bulk_copy_g2s_lead(sm + st * 16384, base + ((size_t)r * 16384) % span, 16384, &bar[st], lead);
  }
  for (int r = reps - 4; r < reps; r++) tc::mbar_wait(&bar[r & 3], (r >> 2) & 1);
  if (threadIdx.x == 0) out[blockIdx.x] = clock64() - t0;
}

// The same for TMA TENSOR loads of the convolutions' activation boxes: box {64 channels, bw pixels, bh rows} of an NHWC fp16
// tensor [B,H,W,C] (every 128-byte box row is its own segment of global memory), SWIZZLE_128B, 2-4 stages.
__global__ void __launch_bounds__(32, 1) fill_box_kernel(const __grid_constant__ CUtensorMap map, int reps, int bw, int bh, int H, int W,
                                                         int cblocks, int stages, long long* out) {
  extern __shared__ unsigned char raw[];
  unsigned char* sm = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t bar[4];
  if (threadIdx.x == 0) { for (int i = 0; i < 4; i++) tc::mbar_init(&bar[i], 1); tc::fence_barrier_init(); }
  __syncwarp();
  const uint32_t lead = tc::elect_one() ? 1u : 0u;
  const uint32_t bytes = 128u * bw * bh, stage_bytes = (bytes + 1023) & ~1023u;
  const int tw = W / 16, th = H / 16;
  const long long t0 = clock64();
  for (int r = 0; r < reps; r++) {
    const int st = r % stages, q = r / stages;
    if (q >= 1) tc::mbar_wait(&bar[st], (q - 1) & 1);
    const int tile = (blockIdx.x * 7 + r / cblocks) % (tw * th);
    tc::mbar_arrive_expect_tx_lead(&bar[st], bytes, lead);
    tc::tma_load_4d_lead(sm + st * stage_bytes, &map, &bar[st], (r % cblocks) * 64, (tile % tw) * 16 - 1, (tile / tw) * 16 - 1,
                         blockIdx.x % 18, lead);
  }
  for (int r = reps - stages; r < reps; r++) tc::mbar_wait(&bar[r % stages], (r / stages) & 1);
  if (threadIdx.x == 0) out[blockIdx.x] = clock64() - t0;
}

// Do TMA operations of DIFFERENT warps of one SM overlap?  `warps` warps, each with its own 4-stage ring: warp 0 loads
// activation boxes {64c,16w,10h} when `mixed`, all others (and warp 0 otherwise) stream bulk copies of `bytes`.
// out[cta * 8 + w] = cycles of warp w for `reps` operations.
__global__ void __launch_bounds__(256, 1) fill_multi_kernel(const __grid_constant__ CUtensorMap map, const unsigned char* src, int reps,
                                                            int bytes, int mixed, int H, int W, long long* out) {
  extern __shared__ unsigned char raw[];
  unsigned char* sm = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t bar[8][4];
  const int warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) { for (int i = 0; i < 4; i++) for (int w = 0; w < 8; w++) tc::mbar_init(&bar[w][i], 1); tc::fence_barrier_init(); }
  __syncthreads();
  const uint32_t lead = tc::elect_one() ? 1u : 0u;
  const bool boxes = mixed && warp == 0;
  const uint32_t nb = boxes ? 20480u : (uint32_t)bytes;
  unsigned char* mine = sm + warp * 4 * 20480;              // the host caps the launch at 2 warps (160 KB of rings)
  const int tw = W / 16, th = H / 8;
  const long long t0 = clock64();
  for (int r = 0; r < reps; r++) {
    const int st = r & 3;
    if (r >= 4) tc::mbar_wait(&bar[warp][st], ((r >> 2) - 1) & 1);
    tc::mbar_arrive_expect_tx_lead(&bar[warp][st], nb, lead);
    if (boxes) {
      const int tile = (blockIdx.x * 5 + r / 2) % (tw * th);
      tc::tma_load_4d_lead(mine + st * 20480, &map, &bar[warp][st], (r & 1) * 64, (tile % tw) * 16 - 1, (tile / tw) * 8 - 1, blockIdx.x % 18, lead);
    } else {
      tc::bulk_copy_g2s_lead(mine + st * 20480, src + ((size_t)(r * 4 + warp) * bytes) % (1 << 20), nb, &bar[warp][st], lead);
    }
  }
  for (int r = reps - 4; r < reps; r++) tc::mbar_wait(&bar[warp][r & 3], (r >> 2) & 1);
  if (lead) out[blockIdx.x * 8 + warp] = clock64() - t0;
}

}  // namespace

// cycles_host[ctas*8]; warps in {1,2}; bytes <= 20480 per bulk copy
extern "C" int tma_multi_warp_probe(int ctas, int warps, int reps, int bytes, int mixed, long long* cycles_host) {
  const int B = 18, H = 64, W = 80, C = 128;
  unsigned char* src = nullptr; __half* act = nullptr; long long* d = nullptr;
  if (warps < 1 || warps > 2 || bytes > 20480 || bytes % 16) return 1;
  if (cudaMalloc(&src, 2 << 20) != cudaSuccess || cudaMalloc(&act, (size_t)B * H * W * C * 2) != cudaSuccess ||
      cudaMalloc(&d, ctas * 8 * sizeof(long long)) != cudaSuccess) return 2;
  cudaMemset(src, 0, 2 << 20); cudaMemset(act, 0, (size_t)B * H * W * C * 2); cudaMemset(d, 0, ctas * 8 * sizeof(long long));
  CUtensorMap map;
  uint64_t dims[4] = {(uint64_t)C, (uint64_t)W, (uint64_t)H, (uint64_t)B};
  uint64_t strides[3] = {(uint64_t)C * 2, (uint64_t)W * C * 2, (uint64_t)H * W * C * 2};
  uint32_t box[4] = {64, 16, 10, 1};
  int rc = tc::make_tmap_f16(&map, act, 4, dims, strides, box);
  const int smem = warps * 4 * 20480 + 1024;
  if (rc == 0) rc = (int)cudaFuncSetAttribute(fill_multi_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  for (int pass = 0; pass < 2 && rc == 0; pass++) {
    fill_multi_kernel<<<ctas, 32 * warps, smem>>>(map, src, reps, bytes, mixed, H, W, d);
    rc = (int)cudaDeviceSynchronize();
  }
  if (rc == 0) rc = (int)cudaMemcpy(cycles_host, d, ctas * 8 * sizeof(long long), cudaMemcpyDeviceToHost);
  cudaFree(src); cudaFree(act); cudaFree(d);
  return rc;
}

// bytes/cycle/SM of TMA tensor loads (box {64, bw, bh, 1}) from an L2-resident NHWC tensor [18, 64, 80, C]
extern "C" int l2_fill_box_probe(int ctas, int reps, int C, int bw, int bh, int stages, long long* cycles_host) {
  const int B = 18, H = 64, W = 80;
  __half* src = nullptr; long long* d = nullptr;
  const size_t total = (size_t)B * H * W * C * 2;
  if (cudaMalloc(&src, total) != cudaSuccess || cudaMalloc(&d, ctas * sizeof(long long)) != cudaSuccess) return 2;
  cudaMemset(src, 0, total);
  CUtensorMap map;
  uint64_t dims[4] = {(uint64_t)C, (uint64_t)W, (uint64_t)H, (uint64_t)B};
  uint64_t strides[3] = {(uint64_t)C * 2, (uint64_t)W * C * 2, (uint64_t)H * W * C * 2};
  uint32_t box[4] = {64, (uint32_t)bw, (uint32_t)bh, 1};
  int rc = tc::make_tmap_f16(&map, src, 4, dims, strides, box);
  const int smem = stages * ((128 * bw * bh + 1023) & ~1023) + 1024;
  if (rc == 0) rc = (int)cudaFuncSetAttribute(fill_box_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  for (int pass = 0; pass < 2 && rc == 0; pass++) {
    fill_box_kernel<<<ctas, 32, smem>>>(map, reps, bw, bh, H, W, C / 64, stages, d);
    rc = (int)cudaDeviceSynchronize();
  }
  if (rc == 0) rc = (int)cudaMemcpy(cycles_host, d, ctas * sizeof(long long), cudaMemcpyDeviceToHost);
  cudaFree(src); cudaFree(d);
  return rc;
}

// bytes/cycle/SM of L2 -> smem bulk copies (after one untimed pass that brings `span` into L2)
extern "C" int l2_fill_probe(int ctas, int reps, int span, int shared_src, long long* cycles_host) {
  unsigned char* src = nullptr; long long* d = nullptr;
  const size_t total = shared_src ? (size_t)span : (size_t)span * ctas;
  if (cudaMalloc(&src, total) != cudaSuccess || cudaMalloc(&d, ctas * sizeof(long long)) != cudaSuccess) return 2;
  cudaMemset(src, 0, total);
  cudaFuncSetAttribute(fill_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 4 * 16384 + 1024);
  int rc = 0;
  for (int pass = 0; pass < 2 && rc == 0; pass++) {
    fill_kernel<<<ctas, 32, 4 * 16384 + 1024>>>(src, reps, span, shared_src, d);
    rc = (int)cudaDeviceSynchronize();
  }
  if (rc == 0) rc = (int)cudaMemcpy(cycles_host, d, ctas * sizeof(long long), cudaMemcpyDeviceToHost);
  cudaFree(src); cudaFree(d);
  return rc;
}

// cycles_host[ctas]; returns CUDA error code
extern "C" int umma_rate_probe(int N, int ctas, int reps, int shift_rows, int sbo_bytes, long long* cycles_host) {
  long long* d = nullptr;
  cudaError_t e = cudaMalloc(&d, ctas * sizeof(long long));
  if (e != cudaSuccess) return (int)e;
  int rc = N == 16 ? run<16>(ctas, reps, shift_rows, sbo_bytes, d) : N == 64 ? run<64>(ctas, reps, shift_rows, sbo_bytes, d)
         : N == 128 ? run<128>(ctas, reps, shift_rows, sbo_bytes, d) : N == 256 ? run<256>(ctas, reps, shift_rows, sbo_bytes, d) : 1;
  if (rc == 0) rc = (int)cudaMemcpy(cycles_host, d, ctas * sizeof(long long), cudaMemcpyDeviceToHost);
  cudaFree(d);
  return rc;
}
