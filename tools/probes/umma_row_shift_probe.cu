// Probe (hardware question for the next convolution redesign): can a K-major, 128B-swizzled UMMA operand START at a
// row that is not a multiple of 8 (start address + s * 128 B) — and does the shared-memory descriptor's
// "matrix base offset" field (bits [49,52)) have to carry s for the swizzle phase to stay right?
//
// If it works, a 3x3 convolution tile of 8 px x 16 rows can read all nine taps from ONE halo box {64c, 16w, 18h}
// (36 KB) by moving the descriptor start (dy * 16 + dx) rows, instead of three column-shifted boxes (60 KB).
//
// Layout written here = what TMA SWIZZLE_128B produces: row r (128 B) at r * 128, its 16-byte chunk j stored at
// chunk position j ^ (r & 7).  Chunk j of row r holds the fp16 values (r, j, 0, ...).  B selects K columns 0 and 8:
// D[m][0] = first value of chunk 0 of the row the MMA took for m, D[m][1] = second value (chunk id, expected 0),
// D[m][2], D[m][3] = the same for chunk 1 (expected row, 1).
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -I nerf_slam_b200/csrc -shared -Xcompiler -fPIC \
//        tools/probes/umma_row_shift_probe.cu -o /tmp/umma_probe.so        (driver: tools/probes/run_umma_probe.py)
#include <cuda_fp16.h>
#include "tc.cuh"

namespace {

__device__ __forceinline__ uint64_t desc_sw128(uint32_t addr, uint32_t sbo_bytes, uint32_t base_offset) {
  uint64_t d = 0;
  d |= (uint64_t)((addr & 0x3FFFF) >> 4);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(sbo_bytes >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)(base_offset & 7) << 49;
  d |= (uint64_t)2 << 61;
  return d;
}

__global__ void __launch_bounds__(128, 1) probe_kernel(int shift_rows, int base_offset, int sbo_bytes, float* out) {
  extern __shared__ unsigned char raw[];
  unsigned char* sm = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~(uintptr_t)1023);
  __half* A = reinterpret_cast<__half*>(sm);                   // 512 rows x 128 B = 64 KB
  __half* B = reinterpret_cast<__half*>(sm + 65536);           // 16 rows x 128 B (one and two 8-row groups)
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < 512 * 8; i += 128) {                   // (row, logical chunk)
    const int r = i >> 3, j = i & 7;
    __half* dst = A + r * 64 + ((j ^ (r & 7)) * 8);
    for (int e = 0; e < 8; e++) dst[e] = __float2half(0.f);
    dst[0] = __float2half((float)r);
    dst[1] = __float2half((float)j);
  }
  for (int i = tid; i < 16 * 64; i += 128) B[i] = __float2half(0.f);
  __syncthreads();
  if (tid == 0) {
    // B[n][k] (row n = output column): n=0 -> k=0, n=1 -> k=1, n=2 -> k=8, n=3 -> k=9 ; swizzled like A
    auto put = [&](int n, int k) { B[n * 64 + (((k >> 3) ^ (n & 7)) * 8) + (k & 7)] = __float2half(1.f); };
    put(0, 0); put(1, 1); put(2, 8); put(3, 9);
    tc::mbar_init(&bar, 1);
    tc::fence_barrier_init();
  }
  if (warp == 0) tc::tmem_alloc<32>(&slot);
  tc::fence_proxy_async();
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem = slot;
  if (tid == 0) {
    const uint32_t a_addr = tc::smem_u32(A) + shift_rows * 128;
    const uint32_t b_addr = tc::smem_u32(B);
    constexpr uint32_t idesc = tc::umma_idesc_f16(128, 16, 0);
    tc::umma_f16(tmem, desc_sw128(a_addr, sbo_bytes, base_offset), desc_sw128(b_addr, 1024, 0), idesc, 0u);
    tc::umma_commit(&bar);
  }
  tc::mbar_wait(&bar, 0);
  tc::tc_fence_after();
  uint32_t r[8];
  tc::tmem_ld_32x8(tmem + ((uint32_t)(warp * 32) << 16), r);
  tc::tmem_ld_wait();
  for (int i = 0; i < 4; i++) out[tid * 4 + i] = __uint_as_float(r[i]);
  tc::tc_fence_before();
  __syncthreads();
  if (warp == 0) tc::tmem_dealloc<32>(tmem);
}

}  // namespace

// out_host: 128 x 4 floats (row read for chunk 0, chunk id, row read for chunk 1, chunk id)
extern "C" int umma_row_shift_probe(int shift_rows, int base_offset, int sbo_bytes, float* out_host) {
  float* d = nullptr;
  cudaError_t e = cudaMalloc(&d, 128 * 4 * sizeof(float));
  if (e != cudaSuccess) return (int)e;
  const int smem = 65536 + 2048 + 1024;
  e = cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  if (e != cudaSuccess) return (int)e;
  probe_kernel<<<1, 128, smem>>>(shift_rows, base_offset, sbo_bytes, d);
  e = cudaDeviceSynchronize();
  if (e == cudaSuccess) e = cudaMemcpy(out_host, d, 128 * 4 * sizeof(float), cudaMemcpyDeviceToHost);
  cudaFree(d);
  return (int)e;
}
