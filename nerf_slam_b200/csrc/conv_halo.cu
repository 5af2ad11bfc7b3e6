// A5 — 3x3 / pad 1 / stride 1 convolutions of the update operator, second generation of the implicit GEMM
// (first generation: conv_igemm.cu, still used for 1x1 / 7x7-as-1x1 / encoder layers).
//
// What bounded conv_igemm.cu (profiles/r02_conv_*.md, ncu of r02 call 3):
//   * every 128-pixel tile re-streams the layer's whole weight set from L2 (q conv: 1.0 MB of weights + 0.42 MB of
//     activations per tile -> 6.5-7.4 TB/s of L2->SM traffic for the two GRU convolutions: L2-bandwidth bound);
//   * the activations of a tile are loaded three times (three column-shifted halo boxes per channel block).
// This kernel
//   * works on SUPER-TILES of 16 x 16 output pixels = two sub-tiles of 16 rows x 8 columns (M = 128 each) that share
//     every weight block: weight traffic per pixel halves (M = 256 per weight block, one CTA, no cluster needed);
//   * loads ONE halo box {64 channels, 18 w, 18 h} (324 rows of 128 B, SWIZZLE_128B) per channel block and super-tile and
//     reaches all nine taps of both sub-tiles through the START ADDRESS of the shared-memory descriptor:
//         row(hh, ww; dy, dx, s) = (hh + dy) * 18 + (ww + dx + 8 s)   ->  start = box + (dy * 18 + dx + 8 s) * 128 B,
//         8-row groups (one image row of the sub-tile) 18 rows = 2304 B apart (SBO).
//     A K-major SWIZZLE_128B operand may start at ANY 128-byte row: the tensor core applies the swizzle to absolute
//     shared-memory address bits, exactly like TMA wrote it (tools/probes/umma_row_shift_probe.cu, run on a B200:
//     all row shifts 0..22 with SBO 1024 / 1280 / 2048 / 2304 read the right rows with base_offset = 0).
//     Activation traffic per 128 pixels: 7 x 41.5 KB / 2 = 145 KB instead of 420 KB at 448 input channels;
//   * issues TMA / tcgen05 in uniform control flow (tc::umma_f16_lead);
//   * balances the tail: when the last wave of super-tiles would leave most SMs idle it is issued as half units
//     (one sub-tile per CTA).
// Epilogue modes 0 (ACT), 1 (ZR), 2 (Q) as in conv_igemm.cu (same epi_chunk, same staging / TMA store scheme).
// TMEM: 2 accumulator buffers x 2 sub-tiles x N columns for N <= 128; N = 256: one buffer (2 x 256 = 512 columns).
#include "conv_common.cuh"

namespace nslam {

constexpr int CH_TH = 16, CH_TW = 8;                 // sub-tile: 16 rows x 8 columns
constexpr int CH_BOX_W = 18, CH_BOX_H = 18;          // halo box of the 16 x 16 super-tile
constexpr int CH_A_BYTES = CH_BOX_W * CH_BOX_H * 128;   // 41472
constexpr int CH_A_STAGE = 41 * 1024;                // stage stride (1024-aligned for the swizzle pattern)
constexpr int CH_A_STAGES = 2;

template <int N>
struct ChSmem {
  static constexpr int W_BLOCK = N * 128;
  static constexpr int PASSES = (N >= 256) ? 2 : 1;
  static constexpr int NOUT64 = (N >= 64) ? (N / PASSES) / 64 : 1;
  static constexpr int OUT_BYTES = (N >= 64) ? NOUT64 * 16384 : 128 * N * 2;
  static constexpr int A = 0;
  static constexpr int W = CH_A_STAGES * CH_A_STAGE;
  static constexpr int BUDGET = 226 * 1024 - W - OUT_BYTES - 3 * N * 4 - 512;
  static constexpr int W_STAGES_RAW = BUDGET / W_BLOCK;
  static constexpr int W_STAGES = W_STAGES_RAW > 9 ? 9 : W_STAGES_RAW;
  static constexpr int OUT = W + W_STAGES * W_BLOCK;
  static constexpr int BIAS = OUT + OUT_BYTES;
  static constexpr int BAR = BIAS + 3 * N * 4;
  static constexpr int TOTAL = BAR + 512;
  static_assert(W_STAGES >= 3, "weight ring");
  static_assert((OUT % 1024) == 0, "staging tiles must keep the 1024-byte swizzle alignment");
};

struct HaloWork {
  int n_super, grid, full, rem, split;      // full waves, remainder super-tiles, remainder issued as half units?
  int n_items;
};

__device__ __forceinline__ void halo_item(const HaloWork& hw, int w, int& st, int& mask) {
  const int base = hw.full * hw.grid;
  if (w < base || !hw.split) { st = w; mask = 3; }
  else { st = base + ((w - base) >> 1); mask = 1 << ((w - base) & 1); }
}

template <int N, int MODE>
__global__ void __launch_bounds__(CG_THREADS_BASE, 1)
conv_halo_kernel(const __grid_constant__ ConvMaps maps, ConvParams p, HaloWork hw) {
  using SM = ChSmem<N>;
  constexpr int WS = SM::W_STAGES, AS = CH_A_STAGES;
  constexpr int NBUF = (N >= 256) ? 1 : 2;
  constexpr int TCOLS_RAW = NBUF * 2 * N;
  // power of two for N in {16,32,64,128,256}; the epilogue reads 32 columns at a time, also from the last (narrow)
  // sub-accumulator: (2 NBUF - 1) N + 32 <= TCOLS
  constexpr int TCOLS = TCOLS_RAW < 128 ? 128 : TCOLS_RAW;
  extern __shared__ unsigned char smem_raw[];
  unsigned char* sm = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t* bars = reinterpret_cast<uint64_t*>(sm + SM::BAR);
  uint64_t* full_w = bars;
  uint64_t* empty_w = bars + WS;
  uint64_t* full_a = bars + 2 * WS;
  uint64_t* empty_a = full_a + AS;
  uint64_t* tm_full = empty_a + AS;
  uint64_t* tm_empty = tm_full + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tm_empty + 2);
  float* sbias = reinterpret_cast<float*>(sm + SM::BIAS);
  static_assert((2 * 9 + 2 * AS + 4) * 8 + 8 <= 512, "barrier block");

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tiles_per_img = p.tiles_h * p.tiles_w;                      // SUPER-tiles per image

  if (warp == 0 && lane == 0) {
    for (int s = 0; s < p.n_src; s++) tc::tma_prefetch_desc(&maps.src[s]);
    for (int s = 0; s < WS; s++) { tc::mbar_init(&full_w[s], 1); tc::mbar_init(&empty_w[s], 1); }
    for (int s = 0; s < AS; s++) { tc::mbar_init(&full_a[s], 1); tc::mbar_init(&empty_a[s], 1); }
    for (int s = 0; s < 2; s++) { tc::mbar_init(&tm_full[s], 1); tc::mbar_init(&tm_empty[s], CG_EPI_WARPS); }
    tc::fence_barrier_init();
  }
  if (warp == 1) tc::tmem_alloc<TCOLS>(tmem_slot);
  for (int i = threadIdx.x; i < N; i += CG_THREADS_BASE) sbias[i] = p.bias ? p.bias[i] : 0.f;
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  // right sub-tiles that lie entirely outside the image (W % 16 in 1..8) are never computed
  auto sub_mask = [&](int st, int mask) {
    const int tt = st % tiles_per_img;
    const int w0 = (tt % p.tiles_w) * 16;
    return (w0 + 8 < p.W) ? mask : (mask & 1);
  };

  if (warp == 0) {
    // ===================== TMA producer (whole warp, elected lane issues) =====================
    const uint32_t lead = tc::elect_one() ? 1u : 0u;
    uint32_t ia = 0, iw = 0;
    auto src_of = [&](int g, int& sidx, int& cb) { sidx = 0; cb = g; while (cb >= p.src_cb[sidx]) { cb -= p.src_cb[sidx]; sidx++; } };
    auto load_a = [&](int g, int w0, int h0, int n) {
      int s, cb;
      src_of(g, s, cb);
      const int sa = ia % AS, pa = (ia / AS) & 1;
      tc::mbar_wait(&empty_a[sa], pa ^ 1);
      tc::mbar_arrive_expect_tx_lead(&full_a[sa], CH_A_BYTES, lead);
      tc::tma_load_4d_lead(sm + SM::A + sa * CH_A_STAGE, &maps.src[s], &full_a[sa], cb * 64, w0 - 1, h0 - 1, n, lead);
      ia++;
    };
    for (int w = blockIdx.x; w < hw.n_items; w += gridDim.x) {
      int st, mask;
      halo_item(hw, w, st, mask);
      if (sub_mask(st, mask) == 0) continue;
      const int n = st / tiles_per_img, tt = st % tiles_per_img;
      const int h0 = (tt / p.tiles_w) * 16, w0 = (tt % p.tiles_w) * 16;
      // The next channel block's box goes out AFTER the first weight blocks of the current one: its ring slot is only
      // free once the MMAs of block g-1 have retired, and a producer that waits for that before it streams block g's
      // weights leaves the weight ring empty at every channel-block boundary.  At tap A_AT the producer is at most WS
      // taps ahead of the MMA warp, so the ring stays full while it waits.
      constexpr int A_AT = WS >= 6 ? 3 : (WS >= 4 ? 2 : 1);
      load_a(0, w0, h0, n);
      for (int g = 0; g < p.cb_total; g++) {
        for (int tap = 0; tap < 9; tap++, iw++) {
          if (tap == A_AT && g + 1 < p.cb_total) load_a(g + 1, w0, h0, n);
          const int sw = iw % WS, pw = (iw / WS) & 1;
          tc::mbar_wait(&empty_w[sw], pw ^ 1);
          tc::mbar_arrive_expect_tx_lead(&full_w[sw], SM::W_BLOCK, lead);
          tc::bulk_copy_g2s_lead(sm + SM::W + sw * SM::W_BLOCK, p.wpacked + (size_t)(tap * p.cb_total + g) * N * 64,
                                 SM::W_BLOCK, &full_w[sw], lead);
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (whole warp, elected lane issues) =====================
    const uint32_t lead = tc::elect_one() ? 1u : 0u;
    constexpr uint32_t idesc = tc::umma_idesc_f16(128, N, 0);
    uint32_t ia = 0, iw = 0, tcount = 0;
    for (int w = blockIdx.x; w < hw.n_items; w += gridDim.x) {
      int st, mask;
      halo_item(hw, w, st, mask);
      mask = sub_mask(st, mask);
      if (mask == 0) continue;
      const int buf = (NBUF == 2) ? (tcount & 1) : 0;
      const int ph = (NBUF == 2) ? ((tcount >> 1) & 1) : (tcount & 1);
      tc::mbar_wait(&tm_empty[buf], ph ^ 1);
      tc::tc_fence_after();
      const uint32_t d0 = tmem_base + buf * (2 * N);
      for (int g = 0; g < p.cb_total; g++, ia++) {
        const int sa = ia % AS, pa = (ia / AS) & 1;
        tc::mbar_wait(&full_a[sa], pa);
        // descriptor of the box's first row with SBO = 18 rows (2304 B): image rows of a sub-tile are 18 box rows apart
        const uint64_t a_desc = tc::umma_desc_sw128_sbo(tc::smem_u32(sm + SM::A + sa * CH_A_STAGE), CH_BOX_W * 128);
#pragma unroll
        for (int tap = 0; tap < 9; tap++, iw++) {
          const int sw = iw % WS, pw = (iw / WS) & 1;
          tc::mbar_wait(&full_w[sw], pw);
          tc::tc_fence_after();
          const uint64_t w_desc = tc::umma_desc_sw128(tc::smem_u32(sm + SM::W + sw * SM::W_BLOCK));
          const int row0 = (tap / 3) * CH_BOX_W + (tap % 3);             // dy * 18 + dx
#pragma unroll
          for (int s = 0; s < 2; s++) {
            if (!((mask >> s) & 1)) continue;
#pragma unroll
            for (int k = 0; k < 4; k++)
              tc::umma_f16_lead(d0 + s * N, a_desc + (uint64_t)((row0 + 8 * s) * 8 + k * 2), w_desc + (uint64_t)(k * 2), idesc,
                                (g | tap | k) ? 1u : 0u, lead);
          }
          tc::umma_commit_lead(&empty_w[sw], lead);
        }
        tc::umma_commit_lead(&empty_a[sa], lead);
      }
      tc::umma_commit_lead(&tm_full[buf], lead);
      tcount++;
    }
  } else {
    // ===================== epilogue: 8 warps; warp pair (w, w+4) shares a TMEM lane quarter and splits the columns of a
    // pass in halves (N >= 64); thread = output pixel of the sub-tile =====================
    const int q = warp & 3;
    const int grp = (warp - 2) >> 2;
    const int row = q * 32 + lane;
    const int hh = row / CH_TW, ww = row % CH_TW;
    const int etid = threadIdx.x - 64;
    const bool works = (N >= 64) || grp == 0;
    uint32_t tcount = 0;
    for (int w = blockIdx.x; w < hw.n_items; w += gridDim.x) {
      int st, mask;
      halo_item(hw, w, st, mask);
      mask = sub_mask(st, mask);
      if (mask == 0) continue;
      const int n = st / tiles_per_img, tt = st % tiles_per_img;
      const int h0 = (tt / p.tiles_w) * 16, w0s = (tt % p.tiles_w) * 16;
      const int buf = (NBUF == 2) ? (tcount & 1) : 0;
      const int ph = (NBUF == 2) ? ((tcount >> 1) & 1) : (tcount & 1);
      tc::mbar_wait(&tm_full[buf], ph);
      tc::tc_fence_after();
      const float* g = p.gctx ? p.gctx + (size_t)n * N : nullptr;
      constexpr int PASSES = SM::PASSES, CPP = N / PASSES;
      constexpr int GC = (N >= 64) ? CPP / 2 : N;
#pragma unroll 1
      for (int s = 0; s < 2; s++) {
        if (!((mask >> s) & 1)) continue;
        const int w0 = w0s + 8 * s;
        const int h = h0 + hh, wpx = w0 + ww;
        const bool valid = (h < p.H) && (wpx < p.W);
        const size_t pix = ((size_t)n * p.H + h) * p.W + wpx;
        const uint32_t taddr = tmem_base + buf * (2 * N) + s * N + ((uint32_t)(q * 32) << 16);
#pragma unroll 1
        for (int pass = 0; pass < PASSES; pass++) {
          // the staging tiles are reused: wait until the previous TMA stores have read them
          if (etid == 0) tma_store_wait_read();
          asm volatile("bar.sync 1, 256;" ::: "memory");
          const int cbeg = pass * CPP + ((N >= 64) ? grp * GC : 0);
          if (works) {
#pragma unroll 1
            for (int c0 = cbeg; c0 < cbeg + GC; c0 += 32) {
              uint32_t r[32];
              tc::tmem_ld_32x32(taddr + c0, r);
              uint4 an[4] = {}, az[4] = {};
              if (MODE == 1) {
                if (c0 >= 128 && valid) {
                  const uint4* np = reinterpret_cast<const uint4*>(p.net + pix * 128 + (c0 - 128));
#pragma unroll
                  for (int i = 0; i < 4; i++) an[i] = np[i];
                }
              } else if (MODE == 2) {
                if (valid) {
                  const uint4* np = reinterpret_cast<const uint4*>(p.net + pix * 128 + c0);
                  const uint4* zp = reinterpret_cast<const uint4*>(p.zbuf + pix * 128 + c0);
#pragma unroll
                  for (int i = 0; i < 4; i++) { an[i] = np[i]; az[i] = zp[i]; }
                }
              }
              tc::tmem_ld_wait();
              float v[32];
              epi_chunk<MODE>(r, v, sbias, g, c0, p.act, valid, an, az);
              const int t64 = (c0 - pass * CPP) / 64;
#pragma unroll
              for (int i = 0; i < 32; i += 8) {
                if (c0 + i >= N) break;
                __half2 h2[4];
#pragma unroll
                for (int j = 0; j < 4; j++) h2[j] = __floats2half2_rn(v[i + 2 * j], v[i + 2 * j + 1]);
                if (N >= 64) {
                  unsigned char* stg = sm + SM::OUT + t64 * 16384 + row * 128;
                  const int chunk = ((c0 % 64) + i) / 8;
                  *reinterpret_cast<uint4*>(stg + ((chunk ^ (row & 7)) * 16)) = *reinterpret_cast<const uint4*>(h2);
                } else {
                  unsigned char* stg = sm + SM::OUT + row * (N * 2);
                  *reinterpret_cast<uint4*>(stg + (c0 + i) * 2) = *reinterpret_cast<const uint4*>(h2);
                }
              }
            }
          }
          tc::fence_proxy_async();
          asm volatile("bar.sync 2, 256;" ::: "memory");
          if (etid == 0) {
            if (MODE == 1) {
              // pass 0: z (columns 0..127) -> out0 ; pass 1: r * net (columns 128..255) -> out1
              tma_store_4d(&maps.out[pass], sm + SM::OUT + 0 * 16384, 0, w0, h0, n);
              tma_store_4d(&maps.out[pass], sm + SM::OUT + 1 * 16384, 64, w0, h0, n);
            } else {
              for (int t = 0; t < SM::NOUT64; t++)
                tma_store_4d(&maps.out[0], sm + SM::OUT + t * 16384, pass * CPP + t * 64, w0, h0, n);
            }
            tma_store_commit();
          }
        }
      }
      // both sub-accumulators of this buffer have been read
      tc::tc_fence_before();
      __syncwarp();
      if (lane == 0) tc::mbar_arrive(&tm_empty[buf]);
      tcount++;
    }
    if (etid == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  }
  tc::tc_fence_before();
  __syncthreads();
  if (warp == 1) tc::tmem_dealloc<TCOLS>(tmem_base);
}

template <int N, int MODE>
static int launch_halo_nm(const ConvMaps& maps, const ConvParams& p, int num_sms, cudaStream_t st) {
  const int smem = ChSmem<N>::TOTAL + 1024;
  static std::atomic<int> configured[NSLAM_MAX_DEVICES];
  int dev = 0;
  cudaGetDevice(&dev);
  dev = (dev >= 0 && dev < NSLAM_MAX_DEVICES) ? dev : 0;
  if (!configured[dev].load()) {
    cudaError_t e = cudaFuncSetAttribute(conv_halo_kernel<N, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return (int)e;
    configured[dev].store(1);
  }
  HaloWork hw;
  hw.n_super = p.B * p.tiles_h * p.tiles_w;
  hw.grid = hw.n_super < num_sms ? hw.n_super : num_sms;
  // tail: the last partial wave of super-tiles as half units when that fills more SMs
  int grid = num_sms;
  hw.full = hw.n_super / grid;
  hw.rem = hw.n_super % grid;
  hw.split = (hw.rem > 0 && 2 * hw.rem <= grid) ? 1 : 0;
  hw.grid = grid;
  hw.n_items = hw.full * grid + (hw.split ? 2 * hw.rem : hw.rem);
  const int launch = hw.n_items < grid ? hw.n_items : grid;
  conv_halo_kernel<N, MODE><<<launch, CG_THREADS_BASE, smem, st>>>(maps, p, hw);
  NSLAM_CHECK_LAUNCH();
  return 0;
}

bool conv_halo_supported(int N, int mode, int KH, int KW, int pad) {
  if (!(KH == 3 && KW == 3 && pad == 1)) return false;
  if (mode == 0) return N == 16 || N == 32 || N == 64 || N == 128 || N == 256;
  return (mode == 1 && N == 256) || (mode == 2 && N == 128);
}

int launch_conv_halo(int N, const ConvMaps& maps, const ConvParams& p, int num_sms, cudaStream_t st) {
  if (p.mode == 0) {
    switch (N) {
      case 16: return launch_halo_nm<16, 0>(maps, p, num_sms, st);
      case 32: return launch_halo_nm<32, 0>(maps, p, num_sms, st);
      case 64: return launch_halo_nm<64, 0>(maps, p, num_sms, st);
      case 128: return launch_halo_nm<128, 0>(maps, p, num_sms, st);
      case 256: return launch_halo_nm<256, 0>(maps, p, num_sms, st);
    }
  } else if (p.mode == 1 && N == 256) {
    return launch_halo_nm<256, 1>(maps, p, num_sms, st);
  } else if (p.mode == 2 && N == 128) {
    return launch_halo_nm<128, 2>(maps, p, num_sms, st);
  }
  return (int)cudaErrorInvalidValue;
}

}  // namespace nslam
