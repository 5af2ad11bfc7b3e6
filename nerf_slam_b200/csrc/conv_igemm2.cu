// A5 — 3x3 convolutions of the update operator with N >= 128 output channels on CTA PAIRS (tcgen05 cta_group::2).
//
// EXPERIMENTAL: written without hardware access at the end of round 1; selected only with NSLAM_CONV_CTA2=1
// (conv_igemm.cu::nslam_conv_igemm_ex), validated by tests/test_gpu_conv.py under the same variable.
//
// Why: with one CTA per tile (conv_igemm.cu) an M128 x N128 x K16 MMA reads A 4 KB + B 4 KB of shared memory in
// 64 tensor-pipe cycles = the 128 B/clk limit, and the weight blocks are re-streamed into shared memory for every
// tile on top of that (profiles/r01_ncu_*: N = 128 convolutions 34-36 % tensor pipe, N = 256 68 %).  A CTA pair
// executes ONE MMA of M = 256: each CTA holds its own pixel tile (A, 128 rows) and HALF of the weight block
// (N/2 rows); the halves are exchanged inside the TPC.  Per CTA: half the weight bytes loaded, half the B bytes read.
//
// Structure = conv_igemm_kernel<N, MODE, HALO = true> with these changes:
//   * cluster (2,1,1); work item = pair of consecutive tiles (2p, 2p+1), CTA rank r takes tile 2p + r
//     (an odd tile count leaves the last pair's second tile out of range: TMA zero-fills its loads and clips
//     its stores, the epilogue's global reads are guarded);
//   * all loads (pixel tiles: 4-D tensor maps; weights: a 2-D tensor map over the packed image, box {64, N/2},
//     no swizzle — the image is already the swizzled smem layout) carry .cta_group::2 and complete on the
//     LEADER's full barriers; the leader's producer posts the expected bytes of both CTAs;
//   * the leader's MMA thread issues tcgen05.mma.cta_group::2 (idesc M = 256) and multicasts its commits to the
//     empty / accumulator-full barriers of both CTAs; the epilogue warps of both CTAs release an accumulator
//     stage on the leader's barrier (16 arrivals);
//   * TMEM is allocated / freed by warp 1 of both CTAs (one collective allocation), teardown behind a cluster barrier.
// Epilogue modes 0 (ACT), 1 (ZR), 2 (Q) as in conv_igemm.cu.
#include "conv_common.cuh"

#include <cstdlib>

namespace nslam {

template <int N>
struct Cg2Smem {
  static constexpr int A_STAGE = 20480;                      // one column-shifted halo tile {64c,16w,10h}
  static constexpr int A_STAGES = 3;
  static constexpr int B_STAGE = (N / 2) * 128;              // this CTA's half of a weight block
  static constexpr int B_STAGES = (N >= 256) ? 6 : 12;
  static constexpr int A = 0;
  static constexpr int B = A_STAGES * A_STAGE;
  static constexpr int OUT = B + B_STAGES * B_STAGE;
  static constexpr int PASSES = (N >= 256) ? 2 : 1;
  static constexpr int NOUT64 = (N / PASSES) / 64;
  static constexpr int BIAS = OUT + NOUT64 * 16384;
  static constexpr int BAR = BIAS + N * 4;
  static constexpr int TOTAL = BAR + 512;
};

constexpr int CG2_W_WARPS = 3;
constexpr int CG2_THREADS = CG_THREADS_BASE + 32 * CG2_W_WARPS;

template <int N, int MODE>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(CG2_THREADS, 1)
conv_igemm2_kernel(const __grid_constant__ ConvMaps maps, const __grid_constant__ CUtensorMap wmap, ConvParams p) {
  static_assert(N == 128 || N == 256, "CTA-pair kernel: N = 128 or 256");
  static_assert(MODE == 0 || MODE == 1 || MODE == 2, "CTA-pair kernel: epilogue modes 0, 1, 2");
  using SM = Cg2Smem<N>;
  constexpr int AS = SM::A_STAGES, BS = SM::B_STAGES;
  constexpr int TCOLS = (N <= 128) ? 256 : 512;              // 2 accumulator stages
  constexpr int ACC_STRIDE = TCOLS / 2;
  extern __shared__ unsigned char smem_raw[];
  unsigned char* sm = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t* bars = reinterpret_cast<uint64_t*>(sm + SM::BAR);
  uint64_t* full_b = bars;                       // used in the leader CTA only
  uint64_t* empty_b = bars + BS;                 // one per CTA, released together by multicast commits
  uint64_t* full_a = bars + 2 * BS;              // leader only
  uint64_t* empty_a = full_a + AS;
  uint64_t* tm_full = empty_a + AS;              // one per CTA (multicast)
  uint64_t* tm_empty = tm_full + 2;              // leader only: epilogue warps of both CTAs arrive
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tm_empty + 2);
  float* sbias = reinterpret_cast<float*>(sm + SM::BIAS);
  static_assert((2 * BS + 2 * AS + 4) * 8 + 8 <= 512, "barrier block");

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int rank = (int)tc::cluster_ctarank();
  const bool leader = rank == 0;
  const int cid = blockIdx.x >> 1, nclusters = gridDim.x >> 1;
  const int tiles_per_img = p.tiles_h * p.tiles_w;
  const int ntiles = p.B * tiles_per_img;
  const int npairs = (ntiles + 1) >> 1;
  const int units = p.cb_total * 3;              // (channel block, dx)

  if (warp == 0 && lane == 0) {
    for (int s = 0; s < p.n_src; s++) tc::tma_prefetch_desc(&maps.src[s]);
    tc::tma_prefetch_desc(&wmap);
    for (int s = 0; s < BS; s++) { tc::mbar_init(&full_b[s], 1); tc::mbar_init(&empty_b[s], 1); }
    for (int s = 0; s < AS; s++) { tc::mbar_init(&full_a[s], 1); tc::mbar_init(&empty_a[s], 1); }
    for (int s = 0; s < 2; s++) { tc::mbar_init(&tm_full[s], 1); tc::mbar_init(&tm_empty[s], 2 * CG_EPI_WARPS); }
    tc::fence_barrier_init();
  }
  if (warp == 1) tc::tmem_alloc_pair<TCOLS>(tmem_slot);
  for (int i = threadIdx.x; i < N; i += CG2_THREADS) sbias[i] = p.bias ? p.bias[i] : 0.f;
  tc::tc_fence_before();
  tc::cluster_sync_all();                        // barriers of BOTH CTAs initialised before any remote signal
  tc::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  // warp roles: 0 = activation (A) producer, 10..12 = weight (B) producers, 1 = MMA issuer (leader), 2..9 = epilogue.
  // All loads of this kernel are tensor-map loads (~500-630 cycles of issue each per warp, whatever their size;
  // profiles/r02_tma_multi_warp_probe_call9.log): 3 + 9 of them per channel block from ONE warp cost 6600 cycles against
  // 2304 (N = 128) / 4608 (N = 256) cycles of MMAs — the reason the first version of this kernel was no faster than the
  // one-CTA kernel.  Four issuing warps: 3 loads per channel block each.
  const int w_part = (warp >= 2 + CG_EPI_WARPS) ? warp - (2 + CG_EPI_WARPS) : -1;
  if (warp == 0 || w_part >= 0) {
    // ===================== TMA producers (both CTAs; whole warp in the loop, elected lane issues) =====================
    const uint32_t lead = tc::elect_one() ? 1u : 0u;
    const uint32_t lead_leader = (lead && leader) ? 1u : 0u;
    auto src_of = [&](int g, int& sidx, int& cb) { sidx = 0; cb = g; while (cb >= p.src_cb[sidx]) { cb -= p.src_cb[sidx]; sidx++; } };
    uint32_t ia = 0, ib = 0;
    for (int pair = cid; pair < npairs; pair += nclusters) {
      const int tile = 2 * pair + rank;        // == ntiles for the missing half of the last pair: image index B, zero-filled
      const int n = tile / tiles_per_img, tt = tile % tiles_per_img;
      const int h0 = (tt / p.tiles_w) * CG_TH, w0 = (tt % p.tiles_w) * CG_TW;
      if (warp == 0) {
        for (int u = 0; u < units; u++, ia++) {
          int s, cb;
          src_of(u / 3, s, cb);
          const int sa = ia % AS, pa = (ia / AS) & 1;
          tc::mbar_wait(&empty_a[sa], pa ^ 1);
          tc::mbar_arrive_expect_tx_lead(&full_a[sa], 2 * SM::A_STAGE, lead_leader);
          tc::tma_load_4d_pair_lead(sm + SM::A + sa * SM::A_STAGE, &maps.src[s], &full_a[sa], cb * 64, w0 + (u % 3) - 1, h0 - 1, n, lead);
        }
      } else {
        for (int u = 0; u < units; u++) {
          const int cbg = u / 3, dx = u % 3;
          for (int dy = 0; dy < 3; dy++, ib++) {
            if ((int)(ib % CG2_W_WARPS) != w_part) continue;
            const int sb = ib % BS, pb = (ib / BS) & 1;
            tc::mbar_wait(&empty_b[sb], pb ^ 1);
            tc::mbar_arrive_expect_tx_lead(&full_b[sb], N * 128, lead_leader);
            const int blk = (dy * 3 + dx) * p.cb_total + cbg;
            tc::tma_load_2d_pair_lead(sm + SM::B + sb * SM::B_STAGE, &wmap, &full_b[sb], 0, blk * N + rank * (N / 2), lead);
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (leader CTA only) =====================
    // whole warp in the loop (uniform control flow), one elected lane issues: see tc::umma_f16_lead
    if (leader) {
      const uint32_t lead = tc::elect_one() ? 1u : 0u;
      constexpr uint32_t idesc = tc::umma_idesc_f16(256, N, 0);
      uint32_t it = 0, ia = 0, tcount = 0;
      for (int pair = cid; pair < npairs; pair += nclusters, tcount++) {
        const int as = tcount & 1, aph = (tcount >> 1) & 1;
        tc::mbar_wait(&tm_empty[as], aph ^ 1);
        tc::tc_fence_after();
        const uint32_t d_tmem = tmem_base + as * ACC_STRIDE;
        for (int u = 0; u < units; u++, ia++) {
          const int sa = ia % AS, pa = (ia / AS) & 1;
          tc::mbar_wait(&full_a[sa], pa);
          const uint64_t a_desc = tc::umma_desc_sw128(tc::smem_u32(sm + SM::A + sa * SM::A_STAGE));
#pragma unroll
          for (int dy = 0; dy < 3; dy++, it++) {
            const int sb = it % BS, pb = (it / BS) & 1;
            tc::mbar_wait(&full_b[sb], pb);
            tc::tc_fence_after();
            const uint64_t b_desc = tc::umma_desc_sw128(tc::smem_u32(sm + SM::B + sb * SM::B_STAGE));
#pragma unroll
            for (int k = 0; k < 4; k++)
              tc::umma_f16_pair_lead(d_tmem, a_desc + (uint64_t)(dy * (CG_TW * 128 / 16) + k * 2), b_desc + (uint64_t)(k * 2), idesc,
                                     (u | dy | k) ? 1u : 0u, lead);
            tc::umma_commit_pair_lead(&empty_b[sb], lead);
          }
          tc::umma_commit_pair_lead(&empty_a[sa], lead);
        }
        tc::umma_commit_pair_lead(&tm_full[as], lead);
      }
    }
  } else {
    // ===================== epilogue (both CTAs, own tile): 8 warps; warp pair (w, w+4) shares a TMEM lane
    // quarter and splits the columns of a pass in halves; thread = output pixel =====================
    const int q = warp & 3;
    const int grp = (warp - 2) >> 2;
    const int row = q * 32 + lane;
    const int hh = row / CG_TW, ww = row % CG_TW;
    const int etid = threadIdx.x - 64;
    uint32_t tcount = 0;
    for (int pair = cid; pair < npairs; pair += nclusters, tcount++) {
      const int tile = 2 * pair + rank;
      const bool tile_ok = tile < ntiles;
      const int n = tile / tiles_per_img, tt = tile % tiles_per_img;
      const int h0 = (tt / p.tiles_w) * CG_TH, w0 = (tt % p.tiles_w) * CG_TW;
      const int h = h0 + hh, w = w0 + ww;
      const bool valid = tile_ok && (h < p.H) && (w < p.W);
      const size_t pix = ((size_t)n * p.H + h) * p.W + w;
      const int as = tcount & 1, aph = (tcount >> 1) & 1;
      tc::mbar_wait(&tm_full[as], aph);
      tc::tc_fence_after();
      if (etid == 0) tma_store_wait_read();      // staging tiles free again (previous tile's TMA stores have read them)
      asm volatile("bar.sync 1, 256;" ::: "memory");
      const uint32_t taddr = tmem_base + as * ACC_STRIDE + ((uint32_t)(q * 32) << 16);
      const float* g = (p.gctx && tile_ok) ? p.gctx + (size_t)n * N : nullptr;
      constexpr int PASSES = SM::PASSES, CPP = N / PASSES, GC = CPP / 2;
#pragma unroll 1
      for (int pass = 0; pass < PASSES; pass++) {
        if (pass > 0) {
          if (etid == 0) tma_store_wait_read();
          asm volatile("bar.sync 1, 256;" ::: "memory");
        }
        const int cbeg = pass * CPP + grp * GC;
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
          unsigned char* st = sm + SM::OUT + t64 * 16384 + row * 128;
#pragma unroll
          for (int i = 0; i < 32; i += 8) {
            __half2 h2[4];
#pragma unroll
            for (int j = 0; j < 4; j++) h2[j] = __floats2half2_rn(v[i + 2 * j], v[i + 2 * j + 1]);
            const int chunk = ((c0 % 64) + i) / 8;
            *reinterpret_cast<uint4*>(st + ((chunk ^ (row & 7)) * 16)) = *reinterpret_cast<const uint4*>(h2);
          }
        }
        if (pass == PASSES - 1) {
          // this CTA's half of the accumulator stage has been read: tell the leader's MMA thread
          tc::tc_fence_before();
          __syncwarp();
          if (lane == 0) tc::mbar_arrive_leader(&tm_empty[as]);
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
    if (etid == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  }
  // neither CTA may exit (or free TMEM) while its partner can still read its shared memory or signal its barriers
  tc::tc_fence_before();
  tc::cluster_sync_all();
  if (warp == 1) tc::tmem_dealloc_pair<TCOLS>(tmem_base);
}

template <int N, int MODE>
static int launch_conv2_nm(const ConvMaps& maps, const CUtensorMap& wmap, const ConvParams& p, int num_sms, cudaStream_t st) {
  const int smem = Cg2Smem<N>::TOTAL + 1024;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(conv_igemm2_kernel<N, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return (int)e;
    configured = true;
  }
  const int ntiles = p.B * p.tiles_h * p.tiles_w;
  const int npairs = (ntiles + 1) / 2;
  int clusters = num_sms / 2;
  if (clusters < 1) clusters = 1;
  if (clusters > npairs) clusters = npairs;
  conv_igemm2_kernel<N, MODE><<<2 * clusters, CG2_THREADS, smem, st>>>(maps, wmap, p);
  NSLAM_CHECK_LAUNCH();
  return 0;
}

bool conv_pairs_enabled() {
  static const bool on = [] { const char* e = std::getenv("NSLAM_CONV_CTA2"); return e && e[0] == '1'; }();
  return on;
}

bool conv_pairs_supported(int N, int mode, bool halo) {
  return halo && ((mode == 0 && (N == 128 || N == 256)) || (mode == 1 && N == 256) || (mode == 2 && N == 128));
}

int launch_conv_pairs(int N, const ConvMaps& maps, const ConvParams& p, int num_sms, cudaStream_t st) {
  // the packed weights [taps * cb_total][N][64] as a 2-D tensor of 128-byte rows; each CTA fetches N/2 rows of a block
  CUtensorMap wmap;
  const uint64_t rows = (uint64_t)p.KH * p.KW * p.cb_total * N;
  uint64_t dims[2] = {64, rows};
  uint64_t strides[1] = {128};
  uint32_t box[2] = {64, (uint32_t)(N / 2)};
  int r = tc::make_tmap_f16(&wmap, p.wpacked, 2, dims, strides, box, false, nullptr, /*swizzle128=*/false);
  if (r) return r;
  if (p.mode == 0 && N == 128) return launch_conv2_nm<128, 0>(maps, wmap, p, num_sms, st);
  if (p.mode == 0 && N == 256) return launch_conv2_nm<256, 0>(maps, wmap, p, num_sms, st);
  if (p.mode == 1 && N == 256) return launch_conv2_nm<256, 1>(maps, wmap, p, num_sms, st);
  if (p.mode == 2 && N == 128) return launch_conv2_nm<128, 2>(maps, wmap, p, num_sms, st);
  return (int)cudaErrorInvalidValue;
}

}  // namespace nslam
