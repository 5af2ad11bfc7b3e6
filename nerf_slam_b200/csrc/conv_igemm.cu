// A5/A1 — convolutions of the update operator as implicit GEMM on tcgen05 (sm_100a).
//
// Replaces the cuDNN convolutions behind UpdateModule / ConvGRU / GraphAgg
// (reference networks/droid_net.py:78-150, networks/modules/gru.py:5-32) and the
// torch.cat / sigmoid / tanh / gating elementwise kernels around them.
//
//   out[n,h,w,:] = epilogue( sum_{taps, sources} A_src[n, h+dy, w+dx, :] . W[tap, src] + bias )
//
// * activations are NHWC fp16; the concatenated conv input (hidden | context | corr | flow, 448
//   channels for the GRU) is never materialised: every 64-channel K-block is fetched by TMA from
//   the SOURCE tensor it belongs to (up to 4 tensor maps);
// * one CTA tile = 8x16 output pixels (M = 128) x N output channels (N = 16..256); for tap
//   (dy,dx) the A tile is the TMA box {64c, 16w, 8h} shifted by (dx-pad, dy-pad): out-of-image
//   rows/cols are zero-filled by TMA, which IS the convolution's zero padding;
// * weights are pre-packed on the host into the exact 128B-swizzled K-major smem image, one
//   contiguous N*128-byte block per (tap, channel-block): a 1-D bulk copy lands it UMMA-ready;
// * 4-stage mbarrier ring, one thread issues tcgen05.mma (M128, N, K16), fp32 accumulators in
//   TMEM, double buffered so the epilogue of tile t overlaps the MMAs of tile t+1;
// * persistent CTAs (grid = #SMs) walk the (image, tile) list;
// * epilogue (8 warps, thread = pixel, the two warps of a TMEM lane quarter split the columns):
//   tcgen05.ld -> +bias (+ per-image global-context vector) -> activation / GRU gating -> fp16 ->
//   swizzled smem -> TMA store (clips partial tiles).  Kernel templated on the epilogue mode.
//
// Epilogue modes:
//   0 ACT   y = act(acc + bias [+ gctx[n]])                          act: 0 none 1 relu 2 sigmoid 3 tanh
//   1 ZR    N = 256: z = sigmoid(acc[0:128]+..), r = sigmoid(acc[128:256]+..);
//           out0 = z, out1 = r * net          (ConvGRU z,r gates, gru.py:27-29)
//   2 Q     q = tanh(acc + ..); out0 = (1 - z) * net + z * q         (gru.py:29-31)
//   3 GLO   y = sigmoid(acc + bias) * net; per-tile column sums (warp transpose-reduce -> shared
//           accumulator) added to gsum[n][c] (fp32)  -> glo = mean (gru.py:23-25); nothing is stored
//   4 ENC   encoder layers (BasicEncoder, extractor.py): y = act(acc + bias) like mode 0, plus (a) optional
//           per-(image, channel) sum / sum of squares of the fp16 outputs for the instance norm that follows
//           and (b) optional stride-2 output: the conv is evaluated at stride 1 and only even rows/columns
//           are stored and counted (a 3x3/s2 conv reads every input pixel anyway).
// Roofline: tensor pipe; FLOPs = 2 * pixels * taps * Cin * Cout.
#include <cstdlib>
#include "conv_common.cuh"

namespace nslam {
constexpr uint32_t CH2_BOX = 18;        // halo box edge of conv_halo.cu (16 + 2)

template <int N, int MODE, bool HALO>
__global__ void __launch_bounds__(CG_THREADS, 1)
conv_igemm_kernel(const __grid_constant__ ConvMaps maps, ConvParams p) {
  using SM = CgSmem<N, HALO>;
  constexpr int CG_STAGES = SM::STAGES;
  constexpr int AS = SM::A_STAGES, BS = SM::B_STAGES;
  constexpr int TCOLS = (N <= 32) ? 64 : (N <= 64 ? 128 : (N <= 128 ? 256 : 512));   // 2 accumulator stages
  constexpr int ACC_STRIDE = TCOLS / 2;
  extern __shared__ unsigned char smem_raw[];
  unsigned char* sm = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t* bars = reinterpret_cast<uint64_t*>(sm + SM::BAR);
  uint64_t* full_b = bars;                       // plain: one ring; HALO: the weight-block ring
  uint64_t* empty_b = bars + BS;
  uint64_t* full_a = bars + 2 * BS;              // HALO only: the shifted-tile ring
  uint64_t* empty_a = full_a + AS;
  uint64_t* tm_full = empty_a + AS;
  uint64_t* tm_empty = tm_full + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tm_empty + 2);
  float* sbias = reinterpret_cast<float*>(sm + SM::BIAS);
  float* sacc = sbias + N;                       // mode 3: per-CTA column sums of the current tile
  static_assert((4 * 4 + 4) * 8 + 8 <= 256, "barrier block");

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tiles_per_img = p.tiles_h * p.tiles_w;
  const int ntiles = p.B * tiles_per_img;
  const int taps = p.KH * p.KW;
  const int nkb = taps * p.cb_total;

  if (warp == 0 && lane == 0) {
    for (int s = 0; s < p.n_src; s++) tc::tma_prefetch_desc(&maps.src[s]);
    for (int s = 0; s < BS; s++) { tc::mbar_init(&full_b[s], HALO ? 1 : 2); tc::mbar_init(&empty_b[s], 1); }
    for (int s = 0; s < AS; s++) { tc::mbar_init(&full_a[s], 1); tc::mbar_init(&empty_a[s], 1); }
    for (int s = 0; s < 2; s++) { tc::mbar_init(&tm_full[s], 1); tc::mbar_init(&tm_empty[s], CG_EPI_WARPS); }
    tc::fence_barrier_init();
  }
  if (warp == 1) tc::tmem_alloc<TCOLS>(tmem_slot);
  for (int i = threadIdx.x; i < N; i += CG_THREADS) { sbias[i] = p.bias ? p.bias[i] : 0.f; sacc[i] = 0.f; sacc[N + i] = 0.f; }
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  // warp roles: 0 and 10 = activation (A) producers, 11 and 12 = weight (B) producers, 1 = MMA issuer, 2..9 = epilogue
  const int a_part = (warp == 0) ? 0 : (warp == 2 + CG_EPI_WARPS) ? 1 : -1;
  const int w_part = (warp >= 3 + CG_EPI_WARPS) ? warp - (3 + CG_EPI_WARPS) : -1;
  if (a_part >= 0 || w_part >= 0) {
    // ===================== TMA producers =====================
    // Every producer warp walks the same sequence of load units and issues those of its own parity: an A warp the
    // activation loads with (counter % CG_A_WARPS) == a_part, a W warp the weight blocks with (counter % CG_W_WARPS) ==
    // w_part.  Whole warp in the loop, elected lane issues (uniform control flow, see tc::umma_f16_lead).
    const uint32_t lead = tc::elect_one() ? 1u : 0u;
    auto src_of = [&](int g, int& sidx, int& cb) { sidx = 0; cb = g; while (cb >= p.src_cb[sidx]) { cb -= p.src_cb[sidx]; sidx++; } };
    if (HALO) {
      // unit = (channel block, dx): one shifted halo tile (A ring) + the three weight blocks of its taps dy = 0..2 (B ring)
      const int units = p.cb_total * 3;
      uint32_t ia = 0, ib = 0;
      for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int n = tile / tiles_per_img, tt = tile % tiles_per_img;
        const int h0 = (tt / p.tiles_w) * CG_TH, w0 = (tt % p.tiles_w) * CG_TW;
        if (a_part >= 0) {
          for (int u = 0; u < units; u++, ia++) {
            if ((int)(ia % CG_A_WARPS) != a_part) continue;
            int s, cb;
            src_of(u / 3, s, cb);
            const int sa = ia % AS, pa = (ia / AS) & 1;
            tc::mbar_wait(&empty_a[sa], pa ^ 1);
            tc::mbar_arrive_expect_tx_lead(&full_a[sa], SM::A_STAGE, lead);
            tc::tma_load_4d_lead(sm + SM::A + sa * SM::A_STAGE, &maps.src[s], &full_a[sa], cb * 64, w0 + (u % 3) - 1, h0 - 1, n, lead);
          }
        } else {
          for (int u = 0; u < units; u++) {
            const int cbg = u / 3, dx = u % 3;
            for (int dy = 0; dy < 3; dy++, ib++) {
              if ((int)(ib % CG_W_WARPS) != w_part) continue;
              const int sb = ib % BS, pb = (ib / BS) & 1;
              tc::mbar_wait(&empty_b[sb], pb ^ 1);
              tc::mbar_arrive_expect_tx_lead(&full_b[sb], N * 128, lead);
              tc::bulk_copy_g2s_lead(sm + SM::B + sb * (N * 128), p.wpacked + (size_t)((dy * 3 + dx) * p.cb_total + cbg) * N * 64,
                                     N * 128, &full_b[sb], lead);
            }
          }
        }
      }
    } else {
      // unit = (tap, channel block): one activation tile and one weight block complete on the SAME barrier (init count 2:
      // one arrive.expect_tx from the A warp, one from the W warp)
      uint32_t it = 0;
      for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int n = tile / tiles_per_img, tt = tile % tiles_per_img;
        const int h0 = (tt / p.tiles_w) * CG_TH, w0 = (tt % p.tiles_w) * CG_TW;
        for (int tap = 0; tap < taps; tap++) {
          const int dy = tap / p.KW - p.pad, dx = tap % p.KW - p.pad;
          int cbg = 0;
          for (int s = 0; s < p.n_src; s++) {
            for (int cb = 0; cb < p.src_cb[s]; cb++, cbg++, it++) {
              const bool mine = (a_part >= 0) ? ((int)(it % CG_A_WARPS) == a_part) : ((int)(it % CG_W_WARPS) == w_part);
              if (!mine) continue;
              const int st = it % CG_STAGES, ph = (it / CG_STAGES) & 1;
              tc::mbar_wait(&empty_b[st], ph ^ 1);
              if (a_part >= 0) {
                tc::mbar_arrive_expect_tx_lead(&full_b[st], 16384, lead);
                tc::tma_load_4d_lead(sm + SM::A + st * 16384, &maps.src[s], &full_b[st], cb * 64, w0 + dx, h0 + dy, n, lead);
              } else {
                tc::mbar_arrive_expect_tx_lead(&full_b[st], N * 128, lead);
                tc::bulk_copy_g2s_lead(sm + SM::B + st * (N * 128), p.wpacked + (size_t)(tap * p.cb_total + cbg) * N * 64,
                                       N * 128, &full_b[st], lead);
              }
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    // The WHOLE warp runs the loop (uniform control flow: loop counters, barrier addresses and descriptors live in
    // uniform registers); only the elected lane issues tcgen05.mma / tcgen05.commit (tc::umma_f16_lead).
    {
      const uint32_t lead = tc::elect_one() ? 1u : 0u;
      constexpr uint32_t idesc = tc::umma_idesc_f16(128, N, 0);
      uint32_t it = 0, ia = 0, tcount = 0;
      for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, tcount++) {
        const int as = tcount & 1, aph = (tcount >> 1) & 1;
        tc::mbar_wait(&tm_empty[as], aph ^ 1);
        const uint32_t d_tmem = tmem_base + as * ACC_STRIDE;
        if (HALO) {
          const int units = p.cb_total * 3;
          for (int u = 0; u < units; u++, ia++) {
            const int sa = ia % AS, pa = (ia / AS) & 1;
            tc::mbar_wait(&full_a[sa], pa);
            // descriptor of the tile's first row; +dy * 16 rows (2048 B) and +k * 32 B are added in descriptor units (16 B)
            const uint64_t a_desc = tc::umma_desc_sw128(tc::smem_u32(sm + SM::A + sa * SM::A_STAGE));
#pragma unroll
            for (int dy = 0; dy < 3; dy++, it++) {
              const int sb = it % BS, pb = (it / BS) & 1;
              tc::mbar_wait(&full_b[sb], pb);
              tc::tc_fence_after();
              const uint64_t b_desc = tc::umma_desc_sw128(tc::smem_u32(sm + SM::B + sb * (N * 128)));
#pragma unroll
              for (int k = 0; k < 4; k++)
                tc::umma_f16_lead(d_tmem, a_desc + (uint64_t)(dy * (CG_TW * 128 / 16) + k * 2), b_desc + (uint64_t)(k * 2), idesc,
                                  (u | dy | k) ? 1u : 0u, lead);
              tc::umma_commit_lead(&empty_b[sb], lead);
            }
            tc::umma_commit_lead(&empty_a[sa], lead);
          }
        } else {
          for (int kb = 0; kb < nkb; kb++, it++) {
            const int st = it % CG_STAGES, ph = (it / CG_STAGES) & 1;
            tc::mbar_wait(&full_b[st], ph);
            tc::tc_fence_after();
            const uint64_t a_desc = tc::umma_desc_sw128(tc::smem_u32(sm + SM::A + st * 16384));
            const uint64_t b_desc = tc::umma_desc_sw128(tc::smem_u32(sm + SM::B + st * (N * 128)));
#pragma unroll
            for (int k = 0; k < 4; k++)
              tc::umma_f16_lead(d_tmem, a_desc + (uint64_t)(k * 2), b_desc + (uint64_t)(k * 2), idesc, (kb | k) ? 1u : 0u, lead);
            tc::umma_commit_lead(&empty_b[st], lead);
          }
        }
        tc::umma_commit_lead(&tm_full[as], lead);
      }
    }
  } else {
    // ===================== epilogue: 8 warps; warp pair (w, w+4) shares a TMEM lane quarter and
    // splits the N columns in halves (N >= 64), thread = output pixel =====================
    const int q = warp & 3;
    const int grp = (warp - 2) >> 2;
    const int row = q * 32 + lane;
    const int hh = row / CG_TW, ww = row % CG_TW;
    const int etid = threadIdx.x - 64;
    constexpr int NG = (N >= 64) ? N / 2 : N;       // columns per group
    const int cbeg = (N >= 64) ? grp * NG : 0;
    const bool works = (N >= 64) || grp == 0;
    uint32_t tcount = 0;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, tcount++) {
      const int n = tile / tiles_per_img, tt = tile % tiles_per_img;
      const int h0 = (tt / p.tiles_w) * CG_TH, w0 = (tt % p.tiles_w) * CG_TW;
      const int h = h0 + hh, w = w0 + ww;
      const bool valid = (h < p.H) && (w < p.W);
      const size_t pix = ((size_t)n * p.H + h) * p.W + w;
      // mode 4 with sub == 2: only even rows/cols are kept; they are compacted to a 4x8 staging tile
      const bool kept = (MODE != 4) || (p.sub == 1) || (((hh | ww) & 1) == 0);
      const int srow = (MODE == 4 && p.sub == 2) ? (hh >> 1) * (CG_TW / 2) + (ww >> 1) : row;
      const int as = tcount & 1, aph = (tcount >> 1) & 1;
      tc::mbar_wait(&tm_full[as], aph);
      tc::tc_fence_after();
      // staging buffers free again once the previous tile's TMA stores have read them
      if (MODE != 3 && etid == 0) tma_store_wait_read();
      asm volatile("bar.sync 1, 256;" ::: "memory");
      const uint32_t taddr = tmem_base + as * ACC_STRIDE + ((uint32_t)(q * 32) << 16);
      const float* g = p.gctx ? p.gctx + (size_t)n * N : nullptr;
      constexpr int PASSES = SM::PASSES, CPP = N / PASSES;          // columns staged per pass
      constexpr int GC = (N >= 64) ? CPP / 2 : N;                   // columns per warp group per pass
#pragma unroll 1
      for (int pass = 0; pass < PASSES; pass++) {
        if (pass > 0) {
          // the staging tiles are reused: wait until the previous pass' TMA stores have read them
          if (etid == 0) tma_store_wait_read();
          asm volatile("bar.sync 1, 256;" ::: "memory");
        }
        const int cbeg = pass * CPP + ((N >= 64) ? grp * GC : 0);
        if (works) {
#pragma unroll 1
          for (int c0 = cbeg; c0 < cbeg + GC; c0 += 32) {
            uint32_t r[32];
            tc::tmem_ld_32x32(taddr + c0, r);
            // operand loads of the gating modes overlap the TMEM read
            uint4 an[4] = {}, az[4] = {};
            if (MODE == 1) {
              if (c0 >= 128 && valid) {
                const uint4* np = reinterpret_cast<const uint4*>(p.net + pix * 128 + (c0 - 128));
#pragma unroll
                for (int i = 0; i < 4; i++) an[i] = np[i];
              }
            } else if (MODE == 2 || MODE == 3) {
              if (valid) {
                const uint4* np = reinterpret_cast<const uint4*>(p.net + pix * 128 + c0);
#pragma unroll
                for (int i = 0; i < 4; i++) an[i] = np[i];
                if (MODE == 2) {
                  const uint4* zp = reinterpret_cast<const uint4*>(p.zbuf + pix * 128 + c0);
#pragma unroll
                  for (int i = 0; i < 4; i++) az[i] = zp[i];
                }
              }
            }
            tc::tmem_ld_wait();
            float v[32];
            epi_chunk<MODE>(r, v, sbias, g, c0, p.act, valid, an, az);
            if (MODE == 3) {
              const float cs = warp_colsum32(v, lane);
              atomicAdd(&sacc[c0 + lane], cs);
            } else {
              // fp16, into the staging tile of this 64-channel group (tile index within the pass)
              const int t64 = (c0 - pass * CPP) / 64;
              if (kept) {
#pragma unroll
                for (int i = 0; i < 32; i += 8) {
                  if (c0 + i >= N) break;
                  __half2 h2[4];
#pragma unroll
                  for (int j = 0; j < 4; j++) h2[j] = __floats2half2_rn(v[i + 2 * j], v[i + 2 * j + 1]);
                  if (N >= 64) {
                    // 128B-swizzled staging tile (matches the SWIZZLE_128B output tensor map)
                    unsigned char* st = sm + SM::OUT + t64 * 16384 + srow * 128;
                    const int chunk = ((c0 % 64) + i) / 8;
                    *reinterpret_cast<uint4*>(st + ((chunk ^ (srow & 7)) * 16)) = *reinterpret_cast<const uint4*>(h2);
                  } else {
                    // narrow outputs: dense rows of N halfs, un-swizzled tensor map
                    unsigned char* st = sm + SM::OUT + srow * (N * 2);
                    *reinterpret_cast<uint4*>(st + (c0 + i) * 2) = *reinterpret_cast<const uint4*>(h2);
                  }
                }
              }
              if (MODE == 4 && p.stats) {
                // statistics of what the next layer will read: the fp16-rounded values of the kept, in-image pixels
                float q2[32];
                const bool cnt = kept && valid;
#pragma unroll
                for (int i = 0; i < 32; i++) {
                  const float hv = cnt ? __half2float(__float2half_rn(v[i])) : 0.f;
                  v[i] = hv; q2[i] = hv * hv;
                }
                const float cs = warp_colsum32(v, lane), cq = warp_colsum32(q2, lane);
                if (c0 + lane < N) { atomicAdd(&sacc[c0 + lane], cs); atomicAdd(&sacc[N + c0 + lane], cq); }
              }
            }
          }
        }
        if (pass == PASSES - 1) {
          // the accumulator stage may be overwritten by the next tile's MMAs
          tc::tc_fence_before();
          __syncwarp();
          if (lane == 0) tc::mbar_arrive(&tm_empty[as]);
        }
        if (MODE != 3) {
          tc::fence_proxy_async();
          asm volatile("bar.sync 2, 256;" ::: "memory");
          if (etid == 0) {
            if (MODE == 1) {
              // pass 0: z (columns 0..127) -> out0 ; pass 1: r * net (columns 128..255) -> out1
              tma_store_4d(&maps.out[pass], sm + SM::OUT + 0 * 16384, 0, w0, h0, n);
              tma_store_4d(&maps.out[pass], sm + SM::OUT + 1 * 16384, 64, w0, h0, n);
            } else if (MODE == 4 && p.sub == 2) {
              for (int t = 0; t < SM::NOUT64; t++)
                tma_store_4d(&maps.out[0], sm + SM::OUT + t * 16384, pass * CPP + t * 64, w0 >> 1, h0 >> 1, n);
            } else {
              for (int t = 0; t < SM::NOUT64; t++)
                tma_store_4d(&maps.out[0], sm + SM::OUT + t * 16384, pass * CPP + t * 64, w0, h0, n);
            }
            tma_store_commit();
          }
        }
      }
      if (MODE != 3) {
        if (MODE == 4 && p.stats) {
          // flush this tile's channel statistics (shared accumulators were completed before the last bar.sync 2)
          for (int i = etid; i < 2 * N; i += 32 * CG_EPI_WARPS) {
            const float sv = sacc[i];
            sacc[i] = 0.f;
            if (sv != 0.f) atomicAdd(p.stats + ((size_t)n * N + (i % N)) * 2 + (i / N), sv);
          }
        }
      } else {
        // flush this tile's column sums (the next tile usually belongs to another image)
        asm volatile("bar.sync 2, 256;" ::: "memory");
        if (etid < N) {
          const float sv = sacc[etid];
          sacc[etid] = 0.f;
          if (sv != 0.f) atomicAdd(p.gsum + (size_t)n * 128 + etid, sv);
        }
      }
    }
    if (MODE != 3 && etid == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  }
  tc::tc_fence_before();
  __syncthreads();
  if (warp == 1) tc::tmem_dealloc<TCOLS>(tmem_base);
}

template <int N, int MODE, bool HALO>
static int launch_conv_mh(const ConvMaps& maps, const ConvParams& p, int num_sms, cudaStream_t st) {
  const int smem = CgSmem<N, HALO>::TOTAL + 1024;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(conv_igemm_kernel<N, MODE, HALO>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return (int)e;
    configured = true;
  }
  const int ntiles = p.B * p.tiles_h * p.tiles_w;
  const int grid = ntiles < num_sms ? ntiles : num_sms;
  conv_igemm_kernel<N, MODE, HALO><<<grid, CG_THREADS, smem, st>>>(maps, p);
  NSLAM_CHECK_LAUNCH();
  return 0;
}
template <int N, int MODE>
static int launch_conv_m(const ConvMaps& maps, const ConvParams& p, bool halo, int num_sms, cudaStream_t st) {
  return halo ? launch_conv_mh<N, MODE, true>(maps, p, num_sms, st) : launch_conv_mh<N, MODE, false>(maps, p, num_sms, st);
}

// only the (N, mode) pairs the update operator uses are instantiated
static int launch_conv(int N, const ConvMaps& maps, const ConvParams& p, bool halo, int num_sms, cudaStream_t st) {
  if (p.mode == 0) {
    switch (N) {
      case 16: return launch_conv_m<16, 0>(maps, p, halo, num_sms, st);
      case 32: return launch_conv_m<32, 0>(maps, p, halo, num_sms, st);
      case 64: return launch_conv_m<64, 0>(maps, p, halo, num_sms, st);
      case 128: return launch_conv_m<128, 0>(maps, p, halo, num_sms, st);
      case 256: return launch_conv_m<256, 0>(maps, p, halo, num_sms, st);
    }
  } else if (p.mode == 1 && N == 256) {
    return launch_conv_m<256, 1>(maps, p, halo, num_sms, st);
  } else if (p.mode == 2 && N == 128) {
    return launch_conv_m<128, 2>(maps, p, halo, num_sms, st);
  } else if (p.mode == 3 && N == 128) {
    return launch_conv_mh<128, 3, false>(maps, p, num_sms, st);
  } else if (p.mode == 4) {
    switch (N) {
      case 32: return launch_conv_m<32, 4>(maps, p, halo, num_sms, st);
      case 64: return launch_conv_m<64, 4>(maps, p, halo, num_sms, st);
      case 128: return launch_conv_m<128, 4>(maps, p, halo, num_sms, st);
      case 256: return launch_conv_m<256, 4>(maps, p, halo, num_sms, st);
    }
  }
  return (int)cudaErrorInvalidValue;
}

}  // namespace nslam

extern "C" {

/* NHWC fp16 convolution (stride 1) on tensor cores.
 *   srcs[i]: [B,H,W,src_channels[i]] fp16 (channels need not be multiples of 64: the tail of the
 *            last 64-block is zero-filled by TMA; the packed weights must be zero there too)
 *   wpacked: produced by nslam_conv_pack_weights; bias [N] fp32 or NULL
 *   out0 (and out1 for mode 1): [B,H,W,out_channels] fp16; N = 16,32,64,128,256 columns per launch
 *   mode/act/gctx/net/zbuf/gsum: see the epilogue modes above. */
int nslam_conv_igemm_ex(const void* const* srcs, const int* src_channels, int n_src, int B, int H, int W,
                        int KH, int KW, int pad, int N, const void* wpacked, const float* bias, int mode,
                        int act, const float* gctx, const void* net, const void* zbuf, float* gsum,
                        void* out0, int out0_channels, void* out1, float* stats, int sub, int num_sms, void* stream) {
  using namespace nslam;
  if (n_src < 1 || n_src > 4 || B <= 0) return (int)cudaErrorInvalidValue;
  ConvMaps maps;
  ConvParams p{};
  p.B = B; p.H = H; p.W = W;
  p.tiles_h = (H + CG_TH - 1) / CG_TH; p.tiles_w = (W + CG_TW - 1) / CG_TW;
  p.n_src = n_src; p.KH = KH; p.KW = KW; p.pad = pad; p.N = N; p.mode = mode; p.act = act;
  p.wpacked = (const __half*)wpacked; p.bias = bias; p.gctx = gctx; p.net = (const __half*)net;
  p.zbuf = (const __half*)zbuf; p.gsum = gsum; p.stats = stats; p.sub = (mode == 4 && sub == 2) ? 2 : 1;
  if (mode == 4 && sub != 1 && sub != 2) return (int)cudaErrorInvalidValue;
  // NSLAM_CONV_HALO=1: the 16x16 super-tile kernel with one halo box per channel block (conv_halo.cu) for 3x3 / pad 1,
  // epilogue modes 0-2 — parity-green but not faster on a B200 (profiles/r02_conv_analysis.md); measurement switch.
  static const bool want_halo2 = [] { const char* e = std::getenv("NSLAM_CONV_HALO"); return e && e[0] == '1'; }();
  const bool halo2 = want_halo2 && !conv_pairs_enabled() && conv_halo_supported(N, mode, KH, KW, pad);
  // first generation, 3x3 / pad 1: column-shifted halo tiles (see CgSmem); everything else: one tile per tap
  const bool halo = (KH == 3 && KW == 3 && pad == 1 && mode != 3);
  int cbt = 0;
  for (int s = 0; s < n_src; s++) {
    const int C = src_channels[s];
    if (C % 8 != 0) return (int)cudaErrorInvalidValue;   // TMA strides must be multiples of 16 bytes
    p.src_cb[s] = (C + 63) / 64;
    cbt += p.src_cb[s];
    uint64_t dims[4] = {(uint64_t)C, (uint64_t)W, (uint64_t)H, (uint64_t)B};
    uint64_t strides[3] = {(uint64_t)C * 2, (uint64_t)W * C * 2, (uint64_t)H * W * C * 2};
    uint32_t box[4] = {64, CG_TW, (uint32_t)(halo ? CG_TH + 2 : CG_TH), 1};
    if (halo2) { box[1] = CH2_BOX; box[2] = CH2_BOX; }
    int r = tc::make_tmap_f16(&maps.src[s], srcs[s], 4, dims, strides, box);
    if (r) return r;
  }
  p.cb_total = cbt;
  if (mode != 3) {
    void* outs[2] = {out0, out1};
    const int nout = (mode == 1) ? 2 : 1;
    for (int o = 0; o < nout; o++) {
      const int C = (mode == 1) ? 128 : out0_channels;
      const int Wo = (W + p.sub - 1) / p.sub, Ho = (H + p.sub - 1) / p.sub;     // stride-2 layers store every other pixel
      uint64_t dims[4] = {(uint64_t)C, (uint64_t)Wo, (uint64_t)Ho, (uint64_t)B};
      uint64_t strides[3] = {(uint64_t)C * 2, (uint64_t)Wo * C * 2, (uint64_t)Ho * Wo * C * 2};
      uint32_t box[4] = {(uint32_t)(C < 64 ? C : 64), (uint32_t)(CG_TW / p.sub), (uint32_t)(CG_TH / p.sub), 1};
      if (halo2) { box[1] = 8; box[2] = 16; }                    // sub-tile of the second-generation kernel
      int r = tc::make_tmap_f16(&maps.out[o], outs[o], 4, dims, strides, box, false, nullptr, /*swizzle128=*/N >= 64);
      if (r) return r;
    }
  }
  if (halo2) {
    p.tiles_h = (H + 15) / 16; p.tiles_w = (W + 15) / 16;        // super-tiles
    return launch_conv_halo(N, maps, p, num_sms, (cudaStream_t)stream);
  }
  if (conv_pairs_enabled() && conv_pairs_supported(N, mode, halo)) return launch_conv_pairs(N, maps, p, num_sms, (cudaStream_t)stream);
  return launch_conv(N, maps, p, halo, num_sms, (cudaStream_t)stream);
}

int nslam_conv_igemm(const void* const* srcs, const int* src_channels, int n_src, int B, int H, int W,
                     int KH, int KW, int pad, int N, const void* wpacked, const float* bias, int mode,
                     int act, const float* gctx, const void* net, const void* zbuf, float* gsum,
                     void* out0, int out0_channels, void* out1, int num_sms, void* stream) {
  return nslam_conv_igemm_ex(srcs, src_channels, n_src, B, H, W, KH, KW, pad, N, wpacked, bias, mode, act, gctx, net, zbuf,
                             gsum, out0, out0_channels, out1, nullptr, 1, num_sms, stream);
}

}  // extern "C"
