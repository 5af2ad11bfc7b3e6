// A14, the inverse-depth covariances exactly as the reference's block computes them, in ONE kernel — the default of the
// live path (cov_mode 1 of nslam_ba_frontend_update, written into the keyframe arenas) and of BAProblem.covariances.
// droid_backends.cov_reference_fixup (torch ops on top of csrc/ba.cu's nslam_ba_cov) is kept as an independent
// cross-check in the tests.
//
// The reference builds, per optimised pose p and depth map k, the 6 x HW block E[p][k] that enters
//   Sigma_z = Q + sum_cols((Q * E^T) L^-1)^2                        (visual_frontend.py:1196-1230).
// Its statement `Ej[range(P), kf0-min:kf1-min, :, :] = Ei[range(P), :, :]` (:1214) is meant to put Ei on the
// diagonal, but an advanced index on dim 0 combined with a slice on dim 1 BROADCASTS the value over dim 0: every pose
// row p receives Ei[q] in column q of the window, and the off-diagonal Ejz blocks of in-window depth maps are
// overwritten.  (tests/golden/ref_covariances.npz records the block's real output; tests/test_cpu_golden.py.)
// Consequently, for a depth map whose frame is optimised (q = kx[k] - kf0 in [0, P)):
//     x_p = Ei[q] for ALL p      =>   x^T (L^-1 L^-T) x = Ei[q]^T ( sum_{p,p'} M[p][p'] ) Ei[q]
// and for the depth maps of fixed frames the intended formula (only Ejz blocks) is what the reference computes.
// nslam_ba_cov (csrc/ba.cu) implements the intended formula for every map; this file reproduces the reference.
#include "common.cuh"
#include "nslam_ba.h"

namespace nslam {

constexpr int COV_TILE = 128;

__global__ void cov_ref_M_kernel(const float* __restrict__ Linv, int n, float* __restrict__ M) {
  const int id = blockIdx.x * blockDim.x + threadIdx.x;
  if (id >= n * n) return;
  const int r = id / n, c = id % n;
  float s = 0.f;
  const int kmax = (r < c ? r : c);  // Linv is lower triangular
  for (int k = 0; k <= kmax; k++) s += Linv[(size_t)r * n + k] * Linv[(size_t)c * n + k];
  M[id] = s;                         // M = L^-1 L^-T: sum_cols((Q x^T) L^-1)^2 = Q^2 x^T M x   (same as csrc/ba.cu)
}

// Msum[a][b] = sum over all pose pairs (p, p') of M[6p + a][6p' + b]
__global__ void cov_ref_msum_kernel(const float* __restrict__ M, int P, float* __restrict__ Msum) {
  const int a = threadIdx.x / 6, c = threadIdx.x % 6;
  const int n = 6 * P;
  float s = 0.f;
  for (int p = 0; p < P; p++)
    for (int q = 0; q < P; q++) s += M[(size_t)(6 * p + a) * n + 6 * q + c];
  Msum[threadIdx.x] = s;
}

// grid (ceil(HW / COV_TILE), K).  ARENA: output row of depth map k is its FRAME id kx[k] (the front end's [buffer,HW]
// arenas) instead of k; `guard` (device int, may be NULL): non-zero = the solve failed, write nothing.
// reference_quirk = false: the intended formula for every map (what nslam_ba_cov computes).
template <bool ARENA>
__global__ void __launch_bounds__(COV_TILE)
ba_cov_ref_kernel(nslam_ba_graph g, nslam_ba_buffers b, const float* __restrict__ M, const float* __restrict__ Msum,
                  float* __restrict__ z_cov, float* __restrict__ depth_cov, bool reference_quirk,
                  const int* __restrict__ guard) {
  if (guard && *guard) return;
  const int k = blockIdx.y;
  const int hw = b.ht * b.wd;
  const int n = 6 * g.P;
  const int p = blockIdx.x * COV_TILE + threadIdx.x;
  if (p >= hw) return;
  const float q = b.Q[(size_t)k * hw + p];
  const int win = g.kx[k] - g.kf0;
  const size_t orow = ARENA ? (size_t)g.kx[k] : (size_t)k;
  float acc = 0.f;
  if (reference_quirk && win >= 0 && win < g.P) {
    // optimised frame: every pose row holds Ei[win] (E rows 0..P-1 are Ei)
    float e[6];
#pragma unroll
    for (int m = 0; m < 6; m++) e[m] = b.Emat[((size_t)win * 6 + m) * hw + p];
#pragma unroll
    for (int m = 0; m < 6; m++) {
      float s = 0.f;
#pragma unroll
      for (int c = 0; c < 6; c++) s += Msum[m * 6 + c] * e[c];
      acc += e[m] * s;
    }
  } else {
    // fixed frame: the Ejz blocks of its edges into the window
    const int r0 = g.row_ptr[k], R = g.row_ptr[k + 1] - r0;
    for (int ra = 0; ra < R; ra++) {
      const float* sa = b.Emat + ((size_t)g.row_erow[r0 + ra] * 6) * hw + p;
      const int pa = g.row_pose[r0 + ra];
      float ea[6];
#pragma unroll
      for (int m = 0; m < 6; m++) ea[m] = sa[(size_t)m * hw];
      for (int rb = 0; rb < R; rb++) {
        const float* sb = b.Emat + ((size_t)g.row_erow[r0 + rb] * 6) * hw + p;
        const int pb = g.row_pose[r0 + rb];
#pragma unroll
        for (int c = 0; c < 6; c++) {
          const float eb = sb[(size_t)c * hw];
          float s = 0.f;
#pragma unroll
          for (int m = 0; m < 6; m++) s += ea[m] * __ldg(M + (size_t)(pa * 6 + m) * n + pb * 6 + c);
          acc += s * eb;
        }
      }
    }
  }
  const float zc = q + q * q * acc;
  z_cov[orow * hw + p] = zc;
  if (depth_cov) {
    const float d = b.disps[(size_t)g.kx[k] * hw + p];
    const float d2 = d * d;
    depth_cov[orow * hw + p] = zc / (d2 * d2);
  }
}

// block-diagonal 6x6 blocks of (L L^T)^-1 = Linv^T Linv, written to rows kf0 + i of the [buffer,6,6] pose-covariance arena
__global__ void pose_cov_arena_kernel(const float* __restrict__ Linv, int P, int kf0, float* __restrict__ sg,
                                      const int* __restrict__ guard) {
  if (guard && *guard) return;
  const int i = blockIdx.x;
  const int r = threadIdx.x / 6, c = threadIdx.x % 6;
  const int n = 6 * P;
  float s = 0.f;
  for (int k = 0; k < n; k++) s += Linv[(size_t)k * n + i * 6 + r] * Linv[(size_t)k * n + i * 6 + c];
  sg[(size_t)(kf0 + i) * 36 + threadIdx.x] = s;
}

}  // namespace nslam

extern "C" int nslam_ba_cov_reference(const nslam_ba_graph* g, const nslam_ba_buffers* b, const float* Linv,
                                      float* Mscratch, float* z_cov, float* depth_cov, void* stream) {
  using namespace nslam;
  const int n = 6 * g->P;
  if (n <= 0) return (int)cudaErrorInvalidValue;
  cudaStream_t st = (cudaStream_t)stream;
  cov_ref_M_kernel<<<(n * n + 255) / 256, 256, 0, st>>>(Linv, n, Mscratch);
  NSLAM_CHECK_LAUNCH();
  float* Msum = Mscratch + (size_t)n * n;          // 36 floats behind the [n,n] matrix
  cov_ref_msum_kernel<<<1, 36, 0, st>>>(Mscratch, g->P, Msum);
  NSLAM_CHECK_LAUNCH();
  const int hw = b->ht * b->wd;
  dim3 grid((hw + COV_TILE - 1) / COV_TILE, g->K);
  ba_cov_ref_kernel<false><<<grid, COV_TILE, 0, st>>>(*g, *b, Mscratch, Msum, z_cov, depth_cov, true, nullptr);
  NSLAM_CHECK_LAUNCH();
  return 0;
}

/* The live path's covariance block (visual_frontend.py:1164-1230) in three launches, results written in place into
 * the keyframe arenas: idepths_cov / depths_cov [buffer,HW] rows kx[k], pose_cov [buffer,6,6] rows kf0..kf0+P-1.
 * mode 1 = the reference's real behaviour (Ei broadcast, see the head of this file), 0 = the intended formula.
 * guard: device int, non-zero = the factorisation failed -> nothing is written. */
extern "C" int nslam_ba_cov_arena(const nslam_ba_graph* g, const nslam_ba_buffers* b, const float* Linv, float* Mscratch,
                                  int mode, const int* guard, float* idepths_cov, float* depths_cov, float* pose_cov,
                                  void* stream) {
  using namespace nslam;
  const int n = 6 * g->P;
  if (n <= 0) return (int)cudaErrorInvalidValue;
  cudaStream_t st = (cudaStream_t)stream;
  cov_ref_M_kernel<<<(n * n + 255) / 256, 256, 0, st>>>(Linv, n, Mscratch);
  NSLAM_CHECK_LAUNCH();
  float* Msum = Mscratch + (size_t)n * n;
  if (mode == 1) {
    cov_ref_msum_kernel<<<1, 36, 0, st>>>(Mscratch, g->P, Msum);
    NSLAM_CHECK_LAUNCH();
  }
  const int hw = b->ht * b->wd;
  dim3 grid((hw + COV_TILE - 1) / COV_TILE, g->K);
  ba_cov_ref_kernel<true><<<grid, COV_TILE, 0, st>>>(*g, *b, Mscratch, Msum, idepths_cov, depths_cov, mode == 1, guard);
  NSLAM_CHECK_LAUNCH();
  if (pose_cov) {
    pose_cov_arena_kernel<<<g->P, 36, 0, st>>>(Linv, g->P, g->kf0, pose_cov, guard);
    NSLAM_CHECK_LAUNCH();
  }
  return 0;
}
