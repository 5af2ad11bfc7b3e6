// Host side of the BA window structure: edge list -> the int32 tables the BA kernels walk (include/nslam_ba.h,
// nslam_ba_graph).  This is the part the reference does on the CPU in C++ on EVERY call — `torch::_unique`
// (src/droid_kernels.cu:1697-1706), the argsort + CSR of `accum_cuda` (:1065-1115) and the enumeration of co-visible
// (i, j, k) triples in `schur_block` (:1349-1399).  Here it runs once per distinct edge set, straight into the packed
// buffer that is uploaded (one call from Python instead of ~60 numpy calls: the GPU idles while a new edge set is
// prepared).  Pure host code (no kernels); orderings are the stable ones of nerf_slam_b200/ba_graph.py, which remains
// as the table-for-table check (tests/test_cpu_graph.py).
#include <algorithm>
#include <cstdint>
#include <vector>

#include "nslam_ba.h"

namespace {

// stable counting sort of `n` items by key in [0, nb): returns positions (order) and fills ptr[nb+1]
void stable_by_key(const std::vector<int64_t>& key, int64_t nb, std::vector<int64_t>& order, std::vector<int64_t>& ptr) {
  const size_t n = key.size();
  ptr.assign((size_t)nb + 1, 0);
  for (size_t i = 0; i < n; i++) ptr[(size_t)key[i] + 1]++;
  for (int64_t b = 0; b < nb; b++) ptr[(size_t)b + 1] += ptr[(size_t)b];
  std::vector<int64_t> cur(ptr.begin(), ptr.end() - 1);
  order.resize(n);
  for (size_t i = 0; i < n; i++) order[(size_t)cur[(size_t)key[i]]++] = (int64_t)i;
}

}  // namespace

extern "C" int nslam_ba_graph_build(const long long* ii, const long long* jj, int E, int kf0, int kf1, int* out,
                                    int capacity, int* meta) {
  const int64_t P = (int64_t)kf1 - kf0;
  if (P <= 0 || E < 0) return 2;
  // depth maps: sorted unique of the window frames and the edges' source frames
  std::vector<int64_t> kx;
  kx.reserve((size_t)P + (size_t)E);
  for (int64_t t = kf0; t < kf1; t++) kx.push_back(t);
  for (int e = 0; e < E; e++) kx.push_back(ii[e]);
  std::sort(kx.begin(), kx.end());
  kx.erase(std::unique(kx.begin(), kx.end()), kx.end());
  const int64_t K = (int64_t)kx.size();
  std::vector<int64_t> kk((size_t)E);
  for (int e = 0; e < E; e++) kk[(size_t)e] = std::lower_bound(kx.begin(), kx.end(), (int64_t)ii[e]) - kx.begin();
  std::vector<int64_t> order, src_ptr;
  stable_by_key(kk, K, order, src_ptr);

  // Schur rows per depth map: self row first (frame in the window), then its edges whose target pose is in the window
  std::vector<int64_t> row_pose, row_erow, row_ptr((size_t)K + 1, 0);
  for (int64_t k = 0; k < K; k++) {
    if (kx[(size_t)k] >= kf0 && kx[(size_t)k] < kf1) { row_pose.push_back(kx[(size_t)k] - kf0); row_erow.push_back(kx[(size_t)k] - kf0); }
    for (int64_t s = src_ptr[(size_t)k]; s < src_ptr[(size_t)k + 1]; s++) {
      const int64_t e = order[(size_t)s];
      if (jj[e] >= kf0 && jj[e] < kf1) { row_pose.push_back(jj[e] - kf0); row_erow.push_back(P + e); }
    }
    row_ptr[(size_t)k + 1] = (int64_t)row_pose.size();
  }
  const int64_t NR = (int64_t)row_pose.size();
  std::vector<int64_t> pair_off((size_t)K + 1, 0);
  int64_t RMAX = 0;
  for (int64_t k = 0; k < K; k++) {
    const int64_t R = row_ptr[(size_t)k + 1] - row_ptr[(size_t)k];
    pair_off[(size_t)k + 1] = pair_off[(size_t)k] + R * R;
    RMAX = std::max(RMAX, R);
  }
  const int64_t NPAIR = pair_off[(size_t)K];

  // contributions to the dense blocks (a, b) of H: four per-edge blocks, then the Schur blocks (negative ids)
  std::vector<int64_t> hkey, hval;
  hkey.reserve((size_t)(4 * E + NPAIR)); hval.reserve((size_t)(4 * E + NPAIR));
  for (int w = 0; w < 4; w++)
    for (int e = 0; e < E; e++) {
      const int64_t a = ii[e] - kf0, b = jj[e] - kf0;
      const bool av = a >= 0 && a < P, bv = b >= 0 && b < P;
      const bool ok = (w == 0) ? av : (w == 3 ? bv : (av && bv));
      if (!ok) continue;
      const int64_t r = (w < 2) ? a : b, c = (w == 0 || w == 2) ? a : b;
      hkey.push_back(r * P + c); hval.push_back((int64_t)w * E + e);
    }
  {
    int64_t blk = 0;
    for (int64_t k = 0; k < K; k++)
      for (int64_t ra = row_ptr[(size_t)k]; ra < row_ptr[(size_t)k + 1]; ra++)
        for (int64_t rb = row_ptr[(size_t)k]; rb < row_ptr[(size_t)k + 1]; rb++) {
          hkey.push_back(row_pose[(size_t)ra] * P + row_pose[(size_t)rb]); hval.push_back(-(blk + 1)); blk++;
        }
  }
  std::vector<int64_t> ho, hc_ptr;
  stable_by_key(hkey, P * P, ho, hc_ptr);
  // contributions to the segments of v
  std::vector<int64_t> vkey, vval;
  for (int e = 0; e < E; e++) { const int64_t a = ii[e] - kf0; if (a >= 0 && a < P) { vkey.push_back(a); vval.push_back(e); } }
  for (int e = 0; e < E; e++) { const int64_t b = jj[e] - kf0; if (b >= 0 && b < P) { vkey.push_back(b); vval.push_back((int64_t)E + e); } }
  for (int64_t r = 0; r < NR; r++) { vkey.push_back(row_pose[(size_t)r]); vval.push_back(-(r + 1)); }
  std::vector<int64_t> vo, vc_ptr;
  stable_by_key(vkey, P, vo, vc_ptr);
  const int64_t NHC = (int64_t)hkey.size(), NVC = (int64_t)vkey.size();

  // packed layout: 13 tables, each padded to a multiple of 4 ints (16-byte aligned device pointers)
  const int64_t lens[13] = {E, E, K, K + 1, E, K + 1, NR, NR, K + 1, P * P + 1, NHC, P + 1, NVC};
  int64_t pos = 0, offs[13];
  for (int t = 0; t < 13; t++) { offs[t] = pos; pos += (lens[t] + 3) / 4 * 4; }
  const int64_t total = pos > 0 ? pos : 4;
  meta[0] = E; meta[1] = (int)P; meta[2] = (int)K; meta[3] = kf0; meta[4] = (int)NR; meta[5] = (int)NPAIR; meta[6] = (int)RMAX;
  meta[7] = (int)NHC; meta[8] = (int)NVC; meta[9] = (int)total;
  for (int t = 0; t < 13; t++) { meta[10 + t] = (int)offs[t]; meta[23 + t] = (int)lens[t]; }
  if (total > capacity) return 1;
  std::fill(out, out + total, 0);
  int* o;
  o = out + offs[0]; for (int e = 0; e < E; e++) o[e] = (int)ii[e];
  o = out + offs[1]; for (int e = 0; e < E; e++) o[e] = (int)jj[e];
  o = out + offs[2]; for (int64_t k = 0; k < K; k++) o[k] = (int)kx[(size_t)k];
  o = out + offs[3]; for (int64_t k = 0; k <= K; k++) o[k] = (int)src_ptr[(size_t)k];
  o = out + offs[4]; for (int e = 0; e < E; e++) o[e] = (int)order[(size_t)e];
  o = out + offs[5]; for (int64_t k = 0; k <= K; k++) o[k] = (int)row_ptr[(size_t)k];
  o = out + offs[6]; for (int64_t r = 0; r < NR; r++) o[r] = (int)row_pose[(size_t)r];
  o = out + offs[7]; for (int64_t r = 0; r < NR; r++) o[r] = (int)row_erow[(size_t)r];
  o = out + offs[8]; for (int64_t k = 0; k <= K; k++) o[k] = (int)pair_off[(size_t)k];
  o = out + offs[9]; for (int64_t b = 0; b <= P * P; b++) o[b] = (int)hc_ptr[(size_t)b];
  o = out + offs[10]; for (int64_t i = 0; i < NHC; i++) o[i] = (int)hval[(size_t)ho[(size_t)i]];
  o = out + offs[11]; for (int64_t b = 0; b <= P; b++) o[b] = (int)vc_ptr[(size_t)b];
  o = out + offs[12]; for (int64_t i = 0; i < NVC; i++) o[i] = (int)vval[(size_t)vo[(size_t)i]];
  return 0;
}

// Host side of add_proximity_factors (visual_frontend.py:712-775 / networks/factor_graph.py:323-387): the order-sensitive
// choice of new edges from the pairwise flow distances — the reference runs three nested Python loops per existing edge
// on values it first copies from the GPU.  d: HOST fp32 [(t-kf0)*(t-kf1)] over the meshgrid (i in [kf0,t), j in [kf1,t)),
// modified in place; ii1/jj1: existing (active + bad + inactive) edges whose neighbourhoods are suppressed first.
// Writes the chosen directed edges (i, j) pairs to es [cap][2]; returns 0 ok, 1 capacity too small.
// Same results as nerf_slam_b200/graph.py::proximity_edges_numpy (bit-exact incl. ties, which take index order, and the
// reference's wrapping of negative flat indices); pinned against the reference's own loops by the recorded traces.
extern "C" int nslam_proximity_edges(float* d, int kf0, int kf1, int t, const long long* ii1, const long long* jj1, int n1,
                                     int rad, int nms, float thresh, int max_factors, int stereo, long long* es, int cap,
                                     int* n_out) {
  const int64_t W = (int64_t)t - kf1, Hh = (int64_t)t - kf0;
  const int64_t n = Hh * W;
  const float INF = __builtin_inff();
  *n_out = 0;
  if (Hh < 0 || W < 0) return 2;
  auto put = [&](int64_t idx) {                 // numpy semantics of d[idx] = inf with a possibly negative index
    if (idx < 0) idx += n;
    if (idx >= 0 && idx < n) d[idx] = INF;
  };
  for (int64_t a = 0; a < Hh; a++)
    for (int64_t b = 0; b < W; b++) {
      const int64_t i = kf0 + a, j = kf1 + b;
      float& v = d[a * W + b];
      if (i - rad < j) v = INF;
      if (v > 100.f) v = INF;
    }
  auto suppress = [&](int64_t i, int64_t j) {
    int64_t r = (i > j ? i - j : j - i) - 2;
    r = r < 0 ? 0 : (r > nms ? nms : r);
    for (int64_t di = -r; di <= r; di++)
      for (int64_t dj = -r; dj <= r; dj++) {
        if ((di < 0 ? -di : di) + (dj < 0 ? -dj : dj) > r) continue;
        const int64_t i1 = i + di, j1 = j + dj;
        if (i1 >= kf0 && i1 < t && j1 >= kf1 && j1 < t) d[(i1 - kf0) * W + (j1 - kf1)] = INF;
      }
  };
  for (int e = 0; e < n1; e++) suppress(ii1[e], jj1[e]);
  int cnt = 0;
  bool overflow = false;
  auto add = [&](int64_t i, int64_t j) {
    if (cnt < cap) { es[2 * cnt] = i; es[2 * cnt + 1] = j; } else overflow = true;
    cnt++;
  };
  for (int64_t i = kf0; i < t; i++) {
    if (stereo) { add(i, i); put((i - kf0) * W + (i - kf1)); }
    for (int64_t j = std::max<int64_t>(i - rad - 1, 0); j < i; j++) { add(i, j); add(j, i); put((i - kf0) * W + (j - kf1)); }
  }
  std::vector<int64_t> order((size_t)n);
  for (int64_t k = 0; k < n; k++) order[(size_t)k] = k;
  std::vector<float> d0(d, d + n);              // the order is fixed by the values BEFORE the selection loop mutates them
  // NaN distances (degenerate frames) sort last, as numpy's argsort places them: with a plain `<` the comparator would
  // not be a strict weak ordering and the sort's behaviour undefined
  std::stable_sort(order.begin(), order.end(), [&](int64_t a, int64_t b) {
    const float x = d0[(size_t)a], y = d0[(size_t)b];
    if (x != x) return false;
    if (y != y) return true;
    return x < y;
  });
  for (int64_t s = 0; s < n; s++) {
    const int64_t k = order[(size_t)s];
    if (d[k] > thresh) continue;
    if (cnt > max_factors) break;
    const int64_t i = kf0 + k / W, j = kf1 + k % W;
    add(i, j); add(j, i);
    suppress(i, j);
  }
  *n_out = cnt;
  return overflow ? 1 : 0;
}
