// TEST INFRASTRUCTURE ONLY. g++ host instantiation of shot_fpfh_b200/csrc/sf_math.cuh, the header the sm_100a
// kernels inline, so that the per-neighbour arithmetic (bin decisions, interpolation weights, winner tables, 3x3
// eigen-solver, FPFH features and NumPy-compatible binning) is checked against the oracle on machines without
// a GPU (tests/test_host_math.py). The loops below mirror the kernels' control flow sequentially; the product
// never loads this library.
#include <cstring>
#include <vector>

#include "../../shot_fpfh_b200/csrc/sf_math.cuh"

using namespace sf;

extern "C" {

double hm_rdist3(double dx, double dy, double dz) { return rdist3(dx, dy, dz); }

void hm_eigh3(const double* m, double* eval, double* evec9) {
  double e[3], v[3][3];
  eigh3(m, e, v);
  for (int c = 0; c < 3; ++c) {
    eval[c] = e[c];
    for (int k = 0; k < 3; ++k) evec9[3 * c + k] = v[c][k];
  }
}

int hm_azimuth_octant(double x, double y) { return azimuth_octant(x, y); }

// Mirrors shot_lrf_kernel: lrf9 row-major, columns [x y z].
void hm_lrf(const double* point, const double* nbrs, int k, double radius, double* lrf9) {
  if (k == 0) {
    for (int i = 0; i < 9; ++i) lrf9[i] = (i % 4 == 0) ? 1.0 : 0.0;
    return;
  }
  double sw = 0, m[6] = {0, 0, 0, 0, 0, 0};
  for (int i = 0; i < k; ++i) {
    const double cx = nbrs[3 * i] - point[0], cy = nbrs[3 * i + 1] - point[1], cz = nbrs[3 * i + 2] - point[2];
    const double w = radius - sqrt(rdist3(cx, cy, cz));
    sw += w;
    m[0] += w * cx * cx; m[1] += w * cx * cy; m[2] += w * cx * cz;
    m[3] += w * cy * cy; m[4] += w * cy * cz; m[5] += w * cz * cz;
  }
  for (int j = 0; j < 6; ++j) m[j] /= sw;
  double eval[3], evec[3][3];
  eigh3(m, eval, evec);
  double x[3] = {evec[2][0], evec[2][1], evec[2][2]}, z[3] = {evec[0][0], evec[0][1], evec[0][2]};
  int neg_x = 0, neg_z = 0;
  for (int i = 0; i < k; ++i) {
    const double cx = nbrs[3 * i] - point[0], cy = nbrs[3 * i + 1] - point[1], cz = nbrs[3 * i + 2] - point[2];
    neg_x += (cx * x[0] + cy * x[1] + cz * x[2]) < 0.0;
    neg_z += (cx * z[0] + cy * z[1] + cz * z[2]) < 0.0;
  }
  if (neg_x > k - neg_x) for (int a = 0; a < 3; ++a) x[a] = -x[a];
  if (neg_z > k - neg_z) for (int a = 0; a < 3; ++a) z[a] = -z[a];
  const double y[3] = {z[1] * x[2] - z[2] * x[1], z[2] * x[0] - z[0] * x[2], z[0] * x[1] - z[1] * x[0]};
  for (int a = 0; a < 3; ++a) {
    lrf9[3 * a + 0] = x[a];
    lrf9[3 * a + 1] = y[a];
    lrf9[3 * a + 2] = z[a];
  }
}

// Mirrors shot_descriptor_kernel for one query: compact winner tables, neighbours taken 32 at a time (key pass,
// then value pass), atomicMax replaced by a sequential max.
void hm_shot_descriptor(const double* point, const double* nbrs, const double* normals, int k, double radius,
                        const double* f, int normalize, int min_nb, float* out352) {
  std::vector<uint32_t> keys(kKeyCount, 0u);
  std::vector<float> vals(kValCount, 0.0f);
  int positive = 0;
  const double inv_radius = 1.0 / radius;
  for (int base = 0; base < k; base += 32) {
    ShotDecision d[32];
    bool active[32];
    for (int lane = 0; lane < 32; ++lane) {
      active[lane] = false;
      const int i = base + lane;
      if (i >= k) continue;
      const double cx = nbrs[3 * i] - point[0], cy = nbrs[3 * i + 1] - point[1], cz = nbrs[3 * i + 2] - point[2];
      const double d2 = rdist3(cx, cy, cz);
      if (!(d2 > 0.0)) continue;
      active[lane] = true;
      ++positive;
      const double rho = sqrt(d2);
      const double X = cx * f[0] + cy * f[3] + cz * f[6];
      const double Y = cx * f[1] + cy * f[4] + cz * f[7];
      const double Z = cx * f[2] + cy * f[5] + cz * f[8];
      double cosine = normals[3 * i] * f[2] + normals[3 * i + 1] * f[5] + normals[3 * i + 2] * f[8];
      cosine = fmin(1.0, fmax(-1.0, cosine));
      d[lane] = shot_decide(X, Y, Z, cosine, rho, radius, inv_radius);
      if (d[lane].key > keys[kKeyOwn + d[lane].own]) keys[kKeyOwn + d[lane].own] = d[lane].key;
      if (d[lane].key > keys[kKeyCos + d[lane].cos_nb]) keys[kKeyCos + d[lane].cos_nb] = d[lane].key;
      if (d[lane].key > keys[kKeyAz + d[lane].az_nb]) keys[kKeyAz + d[lane].az_nb] = d[lane].key;
    }
    for (int lane = 0; lane < 32; ++lane) {
      if (!active[lane]) continue;
      const ShotDecision& e = d[lane];
      const bool win_own = keys[kKeyOwn + e.own] == e.key, win_cos = keys[kKeyCos + e.cos_nb] == e.key,
                 win_az = keys[kKeyAz + e.az_nb] == e.key;
      float a_az = 0.0f;
      if (win_own || win_az) a_az = shot_azimuth(e);
      if (win_own) {
        float own_vol, other_vol;
        shot_elevation(e, own_vol, other_vol);
        vals[kValOwn + e.own] = (1.0f - e.a_cos) + e.own_shell + own_vol + (1.0f - a_az);
        vals[kValRad + e.own] = e.other_shell;
        vals[kValEl + e.own] = other_vol;
      }
      if (win_cos) vals[kValCos + e.cos_nb] = e.a_cos;
      if (win_az) vals[kValAz + e.az_nb] = a_az;
    }
  }
  double sq = 0.0;
  float v[kShotLen];
  for (int g = 0; g < kShotLen / 4; ++g) {  // four bins of one (cosine, azimuth) cell at a time, as the kernel does
    const uint32_t* k = keys.data() + 4 * g;
    const float* x = vals.data() + 4 * g;
    shot_bin_group_compact(k + kKeyOwn, k + kKeyCos, k + kKeyAz, x + kValOwn, x + kValRad, x + kValEl, x + kValCos,
                           x + kValAz, v + 4 * g);
    for (int t = 0; t < 4; ++t) {
      if (v[4 * g + t] != shot_bin_value_compact(keys.data(), vals.data(), 4 * g + t)) v[4 * g + t] = NAN;  // must agree
      sq += double(v[4 * g + t]) * double(v[4 * g + t]);
    }
  }
  const double norm = sqrt(sq);
  const bool keep = positive > min_nb && norm > 0.0;
  const float inv = keep ? (normalize ? float(1.0 / norm) : 1.0f) : 0.0f;
  for (int b = 0; b < kShotLen; ++b) out352[b] = v[b] * inv;
}

// Mirrors spfh_kernel for one point: integer histogram of width 3n or n^3.
void hm_spfh_counts(const double* point, const double* normal, const double* nbrs, const double* nbr_normals, int k,
                    int n_bins, int decorrelated, const double* edges, int* hist) {
  const int width = decorrelated ? 3 * n_bins : n_bins * n_bins * n_bins;
  std::memset(hist, 0, sizeof(int) * width);
  for (int i = 0; i < k; ++i) {
    const double rel[3] = {nbrs[3 * i] - point[0], nbrs[3 * i + 1] - point[1], nbrs[3 * i + 2] - point[2]};
    const double d2 = rdist3(rel[0], rel[1], rel[2]);
    const double* e[3] = {edges, edges + (n_bins + 1), edges + 2 * (n_bins + 1)};
    double scale[3];
    float lo32[3], scale32[3];
    for (int f = 0; f < 3; ++f) {
      scale[f] = double(n_bins) / (e[f][n_bins] - e[f][0]);
      lo32[f] = float(e[f][0]);
      scale32[f] = float(scale[f]);
    }
    // as the kernel: float32-filtered bins, float64 (reciprocal first guess, filtered theta) where unsure
    const float u32[3] = {float(normal[0]), float(normal[1]), float(normal[2])};
    const float u_norm = sqrtf(u32[0] * u32[0] + u32[1] * u32[1] + u32[2] * u32[2]);
    int ia, ip, it;
    const bool counted = fpfh_pair_bins(rel, normal, u32, u_norm, nbr_normals + 3 * i, n_bins, edges, n_bins + 1, scale,
                                        lo32, scale32, true, ia, ip, it);
    if (counted != (d2 > 0.0)) {
      hist[0] = -2000000;
      return;
    }
    if (!counted) continue;
    {  // the filtered path must give the bins of the plain float64 path
      double a2, p2, theta;
      fpfh_features(rel, sqrt(d2), normal, nbr_normals + 3 * i, a2, p2, theta);
      if (ia != histogram_bin(a2, e[0], n_bins) || ip != histogram_bin(p2, e[1], n_bins) ||
          it != histogram_bin(theta, e[2], n_bins)) {
        hist[0] = -1000000;
        return;
      }
    }
    if (decorrelated) {
      if (ia >= 0) ++hist[ia];
      if (ip >= 0) ++hist[n_bins + ip];
      if (it >= 0) ++hist[2 * n_bins + it];
    } else if (ia >= 0 && ip >= 0 && it >= 0) {
      ++hist[(ia * n_bins + ip) * n_bins + it];
    }
  }
}

// One pair through the float32 filter alone: returns 1 and the three bins when the filter is sure, 0 otherwise.
int hm_fpfh_bins_fast(const double* rel, const double* u, const double* nj, int n_bins, const double* edges, int* bins) {
  float lo32[3], scale32[3];
  for (int f = 0; f < 3; ++f) {
    const double* e = edges + f * (n_bins + 1);
    lo32[f] = float(e[0]);
    scale32[f] = float(double(n_bins) / (e[n_bins] - e[0]));
  }
  const float rel32[3] = {float(rel[0]), float(rel[1]), float(rel[2])};
  const float u32[3] = {float(u[0]), float(u[1]), float(u[2])};
  const float nj32[3] = {float(nj[0]), float(nj[1]), float(nj[2])};
  const float u_norm = sqrtf(u32[0] * u32[0] + u32[1] * u32[1] + u32[2] * u32[2]);
  return fpfh_bins_fast(rel32, u32, u_norm, nj32, n_bins, lo32, scale32, bins[0], bins[1], bins[2]) ? 1 : 0;
}
// The same pair through the plain float64 path (np.histogram semantics): bins, or returns 0 when d == 0.
int hm_fpfh_bins_float64(const double* rel, const double* u, const double* nj, int n_bins, const double* edges, int* bins) {
  const double d2 = rdist3(rel[0], rel[1], rel[2]);
  if (!(d2 > 0.0)) return 0;
  double a, p, theta;
  fpfh_features(rel, sqrt(d2), u, nj, a, p, theta);
  bins[0] = histogram_bin(a, edges, n_bins);
  bins[1] = histogram_bin(p, edges + (n_bins + 1), n_bins);
  bins[2] = histogram_bin(theta, edges + 2 * (n_bins + 1), n_bins);
  return 1;
}

// Batch of pairs: stats[0] = pairs the filter was sure about, stats[1] = of those, pairs whose bins differ from the
// float64 path's (or that the float64 path drops) — must be 0; first_bad = index of the first such pair or -1.
void hm_fpfh_fast_check(long n_pairs, const double* rel, const double* u, const double* nj, int n_bins,
                        const double* edges, long* stats, long* first_bad) {
  stats[0] = stats[1] = 0;
  *first_bad = -1;
  for (long i = 0; i < n_pairs; ++i) {
    int fast[3], exact[3];
    if (!hm_fpfh_bins_fast(rel + 3 * i, u + 3 * i, nj + 3 * i, n_bins, edges, fast)) continue;
    ++stats[0];
    const int counted = hm_fpfh_bins_float64(rel + 3 * i, u + 3 * i, nj + 3 * i, n_bins, edges, exact);
    if (!counted || fast[0] != exact[0] || fast[1] != exact[1] || fast[2] != exact[2]) {
      ++stats[1];
      if (*first_bad < 0) *first_bad = i;
    }
  }
}

int hm_histogram_bin(double x, const double* edges, int n) { return histogram_bin(x, edges, n); }
int hm_histogram_bin_scaled(double x, const double* edges, int n) {
  return histogram_bin_scaled(x, edges, n, double(n) / (edges[n] - edges[0]));
}
int hm_theta_bin_float64(double ny, double nx, const double* edges, int n) {  // the unfiltered path
  return histogram_bin(atan2(ny, nx), edges, n);
}
int hm_theta_bin(double ny, double nx, const double* edges, int n) {
  return fpfh_theta_bin(ny, nx, edges, n, double(n) / (edges[n] - edges[0]));
}

}  // extern "C"
