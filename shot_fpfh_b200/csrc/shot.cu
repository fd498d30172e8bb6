// Kernel group S: SHOT local reference frames and 352-bin descriptors, one warp per query point.
//   S1 shot_lrf_kernel        <- get_local_rf, shot.py:16-48 (fan-out shot_parallelization.py:46-84)
//   S2 shot_descriptor_kernel <- compute_single_shot_descriptor, shot.py:175-306 (fan-out :86-133)
//
// S2 implements the reference's actual semantics: its ten `descriptor[idx] += v`
// statements are buffered NumPy fancy-index updates, so per statement and per bin only the neighbour with the
// largest distance that addresses the bin contributes (even with value 0). Each warp owns compact winner tables in
// shared memory (sf_math.cuh: three uint32 key tables raised with the native 32-bit atomicMax, five float value
// tables written by the lanes that still hold a slot's key after a warp barrier); bins then sum their statements.
#include <cub/cub.cuh>

#include "sf_common.cuh"

namespace sf {

// ---- S1 ----------------------------------------------------------------------------------------------------
// Three launches; the (nq, 3, 3) output buffer doubles as scratch between them (9 float64 per query):
//   lrf_moments_kernel  warp per query   : 6 weighted second moments by warp reduction          -> out[0..5], K -> out[6]
//   lrf_eigen_kernel    THREAD per query : LAPACK-path 3x3 eigensolver (a long serial chain: run 32 independent
//                                          ones per warp instead of one per warp)               -> x, z axes in out[0..5]
//   lrf_votes_kernel    warp per query   : sign votes, y = z cross x, final row-major frame      -> out[0..8]
// (Round 1 history: one warp doing all three for one query took 690 us at C2, a warp per batch of 32 queries 240 us
// but with only 3 200 warps in flight for the two gather passes; this split keeps 100k warps available for them.)
__global__ void __launch_bounds__(256)
    lrf_moments_kernel(GridView g, const double* __restrict__ queries, int64_t nq, double radius,
                       const int64_t* __restrict__ offsets, const int32_t* __restrict__ nbr, double* __restrict__ lrf) {
  const int lane = threadIdx.x & 31;
  const int64_t q = (blockIdx.x * int64_t(blockDim.x) + threadIdx.x) >> 5;
  if (q >= nq) return;
  const double qx = __ldg(queries + 3 * q), qy = __ldg(queries + 3 * q + 1), qz = __ldg(queries + 3 * q + 2);
  const int64_t begin = offsets[q], end = offsets[q + 1];
  // weighted covariance, weights (radius - distance), over ALL neighbours incl. the query itself (F5)
  double sw = 0, m[6] = {0, 0, 0, 0, 0, 0};
  for (int64_t i = begin + lane; i < end; i += 32) {
    const double4 p = load_pt(g.pts + __ldg(nbr + i));
    const double cx = p.x - qx, cy = p.y - qy, cz = p.z - qz;
    const double w = radius - sqrt(rdist3(cx, cy, cz));
    sw += w;
    m[0] += w * cx * cx; m[1] += w * cx * cy; m[2] += w * cx * cz;
    m[3] += w * cy * cy; m[4] += w * cy * cz; m[5] += w * cz * cz;
  }
  sw = warp_sum(sw);
#pragma unroll
  for (int k = 0; k < 6; ++k) m[k] = warp_sum(m[k]);
  if (lane < 6) {
    double v = m[0];
#pragma unroll
    for (int k = 1; k < 6; ++k) v = lane == k ? m[k] : v;
    lrf[9 * q + lane] = end > begin ? v / sw : 0.0;
  }
}

__global__ void __launch_bounds__(128)
    lrf_eigen_kernel(int64_t nq, const int64_t* __restrict__ offsets, const int32_t* __restrict__ counts,
                     double* __restrict__ lrf) {
  const int64_t q = blockIdx.x * int64_t(blockDim.x) + threadIdx.x;
  if (q >= nq) return;
  if (counts ? counts[q] == 0 : offsets[q + 1] == offsets[q]) return;  // empty: the votes step writes the identity
  double m[6];
#pragma unroll
  for (int k = 0; k < 6; ++k) m[k] = lrf[9 * q + k];
  double eval[3], evec[3][3];
  eigh3(m, eval, evec);
  lrf[9 * q + 0] = evec[2][0]; lrf[9 * q + 1] = evec[2][1]; lrf[9 * q + 2] = evec[2][2];  // x: largest eigenvalue
  lrf[9 * q + 3] = evec[0][0]; lrf[9 * q + 4] = evec[0][1]; lrf[9 * q + 5] = evec[0][2];  // z: smallest eigenvalue
}

__global__ void __launch_bounds__(256)
    lrf_votes_kernel(GridView g, const double* __restrict__ queries, int64_t nq, const int64_t* __restrict__ offsets,
                     const int32_t* __restrict__ nbr, double* __restrict__ lrf) {
  const int lane = threadIdx.x & 31;
  const int64_t q = (blockIdx.x * int64_t(blockDim.x) + threadIdx.x) >> 5;
  if (q >= nq) return;
  const int64_t begin = offsets[q], end = offsets[q + 1];
  double* out = lrf + 9 * q;
  if (end == begin) {  // shot.py:24-25
    if (lane < 9) out[lane] = (lane % 4 == 0) ? 1.0 : 0.0;
    return;
  }
  const double qx = __ldg(queries + 3 * q), qy = __ldg(queries + 3 * q + 1), qz = __ldg(queries + 3 * q + 2);
  double x[3] = {out[0], out[1], out[2]}, z[3] = {out[3], out[4], out[5]};
  // sign votes (shot.py:40-45): flip when strictly more neighbours project negatively than non-negatively
  int neg_x = 0, neg_z = 0;
  for (int64_t i = begin + lane; i < end; i += 32) {
    const double4 p = load_pt(g.pts + __ldg(nbr + i));
    const double cx = p.x - qx, cy = p.y - qy, cz = p.z - qz;
    neg_x += (cx * x[0] + cy * x[1] + cz * x[2]) < 0.0;
    neg_z += (cx * z[0] + cy * z[1] + cz * z[2]) < 0.0;
  }
  neg_x = warp_sum(neg_x);
  neg_z = warp_sum(neg_z);
  const int k_all = int(end - begin);
  if (neg_x > k_all - neg_x) { x[0] = -x[0]; x[1] = -x[1]; x[2] = -x[2]; }
  if (neg_z > k_all - neg_z) { z[0] = -z[0]; z[1] = -z[1]; z[2] = -z[2]; }
  const double y[3] = {z[1] * x[2] - z[2] * x[1], z[2] * x[0] - z[0] * x[2], z[0] * x[1] - z[1] * x[0]};
  __syncwarp();  // every lane has read the axes before the frame overwrites them
  if (lane < 3) {  // row `lane` of the matrix whose columns are [x y z] (selects, not indexing: stays in registers)
    out[3 * lane + 0] = lane == 0 ? x[0] : (lane == 1 ? x[1] : x[2]);
    out[3 * lane + 1] = lane == 0 ? y[0] : (lane == 1 ? y[1] : y[2]);
    out[3 * lane + 2] = lane == 0 ? z[0] : (lane == 1 ? z[1] : z[2]);
  }
}

// ---- fused single-scale driver: search + moments in ONE pass over the candidates ----------------------------------
// The generic path scans the 27 cells twice (count, fill) and gathers the neighbours twice more for the frame
// (moments, votes). For the single-scale driver (shot_parallelization.py:135-183, the pipeline's default and the
// benchmark's headline) the neighbour list is an internal temporary, so it can be PADDED: query q owns the slots
// [cand_offsets[q], cand_offsets[q+1]) sized by its candidate count (a cell_start lookup, no distance test), the
// one scan writes the hits there, counts them and accumulates the frame's weighted moments while the points are
// in registers. The votes then run inside the descriptor kernel.
__global__ void __launch_bounds__(256)
    candidate_count_kernel(GridView g, const double* __restrict__ queries, int64_t nq, int64_t* __restrict__ cand) {
  const int64_t q = blockIdx.x * int64_t(blockDim.x) + threadIdx.x;
  if (q >= nq) return;
  const int cx = cell_coord(queries[3 * q], g.origin[0], g.inv_cell, g.dims[0]);
  const int cy = cell_coord(queries[3 * q + 1], g.origin[1], g.inv_cell, g.dims[1]);
  const int cz = cell_coord(queries[3 * q + 2], g.origin[2], g.inv_cell, g.dims[2]);
  const int x0 = max(cx - 1, 0), x1 = min(cx + 1, g.dims[0] - 1);
  int total = 0;
  if (x0 <= x1) {
    for (int dz = -1; dz <= 1; ++dz)
      for (int dy = -1; dy <= 1; ++dy) {
        const int yy = cy + dy, zz = cz + dz;
        if (yy < 0 || yy >= g.dims[1] || zz < 0 || zz >= g.dims[2]) continue;
        const int64_t base = (int64_t(zz) * g.dims[1] + yy) * g.dims[0];
        total += __ldg(g.cell_start + base + x1 + 1) - __ldg(g.cell_start + base + x0);
      }
  }
  cand[q] = total;
}

__global__ void __launch_bounds__(256)
    search_moments_kernel(GridView g, const double* __restrict__ queries, int64_t nq, double radius, double r2,
                          const int64_t* __restrict__ cand_offsets, int32_t* __restrict__ nbr,
                          int32_t* __restrict__ counts, double* __restrict__ lrf,
                          unsigned long long* __restrict__ pair_counter) {
  const int lane = threadIdx.x & 31;
  const int64_t q = (blockIdx.x * int64_t(blockDim.x) + threadIdx.x) >> 5;
  if (q >= nq) return;
  const double qx = __ldg(queries + 3 * q), qy = __ldg(queries + 3 * q + 1), qz = __ldg(queries + 3 * q + 2);
  const Runs runs = build_runs(g, qx, qy, qz, lane);
  const int total = runs.pref[9];
  int64_t out = cand_offsets[q];
  int count = 0;
  double sw = 0, m[6] = {0, 0, 0, 0, 0, 0};
  // software pipeline: the candidate of the lane's NEXT round is requested before the current one is processed (the
  // kernel waits on these loads: 0.196 -> 0.173 ms; two rounds ahead spills and is slower, 0.182 ms)
  int pos_next = 0;
  double4 p_next = make_double4(0, 0, 0, 0);
  if (lane < total) {
    pos_next = run_position(runs, lane);
    p_next = load_pt(g.pts + pos_next);
  }
  for (int base = 0; base < total; base += 32) {
    const int v = base + lane;
    bool hit = false;
    const int pos = pos_next;
    const double4 p = p_next;
    if (v + 32 < total) {
      pos_next = run_position(runs, v + 32);
      p_next = load_pt(g.pts + pos_next);
    }
    if (v < total) {
      const double cx = qx - p.x, cy = qy - p.y, cz = qz - p.z;  // second moments do not see the sign
      const double d2 = rdist3(cx, cy, cz);
      hit = d2 <= r2;
      if (hit) {
        const double w = radius - sqrt(d2);
        sw += w;
        m[0] += w * cx * cx; m[1] += w * cx * cy; m[2] += w * cx * cz;
        m[3] += w * cy * cy; m[4] += w * cy * cz; m[5] += w * cz * cz;
      }
    }
    const unsigned mask = __ballot_sync(kFull, hit);
    if (hit) nbr[out + __popc(mask & lanemask_lt())] = pos;
    out += __popc(mask);
    count += __popc(mask);
  }
  sw = warp_sum(sw);
#pragma unroll
  for (int k = 0; k < 6; ++k) m[k] = warp_sum(m[k]);
  if (lane < 6) {
    double v = m[0];
#pragma unroll
    for (int k = 1; k < 6; ++k) v = lane == k ? m[k] : v;
    lrf[9 * q + lane] = count > 0 ? v / sw : 0.0;
  }
  if (lane == 0) {
    counts[q] = count;
    atomicAdd(pair_counter, static_cast<unsigned long long>(count));
  }
}

// ---- S2 ----------------------------------------------------------------------------------------------------
// Per warp: uint32 keys[3][352] + float vals[5][352] = 11 264 B of shared memory (sf_math.cuh, "winner tables,
// compact form"). Per 32 neighbours: (a) float64 decisions, (b) three native 32-bit atomicMax on the key tables,
// (c) warp barrier, (d) the lanes that hold a slot's key compute the transcendental weights they need (only winners
// pay for acosf / atan2f) and store their values. After the last neighbour each lane assembles 11 bins.
constexpr int kShotWarpsPerBlock = 4;
constexpr int kShotSmemPerWarp = (kKeyCount + kValCount) * 4;

// 352 * sizeof(OutT) is a multiple of 16, so every row and every group of four bins is 16-byte aligned
__device__ __forceinline__ void store_group(float* p, float a, float b, float c, float d) {
  *reinterpret_cast<float4*>(p) = make_float4(a, b, c, d);
}
__device__ __forceinline__ void store_group(double* p, float a, float b, float c, float d) {
  reinterpret_cast<double2*>(p)[0] = make_double2(a, b);
  reinterpret_cast<double2*>(p)[1] = make_double2(c, d);
}

template <typename OutT>
__global__ void __launch_bounds__(kShotWarpsPerBlock * 32, 5)
    shot_descriptor_kernel(GridView g, const double* __restrict__ queries, int64_t nq, double radius,
                           const int64_t* __restrict__ offsets, const int32_t* __restrict__ counts,
                           const int32_t* __restrict__ nbr, double* __restrict__ lrf, int fuse_votes, int min_nb,
                           int normalize, OutT* __restrict__ out) {
  // counts == nullptr: classic CSR, neighbours of q are nbr[offsets[q] .. offsets[q+1]). Otherwise a padded list:
  // nbr[offsets[q] .. offsets[q] + counts[q]) (the fused single-scale driver).
  // fuse_votes: lrf[9q + 0..5] holds the RAW eigenvectors (x, z) from lrf_eigen_kernel; the sign votes of
  // shot.py:40-45 run here (the second pass over the same neighbours then hits L1) and the final frame is written
  // back to lrf[9q + 0..8].
  extern __shared__ uint32_t table_mem[];
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  uint32_t* keys = table_mem + warp * (kKeyCount + kValCount);
  float* vals = reinterpret_cast<float*>(keys + kKeyCount);
  const int64_t warps_total = int64_t(gridDim.x) * kShotWarpsPerBlock;
  const double inv_radius = 1.0 / radius;
  for (int64_t q = blockIdx.x * int64_t(kShotWarpsPerBlock) + warp; q < nq; q += warps_total) {
    // values are gated by their keys: only the keys are cleared (16 bytes per lane and store)
    for (int j = lane; j < kKeyCount / 4; j += 32) reinterpret_cast<uint4*>(keys)[j] = make_uint4(0u, 0u, 0u, 0u);
    const double qx = __ldg(queries + 3 * q), qy = __ldg(queries + 3 * q + 1), qz = __ldg(queries + 3 * q + 2);
    const int64_t begin = offsets[q], end = counts ? begin + counts[q] : offsets[q + 1];
    double f[9];
    if (!fuse_votes) {
#pragma unroll
      for (int k = 0; k < 9; ++k) f[k] = lrf[9 * q + k];
    } else if (end == begin) {
#pragma unroll
      for (int k = 0; k < 9; ++k) f[k] = (k % 4 == 0) ? 1.0 : 0.0;  // shot.py:24-25
      __syncwarp();
      if (lane < 9) lrf[9 * q + lane] = (lane % 4 == 0) ? 1.0 : 0.0;
    } else {
      double x[3] = {lrf[9 * q], lrf[9 * q + 1], lrf[9 * q + 2]}, z[3] = {lrf[9 * q + 3], lrf[9 * q + 4], lrf[9 * q + 5]};
      int neg_x = 0, neg_z = 0;
      for (int64_t i = begin + lane; i < end; i += 32) {
        const double4 p = load_pt(g.pts + __ldg(nbr + i));
        const double cx = p.x - qx, cy = p.y - qy, cz = p.z - qz;
        neg_x += (cx * x[0] + cy * x[1] + cz * x[2]) < 0.0;
        neg_z += (cx * z[0] + cy * z[1] + cz * z[2]) < 0.0;
      }
      neg_x = warp_sum(neg_x);
      neg_z = warp_sum(neg_z);
      const int k_all = int(end - begin);
      if (neg_x > k_all - neg_x) { x[0] = -x[0]; x[1] = -x[1]; x[2] = -x[2]; }
      if (neg_z > k_all - neg_z) { z[0] = -z[0]; z[1] = -z[1]; z[2] = -z[2]; }
      const double y[3] = {z[1] * x[2] - z[2] * x[1], z[2] * x[0] - z[0] * x[2], z[0] * x[1] - z[1] * x[0]};
#pragma unroll
      for (int a = 0; a < 3; ++a) { f[3 * a] = x[a]; f[3 * a + 1] = y[a]; f[3 * a + 2] = z[a]; }
      __syncwarp();  // every lane has read the raw axes before the frame overwrites them
      if (lane == 0) {
#pragma unroll
        for (int k = 0; k < 9; ++k) lrf[9 * q + k] = f[k];
      }
    }
    int positive = 0;
    // software pipeline: the gathers of the lane's NEXT neighbour are issued before the current one is processed
    // (the neighbour INDEX runs one round further ahead still, so that a gather never waits for its address)
    int64_t i = begin + lane;
    double4 p_next = make_double4(0, 0, 0, 0), n_next = p_next;
    int s_after = 0;
    if (i < end) {
      const int s = __ldg(nbr + i);
      p_next = load_pt(g.pts + s);
      n_next = load_pt(g.nrm + s);
    }
    if (i + 32 < end) s_after = __ldg(nbr + i + 32);
    __syncwarp();
    for (int64_t base = begin; base < end; base += 32, i += 32) {  // warp-uniform trip count (barriers inside)
      const double4 p = p_next, n = n_next;
      if (i + 32 < end) {
        const int s = s_after;
        p_next = load_pt(g.pts + s);
        n_next = load_pt(g.nrm + s);
        if (i + 64 < end) s_after = __ldg(nbr + i + 64);
      }
      ShotDecision d;
      bool active = false;
      if (i < end) {
        const double cx = p.x - qx, cy = p.y - qy, cz = p.z - qz;
        const double d2 = rdist3(cx, cy, cz);
        if (d2 > 0.0) {  // shot.py:213: neighbours at distance 0 (the query itself, duplicates) are dropped
          active = true;
          ++positive;
          const double rho = sqrt(d2);
          const double X = cx * f[0] + cy * f[3] + cz * f[6];
          const double Y = cx * f[1] + cy * f[4] + cz * f[7];
          const double Z = cx * f[2] + cy * f[5] + cz * f[8];
          double cosine = n.x * f[2] + n.y * f[5] + n.z * f[8];
          cosine = fmin(1.0, fmax(-1.0, cosine));
          d = shot_decide(X, Y, Z, cosine, rho, radius, inv_radius);
          atomicMax(keys + kKeyOwn + d.own, d.key);
          atomicMax(keys + kKeyCos + d.cos_nb, d.key);
          atomicMax(keys + kKeyAz + d.az_nb, d.key);
        }
      }
      __syncwarp();
      if (active) {
        const bool win_own = keys[kKeyOwn + d.own] == d.key;
        const bool win_cos = keys[kKeyCos + d.cos_nb] == d.key;
        const bool win_az = keys[kKeyAz + d.az_nb] == d.key;
        float a_az = 0.0f;
        if (win_own || win_az) a_az = shot_azimuth(d);
        if (win_own) {
          float own_vol, other_vol;
          shot_elevation(d, own_vol, other_vol);
          vals[kValOwn + d.own] = (1.0f - d.a_cos) + d.own_shell + own_vol + (1.0f - a_az);
          vals[kValRad + d.own] = d.other_shell;
          vals[kValEl + d.own] = other_vol;
        }
        if (win_cos) vals[kValCos + d.cos_nb] = d.a_cos;
        if (win_az) vals[kValAz + d.az_nb] = a_az;
      }
      __syncwarp();  // the stores above are ordered before the next round's key updates
    }
    positive = warp_sum(positive);
    // assembly: lane handles the groups {lane, lane + 32, lane + 64} of four consecutive bins (one (cosine, azimuth)
    // cell: the radial / elevation partners are inside the group) -> eight 16-byte table reads and one 16-byte
    // store per group
    constexpr int kGroups = kShotLen / 4, kGroupRounds = (kGroups + 31) / 32;
    float v[kGroupRounds][4];
    double sq = 0.0;
#pragma unroll
    for (int j = 0; j < kGroupRounds; ++j) {
      const int grp = lane + 32 * j;
      if (grp < kGroups) {
        const uint4 ko = *reinterpret_cast<const uint4*>(keys + kKeyOwn + 4 * grp);
        const uint4 kc = *reinterpret_cast<const uint4*>(keys + kKeyCos + 4 * grp);
        const uint4 ka = *reinterpret_cast<const uint4*>(keys + kKeyAz + 4 * grp);
        const float4 vo = *reinterpret_cast<const float4*>(vals + kValOwn + 4 * grp);
        const float4 vr = *reinterpret_cast<const float4*>(vals + kValRad + 4 * grp);
        const float4 ve = *reinterpret_cast<const float4*>(vals + kValEl + 4 * grp);
        const float4 vc = *reinterpret_cast<const float4*>(vals + kValCos + 4 * grp);
        const float4 va = *reinterpret_cast<const float4*>(vals + kValAz + 4 * grp);
        const uint32_t ko_[4] = {ko.x, ko.y, ko.z, ko.w}, kc_[4] = {kc.x, kc.y, kc.z, kc.w},
                       ka_[4] = {ka.x, ka.y, ka.z, ka.w};
        const float vo_[4] = {vo.x, vo.y, vo.z, vo.w}, vr_[4] = {vr.x, vr.y, vr.z, vr.w},
                    ve_[4] = {ve.x, ve.y, ve.z, ve.w}, vc_[4] = {vc.x, vc.y, vc.z, vc.w},
                    va_[4] = {va.x, va.y, va.z, va.w};
        shot_bin_group_compact(ko_, kc_, ka_, vo_, vr_, ve_, vc_, va_, v[j]);
#pragma unroll
        for (int t = 0; t < 4; ++t) sq += double(v[j][t]) * double(v[j][t]);
      }
    }
    sq = warp_sum(sq);
    const double norm = sqrt(sq);
    // shot.py:212, :301-306: zero row when too few neighbours or a zero norm
    const bool keep = positive > min_nb && norm > 0.0;
    const float inv = keep ? (normalize ? float(1.0 / norm) : 1.0f) : 0.0f;
    OutT* row = out + q * kShotLen;
#pragma unroll
    for (int j = 0; j < kGroupRounds; ++j) {
      const int grp = lane + 32 * j;
      if (grp < kGroups) store_group(row + 4 * grp, v[j][0] * inv, v[j][1] * inv, v[j][2] * inv, v[j][3] * inv);
    }
    __syncwarp();  // all lanes have read the tables before the next query clears the keys
  }
}

}  // namespace sf

using namespace sf;

extern "C" int sf_shot_lrf(sf_grid* g, const double* queries, int64_t nq, double radius, const int64_t* offsets,
                           const int32_t* nbr, double* lrf, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  SF_REQUIRE(g != nullptr && g->n > 0, SF_ERR_ARG, "sf_shot_lrf: grid not built");
  SF_REQUIRE(queries && offsets && lrf && nq >= 0, SF_ERR_ARG, "sf_shot_lrf: bad arguments");
  if (nq == 0) return SF_OK;
  const unsigned warp_blocks = unsigned((nq * 32 + 255) / 256);
  lrf_moments_kernel<<<warp_blocks, 256, 0, stream>>>(g->view(), queries, nq, radius, offsets, nbr, lrf);
  lrf_eigen_kernel<<<unsigned((nq + 127) / 128), 128, 0, stream>>>(nq, offsets, nullptr, lrf);
  lrf_votes_kernel<<<warp_blocks, 256, 0, stream>>>(g->view(), queries, nq, offsets, nbr, lrf);
  SF_CUDA(cudaGetLastError());
  return SF_OK;
}

// Optional per-kernel timing of the fused drivers (bench.py's roofline needs the dominant kernel's own duration and
// a driver is a single C call): CUDA events recorded on the launching stream around its three stages.
static bool g_profile = false;
static cudaEvent_t g_events[4] = {nullptr, nullptr, nullptr, nullptr};

namespace sf {
void profile_mark(int i, cudaStream_t stream) {
  if (g_profile) cudaEventRecord(g_events[i], stream);
}
}  // namespace sf

extern "C" int sf_profile_enable(int32_t enable) {
  if (enable && g_events[0] == nullptr)
    for (auto& e : g_events) SF_CUDA(cudaEventCreate(&e));
  g_profile = enable != 0;
  return SF_OK;
}

// ms_out[0..2] = the three stages of the LAST fused driver call (sf_shot_single_scale: search+moments, eigen,
// votes+descriptor; sf_fpfh_cloud: search+weights, SPFH, FPFH). Synchronises.
extern "C" int sf_profile_read(float* ms_out) {
  SF_REQUIRE(g_profile && g_events[0] != nullptr && ms_out != nullptr, SF_ERR_ARG, "sf_profile_read: profiling is off");
  SF_CUDA(cudaEventSynchronize(g_events[3]));
  for (int i = 0; i < 3; ++i) SF_CUDA(cudaEventElapsedTime(ms_out + i, g_events[i], g_events[i + 1]));
  return SF_OK;
}

static int launch_descriptor(sf_grid* g, const double* queries, int64_t nq, double radius, const int64_t* offsets,
                             const int32_t* counts, const int32_t* nbr, double* lrf, int fuse_votes, int min_nb,
                             int normalize, void* out, int out_is_f64, cudaStream_t stream) {
  const size_t smem = size_t(kShotWarpsPerBlock) * kShotSmemPerWarp;
  SF_REQUIRE(reinterpret_cast<uintptr_t>(out) % 16 == 0, SF_ERR_ARG, "SHOT output rows must be 16-byte aligned");
  static bool configured = false;
  if (!configured) {
    SF_CUDA(cudaFuncSetAttribute(shot_descriptor_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)));
    SF_CUDA(cudaFuncSetAttribute(shot_descriptor_kernel<double>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)));
    configured = true;
  }
  // persistent-style launch: 148 SMs x 5 resident blocks (44 KB shared memory each), capped by the work
  const int64_t blocks_needed = (nq + kShotWarpsPerBlock - 1) / kShotWarpsPerBlock;
  const unsigned blocks = unsigned(blocks_needed < 148 * 5 ? blocks_needed : 148 * 5);
  if (out_is_f64)
    shot_descriptor_kernel<double><<<blocks, kShotWarpsPerBlock * 32, smem, stream>>>(
        g->view(), queries, nq, radius, offsets, counts, nbr, lrf, fuse_votes, min_nb, normalize, static_cast<double*>(out));
  else
    shot_descriptor_kernel<float><<<blocks, kShotWarpsPerBlock * 32, smem, stream>>>(
        g->view(), queries, nq, radius, offsets, counts, nbr, lrf, fuse_votes, min_nb, normalize, static_cast<float*>(out));
  SF_CUDA(cudaGetLastError());
  return SF_OK;
}

extern "C" int sf_shot_descriptor(sf_grid* g, const double* queries, int64_t nq, double radius,
                                  const int64_t* offsets, const int32_t* nbr, const double* lrf, int32_t min_nb,
                                  int32_t normalize, void* out, int32_t out_is_f64, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  SF_REQUIRE(g != nullptr && g->n > 0 && g->has_normals, SF_ERR_ARG, "sf_shot_descriptor: grid built without normals");
  SF_REQUIRE(queries && offsets && lrf && out && nq >= 0, SF_ERR_ARG, "sf_shot_descriptor: bad arguments");
  if (nq == 0) return SF_OK;
  return launch_descriptor(g, queries, nq, radius, offsets, nullptr, nbr, const_cast<double*>(lrf), 0, min_nb, normalize,
                           out, out_is_f64, stream);
}

extern "C" int sf_shot_single_scale(sf_grid* g, const double* queries, int64_t nq, double radius, int32_t min_nb,
                                    int32_t normalize, void* out, int32_t out_is_f64, double* lrf_out,
                                    int64_t* pairs_host, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  SF_REQUIRE(g != nullptr && g->n > 0 && g->has_normals, SF_ERR_ARG, "sf_shot_single_scale: grid built without normals");
  SF_REQUIRE(queries && out && nq >= 0, SF_ERR_ARG, "sf_shot_single_scale: bad arguments");
  SF_REQUIRE(radius > 0.0 && radius * 1.0005 <= g->cell, SF_ERR_ARG,
             "sf_shot_single_scale: radius %g exceeds the cell edge %g the grid was built for", radius, g->cell);
  if (pairs_host) *pairs_host = 0;
  if (nq == 0) return SF_OK;
  int64_t *cand = nullptr, *cand_offsets = nullptr;
  int32_t *counts = nullptr, *nbr = nullptr;
  double* lrf = lrf_out;
  void* scan_temp = nullptr;
  size_t scan_bytes = 0;
  cub::DeviceScan::ExclusiveSum(nullptr, scan_bytes, cand, cand_offsets, int(nq + 1), stream);
  SF_CUDA(scratch_alloc(reinterpret_cast<void**>(&cand), size_t(nq + 1) * 8, stream));
  SF_CUDA(scratch_alloc(reinterpret_cast<void**>(&cand_offsets), size_t(nq + 1) * 8, stream));
  SF_CUDA(scratch_alloc(reinterpret_cast<void**>(&counts), size_t(nq) * 4, stream));
  SF_CUDA(scratch_alloc(&scan_temp, scan_bytes + 16, stream));
  if (lrf == nullptr) SF_CUDA(scratch_alloc(reinterpret_cast<void**>(&lrf), size_t(nq) * 9 * 8, stream));
  const GridView view = g->view();
  SF_CUDA(cudaMemsetAsync(cand + nq, 0, 8, stream));
  unsigned long long* pair_counter = nullptr;
  SF_CUDA(scratch_alloc(reinterpret_cast<void**>(&pair_counter), 8, stream));
  SF_CUDA(cudaMemsetAsync(pair_counter, 0, 8, stream));
  candidate_count_kernel<<<unsigned((nq + 255) / 256), 256, 0, stream>>>(view, queries, nq, cand);
  SF_CUDA(cub::DeviceScan::ExclusiveSum(scan_temp, scan_bytes, cand, cand_offsets, int(nq + 1), stream));
  int64_t total = 0;
  SF_CUDA(cudaMemcpyAsync(&total, cand_offsets + nq, 8, cudaMemcpyDeviceToHost, stream));
  SF_CUDA(cudaStreamSynchronize(stream));
  SF_CUDA(scratch_alloc(reinterpret_cast<void**>(&nbr), size_t(total > 0 ? total : 1) * 4, stream));
  const unsigned warp_blocks = unsigned((nq * 32 + 255) / 256);
  profile_mark(0, stream);
  search_moments_kernel<<<warp_blocks, 256, 0, stream>>>(view, queries, nq, radius, radius * radius, cand_offsets, nbr,
                                                        counts, lrf, pair_counter);
  profile_mark(1, stream);
  lrf_eigen_kernel<<<unsigned((nq + 127) / 128), 128, 0, stream>>>(nq, cand_offsets, counts, lrf);
  profile_mark(2, stream);
  int rc = launch_descriptor(g, queries, nq, radius, cand_offsets, counts, nbr, lrf, 1, min_nb, normalize, out, out_is_f64,
                             stream);
  profile_mark(3, stream);
  if (rc == SF_OK && pairs_host != nullptr) {  // neighbour pairs found (logging / algorithmic-byte accounting)
    unsigned long long pairs = 0;
    SF_CUDA(cudaMemcpyAsync(&pairs, pair_counter, 8, cudaMemcpyDeviceToHost, stream));
    SF_CUDA(cudaStreamSynchronize(stream));
    *pairs_host = int64_t(pairs);
  }
  void* to_free[] = {cand, cand_offsets, counts, nbr, scan_temp, pair_counter, lrf_out == nullptr ? lrf : nullptr};
  for (void* p : to_free)
    if (p) cudaFreeAsync(p, stream);
  return rc;
}
