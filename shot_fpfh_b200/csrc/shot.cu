// Kernel group S: SHOT local reference frames and 352-bin descriptors, one warp per query point.
//   S1 shot_lrf_kernel        <- get_local_rf, shot.py:16-48 (fan-out shot_parallelization.py:46-84)
//   S2 shot_descriptor_kernel <- compute_single_shot_descriptor, shot.py:175-306 (fan-out :86-133)
//
// S2 implements the reference's actual semantics: its ten `descriptor[idx] += v`
// statements are buffered NumPy fancy-index updates, so per statement and per bin only the neighbour with the
// largest distance that addresses the bin contributes (even with value 0). Each warp owns a table of 1760
// 64-bit words in shared memory, one word per (statement group, bin): (distance key << 32) | float value, updated
// with atomicMax. The largest key wins and carries its value along; bins then sum their five tables.
#include "sf_common.cuh"

namespace sf {

// ---- S1 ----------------------------------------------------------------------------------------------------
// One warp owns a batch of 32 consecutive queries. Phase A: the 32 lanes stride over the neighbours of query j and
// reduce its 7 weighted moments, which lane j keeps. Phase B: every lane solves ITS OWN query's 3x3 eigenproblem
// (the LAPACK-path solver is a long serial chain: one per lane = 32 in flight, instead of 32 lanes redundantly
// solving one). Phase C: the axes of lane j are broadcast and the warp counts the sign votes of query j.
__global__ void __launch_bounds__(128)
    shot_lrf_kernel(GridView g, const double* __restrict__ queries, int64_t nq, double radius,
                    const int64_t* __restrict__ offsets, const int32_t* __restrict__ nbr, double* __restrict__ lrf) {
  const int lane = threadIdx.x & 31;
  const int64_t q0 = ((blockIdx.x * int64_t(blockDim.x) + threadIdx.x) >> 5) * 32;
  if (q0 >= nq) return;
  const int batch = int(nq - q0 < 32 ? nq - q0 : 32);
  // this lane's query
  const int64_t mine = q0 + (lane < batch ? lane : 0);
  const int64_t my_begin = __ldg(offsets + mine), my_end = __ldg(offsets + mine + 1);
  const double my_qx = __ldg(queries + 3 * mine), my_qy = __ldg(queries + 3 * mine + 1), my_qz = __ldg(queries + 3 * mine + 2);
  // ---- phase A: weighted covariance, weights (radius - distance), over ALL neighbours incl. the query itself (F5)
  double mom[6] = {0, 0, 0, 0, 0, 0};
  for (int j = 0; j < batch; ++j) {
    const int64_t begin = __shfl_sync(kFull, my_begin, j), end = __shfl_sync(kFull, my_end, j);
    const double qx = __shfl_sync(kFull, my_qx, j), qy = __shfl_sync(kFull, my_qy, j), qz = __shfl_sync(kFull, my_qz, j);
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
    for (int k = 0; k < 6; ++k) {
      const double v = warp_sum(m[k]) / sw;
      if (lane == j) mom[k] = v;
    }
  }
  // ---- phase B: one eigen-decomposition per lane
  double x[3] = {1, 0, 0}, z[3] = {0, 0, 1};
  const bool has_neighbours = lane < batch && my_end > my_begin;
  if (has_neighbours) {
    double eval[3], evec[3][3];
    eigh3(mom, eval, evec);
    x[0] = evec[2][0]; x[1] = evec[2][1]; x[2] = evec[2][2];  // largest eigenvalue
    z[0] = evec[0][0]; z[1] = evec[0][1]; z[2] = evec[0][2];  // smallest eigenvalue
  }
  // ---- phase C: sign votes (shot.py:40-45): flip when strictly more neighbours project negatively than not
  int my_neg_x = 0, my_neg_z = 0;
  for (int j = 0; j < batch; ++j) {
    const int64_t begin = __shfl_sync(kFull, my_begin, j), end = __shfl_sync(kFull, my_end, j);
    const double qx = __shfl_sync(kFull, my_qx, j), qy = __shfl_sync(kFull, my_qy, j), qz = __shfl_sync(kFull, my_qz, j);
    const double x0 = __shfl_sync(kFull, x[0], j), x1 = __shfl_sync(kFull, x[1], j), x2 = __shfl_sync(kFull, x[2], j);
    const double z0 = __shfl_sync(kFull, z[0], j), z1 = __shfl_sync(kFull, z[1], j), z2 = __shfl_sync(kFull, z[2], j);
    int neg_x = 0, neg_z = 0;
    for (int64_t i = begin + lane; i < end; i += 32) {
      const double4 p = load_pt(g.pts + __ldg(nbr + i));
      const double cx = p.x - qx, cy = p.y - qy, cz = p.z - qz;
      neg_x += (cx * x0 + cy * x1 + cz * x2) < 0.0;
      neg_z += (cx * z0 + cy * z1 + cz * z2) < 0.0;
    }
    neg_x = warp_sum(neg_x);
    neg_z = warp_sum(neg_z);
    if (lane == j) { my_neg_x = neg_x; my_neg_z = neg_z; }
  }
  if (lane < batch) {
    double* out = lrf + 9 * mine;
    if (!has_neighbours) {  // shot.py:24-25
#pragma unroll
      for (int k = 0; k < 9; ++k) out[k] = (k % 4 == 0) ? 1.0 : 0.0;
    } else {
      const int k_all = int(my_end - my_begin);
      if (my_neg_x > k_all - my_neg_x) { x[0] = -x[0]; x[1] = -x[1]; x[2] = -x[2]; }
      if (my_neg_z > k_all - my_neg_z) { z[0] = -z[0]; z[1] = -z[1]; z[2] = -z[2]; }
      const double y[3] = {z[1] * x[2] - z[2] * x[1], z[2] * x[0] - z[0] * x[2], z[0] * x[1] - z[1] * x[0]};
#pragma unroll
      for (int a = 0; a < 3; ++a) {  // row a of the matrix whose columns are [x y z]
        out[3 * a + 0] = x[a];
        out[3 * a + 1] = y[a];
        out[3 * a + 2] = z[a];
      }
    }
  }
}

// ---- S2 ----------------------------------------------------------------------------------------------------
// Per warp: uint32 keys[3][352] + float vals[5][352] = 11 264 B of shared memory (sf_math.cuh, "winner tables,
// compact form"). Per 32 neighbours: (a) float64 decisions, (b) three native 32-bit atomicMax on the key tables,
// (c) warp barrier, (d) the lanes that hold a slot's key compute the transcendental weights they need (only winners
// pay for acosf / atan2f) and store their values. After the last neighbour each lane assembles 11 bins.
constexpr int kShotWarpsPerBlock = 4;
constexpr int kShotSmemPerWarp = (kKeyCount + kValCount) * 4;

template <typename OutT>
__global__ void __launch_bounds__(kShotWarpsPerBlock * 32, 5)
    shot_descriptor_kernel(GridView g, const double* __restrict__ queries, int64_t nq, double radius,
                           const int64_t* __restrict__ offsets, const int32_t* __restrict__ nbr,
                           const double* __restrict__ lrf, int min_nb, int normalize, OutT* __restrict__ out) {
  extern __shared__ uint32_t table_mem[];
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  uint32_t* keys = table_mem + warp * (kKeyCount + kValCount);
  float* vals = reinterpret_cast<float*>(keys + kKeyCount);
  const int64_t warps_total = int64_t(gridDim.x) * kShotWarpsPerBlock;
  const double inv_radius = 1.0 / radius;
  for (int64_t q = blockIdx.x * int64_t(kShotWarpsPerBlock) + warp; q < nq; q += warps_total) {
#pragma unroll
    for (int j = 0; j < kKeyCount / 32; ++j) keys[lane + 32 * j] = 0u;  // values are gated by their keys: no clearing
    const double qx = __ldg(queries + 3 * q), qy = __ldg(queries + 3 * q + 1), qz = __ldg(queries + 3 * q + 2);
    const int64_t begin = offsets[q], end = offsets[q + 1];
    double f[9];
#pragma unroll
    for (int k = 0; k < 9; ++k) f[k] = __ldg(lrf + 9 * q + k);
    int positive = 0;
    // software pipeline: the gathers of the lane's NEXT neighbour are issued before the current one is processed
    int64_t i = begin + lane;
    double4 p_next = make_double4(0, 0, 0, 0), n_next = p_next;
    if (i < end) {
      const int s = __ldg(nbr + i);
      p_next = load_pt(g.pts + s);
      n_next = load_pt(g.nrm + s);
    }
    __syncwarp();
    for (int64_t base = begin; base < end; base += 32, i += 32) {  // warp-uniform trip count (barriers inside)
      const double4 p = p_next, n = n_next;
      if (i + 32 < end) {
        const int s = __ldg(nbr + i + 32);
        p_next = load_pt(g.pts + s);
        n_next = load_pt(g.nrm + s);
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
    float v[kShotLen / 32];
    double sq = 0.0;
#pragma unroll
    for (int j = 0; j < kShotLen / 32; ++j) {
      v[j] = shot_bin_value_compact(keys, vals, lane + 32 * j);
      sq += double(v[j]) * double(v[j]);
    }
    sq = warp_sum(sq);
    const double norm = sqrt(sq);
    // shot.py:212, :301-306: zero row when too few neighbours or a zero norm
    const bool keep = positive > min_nb && norm > 0.0;
    const float inv = keep ? (normalize ? float(1.0 / norm) : 1.0f) : 0.0f;
    OutT* row = out + q * kShotLen;
#pragma unroll
    for (int j = 0; j < kShotLen / 32; ++j) row[lane + 32 * j] = OutT(v[j] * inv);
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
  const int64_t warps = (nq + 31) / 32;  // one warp per batch of 32 queries
  shot_lrf_kernel<<<unsigned((warps + 3) / 4), 128, 0, stream>>>(g->view(), queries, nq, radius, offsets, nbr, lrf);
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
  const size_t smem = size_t(kShotWarpsPerBlock) * kShotSmemPerWarp;
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
        g->view(), queries, nq, radius, offsets, nbr, lrf, min_nb, normalize, static_cast<double*>(out));
  else
    shot_descriptor_kernel<float><<<blocks, kShotWarpsPerBlock * 32, smem, stream>>>(
        g->view(), queries, nq, radius, offsets, nbr, lrf, min_nb, normalize, static_cast<float*>(out));
  SF_CUDA(cudaGetLastError());
  return SF_OK;
}
