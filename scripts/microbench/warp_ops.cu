// Microbenchmark (tuning only): throughput of the warp-level primitives the SHOT winner selection could use, on one SM
// with 32 resident warps: cycles per warp-instruction per SM sub-partition.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o warp_ops warp_ops.cu && ./warp_ops
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

template <int OP>
__global__ void bench(uint32_t* out, long long* cycles, int iters, uint32_t seed) {
  __shared__ uint32_t table[32 * 352];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  uint32_t* mine = table + warp * 352;
  for (int i = lane; i < 352; i += 32) mine[i] = 0;
  __syncthreads();
  uint32_t x = seed * 2654435761u + threadIdx.x * 40503u, acc = 0;
  const long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
    x = x * 1664525u + 1013904223u;
    const uint32_t bin = (x >> 20) % 40u + 300u;  // ~40 distinct bins, as own bins of a smooth surface
    const uint32_t key = (x >> 3) | 1u;
    if (OP == 0) {  // baseline: the generator only
      acc += bin ^ key;
    } else if (OP == 1) {
      acc += __match_any_sync(0xffffffffu, bin);
    } else if (OP == 2) {
      acc += __reduce_max_sync(0xffffffffu, key);
    } else if (OP == 3) {
      const unsigned m = __match_any_sync(0xffffffffu, bin);
      acc += __reduce_max_sync(m, key);
    } else if (OP == 4) {
      atomicMax(mine + bin, key);
    } else if (OP == 5) {  // optimistic max-store, one pass
      const uint32_t o = mine[bin];
      if (key > o) mine[bin] = key;
      acc += o;
    } else if (OP == 6) {
      acc += __ballot_sync(0xffffffffu, key & 4u);
    } else if (OP == 7) {
      acc += __popc(key);
    } else if (OP == 8) {
      acc += __float_as_uint(rsqrtf(__uint_as_float((key & 0x7fffffu) | 0x3f800000u)));
    } else if (OP == 9) {
      acc += uint32_t(__uint_as_float((key & 0x7fffffu) | 0x3f800000u) * 1000.0f);
    } else if (OP == 10) {
      acc += __shfl_xor_sync(0xffffffffu, key, 5);
    }
  }
  const long long t1 = clock64();
  if (lane == 0) cycles[blockIdx.x * (blockDim.x >> 5) + warp] = t1 - t0;
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc + mine[lane];
}

template <int OP>
double run(int warps_per_sm, int iters) {
  uint32_t* out;
  long long* cyc;
  const int blocks = 148;
  cudaMalloc(&out, blocks * warps_per_sm * 32 * 4);
  cudaMalloc(&cyc, blocks * warps_per_sm * 8);
  bench<OP><<<blocks, warps_per_sm * 32>>>(out, cyc, iters, 1);
  bench<OP><<<blocks, warps_per_sm * 32>>>(out, cyc, iters, 2);
  cudaDeviceSynchronize();
  long long h[148 * 32];
  cudaMemcpy(h, cyc, blocks * warps_per_sm * 8, cudaMemcpyDeviceToHost);
  double mx = 0;
  for (int i = 0; i < blocks * warps_per_sm; ++i) mx = h[i] > mx ? h[i] : mx;
  cudaFree(out);
  cudaFree(cyc);
  // cycles per warp-instruction per sub-partition: warps_per_sm / 4 warps share a scheduler
  return mx / iters / (warps_per_sm / 4.0);
}

int main() {
  const char* names[] = {"baseline (lcg + address)", "match_any", "reduce_max (full mask)", "match_any + reduce_max(mask)",
                         "shared atomicMax", "optimistic max-store pass (LDS, cmp, STS)", "ballot", "popc", "rsqrtf (MUFU)",
                         "float -> uint (F2I)", "shfl_xor"};
  const int iters = 4000;
  for (int w : {4, 32}) {
    printf("%d warps per SM: cycles per loop iteration per sub-partition warp\n", w);
    double r[11];
    r[0] = run<0>(w, iters); r[1] = run<1>(w, iters); r[2] = run<2>(w, iters); r[3] = run<3>(w, iters);
    r[4] = run<4>(w, iters); r[5] = run<5>(w, iters); r[6] = run<6>(w, iters); r[7] = run<7>(w, iters);
    r[8] = run<8>(w, iters); r[9] = run<9>(w, iters); r[10] = run<10>(w, iters);
    for (int i = 0; i < 11; ++i) printf("  %-44s %8.2f   (minus baseline %7.2f)\n", names[i], r[i], r[i] - r[0]);
  }
  return 0;
}
