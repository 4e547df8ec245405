// Random row-gather ceiling: how fast can a B200 read random ROW_BYTES-sized rows of a table much
// larger than L2?  The traversal kernels' HBM traffic is exactly this access pattern, so this is
// the practical roofline next to the streaming-copy peak in MEASURED_PEAKS.json.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o gather_probe gather_probe.cu
//   ./gather_probe [rows_per_warp_in_flight]
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

template <int CPL, int U>   // row = CPL*128 bytes; 8 lanes per row, U x 4 rows in flight per warp
__global__ void __launch_bounds__(128) gather(const float4 *__restrict__ vec, const uint32_t *__restrict__ idx,
                                              size_t n_idx, float *out) {
  const int lane = threadIdx.x & 31, team = lane >> 3, t = lane & 7;
  const size_t warp = (size_t)blockIdx.x * 4 + (threadIdx.x >> 5), nwarps = (size_t)gridDim.x * 4;
  float acc = 0.f;
  for (size_t base = warp * 4 * U; base + 4 * U <= n_idx; base += nwarps * 4 * U) {
    float4 x[U][CPL];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const uint32_t id = __ldg(idx + base + u * 4 + team);
      const float4 *row = vec + (size_t)id * (8 * CPL) + t;
#pragma unroll
      for (int j = 0; j < CPL; ++j) x[u][j] = __ldg(row + 8 * j);
    }
#pragma unroll
    for (int u = 0; u < U; ++u)
#pragma unroll
      for (int j = 0; j < CPL; ++j) acc += x[u][j].x + x[u][j].y + x[u][j].z + x[u][j].w;
  }
  if (acc == 123.456f) out[0] = acc;
}

template <int CPL, int U>
void run(const float4 *vec, const uint32_t *idx, size_t n_idx, float *out, int ctas_per_sm) {
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  const int grid = 148 * ctas_per_sm;
  for (int i = 0; i < 2; ++i) gather<CPL, U><<<grid, 128>>>(vec, idx, n_idx, out);
  cudaEventRecord(e0);
  for (int i = 0; i < 5; ++i) gather<CPL, U><<<grid, 128>>>(vec, idx, n_idx, out);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms;
  cudaEventElapsedTime(&ms, e0, e1);
  ms /= 5;
  printf("row %4d B  rows in flight/warp %2d  warps/SM %2d : %7.3f ms  %7.1f GB/s\n", CPL * 128, 4 * U,
         ctas_per_sm * 4, ms, (double)n_idx * CPL * 128 / ms / 1e6);
}

int main() {
  const size_t n = 1000000, n_idx = 16u << 20;
  const int max_cpl = 30;
  float4 *vec;
  uint32_t *idx;
  float *out;
  cudaMalloc(&vec, n * max_cpl * 128);
  cudaMemset(vec, 0, n * max_cpl * 128);
  cudaMalloc(&idx, n_idx * 4);
  cudaMalloc(&out, 4);
  std::vector<uint32_t> h(n_idx);
  uint64_t s = 88172645463325252ull;
  for (auto &v : h) {
    s ^= s << 13; s ^= s >> 7; s ^= s << 17;
    v = (uint32_t)(s % n);
  }
  cudaMemcpy(idx, h.data(), n_idx * 4, cudaMemcpyHostToDevice);
  for (int c : {4, 6, 8, 12, 16}) {
    run<4, 1>(vec, idx, n_idx, out, c);
    run<4, 2>(vec, idx, n_idx, out, c);
    run<4, 4>(vec, idx, n_idx, out, c);
  }
  for (int c : {4, 6, 8}) {
    run<3, 2>(vec, idx, n_idx / 2, out, c);
    run<30, 1>(vec, idx, n_idx / 8, out, c);
  }
  return 0;
}
