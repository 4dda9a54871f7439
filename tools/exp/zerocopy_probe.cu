// zerocopy_probe.cu -- can a kernel gather 4 KB rows straight out of registered host memory (and scatter back) at a
// rate that beats "CPU memcpy into pinned staging + cudaMemcpy"?  Built ON the GPU box:
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o /tmp/zc tools/exp/zerocopy_probe.cu && /tmp/zc
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cuda_runtime.h>
#include <vector>

__global__ void gather_rows(const float4* const* __restrict__ src, float4* __restrict__ dst, int row_f4) {
  const float4* s = src[blockIdx.x];
  float4* d = dst + size_t(blockIdx.x) * row_f4;
  for (int i = threadIdx.x; i < row_f4; i += blockDim.x) d[i] = s[i];
}
__global__ void scatter_rows(float4* const* __restrict__ dstp, const float4* __restrict__ src, int row_f4) {
  float4* d = dstp[blockIdx.x];
  const float4* s = src + size_t(blockIdx.x) * row_f4;
  for (int i = threadIdx.x; i < row_f4; i += blockDim.x) d[i] = s[i];
}

int main() {
  const int n_handles = 256, rows_per = 12, row_bytes = 4096, n_rows = n_handles * rows_per;
  const size_t blk = 8u << 20;  // 8 MB of "handle memory" each
  std::vector<char*> mem(n_handles);
  for (int h = 0; h < n_handles; ++h) {
    mem[h] = (char*)malloc(blk);
    memset(mem[h], h, blk);
  }
  auto t0 = std::chrono::steady_clock::now();
  for (int h = 0; h < n_handles; ++h)
    if (cudaHostRegister(mem[h], blk, cudaHostRegisterDefault) != cudaSuccess) { printf("register failed\n"); return 1; }
  auto t1 = std::chrono::steady_clock::now();
  printf("cudaHostRegister: %.2f ms per 8 MB block\n", std::chrono::duration<double, std::milli>(t1 - t0).count() / n_handles);
  std::vector<const float4*> hp(n_rows);
  for (int r = 0; r < n_rows; ++r) hp[r] = (const float4*)(mem[r / rows_per] + 1000 * 4096 + (r % rows_per) * 5 * 4096);
  const float4** dp; float4* dst; float4* pinned;
  cudaMalloc(&dp, n_rows * sizeof(void*)); cudaMalloc(&dst, size_t(n_rows) * row_bytes);
  cudaMallocHost(&pinned, size_t(n_rows) * row_bytes);
  cudaMemcpy(dp, hp.data(), n_rows * sizeof(void*), cudaMemcpyHostToDevice);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int threads : {64, 128, 256}) {
    for (int rep = 0; rep < 3; ++rep) {
      cudaEventRecord(e0);
      gather_rows<<<n_rows, threads>>>(dp, dst, row_bytes / 16);
      cudaEventRecord(e1); cudaEventSynchronize(e1);
      float ms; cudaEventElapsedTime(&ms, e0, e1);
      if (rep == 2) printf("gather  %d rows x 4 KB, %3d threads: %.3f ms = %.1f GB/s (%s)\n", n_rows, threads, ms, n_rows * 4096.0 / ms / 1e6, cudaGetErrorString(cudaGetLastError()));
    }
    for (int rep = 0; rep < 3; ++rep) {
      cudaEventRecord(e0);
      scatter_rows<<<n_rows, threads>>>((float4* const*)dp, dst, row_bytes / 16);
      cudaEventRecord(e1); cudaEventSynchronize(e1);
      float ms; cudaEventElapsedTime(&ms, e0, e1);
      if (rep == 2) printf("scatter %d rows x 4 KB, %3d threads: %.3f ms = %.1f GB/s\n", n_rows, threads, ms, n_rows * 4096.0 / ms / 1e6);
    }
  }
  // the present way: one CPU thread's memcpy into pinned staging + one cudaMemcpy
  auto c0 = std::chrono::steady_clock::now();
  for (int rep = 0; rep < 5; ++rep) {
    for (int r = 0; r < n_rows; ++r) memcpy((char*)pinned + size_t(r) * row_bytes, hp[r], row_bytes);
    cudaMemcpy(dst, pinned, size_t(n_rows) * row_bytes, cudaMemcpyHostToDevice);
  }
  auto c1 = std::chrono::steady_clock::now();
  printf("memcpy (1 thread) + cudaMemcpy: %.3f ms\n", std::chrono::duration<double, std::milli>(c1 - c0).count() / 5);
  return 0;
}
