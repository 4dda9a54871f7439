// copy_probe.cu -- what bandwidth can a 100 MB-per-launch streaming kernel reach on this GPU?
// Same traffic as one FD step: per unit read 2 x 4 KB, write 2 x 4 KB; groups rotated so L2 cannot hold it.
#include <cstdio>
#include <cuda_runtime.h>
#include <vector>

template <int WARPS, int UPW>
__global__ void probe(const float4* __restrict__ a, float4* __restrict__ b, float4* __restrict__ c, int n_units) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int u0 = (blockIdx.x * WARPS + warp) * UPW;
  for (int k = 0; k < UPW; ++k) {
    const int u = u0 + k;
    if (u >= n_units) return;
    const float4* pa = a + size_t(u) * 256;
    float4* pb = b + size_t(u) * 256;
    float4* pc = c + size_t(u) * 256;
    float4 x[8], y[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) x[i] = __ldcs(pa + lane + 32 * i);
#pragma unroll
    for (int i = 0; i < 8; ++i) y[i] = __ldcs(pb + lane + 32 * i);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      float4 r = make_float4(x[i].x + y[i].x, x[i].y + y[i].y, x[i].z + y[i].z, x[i].w + y[i].w);
      __stcs(pc + lane + 32 * i, r);
      pb[lane + 32 * i] = x[i];
    }
  }
}

int main() {
  const int U = 6144, G = 8, steps = 200;
  std::vector<float4*> a(G), b(G), c(G);
  for (int g = 0; g < G; ++g) {
    cudaMalloc(&a[g], size_t(U) * 4096);
    cudaMalloc(&b[g], size_t(U) * 4096);
    cudaMalloc(&c[g], size_t(U) * 4096);
    cudaMemset(a[g], 0, size_t(U) * 4096);
    cudaMemset(b[g], 0, size_t(U) * 4096);
  }
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  auto run = [&](const char* name, auto launch) {
    for (int i = 0; i < 20; ++i) launch(i % G);
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    for (int i = 0; i < steps; ++i) launch(i % G);
    cudaEventRecord(e1);
    cudaDeviceSynchronize();
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    printf("%-28s %.2f us/step  %.0f GB/s  (%s)\n", name, ms * 1e3 / steps, U * 16384.0 / (ms * 1e-3 / steps) / 1e9,
           cudaGetErrorString(cudaGetLastError()));
  };
  run("probe<4,1>", [&](int g) { probe<4, 1><<<U / 4, 128>>>(a[g], b[g], c[g], U); });
  run("probe<8,1>", [&](int g) { probe<8, 1><<<U / 8, 256>>>(a[g], b[g], c[g], U); });
  run("probe<1,1>", [&](int g) { probe<1, 1><<<U, 32>>>(a[g], b[g], c[g], U); });
  run("probe<1,3>", [&](int g) { probe<1, 3><<<U / 3, 32>>>(a[g], b[g], c[g], U); });
  run("probe<2,3>", [&](int g) { probe<2, 3><<<U / 6, 64>>>(a[g], b[g], c[g], U); });
  run("probe<4,2>", [&](int g) { probe<4, 2><<<U / 8, 128>>>(a[g], b[g], c[g], U); });
  // build on the GPU box: nvcc -arch=sm_100a -cudart shared -O2 -o copy_probe copy_probe.cu (the binary is never committed)
  run("memcpyD2D 25MB", [&](int g) { cudaMemcpyAsync(c[g], a[g], size_t(U) * 4096, cudaMemcpyDeviceToDevice); });
  return 0;
}
