// fp_probe.cu -- issue throughput of FMUL / FADD / FFMA / mixed on sm_100a (warp-instructions per cycle per SM)
#include <cstdio>
#include <cuda_runtime.h>
template <int MODE>
__global__ void k(float* out, int iters, float a, float b) {
  float x[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) x[i] = a + i + threadIdx.x;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      if (MODE == 0) x[i] = __fmaf_rn(x[i], a, b);
      if (MODE == 1) x[i] = __fmul_rn(x[i], a);
      if (MODE == 2) x[i] = __fadd_rn(x[i], b);
      if (MODE == 3) x[i] = (i & 1) ? __fmul_rn(x[i], a) : __fadd_rn(x[i], b);
      if (MODE == 4) x[i] = (i & 1) ? __fmaf_rn(x[i], a, b) : __fadd_rn(x[i], b);
      if (MODE == 5) x[i] = __fmaf_rn(x[i], a, -0.0f);
      if (MODE == 6) x[i] = __fmaf_rn(x[i], 1.0f, b);
    }
  }
  float s = 0;
#pragma unroll
  for (int i = 0; i < 16; ++i) s += x[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int MODE>
void run(const char* name, int warps_per_sm) {
  float* out;
  cudaMalloc(&out, 148 * 2048 * 4);
  const int iters = 4096;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  k<MODE><<<148, warps_per_sm * 32>>>(out, 16, 1.0001f, 0.5f);
  cudaEventRecord(e0);
  k<MODE><<<148, warps_per_sm * 32>>>(out, iters, 1.0001f, 0.5f);
  cudaEventRecord(e1);
  cudaDeviceSynchronize();
  float ms;
  cudaEventElapsedTime(&ms, e0, e1);
  double inst = double(iters) * 16 * warps_per_sm;  // warp-instructions per SM
  double cyc = ms * 1e-3 * 1.965e9;
  printf("%-22s warps/SM %2d: %.2f warp-inst/cycle/SM (assuming 1965 MHz)\n", name, warps_per_sm, inst / cyc);
  cudaFree(out);
}
int main() {
  for (int w : {4, 16, 32}) {
    run<0>("FFMA", w);
    run<1>("FMUL", w);
    run<2>("FADD", w);
    run<3>("FMUL/FADD mix", w);
    run<4>("FFMA/FADD mix", w);
    run<5>("FFMA(x,a,-0)", w);
    run<6>("FFMA(x,1,b)", w);
  }
  return 0;
}
