// umma_probe.cu -- which shared-memory descriptor conventions does tcgen05.mma kind::tf32 accept?  Built ON the GPU box:
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o /tmp/umma_probe tools/exp/umma_probe.cu && /tmp/umma_probe
// One CTA, D[128 x 32] = A[128 x 8] * B[32 x 8]^T with small integers, checked against the host for a list of candidate
// (layout, LBO, SBO) conventions for A (K-major and MN-major) and B (K-major), all SWIZZLE_NONE.
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
  return uint64_t((saddr & 0x3FFFF) >> 4) | (uint64_t(lbo >> 4) << 16) | (uint64_t(sbo >> 4) << 32) | (uint64_t(1) << 46);
}

struct Cfg {
  int a_mn;                 // 1: A is MN-major
  uint32_t a_lbo, a_sbo, b_lbo, b_sbo;
  int a_layout;             // how the host filled sA: 0 = K-major core (8 rows x 16 B), 1 = MN-major core (8 k x 16 B)
};

__global__ void probe(const float* A /*[128][8]*/, const float* B /*[32][8]*/, float* D /*[128][32]*/, Cfg c, int do_mma) {
  extern __shared__ __align__(1024) unsigned char smem[];
  unsigned char* sA = smem;          // 4096 B
  unsigned char* sB = smem + 4096;   // 1024 B
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 8192);
  uint32_t* slot = reinterpret_cast<uint32_t*>(bar + 1);
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < 128 * 8; i += 128) {
    const int m = i / 8, k = i % 8;
    int off;
    if (c.a_layout == 0)  // K-major: core = 8 rows (m) x 16 B (4 k); k halves 128 B apart; m groups of 8 256 B apart
      off = (m % 8) * 16 + (k % 4) * 4 + (k / 4) * 128 + (m / 8) * 256;
    else                  // MN-major: core = 8 k x 16 B (4 m); m groups of 4 128 B apart
      off = (m % 4) * 4 + k * 16 + (m / 4) * 128;
    *reinterpret_cast<float*>(sA + off) = A[i];
  }
  for (int i = tid; i < 32 * 8; i += 128) {
    const int n = i / 8, k = i % 8;
    const int off = (n % 8) * 16 + (k % 4) * 4 + (k / 4) * 128 + (n / 8) * 256;
    *reinterpret_cast<float*>(sB + off) = B[i];
  }
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(1) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    const uint32_t n_cols = 32;
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot)), "r"(n_cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = *slot;
  if (!do_mma) {  // TMEM store / load round trip: lane * 100 + column
    uint32_t v[32];
    for (int j = 0; j < 32; ++j) v[j] = __float_as_uint(float(tid * 100 + j));
    const uint32_t taddr = tmem + (uint32_t(warp * 32) << 16);
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, "
        "%18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};" ::"r"(taddr),
        "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]), "r"(v[10]),
        "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]), "r"(v[16]), "r"(v[17]), "r"(v[18]), "r"(v[19]), "r"(v[20]),
        "r"(v[21]), "r"(v[22]), "r"(v[23]), "r"(v[24]), "r"(v[25]), "r"(v[26]), "r"(v[27]), "r"(v[28]), "r"(v[29]), "r"(v[30]),
        "r"(v[31])
        : "memory");
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  } else if (tid == 0) {
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | (uint32_t(c.a_mn) << 15) | (4u << 17) | (8u << 24);
    const uint64_t ad = umma_desc(smem_u32(sA), c.a_lbo, c.a_sbo), bd = umma_desc(smem_u32(sB), c.b_lbo, c.b_sbo);
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem),
        "l"(ad), "l"(bd), "r"(idesc), "r"(0)
        : "memory");
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
  }
  if (do_mma) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tWAIT_%=:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@p bra DONE_%=;\n\tbra WAIT_%=;\n\t"
        "DONE_%=:\n\t}\n" ::"r"(smem_u32(bar)),
        "r"(0)
        : "memory");
  }
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  __syncthreads();
  uint32_t v[32];
  const uint32_t taddr = tmem + (uint32_t(warp * 32) << 16);
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, %21, %22, %23, "
      "%24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
        "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]),
        "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]),
        "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
  for (int j = 0; j < 32; ++j) D[tid * 32 + j] = __uint_as_float(v[j]);
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) {
    const uint32_t n_cols = 32;
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(n_cols) : "memory");
  }
}

int main() {
  float hA[128 * 8], hB[32 * 8], hD[128 * 32], want[128 * 32];
  for (int m = 0; m < 128; ++m)
    for (int k = 0; k < 8; ++k) hA[m * 8 + k] = float((m * 7 + k * 3) % 11 - 5);
  for (int n = 0; n < 32; ++n)
    for (int k = 0; k < 8; ++k) hB[n * 8 + k] = float((n * 5 + k) % 7 - 3);
  for (int m = 0; m < 128; ++m)
    for (int n = 0; n < 32; ++n) {
      float s = 0;
      for (int k = 0; k < 8; ++k) s += hA[m * 8 + k] * hB[n * 8 + k];
      want[m * 32 + n] = s;
    }
  float *dA, *dB, *dD;
  cudaMalloc(&dA, sizeof hA), cudaMalloc(&dB, sizeof hB), cudaMalloc(&dD, sizeof hD);
  cudaMemcpy(dA, hA, sizeof hA, cudaMemcpyHostToDevice), cudaMemcpy(dB, hB, sizeof hB, cudaMemcpyHostToDevice);
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 16384);
  {
    Cfg c{0, 128, 256, 128, 256, 0};
    cudaMemset(dD, 0xff, sizeof hD);
    probe<<<1, 128, 16384>>>(dA, dB, dD, c, 0);
    cudaError_t e = cudaDeviceSynchronize();
    cudaMemcpy(hD, dD, sizeof hD, cudaMemcpyDeviceToHost);
    int bad = 0;
    for (int t = 0; t < 128; ++t)
      for (int j = 0; j < 32; ++j) bad += hD[t * 32 + j] != float(t * 100 + j);
    printf("tmem st/ld round trip: %s, %d mismatches (D[5][3] = %g)\n", cudaGetErrorString(e), bad, hD[5 * 32 + 3]);
  }
  const Cfg cands[] = {
      {0, 128, 256, 128, 256, 0},  // K-major A: LBO = k-halves, SBO = 8-row groups
      {0, 256, 128, 256, 128, 0},  // swapped meaning
      {0, 128, 256, 256, 128, 0},
      {0, 256, 128, 128, 256, 0},
      {1, 4096, 128, 128, 256, 1},  // MN-major A as hoa_render_umma.cu stages it
      {1, 128, 4096, 128, 256, 1},
      {1, 128, 128, 128, 256, 1},
      {1, 4096, 128, 256, 128, 1},
      {1, 128, 4096, 256, 128, 1},
  };
  for (const Cfg& c : cands) {
    cudaMemset(dD, 0xff, sizeof hD);
    probe<<<1, 128, 16384>>>(dA, dB, dD, c, 1);
    cudaError_t e = cudaDeviceSynchronize();
    cudaMemcpy(hD, dD, sizeof hD, cudaMemcpyDeviceToHost);
    int bad = 0, zeros = 0;
    for (int i = 0; i < 128 * 32; ++i) bad += hD[i] != want[i], zeros += hD[i] == 0.0f;
    printf("a_mn=%d a(lbo %u sbo %u) b(lbo %u sbo %u): %s, %d / 4096 wrong, %d zeros; D[1][0..3] = %g %g %g %g want %g %g %g %g\n",
           c.a_mn, c.a_lbo, c.a_sbo, c.b_lbo, c.b_sbo, cudaGetErrorString(e), bad, zeros, hD[32], hD[33], hD[34], hD[35],
           want[32], want[33], want[34], want[35]);
    if (e != cudaSuccess) break;
  }
  return 0;
}
