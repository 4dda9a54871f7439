// hoa_render_umma.cu -- the HOA render-matrix contraction on the 5th-generation tensor cores (tcgen05 + TMEM).
//
// Throughput mode 2 of mpeghb200_hoa_renderer_set_tensor_mode for the one dense contraction of the path,
//   out[och][t] = sum_mn M[och][mn] * x[mn][t]          (decoder/impeghd_hoa_renderer.c:231-253),
// as BASELINE.json's north_star asks: "tensor cores only for the batched HOA render-matrix ... contractions".
// Per 128-sample tile of a stream-frame this is D[128 x 32] = A[128 x 32] * B[32 x 32] with
//   A[m][k] = x[k][s0 + m]   (transposed on the way into shared memory: both operands K-major; an MN-major tf32
//                             operand without swizzle computes zeros on sm_100a, tools/exp/umma_probe.cu)
//   B[n][k] = M[n][k]        (K-major), n_out <= 32 and n_coef <= 32 zero-padded,
// issued by one thread as tcgen05.mma.cta_group::1.kind::tf32 (M 128, N 32, K 8), accumulators in tensor memory,
// read back with tcgen05.ld for the clip + store epilogue.
//
// Precision.  A TF32 operand keeps 11 significant bits, so every fp32 value is split EXACTLY into three TF32 terms
// (hi + mid + lo, 11 + 11 + 2 bits) on the CUDA cores, and the six products that matter (hi*hi, hi*mid, mid*hi,
// mid*mid, hi*lo, lo*hi; the others are below 2^-33 of the result) are accumulated in fp32 by the tensor core,
// smallest terms first.  What is left is the difference between the tensor core's fp32 accumulation order and the
// reference's left-to-right one: a few ulp of the sum, i.e. the results are NOT bit-identical, but within the float
// stage tolerance (2^-20 of full scale) by three orders of magnitude and within +-1 LSB of the 24-bit PCM for
// programme-level signals (tests/test_hoa_render_gpu.py::test_umma_mode_*).  Off by default.
//
// Operand staging: shared memory in the canonical no-swizzle UMMA layouts (cute/atom/mma_traits_sm100.hpp,
// make_umma_desc; conventions confirmed on the GPU by tools/exp/umma_probe.cu): core matrices of 8 rows x 16 bytes,
//   byte(row, k) = (row % 8) * 16 + (k % 4) * 4 + ((k % 8) / 4) * 128 + (row / 8) * 256 + (k / 8) * block
// with LBO = 128 (the two K halves of an instruction), SBO = 256 (groups of 8 rows); block = 4096 (A) / 1024 (B).
// The lanes of a warp are mapped (m % 8, k % 4) -> 32 distinct banks per store.  Two A stages: the CUDA cores split
// tile t + 1 while the tensor core works on tile t; 2 CTAs per SM overlap the epilogues.
#include <cstring>

#include "common.h"

namespace mpeghb200 {
namespace {

constexpr int kFrame = 1024;
constexpr int kTileM = 128;                  // samples per tile = UMMA M
constexpr int kN = 32, kK = 32;              // padded output / coefficient counts
constexpr int kKSteps = kK / 8;              // tf32: K = 8 per instruction
constexpr int kABytes = kTileM * kK * 4;     // one split term of the A tile: 16 KB
constexpr int kBBytes = kN * kK * 4;         // one split term of B: 4 KB
constexpr int kBStep = 3 * kN * 8 * 4;       // one k step of the stacked B operand [hi | mid | lo] (96 rows x 8): 3072 B
constexpr int kTmemCols = 128;               // one accumulator set of 96 columns (allocations are powers of two)
constexpr int kThreads = 256;               // 8 warps: warps w and w + 4 share the TMEM lanes 32 (w % 4) ..
constexpr int kATerms = 2;                   // A terms staged: x = hi + mid to within 2^-24 |x| (round-to-nearest split)
constexpr int kSmem = kATerms * kABytes + 3 * kBBytes + 64;  // 44 KB: 4 CTAs per SM

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

// shared-memory matrix descriptor (cute::UMMA::SmemDescriptor): start >> 4 | LBO >> 4 << 16 | SBO >> 4 << 32 |
// version 1 << 46 | layout_type SWIZZLE_NONE (0) << 61
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  return uint64_t((saddr & 0x3FFFF) >> 4) | (uint64_t(lbo_bytes >> 4) << 16) | (uint64_t(sbo_bytes >> 4) << 32) |
         (uint64_t(1) << 46);
}

// instruction descriptor (cute::UMMA::InstrDescriptor): c_format F32 (1) << 4 | a_format TF32 (2) << 7 |
// b_format TF32 (2) << 10 | a_major K (0) << 15 | b_major K (0) << 16 | N >> 3 << 17 | M >> 4 << 24
constexpr uint32_t idesc_n(int n) {
  return (1u << 4) | (2u << 7) | (2u << 10) | (0u << 15) | (0u << 16) | (uint32_t(n >> 3) << 17) | (uint32_t(kTileM >> 4) << 24);
}

__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
      "}\n" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "WAIT_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra DONE_%=;\n\t"
      "bra WAIT_%=;\n\t"
      "DONE_%=:\n\t"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}

// round to the nearest TF32 value (11 significant bits; the low 13 mantissa bits of the result are zero)
__device__ __forceinline__ float tf32_rn(float v) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(v));
  return __uint_as_float(r);
}

// B: the three exact TF32 terms of the padded render matrix in the canonical K-major layout, prepared on the host
struct UmmaB {
  float term[3][kN * kK];
};

// One 128-sample tile of x, split into its three exact TF32 terms, into stage `sA` (K-major canonical layout).
// lane -> (m % 8, k % 4): a warp's store covers 32 distinct banks, its load 4 coefficient rows x 32 contiguous bytes.
// Work item (coefficient quad kq, group of 8 samples g): warp w takes g = w + kWarps * i, so that every address is a
// per-thread base plus a compile-time constant.
constexpr int kWarps = kThreads / 32;
constexpr int kG = (kTileM / 8) / kWarps;  // sample groups per warp: 2
constexpr int kQ = kK / 4;                 // coefficient quads: 8

__device__ __forceinline__ void stage_load(float (&v)[kQ * kG], const float* __restrict__ x, int n_coef, int warp, int lane) {
  const int lm = lane >> 2, lk = lane & 3;
  const float* xt = x + size_t(lk) * kFrame + 8 * warp + lm;
#pragma unroll
  for (int kq = 0; kq < kQ; ++kq)
#pragma unroll
    for (int i = 0; i < kG; ++i)
      v[kq * kG + i] = (4 * kq + lk < n_coef) ? __ldg(xt + size_t(4 * kq) * kFrame + 8 * kWarps * i) : 0.0f;
}

__device__ __forceinline__ void stage_store(unsigned char* sA, const float (&v)[kQ * kG], int n_coef, int warp, int lane) {
  const int lm = lane >> 2, lk = lane & 3;
  unsigned char* st = sA + lm * 16 + lk * 4 + warp * 256;
#pragma unroll
  for (int kq = 0; kq < kQ; ++kq) {
    if (4 * kq < n_coef) {  // quads beyond the coefficient count stay at their initial zeros
#pragma unroll
      for (int i = 0; i < kG; ++i) {
        const float a = v[kq * kG + i];
        // round-to-nearest split: |a - hi| <= 2^-12 |a|, |a - hi - mid| <= 2^-24 |a| -- two TF32 terms carry the whole
        // fp32 value to within half an ulp (a truncating split would need three)
        const float hi = tf32_rn(a);
        const float mid = tf32_rn(__fsub_rn(a, hi));
        unsigned char* q = st + (kq & 1) * 128 + (kWarps * i) * 256 + (kq >> 1) * 4096;
        *reinterpret_cast<float*>(q) = hi;
        *reinterpret_cast<float*>(q + kABytes) = mid;
      }
    }
  }
}

// FAST: n_out == 24 and clip on (the path's own configuration: no per-output predicates, immediate store offsets)
template <bool FAST>
__global__ void __launch_bounds__(kThreads, 4)
    hoa_render_umma_kernel(const float* __restrict__ in, float* __restrict__ out, const UmmaB* __restrict__ Bg, int n_coef,
                           int n_out, int clip, int n_tiles) {
  pdl_prologue();
  extern __shared__ __align__(1024) unsigned char smem[];
  unsigned char* sA = smem;                          // [2 terms][kKSteps][4096 B]
  unsigned char* sB = smem + kATerms * kABytes;      // [kKSteps][96 rows x 8: 3072 B], terms stacked along N
  uint64_t* bar = reinterpret_cast<uint64_t*>(sB + 3 * kBBytes);  // the MMAs of the staged tile completed
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar + 1);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  // ---- one-time set-up: B terms, zero padding of A (coefficients >= n_coef are never written), barrier, TMEM ----
  for (int i = tid; i < 3 * kBBytes / 16; i += kThreads)
    reinterpret_cast<uint4*>(sB)[i] = __ldg(reinterpret_cast<const uint4*>(Bg) + i);
  for (int i = tid; i < kATerms * kABytes / 16; i += kThreads) reinterpret_cast<uint4*>(sA)[i] = make_uint4(0, 0, 0, 0);
  if (tid == 0) {
    mbar_init(bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    const uint32_t n_cols = kTmemCols;
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(n_cols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_d = *tmem_slot;
  const uint32_t a_base = smem_u32(sA), b_base = smem_u32(sB);
  auto tile_x = [&](int tile) { return in + size_t(tile >> 3) * n_coef * kFrame + (tile & 7) * kTileM; };

  // One CTA keeps one tile in flight (stage -> MMA -> epilogue); four CTAs per SM overlap each other's phases, and the
  // global loads of a CTA's next tile are issued before it waits for its MMAs.
  float v[kQ * kG];
  uint32_t phase = 0;
  int tile = blockIdx.x;
  if (tile < n_tiles) stage_load(v, tile_x(tile), n_coef, warp, lane);
  for (; tile < n_tiles; tile += gridDim.x) {
    stage_store(sA, v, n_coef, warp, lane);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // staged tile -> visible to the async proxy
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (tid == 0) {
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      // x = hi + mid to half an ulp (round-to-nearest split), M = hi + mid + lo exactly.  Two MMAs per k step,
      // the B terms stacked along N; the second MMA's columns overlap the first's so that products of equal magnitude
      // share an accumulator:
      //   D[:,  0: 96] (+)= A_hi  * [B_hi | B_mid | B_lo]      -> hi*hi | hi*mid | hi*lo
      //   D[:, 32: 96]  += A_mid * [B_hi | B_mid]                        + mid*hi | + mid*mid
#pragma unroll
      for (int ks = 0; ks < kKSteps; ++ks) {
        // K-major, no swizzle: LBO = the two 16-byte K halves of a core-matrix pair (128 B), SBO = 8-row groups (256 B)
        const uint64_t bd = umma_desc(b_base + ks * kBStep, 128, 256);
        umma_tf32(tmem_d, umma_desc(a_base + ks * 4096, 128, 256), bd, idesc_n(96), ks > 0);
        umma_tf32(tmem_d + 32, umma_desc(a_base + kABytes + ks * 4096, 128, 256), bd, idesc_n(64), 1);
      }
      // arrives on the barrier when every MMA above has completed (implies tcgen05.fence::before_thread_sync)
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
                   : "memory");
    }
    const int next = tile + gridDim.x;
    if (next < n_tiles) stage_load(v, tile_x(next), n_coef, warp, lane);
    mbar_wait(bar, phase);
    phase ^= 1;
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    // ---- epilogue: D row m = TMEM lane m; warps w and w + 4 own lanes 32 (w % 4) .. + 31 and take alternate groups of
    //      four outputs.  Columns [0, 32) hold hi*hi, [32, 64) hi*mid + mid*hi, [64, 96) hi*lo + mid*mid: summed in
    //      fp32, smallest first, clipped, stored ----
    {
      const int m = (warp & 3) * 32 + lane;
      const int c_first = 4 * (warp >> 2);
      const uint32_t taddr = tmem_d + (uint32_t((warp & 3) * 32) << 16) + uint32_t(c_first);
      float* y = out + size_t(tile >> 3) * n_out * kFrame + (tile & 7) * kTileM + m + size_t(c_first) * kFrame;
      uint32_t q[3][3][4];
#pragma unroll
      for (int c = 0; c < 3; ++c)
#pragma unroll
        for (int p = 0; p < 3; ++p)
          asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];"
                       : "=r"(q[c][p][0]), "=r"(q[c][p][1]), "=r"(q[c][p][2]), "=r"(q[c][p][3])
                       : "r"(taddr + uint32_t(64 - 32 * p + 8 * c))
                       : "memory");
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      const int n_left = n_out - c_first;
#pragma unroll
      for (int c = 0; c < 3; ++c)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          if (FAST || 8 * c + j < n_left) {
            float val = __fadd_rn(__fadd_rn(__uint_as_float(q[c][0][j]), __uint_as_float(q[c][1][j])),
                                  __uint_as_float(q[c][2][j]));
            if (FAST || clip) {
              val = (val < 32767.0f) ? val : 32767.0f;
              val = (val > -32768.0f) ? val : -32768.0f;
            }
            __stcs(y + (8 * c + j) * kFrame, val);
          }
        }
    }
    // the accumulator and the A stage are free again once every warp has its results in registers
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
  }
  if (warp == 0) {
    const uint32_t n_cols = kTmemCols;
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_d), "r"(n_cols) : "memory");
  }
}

}  // namespace

// host: M [n_out][n_coef] -> the three TF32 terms in the canonical K-major layout
void umma_prepare_b(const float* M, int n_out, int n_coef, float* terms /* [3][kN * kK] */) {
  std::memset(terms, 0, sizeof(float) * 3 * kN * kK);
  auto trunc = [](float v) {
    uint32_t u;
    std::memcpy(&u, &v, 4);
    u &= 0xFFFFE000u;
    std::memcpy(&v, &u, 4);
    return v;
  };
  for (int n = 0; n < n_out; ++n)
    for (int k = 0; k < n_coef; ++k) {
      const float m = M[n * n_coef + k];
      const float hi = trunc(m), r = m - hi, mid = trunc(r), lo = r - mid;
      const float t3[3] = {hi, mid, lo};
      for (int t = 0; t < 3; ++t) {  // the terms stacked along N: row 32 t + n of a 96-row K-major operand
        const int row = 32 * t + n;
        const int byte = (row % 8) * 16 + (k % 4) * 4 + ((k % 8) / 4) * 128 + (row / 8) * 256 + (k / 8) * kBStep;
        terms[byte / 4] = t3[t];
      }
    }
}

size_t umma_b_bytes() { return sizeof(UmmaB); }

cudaError_t hoa_render_umma_launch(const float* d_in, float* d_out, const void* d_b, int n_coef, int n_out, int clip,
                                   int n_streams, int sm_count, cudaStream_t st) {
  if (n_coef > kK || n_out > kN) return cudaErrorInvalidValue;
  const bool fast = n_out == 24 && clip;
  auto kern = fast ? hoa_render_umma_kernel<true> : hoa_render_umma_kernel<false>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmem);
  if (e != cudaSuccess) return e;
  const int n_tiles = n_streams * (kFrame / kTileM);
  const int grid = n_tiles < 4 * sm_count ? n_tiles : 4 * sm_count;  // 4 CTAs per SM: 4 x 44 KB of shared memory, 4 x 128 = all 512 TMEM columns
  return launch_pdl(kern, dim3(grid), dim3(kThreads), size_t(kSmem), st, d_in, d_out,
                    static_cast<const UmmaB*>(d_b), n_coef, n_out, clip, n_tiles);
}

}  // namespace mpeghb200
