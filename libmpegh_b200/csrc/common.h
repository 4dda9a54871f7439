// common.h -- context, error plumbing and exact-arithmetic device helpers shared by all stages.
#pragma once
#include <cuda_runtime.h>

#include <atomic>
#include <cstdint>
#include <cstdio>
#include <string>
#include <vector>

#include "../../include/mpeghb200.h"
#include "rom_tables.h"

namespace mpeghb200 {

// Device copies of the constant tables and of the tables derived from them at context creation.
struct DeviceTables {
  // FD synthesis
  const float2* tw512 = nullptr;   // [512] (cos, sin) MDCT pre/post twiddle, N = 1024
  const float2* tw64 = nullptr;    // [64]  same for N = 128
  const float* eff = nullptr;      // [256][6] effective radix-4 twiddles (c1,s1,c2,s2,c3,s3) by index j
  const float2* rad2_512 = nullptr;  // [256] effective twiddle of the final radix-2 pass of a 512-point transform
  const float* win_long[2] = {nullptr, nullptr};   // sine_1024, kbd_1024
  const float* win_short[2] = {nullptr, nullptr};  // sine_128, kbd_128
  const float* dct_cos = nullptr;  // [1024] DCT/DST-II pre-twiddle
  const float* dct_sin = nullptr;
  // lane-ordered float4 packings for the long-block fast path (two table entries per 16-byte load,
  // a warp-wide load covers 512 contiguous bytes): index [pair][lane]
  const float4* pre4 = nullptr;    // [8][32] tw512[lane + 32u], u = 2*pair, 2*pair+1
  const float4* post4 = nullptr;   // [8][32] tw512[256H + m16 + 16v], v = 2*pair, 2*pair+1 (lane = 16H + m16)
  const float4* rad2p4 = nullptr;  // [8][32] rad2_512[m16 + 16v]
  const float4* eff8 = nullptr;    // [256][2] (c1,s1,c2,s2), (c3,s3,0,0) by twiddle index j
  float pass1[3][6] = {};          // eff[64m], m = 1..3: warp-uniform, passed by value (constant bank)
};

}  // namespace mpeghb200

struct mpeghb200_ctx {
  int device = -1;
  int sm_count = 0;
  std::string last_error;
  std::atomic<uint64_t> launches{0};
  bool pack_generic = false;  // MPEGHB200_PACK_GENERIC=1: the byte-wise PCM packer for every channel count (comparison)
  int lim_variant = 7;  // MPEGHB200_LIM_VARIANT: bit 0 = 256 threads per stream (else 128), bit 1 = smoother release fast path,
                        // bit 2 = channel loads batched in the output phase for launches of at most 4 streams per SM
  int fd_variant = 7;  // launch shape, see fd_synth_launch (7 = warp-specialised persistent CTAs, 3 per SM)
  mpeghb200::DeviceTables tables;
  std::vector<void*> owned;  // device allocations released by destroy()
};

namespace mpeghb200 {

mpeghb200_status fail(mpeghb200_ctx* ctx, mpeghb200_status code, const std::string& msg);
mpeghb200_status cuda_fail(mpeghb200_ctx* ctx, cudaError_t e, const char* what);

#define MB_CUDA(ctx, expr)                                                        \
  do {                                                                            \
    cudaError_t _e = (expr);                                                      \
    if (_e != cudaSuccess) return ::mpeghb200::cuda_fail((ctx), _e, #expr);       \
  } while (0)

// ---- programmatic dependent launch ------------------------------------------------------------------
// The stages of a frame are consecutive launches on one stream.  Launched with launch_pdl, a kernel may be scheduled
// while its predecessor is still running; its CTAs call pdl_prologue() first thing: allow the next launch in turn,
// then wait until the predecessor has completed and its writes are visible.  What is saved is the launch latency
// and CTA ramp-up between the stages.  Every kernel launched this way MUST call pdl_prologue() (or wait itself)
// before it touches global memory.
extern bool g_pdl_enabled;  // ctx.cu; MPEGHB200_NO_PDL=1 launches every stage fully serialised (comparison / lanes probe)
#ifdef __CUDACC__
__device__ __forceinline__ void pdl_prologue() {
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  asm volatile("griddepcontrol.wait;" ::: "memory");
}
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream,
                              Args... args) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid, cfg.blockDim = block, cfg.dynamicSmemBytes = smem, cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr, cfg.numAttrs = g_pdl_enabled ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
}
#endif

// ---- exact binary32 arithmetic ---------------------------------------------------------------
// The reference is compiled for x86-64 without FMA contraction: every product and every sum is rounded
// separately.  These wrappers keep nvcc from contracting (the files are also built with -fmad=false).
// a - 2b and a + 2b may use a fused multiply-add because 2b is exact: fma(-2, b, a) == a - (2*b) bit for
// bit (barring overflow of 2b, i.e. |b| > 1.7e38, which audio never reaches).
#ifdef __CUDACC__
__device__ __forceinline__ float fmul(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ float fadd(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ float fsub(float a, float b) { return __fsub_rn(a, b); }
__device__ __forceinline__ float sub2(float a, float b) { return __fmaf_rn(-2.0f, b, a); }  // a - 2b
__device__ __forceinline__ float add2(float a, float b) { return __fmaf_rn(2.0f, b, a); }   // a + 2b
// a*b + c*d with three roundings, evaluated left to right like the C expression.
__device__ __forceinline__ float mul_add_mul(float a, float b, float c, float d) {
  return __fadd_rn(__fmul_rn(a, b), __fmul_rn(c, d));
}
#endif

}  // namespace mpeghb200
