// resampler.cu -- output resampler: 60-tap polyphase x2 / x3 interpolation, optional 60-tap decimation by 2.
//
// Stands in for impeghd_resample (decoder/impeghd_resampler.c:93), used for 14.7 .. 32 kHz streams
// (impeghd_resampler_get_sampling_ratio, decoder/impeghd_api.c:132).  One CTA per channel-frame: input + history
// are staged in shared memory, every thread evaluates the polyphase branch sums of its input samples in the
// reference's tap order, the branch sums are recombined exactly like the reference's poly_filt_mem shuffle, and
// the decimator (when present) runs from shared memory.  Filter tables are the reference's ROM
// (impeghd_resampler_filter, impeghd_resampler_filter_3), handed over at creation.
#include <cstring>
#include <new>

#include "common.h"

namespace mpeghb200 {
namespace {

constexpr int kIn = 1024;
constexpr int kStateFloats = 96;  // in_hist [32], up_hist [60], poly_mem [4]

struct RsParams {
  const float* in;
  float* out;
  float* state;
  int up, down;
  float f_up[60], f_down[60];
};

// UP / DOWN at compile time: the tap loops unroll completely, the filter taps become constant-bank operands of the
// multiplies, and a thread takes four consecutive input samples (two consecutive decimator outputs) so that the
// overlapping windows are loaded once, as 16-byte shared-memory loads, instead of once per tap.
template <int UP, int DOWN>
__global__ void __launch_bounds__(256) resample_kernel(const __grid_constant__ RsParams P) {
  constexpr int hist = 59 / UP, taps = 60 / UP;
  constexpr int n_out = kIn * UP / DOWN;
  __shared__ __align__(16) float xin[32 + kIn + 4];
  __shared__ __align__(16) float poly[UP][kIn];  // branch-major: a thread stores its four samples of a branch as one float4
  __shared__ __align__(16) float yh[DOWN == 2 ? 59 + UP * kIn + 5 : 4];
  const int tid = threadIdx.x;
  const size_t u = blockIdx.x;
  const float* x = P.in + u * kIn;
  float* st = P.state + u * kStateFloats;
  float* y = P.out + u * n_out;

  for (int i = tid; i < hist; i += 256) xin[i] = st[i];
  for (int i = tid; i < kIn; i += 256) xin[hist + i] = __ldcs(x + i);
  if (DOWN == 2)
    for (int i = tid; i < 59; i += 256) yh[i] = st[32 + i];
  __syncthreads();
  {
    constexpr int kW = (taps + 3 + 3) / 4;  // float4 loads covering samples n0 .. n0 + taps + 2
    float w[4 * kW];
    const int n0 = 4 * tid;
#pragma unroll
    for (int q = 0; q < kW; ++q) {
      const float4 v = *reinterpret_cast<const float4*>(xin + n0 + 4 * q);
      w[4 * q] = v.x, w[4 * q + 1] = v.y, w[4 * q + 2] = v.z, w[4 * q + 3] = v.w;
    }
    float pb[UP][4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
#pragma unroll
      for (int b = 0; b < UP; ++b) pb[b][j] = 0.0f;
#pragma unroll
      for (int k = 0; k < taps; ++k)
#pragma unroll
        for (int b = 0; b < UP; ++b) pb[b][j] = fadd(pb[b][j], fmul(w[j + k], P.f_up[k * UP + b]));
    }
#pragma unroll
    for (int b = 0; b < UP; ++b)
      *reinterpret_cast<float4*>(&poly[b][n0]) = make_float4(pb[b][0], pb[b][1], pb[b][2], pb[b][3]);
  }
  __syncthreads();
  float* o = (DOWN == 2) ? yh + 59 : y;
  for (int n = tid; n < kIn; n += 256) {
    float mem[2];
    mem[0] = n ? poly[1][n - 1] : st[92];
    mem[1] = (UP == 3) ? (n ? poly[UP - 1][n - 1] : st[93]) : 0.0f;
#pragma unroll
    for (int p = 0; p < UP; ++p) {
      float v = 0.0f;
#pragma unroll
      for (int q = p; q < UP - 1; ++q) v = fadd(v, mem[q]);
#pragma unroll
      for (int q = 0; q <= p; ++q) v = fadd(v, poly[q][n]);
      o[n * UP + p] = v;
    }
  }
  __syncthreads();
  if (DOWN == 2) {
    for (int m = tid; m < n_out / 2; m += 256) {  // outputs 2m and 2m + 1 read yh[4m .. 4m + 61]
      float w[64];
#pragma unroll
      for (int q = 0; q < 16; ++q) {
        const float4 v = *reinterpret_cast<const float4*>(yh + 4 * m + 4 * q);
        w[4 * q] = v.x, w[4 * q + 1] = v.y, w[4 * q + 2] = v.z, w[4 * q + 3] = v.w;
      }
      float mac0 = 0.0f, mac1 = 0.0f;
#pragma unroll
      for (int k = 0; k < 60; ++k) {
        mac0 = fadd(mac0, fmul(w[k], P.f_down[k]));
        mac1 = fadd(mac1, fmul(w[2 + k], P.f_down[k]));
      }
      __stcs(reinterpret_cast<float2*>(y) + m, make_float2(mac0, mac1));
    }
    for (int i = tid; i < 59; i += 256) st[32 + i] = yh[kIn * UP + i];
  }
  for (int i = tid; i < hist; i += 256) st[i] = xin[kIn + i];
  if (tid == 0) {
    st[92] = poly[1][kIn - 1];
    st[93] = (UP == 3) ? poly[UP - 1][kIn - 1] : 0.0f;
  }
}

}  // namespace
}  // namespace mpeghb200

using namespace mpeghb200;

struct mpeghb200_resampler {
  mpeghb200_ctx* ctx = nullptr;
  RsParams P{};
};

extern "C" {

mpeghb200_status mpeghb200_resampler_create(mpeghb200_ctx* ctx, int fac_up, int fac_down, const float* filter_up,
                                            const float* filter_down, mpeghb200_resampler** out) {
  if (!ctx || !out || !filter_up) return MPEGHB200_ERR_BAD_ARG;
  *out = nullptr;
  if (!((fac_up == 2 && fac_down == 1) || (fac_up == 3 && (fac_down == 1 || fac_down == 2))) ||
      (fac_down == 2 && !filter_down))
    return fail(ctx, MPEGHB200_ERR_BAD_ARG, "resampler_create: ratios 2/1, 3/1, 3/2 (decoder/impeghd_api.c:132)");
  auto* r = new (std::nothrow) mpeghb200_resampler();
  if (!r) return MPEGHB200_ERR_NO_MEM;
  r->ctx = ctx;
  r->P.up = fac_up, r->P.down = fac_down;
  std::memcpy(r->P.f_up, filter_up, sizeof(float) * 60);
  if (filter_down) std::memcpy(r->P.f_down, filter_down, sizeof(float) * 60);
  *out = r;
  return MPEGHB200_OK;
}

void mpeghb200_resampler_destroy(mpeghb200_resampler* r) { delete r; }
size_t mpeghb200_resampler_state_floats(const mpeghb200_resampler* r) { return r ? kStateFloats : 0; }
int mpeghb200_resampler_out_len(const mpeghb200_resampler* r) { return r ? kIn * r->P.up / r->P.down : 0; }

mpeghb200_status mpeghb200_resample_dev(mpeghb200_ctx* ctx, const mpeghb200_resampler* r, const float* d_in,
                                        float* d_out, float* d_state, int n_units, void* cuda_stream) {
  if (!ctx || !r) return MPEGHB200_ERR_BAD_ARG;
  if (n_units < 0 || (n_units > 0 && (!d_in || !d_out || !d_state)))
    return fail(ctx, MPEGHB200_ERR_BAD_ARG, "resample_dev: null buffer or negative unit count");
  if (n_units == 0) return MPEGHB200_OK;
  MB_CUDA(ctx, cudaSetDevice(ctx->device));
  RsParams P = r->P;
  P.in = d_in, P.out = d_out, P.state = d_state;
  const cudaStream_t stream = static_cast<cudaStream_t>(cuda_stream);
  if (P.up == 2)
    resample_kernel<2, 1><<<n_units, 256, 0, stream>>>(P);
  else if (P.down == 1)
    resample_kernel<3, 1><<<n_units, 256, 0, stream>>>(P);
  else
    resample_kernel<3, 2><<<n_units, 256, 0, stream>>>(P);
  ctx->launches.fetch_add(1);
  MB_CUDA(ctx, cudaGetLastError());
  return MPEGHB200_OK;
}

}  // extern "C"
