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

__global__ void __launch_bounds__(256) resample_kernel(const __grid_constant__ RsParams P) {
  __shared__ float xin[32 + kIn];
  __shared__ float poly[kIn][3];
  __shared__ float yh[59 + 3 * kIn];
  const int tid = threadIdx.x, up = P.up;
  const int hist = 59 / up, taps = 60 / up;
  const size_t u = blockIdx.x;
  const float* x = P.in + u * kIn;
  float* st = P.state + u * kStateFloats;
  const int n_out = kIn * up / P.down;
  float* y = P.out + u * n_out;

  for (int i = tid; i < hist; i += 256) xin[i] = st[i];
  for (int i = tid; i < kIn; i += 256) xin[hist + i] = __ldcs(x + i);
  if (P.down == 2)
    for (int i = tid; i < 59; i += 256) yh[i] = st[32 + i];
  __syncthreads();
  for (int n = tid; n < kIn; n += 256) {
    float p0 = 0.0f, p1 = 0.0f, p2 = 0.0f;
    for (int k = 0; k < taps; ++k) {
      const float v = xin[n + k];
      p0 = fadd(p0, fmul(v, P.f_up[k * up]));
      p1 = fadd(p1, fmul(v, P.f_up[k * up + 1]));
      if (up == 3) p2 = fadd(p2, fmul(v, P.f_up[k * up + 2]));
    }
    poly[n][0] = p0, poly[n][1] = p1, poly[n][2] = p2;
  }
  __syncthreads();
  float* o = (P.down == 2) ? yh + 59 : y;
  for (int n = tid; n < kIn; n += 256) {
    float mem[2];
    mem[0] = n ? poly[n - 1][1] : st[92];
    mem[1] = n ? poly[n - 1][2] : st[93];
    for (int p = 0; p < up; ++p) {
      float v = 0.0f;
      for (int q = p; q < up - 1; ++q) v = fadd(v, mem[q]);
      for (int q = 0; q <= p; ++q) v = fadd(v, poly[n][q]);
      o[n * up + p] = v;
    }
  }
  __syncthreads();
  if (P.down == 2) {
    for (int n = tid; n < n_out; n += 256) {
      float mac = 0.0f;
#pragma unroll 4
      for (int k = 0; k < 60; ++k) mac = fadd(mac, fmul(yh[2 * n + k], P.f_down[k]));
      __stcs(y + n, mac);
    }
    for (int i = tid; i < 59; i += 256) st[32 + i] = yh[kIn * up + i];
  }
  for (int i = tid; i < hist; i += 256) st[i] = xin[kIn + i];
  if (tid == 0) st[92] = poly[kIn - 1][1], st[93] = poly[kIn - 1][2];
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
  resample_kernel<<<n_units, 256, 0, static_cast<cudaStream_t>(cuda_stream)>>>(P);
  ctx->launches.fetch_add(1);
  MB_CUDA(ctx, cudaGetLastError());
  return MPEGHB200_OK;
}

}  // extern "C"
