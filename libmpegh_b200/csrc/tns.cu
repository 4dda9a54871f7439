// tns.cu -- temporal noise shaping: all-pole filtering of spectral ranges, in place.
//
// Stands in for the filtering half of ia_core_coder_tns_apply (decoder/ia_core_coder_tns.c:204) /
// impeghd_tns_decode_coeff_ar_filter (:83).  The host keeps what depends on libm and on bitstream tables:
// coefficient decoding (double-precision sin, mpeghb200_tns_lpc_from_coef below does it the reference's way)
// and the band -> bin resolution with its clamps (:262-284).  What is left is a strictly serial recurrence per
// filter (y[n] = x[n] - sum lpc[j] y[n-1-j], every product and difference rounded separately), up to 3
// filters per window applied in order.  One WARP runs one chain (= the filters of one window of one channel): the
// warp moves the range through shared memory, one lane does the recurrence (see tns_kernel).  The recurrence
// latency (`order` dependent subtractions per bin), not HBM, bounds this stage; it touches only the filtered bins.
// Worst case (3 filters of order 15 over all 1024 bins) = 3072 x 15 dependent subtractions = 94 us at 4 cycles
// each: the stress stage bench sits at that floor; typical side info (order <= 8, upper bands) takes 20-40 us.
#include <cmath>
#include <cstring>

#include "common.h"

namespace mpeghb200 {
namespace {

// One warp per chain.  The warp stages the filter's bins in shared memory in filtering order (coalesced in either
// direction), lane 0 runs the recurrence there with the last 16 outputs in a register ring, the warp writes the
// range back.  A chain-per-thread layout leaves the GPU with a few dozen warps whose every load is a scattered,
// latency-exposed access; this way there is one warp per chain (thousands), the memory side is coalesced, and what
// remains per bin is the dependent chain of `order` subtractions.
constexpr int kTnsWarps = 4;

__global__ void __launch_bounds__(32 * kTnsWarps)
    tns_kernel(float* __restrict__ coef, const mpeghb200_tns_filter* __restrict__ filters,
               const int32_t* __restrict__ chain_first, int n_chains) {
  pdl_prologue();
  __shared__ __align__(16) float stage[kTnsWarps][1024 + 16];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const int c = blockIdx.x * kTnsWarps + wib;
  if (c >= n_chains) return;  // whole warp
  float* buf = stage[wib];
  for (int f = chain_first[c]; f < chain_first[c + 1]; ++f) {
    const mpeghb200_tns_filter& F = filters[f];
    const int order = F.order, size = F.size, inc = F.inc;
    if (order <= 0 || size <= 0 || size > 1024) continue;
    float* p = coef + size_t(F.unit) * 1024 + F.start + (inc < 0 ? size - 1 : 0);
    for (int i = lane; i < size; i += 32) buf[i] = p[i * inc];
    for (int i = size + lane; i < ((size + 15) & ~15); i += 32) buf[i] = 0.0f;
    __syncwarp();
    if (lane == 0) {
      // the last 16 outputs live in a register ring addressed with compile-time indices (the bin loop is
      // unrolled by 16), so a bin costs its `order` product/difference pairs and nothing else
      float lpc[15], ring[16];
#pragma unroll
      for (int j = 0; j < 15; ++j) lpc[j] = F.lpc[j];
#pragma unroll
      for (int j = 0; j < 16; ++j) ring[j] = 0.0f;
      for (int i0 = 0; i0 < size; i0 += 16) {
        float xb[16];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const float4 v = *reinterpret_cast<const float4*>(buf + i0 + 4 * q);
          xb[4 * q] = v.x, xb[4 * q + 1] = v.y, xb[4 * q + 2] = v.z, xb[4 * q + 3] = v.w;
        }
#pragma unroll
        for (int e = 0; e < 16; ++e) {
          float y = xb[e];
#pragma unroll
          for (int j = 0; j < 15; ++j)
            if (j < order) y = fsub(y, fmul(lpc[j], ring[(e - 1 - j) & 15]));  // y[n-1-j]
          ring[e] = y;
          xb[e] = y;
        }
#pragma unroll
        for (int q = 0; q < 4; ++q)
          *reinterpret_cast<float4*>(buf + i0 + 4 * q) = make_float4(xb[4 * q], xb[4 * q + 1], xb[4 * q + 2], xb[4 * q + 3]);
      }
    }
    __syncwarp();
    for (int i = lane; i < size; i += 32) p[i * inc] = buf[i];
    __syncwarp();
  }
}

}  // namespace
}  // namespace mpeghb200

using namespace mpeghb200;

extern "C" {

mpeghb200_status mpeghb200_tns_lpc_from_coef(int order, int coef_res, const int16_t* coef, float* lpc) {
  if (order < 0 || order > 15 || (coef_res != 3 && coef_res != 4) || !coef || !lpc) return MPEGHB200_ERR_BAD_ARG;
  // decoder/ia_core_coder_tns.c:93-117, same expression types (float / double mix) as the reference
  float tmp[16], b[16], a[16];
  const float pi2 = (float)3.14159265358979323846264338327950288 / 2;
  const float iqfac = (float)(((float)(1 << (coef_res - 1)) - 0.5) / (pi2));
  const float iqfac_m = (float)(((float)(1 << (coef_res - 1)) + 0.5) / (pi2));
  for (int i = 0; i < order; i++) tmp[i + 1] = (float)std::sin(coef[i] / ((coef[i] >= 0) ? iqfac : iqfac_m));
  a[0] = 1;
  for (int m = 1; m <= order; m++) {
    b[0] = a[0];
    for (int i = 1; i < m; i++) {
      const float prod = tmp[m] * a[m - i];
      b[i] = a[i] + prod;
    }
    b[m] = tmp[m];
    for (int i = 0; i <= m; i++) a[i] = b[i];
  }
  for (int i = 0; i < 16; i++) lpc[i] = (i < order) ? a[i + 1] : 0.0f;
  return MPEGHB200_OK;
}

// Host half of ia_core_coder_tns_apply (decoder/ia_core_coder_tns.c:204-300): band clamps, band -> bin lookup,
// coefficient decoding; one chain per window that has at least one effective filter.
mpeghb200_status mpeghb200_tns_resolve(const mpeghb200_tns_channel* ch, const int16_t* sfb_top, int sfb_per_sbk,
                                       int tns_max_band, mpeghb200_tns_filter* filters_out, int filters_cap,
                                       int32_t* chain_len_out, int chains_cap, int* n_filters_out,
                                       int* n_chains_out) {
  if (!ch || !sfb_top || !filters_out || !chain_len_out || !n_filters_out || !n_chains_out || sfb_per_sbk <= 0)
    return MPEGHB200_ERR_BAD_ARG;
  *n_filters_out = 0, *n_chains_out = 0;
  const int n_win = ch->is_long ? 1 : 8, bins = ch->is_long ? 1024 : 128;
  int nbands = ch->max_sfb;
  if (ch->igf_active) {  // ia_core_coder_igf_nbands (:147)
    if (ch->igf_after_tns_synth) {
      if (ch->igf_sfb_start < nbands) nbands = ch->igf_sfb_start;
    } else {
      if (ch->igf_sfb_stop > nbands) nbands = ch->igf_sfb_stop;
    }
  }
  int nf = 0, nc = 0;
  for (int w = 0; w < n_win; ++w) {
    int in_chain = 0;
    const int n_filt = ch->n_filt[w] < 0 ? 0 : (ch->n_filt[w] > 3 ? 3 : ch->n_filt[w]);
    for (int f = 0; f < n_filt; ++f) {
      const auto& t = ch->filt[w][f];
      if (!t.order) continue;
      int start = t.start_band < tns_max_band ? t.start_band : tns_max_band;
      start = start < nbands ? start : nbands;
      if (start > sfb_per_sbk) return MPEGHB200_ERR_BAD_ARG;  // IA_MPEGH_DEC_EXE_FATAL_INVALID_TNS_SFB (:262)
      int stop = t.stop_band;
      if (!ch->igf_active) stop = stop < tns_max_band ? stop : tns_max_band;
      stop = stop < nbands ? stop : nbands;
      if (stop > sfb_per_sbk) return MPEGHB200_ERR_BAD_ARG;
      const int a = start > 0 ? sfb_top[start - 1] : 0, b = stop > 0 ? sfb_top[stop - 1] : 0;
      if (b - a <= 0) continue;
      if (t.order < 0 || t.order > 15 || b > bins) return MPEGHB200_ERR_BAD_ARG;
      if (nf >= filters_cap) return MPEGHB200_ERR_NO_MEM;
      mpeghb200_tns_filter& r = filters_out[nf];
      std::memset(&r, 0, sizeof(r));
      r.unit = ch->unit, r.start = w * bins + a, r.size = b - a, r.inc = t.direction ? -1 : 1, r.order = t.order;
      const mpeghb200_status st = mpeghb200_tns_lpc_from_coef(t.order, ch->coef_res[w], t.coef, r.lpc);
      if (st != MPEGHB200_OK) return st;
      ++nf, ++in_chain;
    }
    if (in_chain) {
      if (nc >= chains_cap) return MPEGHB200_ERR_NO_MEM;
      chain_len_out[nc++] = in_chain;
    }
  }
  *n_filters_out = nf, *n_chains_out = nc;
  return MPEGHB200_OK;
}

mpeghb200_status mpeghb200_tns_dev(mpeghb200_ctx* ctx, float* d_coef, const mpeghb200_tns_filter* d_filters,
                                   const int32_t* d_chain_first, int n_chains, void* cuda_stream) {
  if (!ctx) return MPEGHB200_ERR_BAD_ARG;
  if (n_chains < 0 || (n_chains > 0 && (!d_coef || !d_filters || !d_chain_first)))
    return fail(ctx, MPEGHB200_ERR_BAD_ARG, "tns_dev: null buffer or negative chain count");
  if (n_chains == 0) return MPEGHB200_OK;
  MB_CUDA(ctx, cudaSetDevice(ctx->device));
  MB_CUDA(ctx, launch_pdl(tns_kernel, dim3((n_chains + kTnsWarps - 1) / kTnsWarps), dim3(32 * kTnsWarps), 0,
                          static_cast<cudaStream_t>(cuda_stream), d_coef, d_filters, d_chain_first, n_chains));
  ctx->launches.fetch_add(1);
  MB_CUDA(ctx, cudaGetLastError());
  return MPEGHB200_OK;
}

}  // extern "C"
