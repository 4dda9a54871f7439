// drc_td.cu -- time-domain multi-band DRC: Linkwitz-Riley filter banks, per-band gain curves, band sum.
//
// Stands in for impd_drc_filter_banks_process (decoder/impd_drc_process.c:399) with the banks of
// decoder/impd_drc_filter_bank.c (low/high pair :410, all-pass :367, two/three/four-band trees :486, :511, :550,
// phase-alignment cascade :601) and for the multiply-and-sum of impd_drc_apply_gains_and_add (:68-225), as
// impd_drc_td_process (:529) chains them for one selected DRC set.  SURVEY.md section 8f rank 2.  Coefficient design
// (impd_drc_init_all_filter_banks), gain decoding and interpolation stay on the host.
//
// Every filter is a serial recurrence over the frame, and the all-pass sections of the reference run BACKWARDS over
// it (filter_bank.c:380), so the stages of a tree cannot be fused per sample: a warp owns a channel-frame, keeps its
// band signals in four shared-memory rows, and runs the independent recurrences of a stage on separate lanes (up to
// four: the low/high pairs of stage 3 of the four-band tree) -- through one piece of code, parameters picked by lane
// number, because divergent branches of a warp run one after the other; the low branch of a pair overwrites the
// pair's input row block by block, behind a warp barrier that follows the block loads of both branches.  The sections of the phase-alignment cascade are applied per sample in one loop
// (same order per section as the reference's section-after-section passes).  All 32 lanes move the data in and
// apply the gains and the band sum.  Bound: the recurrence latency (about 16 cycles per sample and stage).
#include "common.h"

namespace mpeghb200 {
namespace {

constexpr int kN = 1024;
constexpr int kRows = 4;
constexpr int kTdWarps = 2;
constexpr int kStateFloats = 17 * 4;

// One branch of impd_drc_apply_low_high_filter: two cascaded sections with the same coefficients.  `out` may be the
// row another lane of the stage (the other branch of the pair) reads as `in`: the loop works in blocks of four
// samples, and every lane of `mask` (the lanes of this stage) holds its block before any of them stores
// (__syncwarp); the block after that is already on its way, so the shared-memory latency stays off the recurrence.
__device__ __forceinline__ void lh_branch(const mpeghb200_drc_biquad f, float* __restrict__ st, const float* in,
                                          float* out, unsigned mask) {
  float s1 = st[0], s2 = st[1], s3 = st[2], s4 = st[3];
  float4 nx = *reinterpret_cast<const float4*>(in);
  for (int i = 0; i < kN; i += 4) {
    const float4 x4 = nx;
    if (i + 4 < kN) nx = *reinterpret_cast<const float4*>(in + i + 4);
    __syncwarp(mask);
    const float xs[4] = {x4.x, x4.y, x4.z, x4.w};
    float os[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const float x = xs[e];
      const float t = fadd(s1, fmul(f.b0, x));
      s1 = fadd(fsub(fmul(f.b1, x), fmul(f.a1, t)), s2);
      s2 = fsub(fmul(f.b2, x), fmul(f.a2, t));
      const float o = fadd(s3, fmul(f.b0, t));
      s3 = fadd(fsub(fmul(f.b1, t), fmul(f.a1, o)), s4);
      s4 = fsub(fmul(f.b2, t), fmul(f.a2, o));
      os[e] = o;
    }
    *reinterpret_cast<float4*>(out + i) = make_float4(os[0], os[1], os[2], os[3]);
  }
  st[0] = s1, st[1] = s2, st[2] = s3, st[3] = s4;
}

// impd_drc_iir_second_order_all_pass_filter: last sample first; state in slots 0 and 2 (x_p[2 ch], y_p[2 ch])
__device__ __forceinline__ void allpass(const mpeghb200_drc_biquad f, float* __restrict__ st, const float* in,
                                        float* out) {
  float s1 = st[0], s2 = st[2];
  // four samples per step, loaded before and stored after the step: `out` may be `in` (the loads of the next step
  // would otherwise wait behind this step's stores, and the shared-memory latency would sit on the recurrence)
  for (int i = kN - 4; i >= 0; i -= 4) {
    const float4 x4 = *reinterpret_cast<const float4*>(in + i);
    const float xs[4] = {x4.x, x4.y, x4.z, x4.w};
    float ys[4];
#pragma unroll
    for (int e = 3; e >= 0; --e) {
      const float x = xs[e];
      const float y = fadd(fmul(f.b0, x), s1);
      s1 = fadd(fsub(fmul(f.b1, x), fmul(f.a1, y)), s2);
      s2 = fsub(fmul(f.b2, x), fmul(f.a2, y));
      ys[e] = y;
    }
    *reinterpret_cast<float4*>(out + i) = make_float4(ys[0], ys[1], ys[2], ys[3]);
  }
  st[0] = s1, st[2] = s2;
}

// Phase-alignment cascade (filter_bank.c:601): sections NC-1 .. 0, in place, last sample first, applied per sample in
// one loop (each section sees its samples in the reference's order).  One lane.
template <int NC>
__device__ __noinline__ void cascade(const mpeghb200_drc_td_bank& bank, float* __restrict__ st, float* row) {
  float s1[NC], s2[NC];
  mpeghb200_drc_biquad f[NC];
#pragma unroll
  for (int k = 0; k < NC; ++k) f[k] = bank.ap[k], s1[k] = st[4 * (8 + k)], s2[k] = st[4 * (8 + k) + 2];
  for (int i = kN - 4; i >= 0; i -= 4) {
    const float4 x4 = *reinterpret_cast<const float4*>(row + i);
    float v[4] = {x4.x, x4.y, x4.z, x4.w};
#pragma unroll
    for (int e = 3; e >= 0; --e) {
#pragma unroll
      for (int k = NC - 1; k >= 0; --k) {
        const float y = fadd(fmul(f[k].b0, v[e]), s1[k]);
        s1[k] = fadd(fsub(fmul(f[k].b1, v[e]), fmul(f[k].a1, y)), s2[k]);
        s2[k] = fsub(fmul(f[k].b2, v[e]), fmul(f[k].a2, y));
        v[e] = y;
      }
    }
    *reinterpret_cast<float4*>(row + i) = make_float4(v[0], v[1], v[2], v[3]);
  }
#pragma unroll
  for (int k = 0; k < NC; ++k) st[4 * (8 + k)] = s1[k], st[4 * (8 + k) + 2] = s2[k];
}

__global__ void __launch_bounds__(32 * kTdWarps)
    drc_td_kernel(float* __restrict__ audio, float* __restrict__ state, const mpeghb200_drc_td_channel* __restrict__ chan,
                  const mpeghb200_drc_td_bank* __restrict__ banks, const float* __restrict__ gains, int n_ch) {
  extern __shared__ __align__(16) float rows_all[];  // [kTdWarps][kRows][kN]
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int c = blockIdx.x * kTdWarps + warp;
  if (c >= n_ch) return;  // whole warp; no CTA-wide barrier below
  float* B[kRows];
#pragma unroll
  for (int r = 0; r < kRows; ++r) B[r] = rows_all + (size_t(warp) * kRows + r) * kN;
  const mpeghb200_drc_td_channel ch = chan[c];
  const mpeghb200_drc_td_bank& bank = banks[ch.bank];
  float* x = audio + size_t(c) * kN;
  float* st = state + size_t(c) * kStateFloats;
  const int nb = bank.n_bands < 1 ? 1 : (bank.n_bands > 4 ? 4 : bank.n_bands);
  const int n_casc = bank.n_cascade < 0 ? 0 : (bank.n_cascade > 9 ? 9 : bank.n_cascade);
  const bool grouped = ch.first_seq >= 0;
  if (!grouped && n_casc == 0) return;  // untouched channel
  for (int i = lane; i < kN / 4; i += 32)
    reinterpret_cast<float4*>(B[0])[i] = __ldcs(reinterpret_cast<const float4*>(x) + i);
  __syncwarp();
  // ---- phase-alignment cascade, sections n_casc-1 .. 0, in place, last sample first (filter_bank.c:601) ----
  if (n_casc > 0 && lane == 0) {
    switch (n_casc) {  // section count at compile time: per-sample tests of it cost more than the sections
      case 1: cascade<1>(bank, st, B[0]); break;
      case 2: cascade<2>(bank, st, B[0]); break;
      case 3: cascade<3>(bank, st, B[0]); break;
      case 4: cascade<4>(bank, st, B[0]); break;
      case 5: cascade<5>(bank, st, B[0]); break;
      case 6: cascade<6>(bank, st, B[0]); break;
      case 7: cascade<7>(bank, st, B[0]); break;
      case 8: cascade<8>(bank, st, B[0]); break;
      default: cascade<9>(bank, st, B[0]); break;
    }
  }
  __syncwarp();
  if (!grouped) {  // outside every channel group: only the cascade applies (impd_drc_process.c:204-207)
    for (int i = lane; i < kN / 4; i += 32) __stcs(reinterpret_cast<float4*>(x) + i, reinterpret_cast<float4*>(B[0])[i]);
    return;
  }
  // ---- band split; row plan: no lane writes a row another lane reads in the same stage ----
  const float* band[4] = {B[0], B[0], B[0], B[0]};
  // Lanes that run concurrently must run the SAME code (separate branches of a warp are serialised): a stage picks
  // its per-lane filter, memory and rows by lane number and calls one function.
  if (nb >= 2) {  // stage 1: B0 -> low B0 (in place, see lh_branch), high B1
    if (lane < 2) lh_branch(bank.f[lane], st + 4 * lane, B[0], lane ? B[1] : B[0], 0x3u);
    __syncwarp();
    band[0] = B[0], band[1] = B[1];
  }
  if (nb == 3) {  // all-pass on the high branch (-> band 2), second split of the low branch
    if (lane == 0) allpass(bank.f[2], st + 8, B[1], B[1]);
    if (lane < 2) lh_branch(bank.f[3 + lane], st + 4 * (3 + lane), B[0], lane ? B[2] : B[0], 0x3u);
    __syncwarp();
    band[0] = B[0], band[1] = B[2], band[2] = B[1];
  } else if (nb == 4) {
    if (lane < 2) {  // each row: one reader / writer
      float* r = lane ? B[1] : B[0];
      allpass(bank.f[2 + lane], st + 4 * (2 + lane), r, r);
    }
    __syncwarp();
    if (lane < 4) {  // lanes 0, 1 split B0 into (B0, B2); lanes 2, 3 split B1 into (B1, B3)
      float* out = lane == 0 ? B[0] : (lane == 1 ? B[2] : (lane == 2 ? B[1] : B[3]));
      lh_branch(bank.f[4 + lane], st + 4 * (4 + lane), lane < 2 ? B[0] : B[1], out, 0xFu);
    }
    __syncwarp();
    band[0] = B[0], band[1] = B[2], band[2] = B[1], band[3] = B[3];
  }
  // ---- gains and band sum, highest band first (impd_drc_process.c:150-155, :211-219) ----
  const float* g = gains + size_t(ch.first_seq) * kN;
  for (int i = lane; i < kN; i += 32) {
    float sum = 0.0f;
#pragma unroll
    for (int k = 3; k >= 0; --k)
      if (k < nb) sum = fadd(sum, fmul(band[k][i], __ldg(g + k * kN + i)));
    __stcs(x + i, sum);
  }
}

}  // namespace
}  // namespace mpeghb200

using namespace mpeghb200;

extern "C" {

size_t mpeghb200_drc_td_state_floats(void) { return kStateFloats; }

mpeghb200_status mpeghb200_drc_td_dev(mpeghb200_ctx* ctx, float* d_audio, float* d_state,
                                      const mpeghb200_drc_td_channel* d_chan, const mpeghb200_drc_td_bank* d_banks,
                                      const float* d_gains, int n_channels, void* cuda_stream) {
  if (!ctx) return MPEGHB200_ERR_BAD_ARG;
  if (n_channels < 0 || (n_channels > 0 && (!d_audio || !d_state || !d_chan || !d_banks || !d_gains)))
    return fail(ctx, MPEGHB200_ERR_BAD_ARG, "drc_td_dev: null buffer or negative channel count");
  if (n_channels == 0) return MPEGHB200_OK;
  MB_CUDA(ctx, cudaSetDevice(ctx->device));
  const int smem = kTdWarps * kRows * kN * int(sizeof(float));
  MB_CUDA(ctx, cudaFuncSetAttribute(drc_td_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  drc_td_kernel<<<(n_channels + kTdWarps - 1) / kTdWarps, 32 * kTdWarps, smem, static_cast<cudaStream_t>(cuda_stream)>>>(
      d_audio, d_state, d_chan, d_banks, d_gains, n_channels);
  ctx->launches.fetch_add(1);
  MB_CUDA(ctx, cudaGetLastError());
  return MPEGHB200_OK;
}

}  // extern "C"
