// ltpf.cu -- long-term post-filter of FD channels, in place on the time signal right after FD synthesis.
//
// Stands in for ia_core_coder_usac_ltpf_process (decoder/ia_core_coder_ltpf.c:395) with ia_core_coder_ltpf_filter
// (:134), ia_core_coder_ltpf_dec_params (:60), ia_core_coder_ltpf_get_lpc (:210) and ia_core_coder_ltpf_get_zir
// (:247); SURVEY.md section 8(f) rank 3.  The frame is filtered in three stretches: [0, 512) with the previous
// frame's parameters, a 128-sample transition (same parameters: plain; fade-in; fade-out; both frames filtering
// with different parameters: new filter minus the zero-input response of the LPC-shaped difference), [640, 1024)
// with the current parameters.
//
// out[n] = in[n] + g * sum_k out[n - pitch + k - 4] c1[fr][k] - 0.95 * (g * sum_k in[n - k] c2[gain_idx][k])
// is recursive with a lag of at least pitch - 3 samples (pitch >= 128 at 48 kHz), so a CTA of 128 threads walks
// each stretch in blocks of min(128, pitch - 3) samples, one sample per thread, every tap sum in the reference's
// order.  The strictly serial pieces of the rare ZIR transition (order-24 Levinson recursion, 128-step all-pole
// recursion, the running fade factor) are done by one thread / one lane with the history in registers.
// One CTA per channel-frame; bytes: 4096 r + 4096 w audio, (pitch_max + 20) * 8 state.
#include <cstring>

#include "common.h"

namespace mpeghb200 {
namespace {

constexpr int kFrame = 1024, kDelay = 512, kTrans = 128, kLpc = 24, kThreads = 128;

struct LtpfCfg {
  int pitch_min, pitch_fr1, pitch_fr2, pitch_res, pitch_max, hist, state_floats;
};

struct LtpfRom {
  float c1[2][8];
  float c2[4][7];
};

// the standard's interpolation / low-pass taps as the reference ROM prints them (decoder/ia_core_coder_rom.c:1247)
const LtpfRom kRom = {
    {{0.0000000f, 0.0304386f, 0.1162701f, 0.2195613f, 0.2674597f, 0.2195613f, 0.1162701f, 0.0304386f},
     {0.0076226f, 0.0676508f, 0.1700032f, 0.2547232f, 0.2547232f, 0.1700032f, 0.0676508f, 0.0076226f}},
    {{0.27150189f, 0.44286013f, 0.23027992f, 0.05759155f, -0.00172290f, -0.00045168f, -0.00005891f},
     {0.27581838f, 0.44682277f, 0.22783915f, 0.05410054f, -0.00353758f, -0.00092331f, -0.00011995f},
     {0.28044685f, 0.45103979f, 0.22519192f, 0.05037740f, -0.00545541f, -0.00141719f, -0.00018336f},
     {0.28543320f, 0.45554676f, 0.22230634f, 0.04638935f, -0.00749011f, -0.00193612f, -0.00024943f}}};

LtpfCfg make_cfg(int sample_rate) {
  LtpfCfg c{};
  // ia_core_coder_ltpf_init (:356), float arithmetic like the reference
  c.pitch_min = int((34.f * ((float(sample_rate) / 2.f) / 12800.f)) + 0.5f) * 2;
  c.pitch_fr1 = 320;
  c.pitch_fr2 = 324 - c.pitch_min;
  c.pitch_max = 54 + 6 * c.pitch_min;
  c.pitch_res = 2;
  c.hist = c.pitch_max + 4;
  c.state_floats = 16 + ((c.hist + 3) / 4) * 4;
  return c;
}

struct Params {
  int pitch, fr, gain_idx;
  float gain;
};

__device__ __forceinline__ Params decode_params(const LtpfCfg& c, int present, int lag_idx, int gain_idx) {
  Params p{0, 0, gain_idx, 0.0f};
  if (present) {
    const int a = (c.pitch_fr2 - c.pitch_min) * c.pitch_res, b = c.pitch_fr1 - c.pitch_fr2;
    if (lag_idx < a) {
      p.pitch = c.pitch_min + lag_idx / c.pitch_res;
      p.fr = lag_idx - (p.pitch - c.pitch_min) * c.pitch_res;
    } else if (lag_idx < a + b * (c.pitch_res >> 1)) {
      p.pitch = c.pitch_fr2 + lag_idx - a;
    } else {
      p.pitch = (lag_idx - a - b) * c.pitch_res + c.pitch_fr1;
    }
    p.gain = fmul(float(gain_idx + 1), 0.0625f);
  }
  return p;
}

// one stretch [n0, n1) of the frame; in / out are the frame starts inside the history-prefixed shared buffers
__device__ void filter_stretch(const float* in, float* out, int n0, int n1, const Params& p, int mode,
                               const float* zir, const float* alpha_tab, const LtpfRom& R) {
  const int t = threadIdx.x;
  if (p.gain == 0.0f) {
    for (int n = n0 + t; n < n1; n += kThreads) out[n] = in[n];
    __syncthreads();
    return;
  }
  float c1[8], c2[7];
#pragma unroll
  for (int k = 0; k < 8; ++k) c1[k] = R.c1[p.fr & 1][k];
#pragma unroll
  for (int k = 0; k < 7; ++k) c2[k] = R.c2[p.gain_idx & 3][k];
  const int block = min(kThreads, p.pitch - 3);  // samples that do not depend on each other
  for (int b0 = n0; b0 < n1; b0 += block) {
    const int n = b0 + t;
    if (t < block && n < n1) {
      float s1 = 0.0f, s2 = 0.0f;
#pragma unroll
      for (int k = 0; k < 8; ++k) s1 = fadd(s1, fmul(out[n - p.pitch + k - 4], c1[k]));
#pragma unroll
      for (int k = 0; k < 7; ++k) s2 = fadd(s2, fmul(in[n - k], c2[k]));
      const float gs1 = fmul(p.gain, s1), gs2 = fmul(0.95f, fmul(p.gain, s2));
      float y;
      if (mode == 0)
        y = fsub(fadd(in[n], gs1), gs2);
      else if (mode == 1)
        y = fsub(fsub(fadd(in[n], gs1), gs2), zir[n - n0]);
      else
        y = fadd(in[n], fmul(alpha_tab[n - n0], fsub(gs1, gs2)));
      out[n] = y;
    }
    __syncthreads();
  }
}

__global__ void __launch_bounds__(kThreads)
    ltpf_kernel(float* __restrict__ audio, const mpeghb200_ltpf_side* __restrict__ sides, float* __restrict__ state,
                const LtpfCfg C, const LtpfRom R) {
  pdl_prologue();
  extern __shared__ __align__(16) float sm[];
  float* bin = sm;                      // [8 + 1024]: 6 history samples right-aligned in the first 8
  float* bout = sm + 8 + kFrame;        // [hist4 + 1024]
  const int hist4 = ((C.hist + 3) / 4) * 4;
  float* in = bin + 8;
  float* out = bout + hist4;
  float* hbase = out - C.hist;  // the carried outputs end right in front of the frame
  __shared__ float zir[kTrans], alpha_tab[kTrans], r[kLpc + 1], a[kLpc + 1], zbuf[kLpc + kTrans];
  const int t = threadIdx.x;
  const mpeghb200_ltpf_side sd = sides[blockIdx.x];
  float* x = audio + size_t(sd.unit) * kFrame;
  float* st = state + size_t(sd.unit) * C.state_floats;

  Params prev;
  prev.pitch = __float_as_int(st[0]), prev.fr = __float_as_int(st[1]), prev.gain_idx = __float_as_int(st[2]);
  prev.gain = st[3];
  const Params cur = decode_params(C, sd.data_present, sd.pitch_lag_idx, sd.gain_idx);
  if (t < 6) bin[2 + t] = st[4 + t];
  for (int i = t; i < C.hist; i += kThreads) hbase[i] = st[16 + i];
#pragma unroll
  for (int q = 0; q < 2; ++q) reinterpret_cast<float4*>(in)[t + kThreads * q] = reinterpret_cast<const float4*>(x)[t + kThreads * q];
  __syncthreads();

  // [0, 512): previous parameters
  filter_stretch(in, out, 0, kDelay, prev, 0, nullptr, nullptr, R);

  // transition [512, 640)
  int mode;
  Params tp = cur;
  if (cur.gain == prev.gain && cur.pitch == prev.pitch && cur.fr == prev.fr) {
    mode = 0;
  } else if (prev.gain == 0.0f) {
    mode = 2;
  } else if (cur.gain == 0.0f) {
    mode = 3;
    tp = prev;
  } else {
    mode = 1;
    // LPC of the 256 outputs before the transition: one autocorrelation lag per thread (:218-227)
    if (t <= kLpc) {
      const float* at = out + kDelay;
      float s = 0.0f;
      for (int j = 0; j < kFrame / 4 - t; ++j) s = fadd(s, fmul(at[j - kFrame / 4], at[j - kFrame / 4 + t]));
      r[t] = s;
    }
    __syncthreads();
    if (t < 32) {  // Levinson-Durbin (:229-262) on warp 0: lane j holds rr[j] and aa[j]
      // The inner sums are strictly ordered additions; what the warp buys is that the products, the table values
      // and the symmetric coefficient update are there in one step each instead of as dependent local-memory
      // accesses of one thread (this recursion used to cost as much as the 128-step zero-input response below).
      // Every lane forms the same sum, so the early exit is uniform.
      constexpr unsigned kAll = 0xffffffffu;
      float rrl = (t <= kLpc) ? r[t] : 0.0f;
      if (t == 0) rrl = fmul(fmaxf(rrl, 100.0f), 1.0001f);
      const float rr0 = __shfl_sync(kAll, rrl, 0), rr1 = __shfl_sync(kAll, rrl, 1);
      float rci = __fdiv_rn(-rr1, rr0);
      float aal = (t == 0) ? 1.0f : (t == 1 ? rci : 0.0f);
      float sigma2 = fadd(rr0, fmul(rr1, rci));
#pragma unroll
      for (int i = 2; i <= kLpc; ++i) {
        const float p = fmul(__shfl_sync(kAll, rrl, (i - t) & 31), aal);  // rr[i - j] * aa[j] on lane j
        float sum = 0.0f;
#pragma unroll
        for (int j = 0; j < i; ++j) sum = fadd(sum, __shfl_sync(kAll, p, j));
        rci = __fdiv_rn(-sum, sigma2);
        sigma2 = fmul(sigma2, fsub(1.0f, fmul(rci, rci)));
        if (sigma2 <= 1.0E-09f) {
          if (t >= i) aal = 0.0f;
          break;
        }
        const float other = __shfl_sync(kAll, aal, (i - t) & 31);  // aa[i - j], old value
        if (t >= 1 && t < i) aal = fadd(aal, fmul(rci, other));
        if (t == i) aal = rci;
      }
      if (t <= kLpc) a[t] = aal;
    }
    // start values of the zero-input response: the 24 samples before the transition (:270-288)
    if (t >= 32 && t < 32 + kLpc) {
      const int n = t - 32;
      const float* i2 = in + kDelay;
      const float* o2 = out + kDelay;
      float s1 = 0.0f, s2 = 0.0f;
#pragma unroll
      for (int k = 0; k < 8; ++k) s1 = fadd(s1, fmul(o2[n - kLpc - cur.pitch + k - 4], R.c1[cur.fr & 1][k]));
#pragma unroll
      for (int k = 0; k < 7; ++k) s2 = fadd(s2, fmul(i2[n - kLpc - k], R.c2[cur.gain_idx & 3][k]));
      zbuf[n] = fsub(fsub(i2[n - kLpc], fmul(0.95f, fmul(cur.gain, s2))), fsub(o2[n - kLpc], fmul(cur.gain, s1)));
    }
    __syncthreads();
    if (t == 0) {  // all-pole recursion over the transition, the last 24 values in a register ring (:290-297)
      // ring[e % 24] holds the value produced at step e; with chunks of 24 fully unrolled steps every ring index
      // is a compile-time constant, so the 24 history values stay in registers
      float lpc[kLpc], ring[kLpc];
#pragma unroll
      for (int k = 0; k < kLpc; ++k) lpc[k] = a[k + 1], ring[k] = zbuf[k];
      for (int n0 = 0; n0 < kTrans; n0 += kLpc) {
#pragma unroll
        for (int e = 0; e < kLpc; ++e) {
          if (n0 + e < kTrans) {
            float v = 0.0f;
#pragma unroll
            for (int k = 0; k < kLpc; ++k) v = fsub(v, fmul(lpc[k], ring[(e - 1 - k + 2 * kLpc) % kLpc]));
            ring[e] = v;
            zbuf[kLpc + n0 + e] = v;
          }
        }
      }
      float alpha = 1.0f;
      const float step = __fdiv_rn(1.0f, float(kTrans / 2));
      for (int n = 0; n < kTrans / 2; ++n) zir[n] = zbuf[kLpc + n];
      for (int n = kTrans / 2; n < kTrans; ++n) {
        zir[n] = fmul(zbuf[kLpc + n], alpha);
        alpha = fsub(alpha, step);
      }
    }
    __syncthreads();
  }
  if (mode >= 2) {  // the running fade factor, accumulated like the reference does (:146-156, :196-199)
    if (t == 0) {
      float alpha = mode == 2 ? 0.0f : 1.0f;
      const float step = mode == 2 ? __fdiv_rn(1.0f, float(kTrans)) : __fdiv_rn(-1.0f, float(kTrans));
      for (int n = 0; n < kTrans; ++n) {
        alpha_tab[n] = alpha;
        alpha = fadd(alpha, step);
      }
    }
    __syncthreads();
  }
  filter_stretch(in, out, kDelay, kDelay + kTrans, tp, mode, zir, alpha_tab, R);

  // [640, 1024): current parameters
  filter_stretch(in, out, kDelay + kTrans, kFrame, cur, 0, nullptr, nullptr, R);

  // output and state: last 6 inputs, last hist outputs, the parameters of this frame
#pragma unroll
  for (int q = 0; q < 2; ++q) reinterpret_cast<float4*>(x)[t + kThreads * q] = reinterpret_cast<const float4*>(out)[t + kThreads * q];
  if (t < 6) st[4 + t] = in[kFrame - 6 + t];
  for (int i = t; i < C.hist; i += kThreads) st[16 + i] = hbase[kFrame + i];
  if (t == 0) {
    st[0] = __int_as_float(cur.pitch), st[1] = __int_as_float(cur.fr), st[2] = __int_as_float(int(sd.gain_idx));
    st[3] = cur.gain;
  }
}

}  // namespace
}  // namespace mpeghb200

using namespace mpeghb200;

extern "C" {

size_t mpeghb200_ltpf_state_floats(int sample_rate) { return sample_rate > 0 ? size_t(make_cfg(sample_rate).state_floats) : 0; }

mpeghb200_status mpeghb200_ltpf_dev(mpeghb200_ctx* ctx, int sample_rate, float* d_audio,
                                    const mpeghb200_ltpf_side* d_side, float* d_state, int n_units, void* cuda_stream) {
  if (!ctx) return MPEGHB200_ERR_BAD_ARG;
  if (n_units < 0 || (n_units > 0 && (!d_audio || !d_side || !d_state)))
    return fail(ctx, MPEGHB200_ERR_BAD_ARG, "ltpf_dev: null buffer or negative unit count");
  if (sample_rate < 8000 || sample_rate > 96000) return fail(ctx, MPEGHB200_ERR_UNSUPPORTED, "ltpf_dev: sampling rate");
  if (n_units == 0) return MPEGHB200_OK;
  MB_CUDA(ctx, cudaSetDevice(ctx->device));
  const LtpfCfg C = make_cfg(sample_rate);
  if (C.pitch_min < 8) return fail(ctx, MPEGHB200_ERR_UNSUPPORTED, "ltpf_dev: sampling rate too low for a block walk");
  const size_t smem = sizeof(float) * (8 + kFrame + ((C.hist + 3) / 4) * 4 + kFrame);
  MB_CUDA(ctx, launch_pdl(ltpf_kernel, dim3(n_units), dim3(kThreads), smem, static_cast<cudaStream_t>(cuda_stream), d_audio,
                          d_side, d_state, C, kRom));
  ctx->launches.fetch_add(1);
  MB_CUDA(ctx, cudaGetLastError());
  return MPEGHB200_OK;
}

}  // extern "C"
