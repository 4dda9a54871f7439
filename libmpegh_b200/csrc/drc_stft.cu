// drc_stft.cu -- DRC in the STFT domain, fused: domain switcher time -> frequency, sub-band gains, frequency -> time.
//
// Stands in for the chain at decoder/ia_core_coder_decode_main.c:2128-2141: impeghd_domain_switcher_process
// (decoder/ia_core_coder_process.c:505; t2f :556-585, f2t :589-619) around impd_drc_apply_gains_subband
// (decoder/impd_drc_process.c:243, see gain_subband_kernel in mix.cu for the gain rules).  SURVEY.md section 8f rank 2.
// Per channel-frame and slot of 256 samples: sine-windowed [previous 256 | current 256] -> 512-point e^{-j}
// transform -> bins 0..255 + Nyquist times the DRC gain(s) -> Hermitian extension -> 512-point e^{+j} transform
// -> /512 -> windowed overlap-add.  Channels without DRC still take the round trip (the reference transforms all).
//
// One warp per channel-frame; the spectrum never leaves registers / the warp's shared buffer.  The two transforms
// are the format converter's (fft_exact.cuh: fft512_tail + a fused last radix-2 pass), with the families swapped.
// HBM traffic per channel-frame: 4 KB in, 4 KB out, 2 KB of state read and written.
#include "fft_exact.cuh"

namespace mpeghb200 {
namespace {

using namespace fftx;

constexpr int kDsWarps = 4;
constexpr int kDsBuf = 544;  // float2 per warp: 512 + padding

struct DsParams {
  float* audio;  // [n_ch][1024], in place
  float* state;  // [n_ch][512]: previous input slot, overlap of the output
  const mpeghb200_drc_sb_channel* chan;  // [n_sets][n_ch]
  const float* gains;                    // [n_seq][4]
  const float* weights;                  // [rows][256]
  float* spec;                           // stand-alone transforms: [n_ch][2048] = re [4][256] | im [4][256], im[slot][0] = Nyquist
  const float* win;                      // [256]
  const float4* eff8;
  const float2* rad2;
  int n_ch, n_sets;
};

// MODE 0: the fused step.  MODE 1 / 2: one half of the domain switcher on its own (time -> frequency into P.spec,
// frequency -> time out of it), for callers that run the gain application as a separate step (the drop-in library).
template <int MODE>
__global__ void __launch_bounds__(32 * kDsWarps) drc_stft_kernel(const __grid_constant__ DsParams P) {
  __shared__ __align__(16) float2 work[kDsWarps][kDsBuf];
  __shared__ __align__(16) float twf[500];
  __shared__ float2 rad2_s[256];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const RowTw tw = fill_row_tw(twf, P.eff8, 498);
  for (int i = tid; i < 256; i += blockDim.x) rad2_s[i] = __ldg(P.rad2 + i);
  __syncthreads();
  const int c = blockIdx.x * kDsWarps + warp;
  if (c >= P.n_ch) return;  // whole warp; no CTA-wide barrier below
  float2* buf = work[warp];
  float* x = P.audio + size_t(c) * 1024;
  float* st = P.state + size_t(c) * 512;
  float* sp = MODE == 0 ? nullptr : P.spec + size_t(c) * 2048;
  const int r3 = rev32(lane);
  float wv[8], wr[8], hist[8], ohist[8];  // the lane's sample positions are i = r3 + 32 u (see format_conv.cu)
#pragma unroll
  for (int u = 0; u < 8; ++u) {
    wv[u] = __ldg(P.win + r3 + 32 * u);
    wr[u] = __ldg(P.win + 255 - r3 - 32 * u);
    hist[u] = MODE != 2 ? st[r3 + 32 * u] : 0.0f;
    ohist[u] = MODE != 1 ? st[256 + r3 + 32 * u] : 0.0f;
  }
  for (int slot = 0; slot < 4; ++slot) {
    // ---- time -> frequency (process.c:556-585) ----
    float2 v[16];
    float* xs = x + slot * 256;
    float re[8], im[8], nyq = 0.0f;  // bins k = lane + 32 u
    if (MODE != 2) {
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const float cur = xs[r3 + 32 * u];
        v[u] = make_float2(fmul(hist[u], wv[u]), 0.0f);
        v[8 + u] = make_float2(fmul(wr[u], cur), 0.0f);
        hist[u] = cur;
      }
      fft512_tail<false>(v, buf, lane, tw);
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const int k = lane + 32 * u;
        const float2 x0 = buf[phys(k)];
        float2 t = buf[phys(k + 256)];
        const float2 cs = rad2_s[k];
        rot<false>(t, cs.x, cs.y);
        re[u] = fadd(x0.x, t.x), im[u] = fadd(x0.y, t.y);
        if (k == 0) nyq = fsub(x0.x, t.x);
      }
      __syncwarp();
    }
    if (MODE == 1) {  // the reference's packed layout: Im of bin 0 is dropped, its slot carries Re of bin 256
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const int k = lane + 32 * u;
        sp[slot * 256 + k] = re[u];
        sp[1024 + slot * 256 + k] = k == 0 ? nyq : im[u];
      }
      continue;
    }
    if (MODE == 2) {
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const int k = lane + 32 * u;
        re[u] = sp[slot * 256 + k];
        im[u] = sp[1024 + slot * 256 + k];
        if (k == 0) nyq = im[u];
      }
    }
    // ---- sub-band gains, one selected DRC set after the other (impd_drc_process.c:640-651, :243-385) ----
    for (int set = 0; MODE == 0 && set < P.n_sets; ++set) {
      const mpeghb200_drc_sb_channel ch = P.chan[size_t(set) * P.n_ch + c];
      if (ch.first_seq < 0) continue;
      const int nb = ch.band_count < 1 ? 1 : (ch.band_count > 4 ? 4 : ch.band_count);
      if (nb == 1) {
        const float g = __ldg(P.gains + size_t(ch.first_seq) * 4 + slot);
#pragma unroll
        for (int u = 0; u < 8; ++u) re[u] = fmul(re[u], g), im[u] = fmul(im[u], g);
        nyq = fmul(nyq, g);
      } else {
        float gl[4];
#pragma unroll
        for (int b = 0; b < 4; ++b) gl[b] = (b < nb) ? __ldg(P.gains + size_t(ch.first_seq + b) * 4 + slot) : 0.0f;
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          float g = 0.0f;
#pragma unroll
          for (int b = 0; b < 4; ++b)
            if (b < nb) g = fadd(g, fmul(__ldg(P.weights + size_t(ch.weight_row + b) * 256 + lane + 32 * u), gl[b]));
          re[u] = fmul(re[u], g);
          if (lane + 32 * u != 0) im[u] = fmul(im[u], g);  // im of bin 0 is identically the packed Nyquist slot
        }
        float gn = 0.0f;
#pragma unroll
        for (int b = 0; b < 4; ++b)
          if (b < nb) gn = fadd(gn, fmul(__ldg(P.weights + size_t(ch.weight_row + b) * 256 + 255), gl[b]));
        nyq = fmul(nyq, gn);
      }
    }
    // ---- frequency -> time (process.c:589-619) ----
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const int k = lane + 32 * u;
      if (k == 0) {
        buf[phys(0)] = make_float2(re[u], 0.0f);
        buf[phys(256)] = make_float2(nyq, 0.0f);
      } else {
        buf[phys(k)] = make_float2(re[u], im[u]);
        buf[phys(512 - k)] = make_float2(re[u], -im[u]);
      }
    }
    __syncwarp();
#pragma unroll
    for (int t = 0; t < 16; ++t) v[t] = buf[phys(r3 + 32 * t)];
    __syncwarp();
    fft512_tail<true>(v, buf, lane, tw);
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const int i = r3 + 32 * u;
      const float2 x0 = buf[phys(i)], t = buf[phys(i + 256)], cs = rad2_s[i];
      const float tr = fadd(fmul(t.x, cs.x), fmul(t.y, cs.y));  // Re(w t), e^{+j} family
      const float a = fmul(fadd(x0.x, tr), 0.001953125f);      // == x / 512 bit for bit
      xs[i] = fadd(fmul(a, wv[u]), fmul(ohist[u], wr[u]));
      ohist[u] = fmul(fsub(x0.x, tr), 0.001953125f);
    }
    __syncwarp();
  }
#pragma unroll
  for (int u = 0; u < 8; ++u) {
    if (MODE != 2) st[r3 + 32 * u] = hist[u];
    if (MODE != 1) st[256 + r3 + 32 * u] = ohist[u];
  }
}

}  // namespace
}  // namespace mpeghb200

using namespace mpeghb200;

extern "C" {

static mpeghb200_status stft_launch(mpeghb200_ctx* ctx, int mode, float* d_audio, float* d_spec, float* d_state,
                                    const mpeghb200_drc_sb_channel* d_chan, int n_sets, const float* d_gains,
                                    const float* d_weights, int n_channels, void* cuda_stream) {
  MB_CUDA(ctx, cudaSetDevice(ctx->device));
  static float* d_win[64] = {};
  if (ctx->device >= 64) return fail(ctx, MPEGHB200_ERR_UNSUPPORTED, "drc_stft_dev: device index above 63");
  if (!d_win[ctx->device]) {
    const std::vector<float>& win = rom_table(ROM_SINE_256);
    float* p = nullptr;
    MB_CUDA(ctx, cudaMalloc(&p, 256 * sizeof(float)));
    MB_CUDA(ctx, cudaMemcpy(p, win.data(), 256 * sizeof(float), cudaMemcpyHostToDevice));
    d_win[ctx->device] = p;
  }
  DsParams P{};
  P.audio = d_audio, P.state = d_state, P.chan = d_chan, P.gains = d_gains, P.weights = d_weights, P.spec = d_spec;
  P.win = d_win[ctx->device], P.eff8 = ctx->tables.eff8, P.rad2 = ctx->tables.rad2_512;
  P.n_ch = n_channels, P.n_sets = n_sets;
  const dim3 grid((n_channels + kDsWarps - 1) / kDsWarps), block(32 * kDsWarps);
  cudaStream_t st = static_cast<cudaStream_t>(cuda_stream);
  if (mode == 0)
    drc_stft_kernel<0><<<grid, block, 0, st>>>(P);
  else if (mode == 1)
    drc_stft_kernel<1><<<grid, block, 0, st>>>(P);
  else
    drc_stft_kernel<2><<<grid, block, 0, st>>>(P);
  ctx->launches.fetch_add(1);
  MB_CUDA(ctx, cudaGetLastError());
  return MPEGHB200_OK;
}

mpeghb200_status mpeghb200_drc_stft_dev(mpeghb200_ctx* ctx, float* d_audio, float* d_state,
                                        const mpeghb200_drc_sb_channel* d_chan, int n_sets, const float* d_gains,
                                        const float* d_weights, int n_channels, void* cuda_stream) {
  if (!ctx) return MPEGHB200_ERR_BAD_ARG;
  if (n_channels < 0 || n_sets < 0 || n_sets > 3 || (n_channels > 0 && (!d_audio || !d_state)) ||
      (n_channels > 0 && n_sets > 0 && (!d_chan || !d_gains)))
    return fail(ctx, MPEGHB200_ERR_BAD_ARG, "drc_stft_dev: null buffer, negative channel count or more than 3 sets");
  if (n_channels == 0) return MPEGHB200_OK;
  return stft_launch(ctx, 0, d_audio, nullptr, d_state, d_chan, n_sets, d_gains, d_weights, n_channels, cuda_stream);
}

mpeghb200_status mpeghb200_stft_dev(mpeghb200_ctx* ctx, float* d_audio, float* d_spec, float* d_state, int to_frequency,
                                    int n_channels, void* cuda_stream) {
  if (!ctx) return MPEGHB200_ERR_BAD_ARG;
  if (n_channels < 0 || (n_channels > 0 && (!d_audio || !d_spec || !d_state)))
    return fail(ctx, MPEGHB200_ERR_BAD_ARG, "stft_dev: null buffer or negative channel count");
  if (n_channels == 0) return MPEGHB200_OK;
  return stft_launch(ctx, to_frequency ? 1 : 2, d_audio, d_spec, d_state, nullptr, 0, nullptr, nullptr, n_channels,
                     cuda_stream);
}

}  // extern "C"
