// mix.cu -- the element-wise steps that sit between the renderers and the limiter:
//   bus add       time_sample_vector[ch][s] = a[ch][s] + b[ch][s]: bed + objects, bed + HOA, HOA + objects
//                 (decoder/ia_core_coder_decode_main.c:2185-2195, :2203-2213, :2276-2302)
//   channel mask  a member channel of a switched-off metadata group is replaced by zeros, the others pass
//                 (impeghd_mdp_dec_process, decoder/ia_core_coder_decode_main.c:906-1070: the group / preset logic
//                 that decides on or off stays with the host, the kernel applies the per-channel verdict)
//   gain apply    audio[ch][s] *= gain[seq(ch)][s]: the time-domain application of single-band DRC gain sequences
//                 (impd_drc_apply_gains_and_add, decoder/impd_drc_process.c:68-175 with impd_apply_gains == 1,
//                 SURVEY.md section 8f rank 2); decoding, selecting and interpolating the gains (spline nodes,
//                 libm exp/log) stays with the host, which hands over one 1024-sample curve per sequence and frame
// All are pure HBM streams: float4 per thread.
#include "common.h"

namespace mpeghb200 {
namespace {

__global__ void __launch_bounds__(256)
    bus_add_kernel(const float4* __restrict__ a, const float4* __restrict__ b, float4* __restrict__ out, size_t n4) {
  for (size_t i = blockIdx.x * size_t(blockDim.x) + threadIdx.x; i < n4; i += size_t(gridDim.x) * blockDim.x) {
    const float4 x = a[i], y = b[i];
    out[i] = make_float4(fadd(x.x, y.x), fadd(x.y, y.y), fadd(x.z, y.z), fadd(x.w, y.w));
  }
}

__global__ void __launch_bounds__(256)
    channel_mask_kernel(float4* __restrict__ audio, const uint8_t* __restrict__ on, int n_channels) {
  const int c = blockIdx.x;
  if (c < n_channels && !on[c]) audio[size_t(c) * 256 + threadIdx.x] = make_float4(0.f, 0.f, 0.f, 0.f);
}

__global__ void __launch_bounds__(256)
    gain_apply_kernel(float4* __restrict__ audio, const float4* __restrict__ gains, const int32_t* __restrict__ seq_of_ch,
                      int n_channels) {
  const int c = blockIdx.x;
  if (c >= n_channels) return;
  const int q = seq_of_ch[c];
  if (q < 0) return;  // channel without a DRC channel group
  float4 x = audio[size_t(c) * 256 + threadIdx.x];
  const float4 g = gains[size_t(q) * 256 + threadIdx.x];
  x.x = fmul(x.x, g.x), x.y = fmul(x.y, g.y), x.z = fmul(x.z, g.z), x.w = fmul(x.w, g.w);
  audio[size_t(c) * 256 + threadIdx.x] = x;
}

// Planar <-> interleaved.  The reference's limiter works on an interleaved copy of the frame ([sample][channel],
// impegh_dec_peak_limiter, decoder/ia_core_coder_decode_main.c:1429-1460); the kernels are planar.  A CTA moves a tile of
// 128 samples of one stream through shared memory so that both sides are contiguous in global memory.
constexpr int kIlvTile = 128;
__global__ void __launch_bounds__(256)
    interleave_kernel(const float* __restrict__ in, float* __restrict__ out, int n_ch, int n_samples, int to_planar) {
  extern __shared__ float tile[];  // [n_ch][kIlvTile + 1]
  const size_t s = blockIdx.y;
  const int j0 = blockIdx.x * kIlvTile;
  const int nj = min(kIlvTile, n_samples - j0);
  const float* src = in + s * size_t(n_ch) * n_samples;
  float* dst = out + s * size_t(n_ch) * n_samples;
  const int total = nj * n_ch;
  if (to_planar) {
    for (int e = threadIdx.x; e < total; e += blockDim.x) {  // e = j * n_ch + c, contiguous in the source
      const int j = e / n_ch, c = e - j * n_ch;
      tile[c * (kIlvTile + 1) + j] = src[size_t(j0) * n_ch + e];
    }
    __syncthreads();
    for (int e = threadIdx.x; e < total; e += blockDim.x) {  // e = c * nj + j, rows contiguous in the destination
      const int c = e / nj, j = e - c * nj;
      dst[size_t(c) * n_samples + j0 + j] = tile[c * (kIlvTile + 1) + j];
    }
  } else {
    for (int e = threadIdx.x; e < total; e += blockDim.x) {
      const int c = e / nj, j = e - c * nj;
      tile[c * (kIlvTile + 1) + j] = src[size_t(c) * n_samples + j0 + j];
    }
    __syncthreads();
    for (int e = threadIdx.x; e < total; e += blockDim.x) {
      const int j = e / n_ch, c = e - j * n_ch;
      dst[size_t(j0) * n_ch + e] = tile[c * (kIlvTile + 1) + j];
    }
  }
}

// DRC gains in the STFT domain: 4 time slots x 256 bands per channel-frame, re / im planes; im[0] of a slot is the
// Nyquist bin.  impd_drc_apply_gains_subband (decoder/impd_drc_process.c:243-385): a single-band group multiplies
// everything by the slot's gain; a multi-band group forms the band gain as the ordered sum of overlap weight x
// sequence gain (:350-357) and gives the Nyquist bin the gain of band 255 (:366-369).
__global__ void __launch_bounds__(256)
    gain_subband_kernel(float* __restrict__ re, float* __restrict__ im, const mpeghb200_drc_sb_channel* __restrict__ chan,
                        const float* __restrict__ gains, const float* __restrict__ weights, int n_channels) {
  const int c = blockIdx.x, k = threadIdx.x;
  if (c >= n_channels) return;
  const mpeghb200_drc_sb_channel ch = chan[c];
  if (ch.first_seq < 0) return;  // channel without a DRC channel group
  float* r = re + size_t(c) * 1024;
  float* q = im + size_t(c) * 1024;
  const int nb = ch.band_count < 1 ? 1 : (ch.band_count > 4 ? 4 : ch.band_count);
  float w[4] = {0.0f, 0.0f, 0.0f, 0.0f}, w_last[4] = {0.0f, 0.0f, 0.0f, 0.0f};
  if (nb > 1) {
#pragma unroll
    for (int b = 0; b < 4; ++b)
      if (b < nb) {
        w[b] = __ldg(weights + size_t(ch.weight_row + b) * 256 + k);
        w_last[b] = __ldg(weights + size_t(ch.weight_row + b) * 256 + 255);
      }
  }
#pragma unroll
  for (int s = 0; s < 4; ++s) {
    float g, g_nyq;
    if (nb == 1) {
      g = g_nyq = __ldg(gains + size_t(ch.first_seq) * 4 + s);
    } else {
      g = 0.0f, g_nyq = 0.0f;
#pragma unroll
      for (int b = 0; b < 4; ++b)
        if (b < nb) {
          const float gl = __ldg(gains + size_t(ch.first_seq + b) * 4 + s);
          g = fadd(g, fmul(w[b], gl));
          g_nyq = fadd(g_nyq, fmul(w_last[b], gl));
        }
    }
    r[s * 256 + k] = fmul(r[s * 256 + k], g);
    q[s * 256 + k] = fmul(q[s * 256 + k], k == 0 ? g_nyq : g);
  }
}

}  // namespace
}  // namespace mpeghb200

using namespace mpeghb200;

extern "C" {

mpeghb200_status mpeghb200_bus_add_dev(mpeghb200_ctx* ctx, const float* d_a, const float* d_b, float* d_out,
                                       size_t n_floats, void* cuda_stream) {
  if (!ctx) return MPEGHB200_ERR_BAD_ARG;
  if (n_floats == 0) return MPEGHB200_OK;
  if (!d_a || !d_b || !d_out || (n_floats & 3))
    return fail(ctx, MPEGHB200_ERR_BAD_ARG, "bus_add_dev: null buffer, or a length that is not a multiple of 4");
  MB_CUDA(ctx, cudaSetDevice(ctx->device));
  const size_t n4 = n_floats / 4;
  const size_t want = (n4 + 255) / 256;
  const int grid = int(want < size_t(ctx->sm_count) * 8 ? want : size_t(ctx->sm_count) * 8);
  bus_add_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(cuda_stream)>>>(
      reinterpret_cast<const float4*>(d_a), reinterpret_cast<const float4*>(d_b), reinterpret_cast<float4*>(d_out), n4);
  ctx->launches.fetch_add(1);
  MB_CUDA(ctx, cudaGetLastError());
  return MPEGHB200_OK;
}

mpeghb200_status mpeghb200_interleave_dev(mpeghb200_ctx* ctx, const float* d_in, float* d_out, int n_ch, int n_samples,
                                          int to_planar, int n_streams, void* cuda_stream) {
  if (!ctx) return MPEGHB200_ERR_BAD_ARG;
  if (n_streams < 0 || n_ch < 1 || n_ch > 64 || n_samples < 1 || (n_streams > 0 && (!d_in || !d_out || d_in == d_out)))
    return fail(ctx, MPEGHB200_ERR_BAD_ARG, "interleave_dev: 1..64 channels, distinct non-null buffers");
  if (n_streams == 0) return MPEGHB200_OK;
  MB_CUDA(ctx, cudaSetDevice(ctx->device));
  const size_t smem = size_t(n_ch) * (kIlvTile + 1) * sizeof(float);
  interleave_kernel<<<dim3((n_samples + kIlvTile - 1) / kIlvTile, n_streams), 256, smem,
                      static_cast<cudaStream_t>(cuda_stream)>>>(d_in, d_out, n_ch, n_samples, to_planar != 0);
  ctx->launches.fetch_add(1);
  MB_CUDA(ctx, cudaGetLastError());
  return MPEGHB200_OK;
}

mpeghb200_status mpeghb200_channel_mask_dev(mpeghb200_ctx* ctx, float* d_audio, const uint8_t* d_on, int n_channels,
                                            void* cuda_stream) {
  if (!ctx) return MPEGHB200_ERR_BAD_ARG;
  if (n_channels < 0 || (n_channels > 0 && (!d_audio || !d_on)))
    return fail(ctx, MPEGHB200_ERR_BAD_ARG, "channel_mask_dev: null buffer or negative channel count");
  if (n_channels == 0) return MPEGHB200_OK;
  MB_CUDA(ctx, cudaSetDevice(ctx->device));
  channel_mask_kernel<<<n_channels, 256, 0, static_cast<cudaStream_t>(cuda_stream)>>>(
      reinterpret_cast<float4*>(d_audio), d_on, n_channels);
  ctx->launches.fetch_add(1);
  MB_CUDA(ctx, cudaGetLastError());
  return MPEGHB200_OK;
}

mpeghb200_status mpeghb200_gain_apply_dev(mpeghb200_ctx* ctx, float* d_audio, const float* d_gains,
                                          const int32_t* d_seq_of_ch, int n_channels, void* cuda_stream) {
  if (!ctx) return MPEGHB200_ERR_BAD_ARG;
  if (n_channels < 0 || (n_channels > 0 && (!d_audio || !d_gains || !d_seq_of_ch)))
    return fail(ctx, MPEGHB200_ERR_BAD_ARG, "gain_apply_dev: null buffer or negative channel count");
  if (n_channels == 0) return MPEGHB200_OK;
  MB_CUDA(ctx, cudaSetDevice(ctx->device));
  gain_apply_kernel<<<n_channels, 256, 0, static_cast<cudaStream_t>(cuda_stream)>>>(
      reinterpret_cast<float4*>(d_audio), reinterpret_cast<const float4*>(d_gains), d_seq_of_ch, n_channels);
  ctx->launches.fetch_add(1);
  MB_CUDA(ctx, cudaGetLastError());
  return MPEGHB200_OK;
}

mpeghb200_status mpeghb200_gain_subband_dev(mpeghb200_ctx* ctx, float* d_re, float* d_im,
                                            const mpeghb200_drc_sb_channel* d_chan, const float* d_gains,
                                            const float* d_weights, int n_channels, void* cuda_stream) {
  if (!ctx) return MPEGHB200_ERR_BAD_ARG;
  if (n_channels < 0 || (n_channels > 0 && (!d_re || !d_im || !d_chan || !d_gains)))
    return fail(ctx, MPEGHB200_ERR_BAD_ARG, "gain_subband_dev: null buffer or negative channel count");
  if (n_channels == 0) return MPEGHB200_OK;
  MB_CUDA(ctx, cudaSetDevice(ctx->device));
  gain_subband_kernel<<<n_channels, 256, 0, static_cast<cudaStream_t>(cuda_stream)>>>(d_re, d_im, d_chan, d_gains,
                                                                                      d_weights, n_channels);
  ctx->launches.fetch_add(1);
  MB_CUDA(ctx, cudaGetLastError());
  return MPEGHB200_OK;
}

}  // extern "C"
