// stereo.cu -- joint-stereo tools of a channel-pair element on the dequantised spectra, in place.
//
// Stands in for impeghd_ms_stereo (decoder/ia_core_coder_ext_ch_ele.c:1427), ia_core_coder_cplx_pred_upmixing
// (:561) with its MDST estimate of the downmix (ia_core_coder_estimate_dmx_im :495, ia_core_coder_filter_and_add
// :412), the previous-frame downmix ia_core_coder_cplx_prev_mdct_dmx (:703) and the spectrum save
// ia_core_coder_usac_cplx_save_prev (:93).
//
// One CTA per pair, 256 threads, four consecutive bins per thread (one float4 of each spectrum).  Everything
// the reference walks band by band is a per-bin decision here: the band of a bin comes from a 1 KB bin -> sfb
// table built at create time, the per-(window, band) flags and prediction coefficients from the pair's side
// record staged in shared memory.  The only neighbour access is the 7-tap MDST filter over the downmix, which
// reads the downmix of the window (and of the previous window / previous frame) from shared memory with the
// reference's mirrored edges.  Every product and sum is a separately rounded binary32 operation in the
// reference's association order, including the "+ 0" of accumulating into a cleared buffer (it turns -0 into +0).
//
// Bytes per pair-frame: read 2 x 4096 + write 2 x 4096 spectra, 664 B side info, and for complex prediction with
// use_prev_frame the last window of the two saved spectra (<= 8192 B); the save kernel writes them (<= 8192 B).
#include <cstring>
#include <new>

#include "common.h"
#include "bands.h"

namespace mpeghb200 {
namespace {

constexpr int kThreads = 256;
constexpr int kFrame = 1024;
constexpr int kStateFloats = 2 * kFrame;  // saved left, saved right spectrum (coef_save_float)

// MDST filter coefficients of ISO/IEC 23003-3 (complex stereo prediction), as printed there with six decimals;
// [window class][previous shape][shape][tap] and [class >> 1][previous shape][tap].
// class 0 ONLY_LONG / EIGHT_SHORT, 1 LONG_START, 2 LONG_STOP, 3 STOP_START.  tests/test_rom_tables.py pins them
// against the reference ROM (decoder/ia_core_coder_rom.c:953-1022) via mpeghb200_stereo_mdst_rom.
const float kMdstCurr[4][2][2][7] = {
    {{{-0.000000f, -0.000000f, 0.500000f, 0.000000f, -0.500000f, 0.000000f, 0.000000f},
      {0.045748f, 0.057238f, 0.540714f, 0.000000f, -0.540714f, -0.057238f, -0.045748f}},
     {{0.045748f, -0.057238f, 0.540714f, 0.000000f, -0.540714f, 0.057238f, -0.045748f},
      {0.091497f, -0.000000f, 0.581427f, 0.000000f, -0.581427f, 0.000000f, -0.091497f}}},
    {{{0.102658f, 0.103791f, 0.567149f, 0.000000f, -0.567149f, -0.103791f, -0.102658f},
      {0.104763f, 0.105207f, 0.567861f, 0.000000f, -0.567861f, -0.105207f, -0.104763f}},
     {{0.148406f, 0.046553f, 0.607863f, 0.000000f, -0.607863f, -0.046553f, -0.148406f},
      {0.150512f, 0.047969f, 0.608574f, 0.000000f, -0.608574f, -0.047969f, -0.150512f}}},
    {{{0.102658f, -0.103791f, 0.567149f, 0.000000f, -0.567149f, 0.103791f, -0.102658f},
      {0.148406f, -0.046553f, 0.607863f, 0.000000f, -0.607863f, 0.046553f, -0.148406f}},
     {{0.104763f, -0.105207f, 0.567861f, 0.000000f, -0.567861f, 0.105207f, -0.104763f},
      {0.150512f, -0.047969f, 0.608574f, 0.000000f, -0.608574f, 0.047969f, -0.150512f}}},
    {{{0.205316f, -0.000000f, 0.634298f, 0.000000f, -0.634298f, 0.000000f, -0.205316f},
      {0.207421f, 0.001416f, 0.635010f, 0.000000f, -0.635010f, -0.001416f, -0.207421f}},
     {{0.207421f, -0.001416f, 0.635010f, 0.000000f, -0.635010f, 0.001416f, -0.207421f},
      {0.209526f, -0.000000f, 0.635722f, 0.000000f, -0.635722f, 0.000000f, -0.209526f}}}};
const float kMdstPrev[2][2][7] = {
    {{-0.000000f, 0.106103f, 0.250000f, 0.318310f, 0.250000f, 0.106103f, -0.000000f},
     {0.059509f, 0.123714f, 0.186579f, 0.213077f, 0.186579f, 0.123714f, 0.059509f}},
    {{0.038498f, 0.039212f, 0.039645f, 0.039790f, 0.039645f, 0.039212f, 0.038498f},
     {0.026142f, 0.026413f, 0.026577f, 0.026631f, 0.026577f, 0.026413f, 0.026142f}}};

struct MdstRom {
  float curr[4][2][2][7];
  float prev[2][2][7];
};

struct StereoTables {
  const uint8_t* bin2sfb_long;   // [1024]
  const uint8_t* bin2sfb_short;  // [128]
  MdstRom rom;                   // by value: lands in the constant bank, indexed uniformly per CTA
};

__device__ __forceinline__ int mdst_class(int window_sequence) {
  return window_sequence == 0 || window_sequence == 2 ? 0 : window_sequence == 1 ? 1 : window_sequence == 3 ? 2 : 3;
}

// tap sum of the 7-tap filter around bin b of a window of `len` bins, mirrored edges (-k -> k-1, len-1+k -> len-k)
__device__ __forceinline__ float tap_sum(const float* __restrict__ in, int len, const float (&f)[7], int b) {
  float x[7];
#pragma unroll
  for (int t = 0; t < 7; ++t) {
    int j = b + t - 3;
    j = j < 0 ? -j - 1 : (j >= len ? 2 * len - 1 - j : j);
    x[t] = in[j];
  }
  const float a = fadd(fmul(f[6], x[0]), fmul(f[5], x[1]));
  const float bb = fadd(fmul(f[4], x[2]), fmul(f[3], x[3]));
  const float c = fadd(fmul(f[2], x[4]), fmul(f[1], x[5]));
  const float d = fmul(f[0], x[6]);
  return fadd(fadd(a, bb), fadd(c, d));
}

__global__ void __launch_bounds__(kThreads)
    stereo_kernel(float* __restrict__ coef, const mpeghb200_stereo_pair* __restrict__ pairs,
                  const float* __restrict__ state, const StereoTables T) {
  pdl_prologue();
  __shared__ mpeghb200_stereo_pair sd;
  __shared__ __align__(16) float dmx_re[kFrame];
  __shared__ __align__(16) float dmx_prev[kFrame];
  const int tid = threadIdx.x;
  {
    const uint32_t* src = reinterpret_cast<const uint32_t*>(pairs + blockIdx.x);
    uint32_t* dst = reinterpret_cast<uint32_t*>(&sd);
    for (int i = tid; i < int(sizeof(sd) / 4); i += kThreads) dst[i] = src[i];
  }
  __syncthreads();
  const int tool = sd.tool;
  if (tool != 1 && tool != 3) return;
  const bool is_short = sd.is_short != 0;
  const int bins = is_short ? 128 : 1024;
  const int i0 = 4 * tid;                 // first of this thread's four bins
  const int w = is_short ? i0 >> 7 : 0;   // window
  const int b0 = i0 - w * bins;           // bin inside the window
  const int sbase = is_short ? 16 * w : 0;
  const uint8_t* b2s = is_short ? T.bin2sfb_short : T.bin2sfb_long;
  float4* pl = reinterpret_cast<float4*>(coef + size_t(sd.left) * kFrame) + tid;
  float4* pr = reinterpret_cast<float4*>(coef + size_t(sd.right) * kFrame) + tid;
  const float4 l4 = *pl, r4 = *pr;
  float l[4] = {l4.x, l4.y, l4.z, l4.w}, r[4] = {r4.x, r4.y, r4.z, r4.w};
  int sfb[4];
  {
    const uchar4 s4 = *reinterpret_cast<const uchar4*>(b2s + b0);
    sfb[0] = s4.x, sfb[1] = s4.y, sfb[2] = s4.z, sfb[3] = s4.w;
  }

  if (tool == 1) {  // M/S on the masked bands of [first_band, last_band)
    bool any = false;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      if (sfb[k] >= sd.first_band && sfb[k] < sd.last_band && sd.used[sbase + sfb[k]]) {
        const float a = l[k], c = r[k];
        l[k] = fadd(a, c), r[k] = fsub(a, c);
        any = true;
      }
    }
    if (any) {
      *pl = make_float4(l[0], l[1], l[2], l[3]);
      *pr = make_float4(r[0], r[1], r[2], r[3]);
    }
    return;
  }

  // ---- complex prediction ---------------------------------------------------------------------------------
  const float factor = sd.pred_dir ? -1.0f : 1.0f;
  const bool cplx = sd.complex_coef != 0;
  float im[4] = {0.f, 0.f, 0.f, 0.f};
  if (cplx) {
    // downmix spectrum: predicted bands carry it in the left channel, the others are L/R coded
#pragma unroll
    for (int k = 0; k < 4; ++k)
      dmx_re[i0 + k] = sd.used[sbase + sfb[k]] == 1 ? l[k] : fmul(0.5f, fadd(l[k], fmul(factor, r[k])));
    if (sd.use_prev_frame && i0 < bins) {  // previous frame's downmix from the saved spectra, or zeros
      float4 p = make_float4(0.f, 0.f, 0.f, 0.f);
      if (sd.prev_dmx) {
        const float* sv = state + size_t(sd.slot) * kStateFloats + (kFrame - bins);
        const float4 a = *reinterpret_cast<const float4*>(sv + i0);
        const float4 c = *reinterpret_cast<const float4*>(sv + kFrame + i0);
        p.x = fmul(0.5f, fadd(a.x, fmul(factor, c.x)));
        p.y = fmul(0.5f, fadd(a.y, fmul(factor, c.y)));
        p.z = fmul(0.5f, fadd(a.z, fmul(factor, c.z)));
        p.w = fmul(0.5f, fadd(a.w, fmul(factor, c.w)));
      }
      *reinterpret_cast<float4*>(dmx_prev + i0) = p;
    }
    __syncthreads();
    const int cls = mdst_class(sd.window_sequence);
    float fc[7], fp[7];
#pragma unroll
    for (int t = 0; t < 7; ++t) {
      fc[t] = T.rom.curr[cls][sd.window_shape_prev & 1][sd.window_shape & 1][t];
      fp[t] = T.rom.prev[cls >> 1][sd.window_shape_prev & 1][t];
    }
    const float* cur = dmx_re + w * bins;
    const float* prv = w > 0 ? dmx_re + (w - 1) * bins : (sd.use_prev_frame ? dmx_prev : nullptr);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int b = b0 + k;
      float v = fadd(0.0f, tap_sum(cur, bins, fc, b));  // even bins: + s * 1
      if (prv) {
        const float s = tap_sum(prv, bins, fp, b);
        v = (b & 1) ? fadd(v, s) : fadd(v, fmul(s, -1.0f));
      }
      im[k] = v;
    }
  }
  bool any = false;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int e = sbase + sfb[k];
    bool on = sd.used[e] != 0;
    if (!cplx) on = on && sfb[k] >= sd.first_band && sfb[k] < sd.last_band;
    if (on) {
      const float a_re = fmul(float(sd.alpha_re[e]), 0.1f);
      float mix = fsub(r[k], fmul(a_re, l[k]));
      if (cplx) mix = fsub(mix, fmul(fmul(float(sd.alpha_im[e]), 0.1f), im[k]));
      r[k] = fmul(factor, fsub(l[k], mix));
      l[k] = fadd(l[k], mix);
      any = true;
    }
  }
  if (any) {
    *pl = make_float4(l[0], l[1], l[2], l[3]);
    *pr = make_float4(r[0], r[1], r[2], r[3]);
  }
}

__global__ void __launch_bounds__(kThreads)
    stereo_save_kernel(const float* __restrict__ coef, const mpeghb200_stereo_pair* __restrict__ pairs,
                       float* __restrict__ state) {
  pdl_prologue();
  const mpeghb200_stereo_pair* sd = pairs + blockIdx.x;
  const int save = sd->save;
  if (save != 1 && save != 2) return;
  const int bins = sd->is_short ? 128 : 1024;
  const int i0 = (kFrame - bins) + 4 * threadIdx.x;
  if (i0 >= kFrame) return;
  float* sv = state + size_t(sd->slot) * kStateFloats;
  float4 a = make_float4(0.f, 0.f, 0.f, 0.f), c = a;
  if (save == 1) {
    a = *reinterpret_cast<const float4*>(coef + size_t(sd->left) * kFrame + i0);
    c = *reinterpret_cast<const float4*>(coef + size_t(sd->right) * kFrame + i0);
  }
  *reinterpret_cast<float4*>(sv + i0) = a;
  *reinterpret_cast<float4*>(sv + kFrame + i0) = c;
}

}  // namespace
}  // namespace mpeghb200

using namespace mpeghb200;

static_assert(sizeof(mpeghb200_stereo_pair) == 664, "mpeghb200_stereo_pair layout");

static StereoTables stereo_tables(const mpeghb200_bands* b) {
  StereoTables T{};
  T.bin2sfb_long = b->d_bin2sfb_long;
  T.bin2sfb_short = b->d_bin2sfb_short;
  std::memcpy(T.rom.curr, kMdstCurr, sizeof(kMdstCurr));
  std::memcpy(T.rom.prev, kMdstPrev, sizeof(kMdstPrev));
  return T;
}

extern "C" {

void mpeghb200_stereo_mdst_rom(float* curr112, float* prev28) {
  if (curr112) std::memcpy(curr112, kMdstCurr, sizeof(kMdstCurr));
  if (prev28) std::memcpy(prev28, kMdstPrev, sizeof(kMdstPrev));
}

size_t mpeghb200_stereo_state_floats(void) { return kStateFloats; }

mpeghb200_status mpeghb200_bands_create(mpeghb200_ctx* ctx, const int16_t* sfb_top_long, int n_long,
                                        const int16_t* sfb_top_short, int n_short, mpeghb200_bands** out) {
  if (!ctx || !out || !sfb_top_long || !sfb_top_short) return MPEGHB200_ERR_BAD_ARG;
  *out = nullptr;
  // the band tables must tile the window exactly (they do for every sampling rate of the reference ROM); side
  // records address at most 64 long bands and 16 bands per short window
  if (n_long < 1 || n_long > 64 || n_short < 1 || n_short > 16 || sfb_top_long[n_long - 1] != 1024 ||
      sfb_top_short[n_short - 1] != 128)
    return fail(ctx, MPEGHB200_ERR_BAD_ARG, "bands_create: band tables must end at 1024 / 128 bins");
  std::vector<uint8_t> h(1024 + 128 + 2 * 64 + 2 * 16);
  auto fill = [&](const int16_t* top, int n, uint8_t* dst, int len) {
    int b = 0;
    for (int s = 0; s < n; ++s) {
      if (top[s] <= b || top[s] > len) return false;
      for (; b < top[s]; ++b) dst[b] = uint8_t(s);
    }
    return true;
  };
  if (!fill(sfb_top_long, n_long, h.data(), 1024) || !fill(sfb_top_short, n_short, h.data() + 1024, 128))
    return fail(ctx, MPEGHB200_ERR_BAD_ARG, "bands_create: band edges must increase");
  std::memcpy(h.data() + 1152, sfb_top_long, sizeof(int16_t) * n_long);
  std::memcpy(h.data() + 1152 + 128, sfb_top_short, sizeof(int16_t) * n_short);
  auto* s = new (std::nothrow) mpeghb200_bands();
  if (!s) return MPEGHB200_ERR_NO_MEM;
  s->ctx = ctx;
  if (cudaSetDevice(ctx->device) != cudaSuccess || cudaMalloc(&s->d_blob, h.size()) != cudaSuccess ||
      cudaMemcpy(s->d_blob, h.data(), h.size(), cudaMemcpyHostToDevice) != cudaSuccess) {
    cudaError_t e = cudaGetLastError();
    if (s->d_blob) cudaFree(s->d_blob);
    delete s;
    return cuda_fail(ctx, e, "bands_create: table upload");
  }
  s->d_bin2sfb_long = s->d_blob;
  s->d_bin2sfb_short = s->d_blob + 1024;
  s->d_sfb_long = reinterpret_cast<const int16_t*>(s->d_blob + 1152);
  s->d_sfb_short = reinterpret_cast<const int16_t*>(s->d_blob + 1152 + 128);
  s->n_long = n_long, s->n_short = n_short;
  *out = s;
  return MPEGHB200_OK;
}

void mpeghb200_bands_destroy(mpeghb200_bands* s) {
  if (!s) return;
  if (s->d_blob) cudaFree(s->d_blob);
  delete s;
}

mpeghb200_status mpeghb200_stereo_dev(mpeghb200_ctx* ctx, const mpeghb200_bands* s, float* d_coef,
                                      const mpeghb200_stereo_pair* d_pairs, const float* d_state, int n_pairs,
                                      void* cuda_stream) {
  if (!ctx || !s) return MPEGHB200_ERR_BAD_ARG;
  if (n_pairs < 0 || (n_pairs > 0 && (!d_coef || !d_pairs || !d_state)))
    return fail(ctx, MPEGHB200_ERR_BAD_ARG, "stereo_dev: null buffer or negative pair count");
  if (n_pairs == 0) return MPEGHB200_OK;
  MB_CUDA(ctx, cudaSetDevice(ctx->device));
  MB_CUDA(ctx, launch_pdl(stereo_kernel, dim3(n_pairs), dim3(kThreads), 0, static_cast<cudaStream_t>(cuda_stream), d_coef,
                          d_pairs, d_state, stereo_tables(s)));
  ctx->launches.fetch_add(1);
  MB_CUDA(ctx, cudaGetLastError());
  return MPEGHB200_OK;
}

mpeghb200_status mpeghb200_stereo_save_dev(mpeghb200_ctx* ctx, const mpeghb200_bands* s, const float* d_coef,
                                           const mpeghb200_stereo_pair* d_pairs, float* d_state, int n_pairs,
                                           void* cuda_stream) {
  if (!ctx || !s) return MPEGHB200_ERR_BAD_ARG;
  if (n_pairs < 0 || (n_pairs > 0 && (!d_coef || !d_pairs || !d_state)))
    return fail(ctx, MPEGHB200_ERR_BAD_ARG, "stereo_save_dev: null buffer or negative pair count");
  if (n_pairs == 0) return MPEGHB200_OK;
  MB_CUDA(ctx, cudaSetDevice(ctx->device));
  MB_CUDA(ctx, launch_pdl(stereo_save_kernel, dim3(n_pairs), dim3(kThreads), 0, static_cast<cudaStream_t>(cuda_stream),
                          d_coef, d_pairs, d_state));
  ctx->launches.fetch_add(1);
  MB_CUDA(ctx, cudaGetLastError());
  return MPEGHB200_OK;
}

}  // extern "C"
