// mct.cu -- multichannel coding tool: the inverse rotation / prediction boxes, applied to pairs of channel spectra.
//
// Stands in for impeghd_mc_process (decoder/impeghd_multichannel.c:899) with impeghd_apply_mc_rotation (:781) and
// impeghd_apply_mc_prediction (:817), and for the window loop around it (decoder/ia_core_coder_process.c:1043-1075).
// the step immediately ahead of the seam at process.c:1092.  Host keeps the bitstream
// side (impeghd_mc_parse, :386), stereo filling (:104, :698) and the band -> bin resolution, which
// mpeghb200_mct_resolve below does the reference's way; the same box applies unchanged to the four IGF source-tile
// copies of a channel (process.c:1059-1072) when the caller lists them as further boxes.
//
// The boxes of one stream-frame form a group and are applied in order (a channel may sit in several boxes); bins
// are independent, so a CTA owns a 256-bin slice of a group, a thread owns the same four bins of every box, and no
// ordering between threads is needed.  HBM traffic: 2 x 4 KB read + written per box (L1/L2 hits when a channel repeats).
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "common.h"

namespace mpeghb200 {
namespace {

__constant__ float c_mct_sin[65];
__constant__ float c_mct_cos[65];

// impeghd_mc_index_to_sin_alpha / _cos_alpha (decoder/impeghd_multichannel_rom.c:129,141): sin / cos of
// (idx - 32) pi / 64 as six-decimal literals.
void mct_tables(float* sn, float* cs) {
  for (int i = 0; i < 65; ++i) {
    char buf[32];
    const double a = (i - 32) * 3.14159265358979323846 / 64.0;
    double s = std::sin(a), c = std::cos(a);
    if (std::fabs(s) < 5e-7) s = 0.0;
    if (std::fabs(c) < 5e-7) c = 0.0;
    std::snprintf(buf, sizeof buf, "%.6f", s);
    sn[i] = std::strtof(buf, nullptr);
    std::snprintf(buf, sizeof buf, "%.6f", c);
    cs[i] = std::strtof(buf, nullptr);
  }
}

constexpr int kNone = -32768;

constexpr int kSlice = 256;    // bins per CTA: a group is spread over 1024 / kSlice CTAs (bins are independent)
constexpr int kThreads = kSlice / 4;

__global__ void __launch_bounds__(kThreads)
    mct_kernel(float* __restrict__ spec, const mpeghb200_mct_box* __restrict__ boxes,
               const int32_t* __restrict__ group_first, int n_groups) {
  __shared__ short cmap[kSlice];  // coefficient index of every bin of the current box, kNone = not in any segment
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int g = blockIdx.x, bin0 = blockIdx.y * kSlice;
  // The boxes of a group run one after the other, each a short dependent chain (record -> bin map -> rows): what
  // bounds the kernel is latency.  The records of up to kBatch boxes are staged in shared memory with one
  // cooperative copy and the row slices they touch are requested into L2 up front.
  constexpr int kBatch = 16;
  __shared__ __align__(16) mpeghb200_mct_box sbox[kBatch];
  static_assert(sizeof(mpeghb200_mct_box) % 16 == 0, "records are copied as int4");
  const int b_begin = group_first[g], b_end = group_first[g + 1];
  for (int b0 = b_begin; b0 < b_end; b0 += kBatch) {
  const int nb = (b_end - b0 < kBatch) ? b_end - b0 : kBatch;
  __syncthreads();
  {
    const int4* src = reinterpret_cast<const int4*>(boxes + b0);
    int4* dst = reinterpret_cast<int4*>(sbox);
    for (int i = tid; i < nb * int(sizeof(mpeghb200_mct_box) / 16); i += kThreads) dst[i] = __ldg(src + i);
  }
  __syncthreads();
  for (int i = tid; i < nb * 2 * (kSlice / 32); i += kThreads) {  // 128-byte lines of the slices of both rows
    const int bx = i / (2 * (kSlice / 32)), rem = i - bx * 2 * (kSlice / 32);
    const int unit = (rem & 1) ? sbox[bx].unit_r : sbox[bx].unit_l;
    asm volatile("prefetch.global.L2 [%0];" ::"l"(spec + size_t(unit) * 1024 + bin0 + 32 * (rem >> 1)));
  }
  for (int bi = 0; bi < nb; ++bi) {
    const mpeghb200_mct_box& B = sbox[bi];
    *reinterpret_cast<short4*>(cmap + 4 * tid) = make_short4(kNone, kNone, kNone, kNone);
    __syncthreads();
    const int n_seg = B.n_seg < MPEGHB200_MCT_MAX_SEG ? B.n_seg : MPEGHB200_MCT_MAX_SEG;
    for (int s = warp; s < n_seg; s += kThreads / 32) {
      int start = B.start[s], end = start + B.len[s];
      const short c = B.coef[s];
      if (start < 0 || end > 1024) continue;
      start = (start > bin0 ? start : bin0) - bin0;
      end = (end < bin0 + kSlice ? end : bin0 + kSlice) - bin0;
      for (int i = start + lane; i < end; i += 32) cmap[i] = c;
    }
    __syncthreads();
    const short4 c4 = *reinterpret_cast<const short4*>(cmap + 4 * tid);
    if (c4.x != kNone || c4.y != kNone || c4.z != kNone || c4.w != kNone) {
      float* pl = spec + size_t(B.unit_l) * 1024 + bin0 + 4 * tid;
      float* pr = spec + size_t(B.unit_r) * 1024 + bin0 + 4 * tid;
      const float4 l4 = *reinterpret_cast<const float4*>(pl), r4 = *reinterpret_cast<const float4*>(pr);
      float l[4] = {l4.x, l4.y, l4.z, l4.w}, r[4] = {r4.x, r4.y, r4.z, r4.w};
      const short cc[4] = {c4.x, c4.y, c4.z, c4.w};
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        if (cc[e] == kNone) continue;
        const float d = l[e], q = r[e];
        if (B.op == 2) {
          const int idx = cc[e] < 0 ? 0 : (cc[e] > 64 ? 64 : cc[e]);
          const float cs = c_mct_cos[idx], sn = c_mct_sin[idx];
          l[e] = fsub(fmul(d, cs), fmul(q, sn));  // :792
          r[e] = fadd(fmul(d, sn), fmul(q, cs));  // :794
        } else {
          const float alpha = fmul(float(int(cc[e])), 0.1f);  // :824
          const float ad = fmul(alpha, d);
          l[e] = fadd(fadd(d, ad), q);
          r[e] = (B.op == 0) ? fsub(fsub(d, ad), q) : fadd(fadd(-d, ad), q);
        }
      }
      *reinterpret_cast<float4*>(pl) = make_float4(l[0], l[1], l[2], l[3]);
      *reinterpret_cast<float4*>(pr) = make_float4(r[0], r[1], r[2], r[3]);
    }
    __syncthreads();  // cmap is rewritten by the next box
  }
  }
}

}  // namespace
}  // namespace mpeghb200

using namespace mpeghb200;

extern "C" {

void mpeghb200_mct_rom(float* sin_alpha, float* cos_alpha) {
  if (sin_alpha && cos_alpha) mct_tables(sin_alpha, cos_alpha);
}

// Host half of impeghd_mc_process (decoder/impeghd_multichannel.c:899-955) and of the window loop of
// decoder/ia_core_coder_process.c:1043-1075: one segment per masked band (two scale-factor bands wide) and window,
// or one per window for a full-band box.
mpeghb200_status mpeghb200_mct_resolve(int signal_type, int pred_dir, int fullband, int n_windows, int bins_per_window,
                                       int bands_per_window, int total_sfb, const int16_t* sfb_offset,
                                       const int32_t* coef_idx, const int32_t* mask, int unit_l, int unit_r,
                                       mpeghb200_mct_box* out) {
  if (!out || !coef_idx || n_windows < 1 || n_windows > 8 || bins_per_window < 1 ||
      n_windows * bins_per_window > 1024 || bands_per_window < 0 || unit_l < 0 || unit_r < 0 || unit_l == unit_r ||
      (!fullband && (!mask || !sfb_offset || total_sfb < 1)))
    return MPEGHB200_ERR_BAD_ARG;
  std::memset(out, 0, sizeof(*out));
  out->unit_l = unit_l, out->unit_r = unit_r;
  out->op = signal_type == 0 ? (pred_dir ? 1 : 0) : 2;
  int n = 0;
  for (int w = 0; w < n_windows; ++w) {
    const int32_t* c = coef_idx + w * bands_per_window;  // mct_band_offset += mct_bands_per_win (process.c:1073)
    const int32_t* m = mask ? mask + w * bands_per_window : nullptr;
    const int base = w * bins_per_window;
    auto push = [&](int start, int len, int coef) -> bool {
      if (len <= 0) return true;
      if (n >= MPEGHB200_MCT_MAX_SEG || start < 0 || start + len > bins_per_window || coef < -32767 || coef > 32767 ||
          (out->op == 2 && (coef < 0 || coef > 64)))
        return false;
      out->start[n] = int16_t(base + start), out->len[n] = int16_t(len), out->coef[n] = int16_t(coef);
      ++n;
      return true;
    };
    if (fullband) {
      if (!push(0, bins_per_window, c[0])) return MPEGHB200_ERR_BAD_ARG;
      continue;
    }
    int sfb = -1;
    for (int band = 0; band < bands_per_window; ++band) {
      if (m[band] == 1) {
        const int start = (sfb < 0) ? 0 : sfb_offset[sfb];
        const int stop = (sfb + 2 < total_sfb) ? sfb_offset[sfb + 2] : sfb_offset[sfb + 1];
        if (!push(start, stop - start, c[band])) return MPEGHB200_ERR_BAD_ARG;
      }
      sfb += 2;
      if (sfb >= total_sfb - 1) break;
    }
  }
  out->n_seg = n;
  return MPEGHB200_OK;
}

mpeghb200_status mpeghb200_mct_dev(mpeghb200_ctx* ctx, float* d_spec, const mpeghb200_mct_box* d_boxes,
                                   const int32_t* d_group_first, int n_groups, void* cuda_stream) {
  if (!ctx) return MPEGHB200_ERR_BAD_ARG;
  if (n_groups < 0 || (n_groups > 0 && (!d_spec || !d_boxes || !d_group_first)))
    return fail(ctx, MPEGHB200_ERR_BAD_ARG, "mct_dev: null buffer or negative group count");
  if (n_groups == 0) return MPEGHB200_OK;
  MB_CUDA(ctx, cudaSetDevice(ctx->device));
  static bool tables_done[64] = {};
  if (ctx->device >= 64 || !tables_done[ctx->device]) {
    float sn[65], cs[65];
    mct_tables(sn, cs);
    MB_CUDA(ctx, cudaMemcpyToSymbol(c_mct_sin, sn, sizeof sn));
    MB_CUDA(ctx, cudaMemcpyToSymbol(c_mct_cos, cs, sizeof cs));
    if (ctx->device < 64) tables_done[ctx->device] = true;
  }
  mct_kernel<<<dim3(n_groups, 1024 / kSlice), kThreads, 0, static_cast<cudaStream_t>(cuda_stream)>>>(d_spec, d_boxes, d_group_first, n_groups);
  ctx->launches.fetch_add(1);
  MB_CUDA(ctx, cudaGetLastError());
  return MPEGHB200_OK;
}

}  // extern "C"
