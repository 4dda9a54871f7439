// hoa_render.cu -- HOA loudspeaker rendering: near-field compensation IIRs, render-matrix apply, clip.
//
// Stands in for impeghd_hoa_ren_block_process (decoder/impeghd_hoa_renderer.c:181) and
// impeghd_hoa_ren_nfc_filter_apply (decoder/impeghd_hoa_nfc_filtering.c:201) with its second/first order
// cells (:64, :99).  The render matrix (impeghd_hoa_ren_renderer_init, renderer.c:399: robust panning + SVD)
// and the NFC coefficients (impeghd_hoa_ren_nfc_set_filter_params, nfc_filtering.c:127) are designed once per
// configuration by the reference's host init and handed over.
//
// out[och][t] = sum_mn M[och][mn] * x[mn][t], accumulated left to right from 0 with separately rounded
// products and sums exactly like the reference (so no tensor cores and no FMA on this path: a TF32/BF16
// product is 2^-11/2^-8 relative, far outside the +-1 LSB@24 bit gate, and any re-association changes LSBs).
// NFC off: hoa_render_kernel, one CTA per (stream, 128-sample chunk), a thread per sample: its n_coef inputs sit in
//   registers, M comes from the kernel-parameter constant bank (warp-uniform operand, no load instruction),
//   2 FP instructions per MAC.
// NFC on:  hoa_render_nfc_kernel, one CTA per stream: filter warps (a lane per coefficient, cascade state in
//   registers over the whole frame) run one 64-sample stage ahead of the mixing warps through a ring of tiles.
// Algorithmic bytes per stream-frame: (n_coef + n_out) * 4096 (+ NFC state). For order 4 -> 22.2 that is
// 200 KB against 1.23 M unfused mul+add (+ ~20 k warp-instructions of cascades): the stage is FP32-issue bound
// (tools/exp/fp_probe: 3.1-3.4 FMUL/FADD warp-instructions per cycle per SM), not HBM bound.
#include <cstring>
#include <new>
#include <vector>

#include "common.h"

namespace mpeghb200 {
namespace {

constexpr int kFrame = 1024;
constexpr int kChunk = 128;
constexpr int kPitch = kChunk + 1;
constexpr int kMaxOut = 24;
constexpr int kMaxCoef = 49;

__device__ __forceinline__ void cp_async4(void* smem_dst, const void* src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(unsigned(__cvta_generic_to_shared(smem_dst))), "l"(src)
               : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

struct HoaParams {
  int n_out, n_coef, use_nfc, clip;
  float M[kMaxOut * kMaxCoef];
};

// NC = number of HOA coefficients (compile time: the sample's inputs live in registers);
// NO = number of outputs when known at compile time, 0 = run-time loop
// NFC off: every 128-sample chunk of every stream is independent, one CTA each; a thread owns one sample, reads its
// NC inputs straight from global memory (a warp = 128 consecutive bytes of a row) and writes n_out outputs.
template <int NC, int NO>
__global__ void __launch_bounds__(kChunk)
    hoa_render_kernel(const float* __restrict__ in, float* __restrict__ out, const __grid_constant__ HoaParams P) {
  pdl_prologue();
  const int t = threadIdx.x;
  const size_t s = blockIdx.x / (kFrame / kChunk);
  const int ck = blockIdx.x % (kFrame / kChunk);
  const float* x = in + s * size_t(NC) * kFrame + ck * kChunk + t;
  const int n_out = NO ? NO : P.n_out;
  float* yo = out + s * size_t(n_out) * kFrame + ck * kChunk + t;

  float xin[NC];
#pragma unroll
  for (int mn = 0; mn < NC; ++mn) xin[mn] = __ldcs(x + size_t(mn) * kFrame);
  if (NO) {
#pragma unroll
    for (int och = 0; och < (NO ? NO : 1); ++och) {
      float val = 0.0f;
#pragma unroll
      for (int mn = 0; mn < NC; ++mn) val = fadd(val, fmul(P.M[och * NC + mn], xin[mn]));
      if (P.clip) {
        val = (val < 32767.0f) ? val : 32767.0f;
        val = (val > -32768.0f) ? val : -32768.0f;
      }
      __stcs(yo + size_t(och) * kFrame, val);
    }
  } else {
    for (int och = 0; och < n_out; ++och) {
      float val = 0.0f;
      const float* Mr = P.M + och * NC;
#pragma unroll
      for (int mn = 0; mn < NC; ++mn) val = fadd(val, fmul(Mr[mn], xin[mn]));
      if (P.clip) {
        val = (val < 32767.0f) ? val : 32767.0f;
        val = (val > -32768.0f) ? val : -32768.0f;
      }
      __stcs(yo + size_t(och) * kFrame, val);
    }
  }
}

template <int NC, int NO>
cudaError_t launch(const float* in, float* out, float*, const float*, const HoaParams& P, int n_streams,
                   cudaStream_t stream) {
  return launch_pdl(hoa_render_kernel<NC, NO>, dim3(n_streams * (kFrame / kChunk)), dim3(kChunk), 0, stream, in, out, P);
}

// One filter lane's cascade over a chunk row, cell counts at compile time (the reference gives every coefficient
// of an order-N renderer the same N / 2 second-order cells and N % 2 first-order cell, so a warp is uniform in
// practice): nothing predicated, the cells of consecutive samples overlap in the pipeline.  The gain multiply is
// unconditional: x * 1.0f is x bit for bit, the reference skips it only when the gain is 1.
// A cascade is split over two filter warps a stage apart: the first runs the gain and cells [0, K1) (K0 = 0, GAIN),
// the second cells [K0, N2) and the first-order cell.
template <int K0, int K1, bool N1, bool GAIN, int LEN>
__device__ __forceinline__ void nfc_cascade(float* row, const float* __restrict__ cg, float (&st)[14]) {
  // only the coefficients this instantiation uses, re-read (L1 hits) per call instead of held across the mix
  float c[17];
  if (GAIN) c[0] = __ldg(cg);
#pragma unroll
  for (int i = 3 + 4 * K0; i < 3 + 4 * K1; ++i) c[i] = __ldg(cg + i);
  if (N1) {
    c[15] = __ldg(cg + 15);
    c[16] = __ldg(cg + 16);
  }
  for (int j0 = 0; j0 < LEN; j0 += 8) {
    float v[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) v[u] = row[j0 + u];  // off the dependent chain
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      float w = GAIN ? fmul(v[u], c[0]) : v[u];
#pragma unroll
      for (int k = K0; k < K1; ++k) {
        const float a1 = c[3 + 4 * k], a2 = c[4 + 4 * k], b1 = c[5 + 4 * k], b2 = c[6 + 4 * k];
        const float x0 = w;
        // ((x2*b2 - y1*a1) - y2*a2) + (x1*b1 + x0*b0), b0 = 1 (decoder/impeghd_hoa_nfc_filtering.c:72-77)
        w = fadd(fsub(fsub(fmul(st[4 * k + 1], b2), fmul(st[4 * k + 2], a1)), fmul(st[4 * k + 3], a2)),
                 fadd(fmul(st[4 * k], b1), x0));
        st[4 * k + 3] = st[4 * k + 2];
        st[4 * k + 2] = w;
        st[4 * k + 1] = st[4 * k];
        st[4 * k] = x0;
      }
      if (N1) {
        const float x0 = w;
        w = fadd(fsub(fmul(st[12], c[16]), fmul(st[13], c[15])), x0);  // :108-110
        st[13] = w;
        st[12] = x0;
      }
      v[u] = w;
    }
#pragma unroll
    for (int u = 0; u < 8; ++u) row[j0 + u] = v[u];
  }
}

// ---- NFC on: the cascades run ahead of the mix, on their own warps ------------------------------------------
// The CTA of a stream has 64 mixing threads + two groups of ceil(NC / 32) filter warps (a lane per coefficient) and
// a four-deep ring of 64-sample tiles.  In step k the mixing warps apply M to stage k (a thread per sample), filter group B runs the second half of the cascade over stage k + 1, group A the gain and the
// first half over stage k + 2, and the raw stage k + 3 is on its way from HBM by cp.async (no registers, no exposed
// latency).  One barrier per step.  A filter warp issues one sample per instruction, so with the whole cascade on
// one warp that warp was the critical path of every step; split, it is about as long as a mixing warp's share.
// Same expressions, same order: bit-identical.
constexpr int kSub = 64;
constexpr int kSubPitch = kSub + 1;  // odd: the filter lanes walk their rows without bank conflicts

template <int NC, int NO>
__global__ void __launch_bounds__(kSub + 64 * ((NC + 31) / 32), NC <= 25 ? 7 : 1)
    hoa_render_nfc_kernel(const float* __restrict__ in, float* __restrict__ out, float* __restrict__ nfc_state,
                          const float* __restrict__ nfc, const __grid_constant__ HoaParams P) {
  pdl_prologue();
  extern __shared__ float tile[];  // [4][NC][kSubPitch]
  constexpr int kBuf = NC * kSubPitch;
  constexpr int kSubs = kFrame / kSub;
  const int t = threadIdx.x;
  const size_t s = blockIdx.x;
  const float* x = in + s * size_t(NC) * kFrame;
  const int n_out = NO ? NO : P.n_out;
  float* y = out + s * size_t(n_out) * kFrame;
  constexpr int kFilt = 32 * ((NC + 31) / 32);  // lanes of one filter group
  const bool mixer = t < kSub;
  const bool grp_b = t >= kSub + kFilt;
  const int tn = t - kSub - (grp_b ? kFilt : 0);  // filter lane: coefficient index
  const bool filt = !mixer && tn < NC;

  float st[14];
  const float* cg = nfc + (filt ? tn : 0) * 17;
  if (filt) {
#pragma unroll
    for (int i = 0; i < 14; ++i) st[i] = nfc_state[(s * NC + tn) * 14 + i];
  }
  const int n2 = filt ? int(__ldg(cg + 1)) : 0;
  const bool n1 = filt && __ldg(cg + 2) != 0.0f;

  // cell counts of this warp's lanes: uniform (always, with reference-designed coefficients) or mixed; a CTA
  // with a mixed warp anywhere runs every cascade whole, predicated, in group A
  int key = -1;
  if (!mixer) {
    const int mine = filt ? n2 * 2 + int(n1) : -1;
    const int first = __shfl_sync(0xffffffffu, mine, 0);
    key = __all_sync(0xffffffffu, !filt || mine == first) ? first : -1;
  }
  const bool split = __syncthreads_and(mixer || key >= 0) != 0;
  // group A: gain + cells [0, (n2 + 1) / 2); group B: the remaining second-order cells + the first-order cell
  auto cascade = [&](float* row) {
    if (split && !grp_b) {
      switch (key >> 1) {
        case 0: nfc_cascade<0, 0, false, true, kSub>(row, cg, st); return;
        case 1: nfc_cascade<0, 1, false, true, kSub>(row, cg, st); return;
        case 2: nfc_cascade<0, 1, false, true, kSub>(row, cg, st); return;
        default: nfc_cascade<0, 2, false, true, kSub>(row, cg, st); return;
      }
    }
    if (split) {
      switch (key) {
        case 1: nfc_cascade<0, 0, true, false, kSub>(row, cg, st); return;   // n2 = 0 + first order
        case 3: nfc_cascade<1, 1, true, false, kSub>(row, cg, st); return;   // n2 = 1 + first order
        case 4: nfc_cascade<1, 2, false, false, kSub>(row, cg, st); return;  // n2 = 2
        case 5: nfc_cascade<1, 2, true, false, kSub>(row, cg, st); return;
        case 6: nfc_cascade<2, 3, false, false, kSub>(row, cg, st); return;  // n2 = 3
        case 7: nfc_cascade<2, 3, true, false, kSub>(row, cg, st); return;
        default: return;  // n2 <= 1 without a first-order cell: nothing left for group B
      }
    }
    if (grp_b) return;
    float c[17];  // mixed cell counts: predicated cells
#pragma unroll
    for (int i = 0; i < 17; ++i) c[i] = __ldg(cg + i);
    for (int j = 0; j < kSub; ++j) {
      float w = fmul(row[j], c[0]);
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        if (k < n2) {
          const float a1 = c[3 + 4 * k], a2 = c[4 + 4 * k], b1 = c[5 + 4 * k], b2 = c[6 + 4 * k];
          const float x0 = w;
          w = fadd(fsub(fsub(fmul(st[4 * k + 1], b2), fmul(st[4 * k + 2], a1)), fmul(st[4 * k + 3], a2)),
                   fadd(fmul(st[4 * k], b1), x0));
          st[4 * k + 3] = st[4 * k + 2];
          st[4 * k + 2] = w;
          st[4 * k + 1] = st[4 * k];
          st[4 * k] = x0;
        }
      }
      if (n1) {
        const float x0 = w;
        w = fadd(fsub(fmul(st[12], c[16]), fmul(st[13], c[15])), x0);
        st[13] = w;
        st[12] = x0;
      }
      row[j] = w;
    }
  };
  // raw stage k -> tile k % 4, by the mixing threads (a warp copies 32 consecutive samples of a row)
  auto fetch = [&](int k) {
    float* b = tile + (k & 3) * kBuf;
    const float* src = x + k * kSub;
#pragma unroll 5
    for (int mn = 0; mn < NC; ++mn) cp_async4(b + mn * kSubPitch + t, src + size_t(mn) * kFrame + t);
  };
  // one output row of the mix: och and the matrix row are compile-time or warp-uniform in every caller
  auto mix_row = [&](const float (&xin)[NC], const float* Mr, float* yo) {
    float val = 0.0f;
#pragma unroll
    for (int mn = 0; mn < NC; ++mn) val = fadd(val, fmul(Mr[mn], xin[mn]));
    if (P.clip) {
      val = (val < 32767.0f) ? val : 32767.0f;
      val = (val > -32768.0f) ? val : -32768.0f;
    }
    __stcs(yo, val);
  };

  auto row_of = [&](int k) { return tile + (k & 3) * kBuf + tn * kSubPitch; };
  if (mixer) {
    fetch(0);
    cp_async_wait_all();
  }
  __syncthreads();
  if (mixer) {
    fetch(1);
    cp_async_wait_all();
  } else if (filt && !grp_b) {
    cascade(row_of(0));
  }
  __syncthreads();
  if (mixer) {
    fetch(2);
    cp_async_wait_all();
  } else if (filt) {
    cascade(row_of(grp_b ? 0 : 1));
  }
  __syncthreads();

  for (int k = 0; k < kSubs; ++k) {
    // here: tile k & 3 holds the filtered stage k, the next one stage k + 1 after group A, the next the raw stage k + 2
    if (mixer) {
      if (k + 3 < kSubs) fetch(k + 3);  // into the tile the mix of stage k - 1 has finished with
      const float* cur = tile + (k & 3) * kBuf;
      float xin[NC];
#pragma unroll
      for (int mn = 0; mn < NC; ++mn) xin[mn] = cur[mn * kSubPitch + t];
      float* yo = y + k * kSub + t;
      if (NO) {
#pragma unroll
        for (int och = 0; och < (NO ? NO : 1); ++och) mix_row(xin, P.M + och * NC, yo + size_t(och) * kFrame);
      } else {
        for (int och = 0; och < n_out; ++och) mix_row(xin, P.M + och * NC, yo + size_t(och) * kFrame);
      }
      cp_async_wait_all();
    } else if (filt && k + (grp_b ? 1 : 2) < kSubs) {
      cascade(row_of(k + (grp_b ? 1 : 2)));
    }
    __syncthreads();
  }
  if (filt) {  // each group writes back the state of the cells it ran
    const int ka = split ? (n2 + 1) / 2 : 3;
#pragma unroll
    for (int i = 0; i < 12; ++i)
      if ((i / 4 < ka) != grp_b) nfc_state[(s * NC + tn) * 14 + i] = st[i];
    if (grp_b == split) {
      nfc_state[(s * NC + tn) * 14 + 12] = st[12];
      nfc_state[(s * NC + tn) * 14 + 13] = st[13];
    }
  }
}

template <int NC, int NO>
cudaError_t launch_nfc(const float* in, float* out, float* st, const float* nfc, const HoaParams& P, int n_streams,
                       cudaStream_t stream) {
  auto k = hoa_render_nfc_kernel<NC, NO>;
  const size_t smem = 4 * NC * kSubPitch * sizeof(float);
  if (smem > 48 * 1024) cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem));
  return launch_pdl(k, dim3(n_streams), dim3(kSub + 64 * ((NC + 31) / 32)), smem, stream, in, out, st, nfc, P);
}

// ---- tensor-core variant (opt-in, mpeghb200_hoa_renderer_set_tensor_mode) -------------------------------------
// BASELINE.json's north_star allows tensor cores for this one contraction, within max |err| <= 2^-20 of full
// scale instead of bit-identity.  The render matrix is applied as out^T[sample][och] = x^T[sample][mn] * M^T[mn][och]
// with mma.sync m16n8k8 TF32 and the 3xTF32 split: every operand v = hi + lo with hi = tf32(v), lo = tf32(v - hi),
// and hi*hi + (hi*lo + lo*hi) accumulated in fp32 (the cross terms in their own accumulators, so the big one sees
// only ceil(n_coef / 8) additions).  Relative error per product ~2^-21.  NFC stays the exact serial cascade.
// Samples are the M dimension (a 32-sample slice of the chunk per warp = two m16 tiles), the <= 24 outputs the N
// dimension (three n8 tiles, no padding for 22.2), the coefficients K padded to a multiple of 8 with zero rows.
constexpr int kPitchTc = 137;  // odd (NFC row walkers) and = 9 mod 32 (A fragments: k * 9 + m, two-way at worst)

__device__ __forceinline__ uint32_t to_tf32(float v) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(v));
  return r;
}
__device__ __forceinline__ void mma_tf32(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm(  // not volatile: a pure function of its operands, the scheduler may interleave independent chains
      "mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

template <int NC>
__global__ void __launch_bounds__(kChunk, NC <= 25 ? 4 : 2)
    hoa_render_tc_kernel(const float* __restrict__ in, float* __restrict__ out, float* __restrict__ nfc_state,
                         const float* __restrict__ nfc, const __grid_constant__ HoaParams P) {
  pdl_prologue();
  constexpr int KS = (NC + 7) / 8;
  extern __shared__ float tile[];  // [KS * 8][kPitchTc], rows >= NC stay zero
  const int t = threadIdx.x, lane = t & 31, warp = t >> 5, g = lane >> 2, tig = lane & 3;
  const size_t s = blockIdx.x;
  const float* x = in + s * size_t(NC) * kFrame;
  const int n_out = P.n_out;
  float* y = out + s * size_t(n_out) * kFrame;

  // B fragments of M^T, split once: b0 = (k = tig, n = g), b1 = (k = tig + 4, n = g)
  uint32_t bh[KS][3][2], bl[KS][3][2];
#pragma unroll
  for (int ks = 0; ks < KS; ++ks)
#pragma unroll
    for (int nt = 0; nt < 3; ++nt)
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int k = ks * 8 + tig + 4 * h, och = nt * 8 + g;
        const float m = (k < NC && och < n_out) ? P.M[och * NC + k] : 0.0f;
        bh[ks][nt][h] = to_tf32(m);
        bl[ks][nt][h] = to_tf32(fsub(m, __uint_as_float(bh[ks][nt][h])));
      }
  for (int r = NC; r < KS * 8; ++r) tile[r * kPitchTc + t] = 0.0f;

  float c[17], st[14];
  const bool nfc_thread = P.use_nfc && t < NC;
  if (nfc_thread) {
#pragma unroll
    for (int i = 0; i < 17; ++i) c[i] = __ldg(nfc + t * 17 + i);
#pragma unroll
    for (int i = 0; i < 14; ++i) st[i] = nfc_state[(s * NC + t) * 14 + i];
  }
  const int n2 = nfc_thread ? int(c[1]) : 0;
  const bool n1 = nfc_thread && c[2] != 0.0f;

  for (int ck = 0; ck < kFrame / kChunk; ++ck) {
#pragma unroll 5
    for (int mn = 0; mn < NC; ++mn) tile[mn * kPitchTc + t] = __ldcs(x + size_t(mn) * kFrame + ck * kChunk + t);
    __syncthreads();
    if (nfc_thread) {  // the exact cascade of hoa_render_kernel
      float* row = tile + t * kPitchTc;
      const bool scale = c[0] != 1.0f;
      for (int j = 0; j < kChunk; ++j) {
        float v = row[j];
        if (scale) v = fmul(v, c[0]);
#pragma unroll
        for (int k = 0; k < 3; ++k) {
          if (k < n2) {
            const float a1 = c[3 + 4 * k], a2 = c[4 + 4 * k], b1 = c[5 + 4 * k], b2 = c[6 + 4 * k];
            const float x0 = v;
            v = fadd(fsub(fsub(fmul(st[4 * k + 1], b2), fmul(st[4 * k + 2], a1)), fmul(st[4 * k + 3], a2)),
                     fadd(fmul(st[4 * k], b1), x0));
            st[4 * k + 3] = st[4 * k + 2];
            st[4 * k + 2] = v;
            st[4 * k + 1] = st[4 * k];
            st[4 * k] = x0;
          }
        }
        if (n1) {
          const float x0 = v;
          v = fadd(fsub(fmul(st[12], c[16]), fmul(st[13], c[15])), x0);
          st[13] = v;
          st[12] = x0;
        }
        row[j] = v;
      }
    }
    if (P.use_nfc) __syncthreads();
#pragma unroll 1
    for (int mt = 0; mt < 2; ++mt) {
      const int m0 = warp * 32 + mt * 16;
      float hi[3][4], lo[3][4], lo2[3][4];
#pragma unroll
      for (int nt = 0; nt < 3; ++nt)
#pragma unroll
        for (int j = 0; j < 4; ++j) hi[nt][j] = lo[nt][j] = lo2[nt][j] = 0.0f;
#pragma unroll
      for (int ks = 0; ks < KS; ++ks) {
        const float* a_lo_k = tile + (ks * 8 + tig) * kPitchTc + m0 + g;
        const float af[4] = {a_lo_k[0], a_lo_k[8], a_lo_k[4 * kPitchTc], a_lo_k[4 * kPitchTc + 8]};
        uint32_t ah[4], al[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          ah[j] = to_tf32(af[j]);
          al[j] = to_tf32(fsub(af[j], __uint_as_float(ah[j])));
        }
// nine independent accumulator chains per k step (an MMA waits for the one it depends on)
#pragma unroll
        for (int nt = 0; nt < 3; ++nt) mma_tf32(lo[nt], al, bh[ks][nt][0], bh[ks][nt][1]);
#pragma unroll
        for (int nt = 0; nt < 3; ++nt) mma_tf32(lo2[nt], ah, bl[ks][nt][0], bl[ks][nt][1]);
#pragma unroll
        for (int nt = 0; nt < 3; ++nt) mma_tf32(hi[nt], ah, bh[ks][nt][0], bh[ks][nt][1]);
      }
      // c0 = (m = g, n = 2 tig), c1 = (g, 2 tig + 1), c2 = (g + 8, 2 tig), c3 = (g + 8, 2 tig + 1)
#pragma unroll
      for (int nt = 0; nt < 3; ++nt)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int och = nt * 8 + 2 * tig + (j & 1);
          float val = fadd(hi[nt][j], fadd(lo[nt][j], lo2[nt][j]));
          if (P.clip) {
            val = (val < 32767.0f) ? val : 32767.0f;
            val = (val > -32768.0f) ? val : -32768.0f;
          }
          if (och < n_out) __stcs(y + size_t(och) * kFrame + ck * kChunk + m0 + g + (j >> 1) * 8, val);
        }
    }
    __syncthreads();
  }
  if (nfc_thread) {
#pragma unroll
    for (int i = 0; i < 14; ++i) nfc_state[(s * NC + t) * 14 + i] = st[i];
  }
}

// NFC off: no shared memory and no barriers either.  A CTA takes 256 samples of one stream, a warp 64 of them = four
// m16 tiles; the A fragments come straight from global memory (a fragment register of a warp = 4 rows x 8 consecutive
// samples = four full 32-byte sectors), the next tile's 16 loads are in flight under the current tile's 36 MMAs.
constexpr int kTcSpan = 256;
template <int NC>
__global__ void __launch_bounds__(kChunk)
    hoa_render_tc_direct_kernel(const float* __restrict__ in, float* __restrict__ out, const __grid_constant__ HoaParams P) {
  pdl_prologue();
  constexpr int KS = (NC + 7) / 8;
  const int t = threadIdx.x, lane = t & 31, warp = t >> 5, g = lane >> 2, tig = lane & 3;
  const size_t s = blockIdx.x / (kFrame / kTcSpan);
  const int base = (blockIdx.x % (kFrame / kTcSpan)) * kTcSpan + warp * 64;
  const float* x = in + s * size_t(NC) * kFrame + base + g;
  const int n_out = P.n_out;
  float* y = out + s * size_t(n_out) * kFrame + base + g;

  uint32_t bh[KS][3][2], bl[KS][3][2];
#pragma unroll
  for (int ks = 0; ks < KS; ++ks)
#pragma unroll
    for (int nt = 0; nt < 3; ++nt)
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int k = ks * 8 + tig + 4 * h, och = nt * 8 + g;
        const float m = (k < NC && och < n_out) ? P.M[och * NC + k] : 0.0f;
        bh[ks][nt][h] = to_tf32(m);
        bl[ks][nt][h] = to_tf32(fsub(m, __uint_as_float(bh[ks][nt][h])));
      }

  auto load_tile = [&](int mt, float (&af)[KS][4]) {
#pragma unroll
    for (int ks = 0; ks < KS; ++ks)
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int k = ks * 8 + tig + 4 * h;
        const bool ok = (ks * 8 + 4 * h + 3 < NC) || k < NC;  // compile-time true except in the padded last step
        af[ks][2 * h] = ok ? __ldcs(x + size_t(k) * kFrame + mt * 16) : 0.0f;
        af[ks][2 * h + 1] = ok ? __ldcs(x + size_t(k) * kFrame + mt * 16 + 8) : 0.0f;
      }
  };
  float cur[KS][4], nxt[KS][4];
  load_tile(0, cur);
#pragma unroll 1
  for (int mt = 0; mt < 4; ++mt) {
    if (mt + 1 < 4) load_tile(mt + 1, nxt);
    float hi[3][4], lo[3][4], lo2[3][4];
#pragma unroll
    for (int nt = 0; nt < 3; ++nt)
#pragma unroll
      for (int j = 0; j < 4; ++j) hi[nt][j] = lo[nt][j] = lo2[nt][j] = 0.0f;
#pragma unroll
    for (int ks = 0; ks < KS; ++ks) {
      uint32_t ah[4], al[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        ah[j] = to_tf32(cur[ks][j]);
        al[j] = to_tf32(fsub(cur[ks][j], __uint_as_float(ah[j])));
      }
// nine independent accumulator chains per k step (an MMA waits for the one it depends on)
#pragma unroll
      for (int nt = 0; nt < 3; ++nt) mma_tf32(lo[nt], al, bh[ks][nt][0], bh[ks][nt][1]);
#pragma unroll
      for (int nt = 0; nt < 3; ++nt) mma_tf32(lo2[nt], ah, bl[ks][nt][0], bl[ks][nt][1]);
#pragma unroll
      for (int nt = 0; nt < 3; ++nt) mma_tf32(hi[nt], ah, bh[ks][nt][0], bh[ks][nt][1]);
    }
#pragma unroll
    for (int nt = 0; nt < 3; ++nt)
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int och = nt * 8 + 2 * tig + (j & 1);
        float val = fadd(hi[nt][j], fadd(lo[nt][j], lo2[nt][j]));
        if (P.clip) {
          val = (val < 32767.0f) ? val : 32767.0f;
          val = (val > -32768.0f) ? val : -32768.0f;
        }
        if (och < n_out) __stcs(y + size_t(och) * kFrame + mt * 16 + (j >> 1) * 8, val);
      }
#pragma unroll
    for (int ks = 0; ks < KS; ++ks)
#pragma unroll
      for (int j = 0; j < 4; ++j) cur[ks][j] = nxt[ks][j];
  }
}

template <int NC>
cudaError_t launch_tc(const float* in, float* out, float* st, const float* nfc, const HoaParams& P, int n_streams,
                      cudaStream_t stream) {
  if (!P.use_nfc)
    return launch_pdl(hoa_render_tc_direct_kernel<NC>, dim3(n_streams * (kFrame / kTcSpan)), dim3(kChunk), 0, stream, in,
                      out, P);
  return launch_pdl(hoa_render_tc_kernel<NC>, dim3(n_streams), dim3(kChunk),
                    ((NC + 7) / 8) * 8 * kPitchTc * sizeof(float), stream, in, out, st, nfc, P);
}

}  // namespace
}  // namespace mpeghb200

using namespace mpeghb200;

namespace mpeghb200 {
// hoa_render_umma.cu: the contraction on tcgen05 / TMEM
void umma_prepare_b(const float* M, int n_out, int n_coef, float* terms);
size_t umma_b_bytes();
cudaError_t hoa_render_umma_launch(const float* d_in, float* d_out, const void* d_b, int n_coef, int n_out, int clip,
                                   int n_streams, int sm_count, cudaStream_t st);
}  // namespace mpeghb200

struct mpeghb200_hoa_renderer {
  mpeghb200_ctx* ctx = nullptr;
  HoaParams P{};
  float* d_nfc = nullptr;  // [n_coef][17]
  int tensor = 0;          // 0: bit-exact fp32; 1: legacy mma.sync 3xTF32; 2: tcgen05 + TMEM, exact 3-term TF32 split
  void* d_umma_b = nullptr;
};

extern "C" {

mpeghb200_status mpeghb200_hoa_renderer_create(mpeghb200_ctx* ctx, const float* matrix, int n_out, int n_coef,
                                               const float* nfc_coeffs, int clip, mpeghb200_hoa_renderer** out) {
  if (!ctx || !out || !matrix) return MPEGHB200_ERR_BAD_ARG;
  *out = nullptr;
  int order = 0;
  while ((order + 1) * (order + 1) < n_coef) ++order;
  if (n_out <= 0 || n_out > kMaxOut || (order + 1) * (order + 1) != n_coef || order < 1 || order > 6)
    return fail(ctx, MPEGHB200_ERR_BAD_ARG, "hoa_renderer_create: 1..24 outputs, (N+1)^2 coefficients, N = 1..6");
  MB_CUDA(ctx, cudaSetDevice(ctx->device));
  auto* r = new (std::nothrow) mpeghb200_hoa_renderer();
  if (!r) return MPEGHB200_ERR_NO_MEM;
  r->ctx = ctx;
  r->P.n_out = n_out;
  r->P.n_coef = n_coef;
  r->P.use_nfc = nfc_coeffs != nullptr;
  r->P.clip = clip;
  std::memcpy(r->P.M, matrix, sizeof(float) * n_out * n_coef);
  if (nfc_coeffs) {
    for (int mn = 0; mn < n_coef; ++mn) {
      const float* c = nfc_coeffs + mn * 17;
      if (c[1] < 0 || c[1] > 3 || (c[2] != 0.0f && c[2] != 1.0f)) {
        delete r;
        return fail(ctx, MPEGHB200_ERR_BAD_ARG, "hoa_renderer_create: NFC cell counts out of range");
      }
    }
    cudaError_t e = cudaMalloc(&r->d_nfc, sizeof(float) * n_coef * 17);
    if (e == cudaSuccess)
      e = cudaMemcpy(r->d_nfc, nfc_coeffs, sizeof(float) * n_coef * 17, cudaMemcpyHostToDevice);
    if (e != cudaSuccess) {
      cudaFree(r->d_nfc);
      delete r;
      return cuda_fail(ctx, e, "hoa_renderer_create upload");
    }
  }
  *out = r;
  return MPEGHB200_OK;
}

void mpeghb200_hoa_renderer_destroy(mpeghb200_hoa_renderer* r) {
  if (!r) return;
  cudaSetDevice(r->ctx->device);
  cudaFree(r->d_nfc);
  cudaFree(r->d_umma_b);
  delete r;
}

mpeghb200_status mpeghb200_hoa_renderer_set_tensor_mode(mpeghb200_hoa_renderer* r, int on) {
  if (!r) return MPEGHB200_ERR_BAD_ARG;
  if (on < 0 || on > 2) return fail(r->ctx, MPEGHB200_ERR_BAD_ARG, "hoa_renderer_set_tensor_mode: mode 0, 1 or 2");
  if (on == 2) {
    if (r->P.use_nfc || r->P.n_coef > 32 || r->P.n_out > 32)
      return fail(r->ctx, MPEGHB200_ERR_UNSUPPORTED,
                  "hoa_renderer_set_tensor_mode(2): the tcgen05 contraction covers NFC-off renderers of order <= 4 "
                  "(25 coefficients) to at most 32 loudspeakers");
    if (!r->d_umma_b) {
      MB_CUDA(r->ctx, cudaSetDevice(r->ctx->device));
      std::vector<float> terms(umma_b_bytes() / sizeof(float));
      umma_prepare_b(r->P.M, r->P.n_out, r->P.n_coef, terms.data());
      MB_CUDA(r->ctx, cudaMalloc(&r->d_umma_b, umma_b_bytes()));
      MB_CUDA(r->ctx, cudaMemcpy(r->d_umma_b, terms.data(), umma_b_bytes(), cudaMemcpyHostToDevice));
    }
  }
  r->tensor = on;
  return MPEGHB200_OK;
}

size_t mpeghb200_hoa_nfc_state_floats(const mpeghb200_hoa_renderer* r) { return r ? size_t(r->P.n_coef) * 14 : 0; }
int mpeghb200_hoa_renderer_num_out(const mpeghb200_hoa_renderer* r) { return r ? r->P.n_out : 0; }
int mpeghb200_hoa_renderer_num_coeffs(const mpeghb200_hoa_renderer* r) { return r ? r->P.n_coef : 0; }

mpeghb200_status mpeghb200_hoa_render_dev(mpeghb200_ctx* ctx, const mpeghb200_hoa_renderer* r, const float* d_in,
                                          float* d_out, float* d_nfc_state, int n_streams, void* cuda_stream) {
  if (!ctx || !r) return MPEGHB200_ERR_BAD_ARG;
  if (n_streams < 0 || (n_streams > 0 && (!d_in || !d_out || (r->P.use_nfc && !d_nfc_state))))
    return fail(ctx, MPEGHB200_ERR_BAD_ARG, "hoa_render_dev: null buffer or negative stream count");
  if (n_streams == 0) return MPEGHB200_OK;
  MB_CUDA(ctx, cudaSetDevice(ctx->device));
  cudaStream_t st = static_cast<cudaStream_t>(cuda_stream);
  const HoaParams& P = r->P;
  cudaError_t e;
  if (r->tensor == 2) {
    e = hoa_render_umma_launch(d_in, d_out, r->d_umma_b, P.n_coef, P.n_out, P.clip, n_streams, ctx->sm_count, st);
    ctx->launches.fetch_add(1);
    MB_CUDA(ctx, e);
    return MPEGHB200_OK;
  }
  if (r->tensor) {
    switch (P.n_coef) {
      case 4: e = launch_tc<4>(d_in, d_out, d_nfc_state, r->d_nfc, P, n_streams, st); break;
      case 9: e = launch_tc<9>(d_in, d_out, d_nfc_state, r->d_nfc, P, n_streams, st); break;
      case 16: e = launch_tc<16>(d_in, d_out, d_nfc_state, r->d_nfc, P, n_streams, st); break;
      case 25: e = launch_tc<25>(d_in, d_out, d_nfc_state, r->d_nfc, P, n_streams, st); break;
      case 36: e = launch_tc<36>(d_in, d_out, d_nfc_state, r->d_nfc, P, n_streams, st); break;
      default: e = launch_tc<49>(d_in, d_out, d_nfc_state, r->d_nfc, P, n_streams, st); break;
    }
    ctx->launches.fetch_add(1);
    MB_CUDA(ctx, e);
    return MPEGHB200_OK;
  }
  if (P.use_nfc) {
    switch (P.n_coef) {
      case 4: e = launch_nfc<4, 0>(d_in, d_out, d_nfc_state, r->d_nfc, P, n_streams, st); break;
      case 9: e = launch_nfc<9, 0>(d_in, d_out, d_nfc_state, r->d_nfc, P, n_streams, st); break;
      case 16: e = launch_nfc<16, 0>(d_in, d_out, d_nfc_state, r->d_nfc, P, n_streams, st); break;
      case 25:
        e = (P.n_out == 24) ? launch_nfc<25, 24>(d_in, d_out, d_nfc_state, r->d_nfc, P, n_streams, st)
                            : launch_nfc<25, 0>(d_in, d_out, d_nfc_state, r->d_nfc, P, n_streams, st);
        break;
      case 36: e = launch_nfc<36, 0>(d_in, d_out, d_nfc_state, r->d_nfc, P, n_streams, st); break;
      default: e = launch_nfc<49, 0>(d_in, d_out, d_nfc_state, r->d_nfc, P, n_streams, st); break;
    }
    ctx->launches.fetch_add(1);
    MB_CUDA(ctx, e);
    return MPEGHB200_OK;
  }
  switch (P.n_coef) {
    case 4: e = launch<4, 0>(d_in, d_out, d_nfc_state, r->d_nfc, P, n_streams, st); break;
    case 9: e = launch<9, 0>(d_in, d_out, d_nfc_state, r->d_nfc, P, n_streams, st); break;
    case 16: e = launch<16, 0>(d_in, d_out, d_nfc_state, r->d_nfc, P, n_streams, st); break;
    case 25:
      e = (P.n_out == 24) ? launch<25, 24>(d_in, d_out, d_nfc_state, r->d_nfc, P, n_streams, st)
                          : launch<25, 0>(d_in, d_out, d_nfc_state, r->d_nfc, P, n_streams, st);
      break;
    case 36: e = launch<36, 0>(d_in, d_out, d_nfc_state, r->d_nfc, P, n_streams, st); break;
    default: e = launch<49, 0>(d_in, d_out, d_nfc_state, r->d_nfc, P, n_streams, st); break;
  }
  ctx->launches.fetch_add(1);
  MB_CUDA(ctx, e);
  return MPEGHB200_OK;
}

}  // extern "C"
