// binaural_generic.cu -- the binaural renderer for the transform lengths binaural.cu has no tuned kernel for:
// N = 2^k <= 1024 (direct filters of up to 512 taps: the reference transforms with impeghd_rad2_cplx_ifft / _fft
// directly, decoder/impeghd_ifft.c:1369-1372, impeghd_fft.c:1519-1522) and N = 16384 (4097 .. 8192 taps: sixteen
// 1024-point row transforms + inter-stage twiddle; the e^{+j} direction has NO column transform in the reference,
// decoder/impeghd_ifft.c:1318-1359, which is reproduced as it is -- oracle/oracle_fft.c pins it).  N = 4096 does not
// exist here either: the reference's result for it depends on stale scratch memory (see binaural.cu).
//
// One CTA per stream-block, every buffer in global memory (a per-stream workspace behind the carried state): these
// sizes are rare and the point is the reference's exact arithmetic, not speed.  The transform is the pass structure of
// impeghd_rad2_cplx_ifft / impeghd_rad2_cplx_fft (radix-4 DIT on base-4 digit-reversed input, twiddles from the
// 1024-point quarter-wave table with folded forms in four index regions, a final radix-2 pass when log2 n is odd),
// butterflies spread over the CTA's threads with a barrier between passes.  The block algorithm is that of
// binaural.cu / impeghd_binaural_struct_process_ld_ola (decoder/impeghd_binaural_renderer.c:543), same state layout.
#include <cstring>
#include <vector>

#include "common.h"
#include "rom_tables.h"

namespace mpeghb200 {
namespace bingen {

constexpr int kThreads = 256;

struct cpx {
  float r, i;
};

__device__ __forceinline__ unsigned rev_base4_16(unsigned v) {
  unsigned r = 0;
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    r = (r << 2) | (v & 3u);
    v >>= 2;
  }
  return r;
}

// e^{+j} family (INV): (xr*wc + xi*ws, -(xr*ws) + xi*wc); e^{-j}: (xr*wc - xi*ws, xr*ws + xi*wc)
template <bool INV>
__device__ __forceinline__ cpx rot_std(cpx x, float wc, float ws) {
  cpx o;
  if (INV) {
    o.r = fadd(fmul(x.r, wc), fmul(x.i, ws));
    o.i = fadd(-fmul(x.r, ws), fmul(x.i, wc));
  } else {
    o.r = fsub(fmul(x.r, wc), fmul(x.i, ws));
    o.i = fadd(fmul(x.r, ws), fmul(x.i, wc));
  }
  return o;
}
template <bool INV>
__device__ __forceinline__ cpx rot_fold(cpx x, float a, float b) {
  cpx o;
  if (INV) {
    o.r = fsub(fmul(x.r, b), fmul(x.i, a));
    o.i = fadd(fmul(x.r, a), fmul(x.i, b));
  } else {
    o.r = fadd(fmul(x.r, b), fmul(x.i, a));
    o.i = fadd(-fmul(x.r, a), fmul(x.i, b));
  }
  return o;
}

// a - 2b: one rounding of the product 2b (exact) and one of the difference; the first pass of the e^{-j} transform
// forms it as (a - b) - b instead (decoder/impeghd_fft.c:352-386)
template <bool TWO_STEP>
__device__ __forceinline__ float sub2x(float a, float b) {
  return TWO_STEP ? fsub(fsub(a, b), b) : fsub(a, fmul(b, 2.0f));
}
template <bool TWO_STEP>
__device__ __forceinline__ float add2x(float a, float b) {
  return TWO_STEP ? fadd(fadd(a, b), b) : fadd(a, fmul(b, 2.0f));
}

template <bool INV, bool TWO_STEP>
__device__ __forceinline__ void bfly4(cpx& p0, cpx& p1, cpx& p2, cpx& p3, bool region4) {
  cpx x0 = p0, x1 = p1, x2 = p2, x3 = p3;
  x0.r = fadd(x0.r, x2.r);
  x0.i = fadd(x0.i, x2.i);
  x2.r = sub2x<TWO_STEP>(x0.r, x2.r);
  x2.i = sub2x<TWO_STEP>(x0.i, x2.i);
  x1.r = fadd(x1.r, x3.r);
  if (TWO_STEP) {
    x1.i = fadd(x1.i, x3.i);
    x3.r = sub2x<true>(x1.r, x3.r);
    x3.i = sub2x<true>(x1.i, x3.i);
  } else {
    x3.r = sub2x<false>(x1.r, x3.r);
    if (!region4) {
      x1.i = fadd(x1.i, x3.i);
      x3.i = sub2x<false>(x1.i, x3.i);
    } else {  // x3.i arrives negated from the region-4 twiddle form
      x1.i = fsub(x1.i, x3.i);
      x3.i = add2x<false>(x1.i, x3.i);
    }
  }
  x0.r = fadd(x0.r, x1.r);
  x0.i = fadd(x0.i, x1.i);
  x1.r = sub2x<TWO_STEP>(x0.r, x1.r);
  x1.i = sub2x<TWO_STEP>(x0.i, x1.i);
  float dr, di;
  if (INV) {
    x2.r = fsub(x2.r, x3.i);
    x2.i = fadd(x2.i, x3.r);
    dr = add2x<false>(x2.r, x3.i);
    di = sub2x<false>(x2.i, x3.r);
  } else {
    x2.r = fadd(x2.r, x3.i);
    x2.i = fsub(x2.i, x3.r);
    dr = sub2x<TWO_STEP>(x2.r, x3.i);
    di = add2x<TWO_STEP>(x2.i, x3.r);
  }
  p0 = x0, p1 = x2, p2 = x1, p3.r = dr, p3.i = di;
}

// n = 2^lg <= 1024 points, in place on re / im (stride 1), y = n cpx of scratch.  Called by every thread of the CTA;
// ends with a barrier.  w = ia_fft_twiddle_table_float (cos at [j], sin at [j + 257]).
template <bool INV>
__device__ void cta_pow2(float* re, float* im, cpx* y, int n, const float* __restrict__ w) {
  const int tid = threadIdx.x, nt = blockDim.x;
  int lg = 0;
  while ((1 << lg) < n) lg++;
  const bool odd = lg & 1;
  const int npass = lg >> 1;
  const int q = n / 4;
  __syncthreads();
  for (int b = tid; b < q; b += nt) {
    unsigned h = rev_base4_16(unsigned(4 * b)) >> (15 - lg);
    if (odd) h = (h + 1) & ~1u;
    const int c = int(h >> 1);
    cpx x[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) x[i].r = re[c + i * q], x[i].i = im[c + i * q];
    bfly4<INV, !INV>(x[0], x[1], x[2], x[3], false);
#pragma unroll
    for (int i = 0; i < 4; ++i) y[4 * b + i] = x[i];
  }
  __syncthreads();
  int del = 4, ns = 64, groups = n >> 4;
  for (int p = 1; p < npass; ++p) {
    const int full = ns * del;  // 256
    const int lim1 = full / 4 + full / 8 - full / 16 + full / 32 - full / 64 + full / 128 - full / 256;
    const int total = del * groups;
    for (int e = tid; e < total; e += nt) {
      const int m = e / groups, g = e - m * groups;
      const int j = m * ns;
      cpx* d = y + g * 4 * del + m;
      cpx x0 = d[0], x1 = d[del], x2 = d[2 * del], x3 = d[3 * del];
      bool region4 = false;
      if (m > 0) {
        x1 = rot_std<INV>(x1, __ldg(w + j), __ldg(w + j + 257));
        if (j <= full / 2)
          x2 = rot_std<INV>(x2, __ldg(w + 2 * j), __ldg(w + 2 * j + 257));
        else
          x2 = rot_fold<INV>(x2, __ldg(w + 2 * j - 256), __ldg(w + 2 * j + 1));
        if (j <= lim1)
          x3 = rot_std<INV>(x3, __ldg(w + 3 * j), __ldg(w + 3 * j + 257));
        else if (j <= 2 * lim1)
          x3 = rot_fold<INV>(x3, __ldg(w + 3 * j - 256), __ldg(w + 3 * j + 1));
        else {
          const float a = __ldg(w + 3 * j - 512), bq = __ldg(w + 3 * j - 512 + 257);
          cpx t;
          if (INV) {
            t.r = fsub(-fmul(x3.r, a), fmul(x3.i, bq));
            t.i = fadd(-fmul(x3.r, bq), fmul(x3.i, a));
          } else {
            t.r = fadd(-fmul(x3.r, a), fmul(x3.i, bq));
            t.i = fadd(fmul(x3.r, bq), fmul(x3.i, a));
          }
          x3 = t;
          region4 = true;
        }
      }
      bfly4<INV, false>(x0, x1, x2, x3, region4);
      d[0] = x0, d[del] = x1, d[2 * del] = x2, d[3 * del] = x3;
    }
    __syncthreads();
    ns >>= 2, del <<= 2, groups >>= 2;
  }
  if (odd) {
    const int half = del / 2, step = ns << 1;
    for (int i = tid; i < del; i += nt) {
      const int k = (i < half) ? i : i - half;
      const float wc = __ldg(w + k * step), ws = __ldg(w + k * step + 257);
      const cpx x0 = y[i], x1 = y[i + del];
      cpx t;
      if (i < half)
        t = rot_std<INV>(x1, wc, ws);
      else if (INV) {
        t.r = fsub(fmul(x1.r, ws), fmul(x1.i, wc));
        t.i = fadd(fmul(x1.r, wc), fmul(x1.i, ws));
      } else {
        t.r = fadd(fmul(x1.r, ws), fmul(x1.i, wc));
        t.i = fadd(-fmul(x1.r, wc), fmul(x1.i, ws));
      }
      y[i + del].r = fsub(x0.r, t.r), y[i + del].i = fsub(x0.i, t.i);
      y[i].r = fadd(x0.r, t.r), y[i].i = fadd(x0.i, t.i);
    }
    __syncthreads();
  }
  for (int i = tid; i < n; i += nt) re[i] = y[i].r, im[i] = y[i].i;
  __syncthreads();
}

struct Tab {
  const float* w;        // [514]
  const float* mix_cos;  // [16384]
  const float* mix_sin;
};

// N-point transform of re / im in place.  ws: 2N floats of scratch (cpx y[N] for N <= 1024; for N = 16384 the sixteen
// rows, each transformed with ws2 = another 2 * 1024 floats).
template <bool INV>
__device__ void cta_transform(float* re, float* im, float* ws, float* ws2, int N, const Tab& T) {
  const int tid = threadIdx.x, nt = blockDim.x;
  if (N <= 1024) {
    cta_pow2<INV>(re, im, reinterpret_cast<cpx*>(ws), N, T.w);
    return;
  }
  // N = 16384 = 1024 x 16 (impeghd_cplx_ifft_8k / impeghd_cplx_fft_16k with dim2 = 16, fac = 1)
  float* rr = ws;           // [16][1024]
  float* ri = ws + 16384;   // [16][1024]
  __syncthreads();
  for (int e = tid; e < 16384; e += nt) {
    const int i = e >> 10, j = e & 1023;
    rr[e] = re[16 * j + i], ri[e] = im[16 * j + i];
  }
  __syncthreads();
  for (int i = 0; i < 16; ++i) cta_pow2<INV>(rr + 1024 * i, ri + 1024 * i, reinterpret_cast<cpx*>(ws2), 1024, T.w);
  // inter-stage twiddle, then (e^{-j} only) the 16-point column transform, transposed write-back
  for (int i = tid; i < 1024; i += nt) {
    float cr[16], ci[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      const float c = __ldg(T.mix_cos + i * 16 + j), s = __ldg(T.mix_sin + i * 16 + j);
      const float a = rr[1024 * j + i], b = ri[1024 * j + i];
      if (INV)
        cr[j] = fsub(fmul(a, c), fmul(b, s)), ci[j] = fadd(fmul(a, s), fmul(b, c));
      else
        cr[j] = fadd(fmul(a, c), fmul(b, s)), ci[j] = fsub(fmul(b, c), fmul(a, s));
    }
    if (!INV) {
      // impeghd_rad2_cplx_fft on 16 points: first radix-4 pass on digit-reversed input, one twiddled radix-4 pass
      cpx y[16];
#pragma unroll
      for (int b = 0; b < 4; ++b) {
        // rev_base4_16(4 b) >> 11 for lg = 4: c = b-th digit-reversed start; q = 4
        const int c = int((rev_base4_16(unsigned(4 * b)) >> 11) >> 1);
        cpx x[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) x[k].r = cr[c + 4 * k], x[k].i = ci[c + 4 * k];
        bfly4<false, true>(x[0], x[1], x[2], x[3], false);
#pragma unroll
        for (int k = 0; k < 4; ++k) y[4 * b + k] = x[k];
      }
      {
        const int del = 4, ns = 64, full = 256;
        const int lim1 = full / 4 + full / 8 - full / 16 + full / 32 - full / 64 + full / 128 - full / 256;
#pragma unroll
        for (int m = 0; m < 4; ++m) {
          const int j = m * ns;
          cpx x0 = y[m], x1 = y[del + m], x2 = y[2 * del + m], x3 = y[3 * del + m];
          bool region4 = false;
          if (m > 0) {
            x1 = rot_std<false>(x1, __ldg(T.w + j), __ldg(T.w + j + 257));
            if (j <= full / 2)
              x2 = rot_std<false>(x2, __ldg(T.w + 2 * j), __ldg(T.w + 2 * j + 257));
            else
              x2 = rot_fold<false>(x2, __ldg(T.w + 2 * j - 256), __ldg(T.w + 2 * j + 1));
            if (j <= lim1)
              x3 = rot_std<false>(x3, __ldg(T.w + 3 * j), __ldg(T.w + 3 * j + 257));
            else if (j <= 2 * lim1)
              x3 = rot_fold<false>(x3, __ldg(T.w + 3 * j - 256), __ldg(T.w + 3 * j + 1));
            else {
              const float a = __ldg(T.w + 3 * j - 512), bq = __ldg(T.w + 3 * j - 512 + 257);
              cpx t;
              t.r = fadd(-fmul(x3.r, a), fmul(x3.i, bq));
              t.i = fadd(fmul(x3.r, bq), fmul(x3.i, a));
              x3 = t;
              region4 = true;
            }
          }
          bfly4<false, false>(x0, x1, x2, x3, region4);
          y[m] = x0, y[del + m] = x1, y[2 * del + m] = x2, y[3 * del + m] = x3;
        }
      }
#pragma unroll
      for (int j = 0; j < 16; ++j) cr[j] = y[j].r, ci[j] = y[j].i;
    }
#pragma unroll
    for (int j = 0; j < 16; ++j) re[1024 * j + i] = cr[j], im[1024 * j + i] = ci[j];
  }
  __syncthreads();
}

// Filter preparation (impeghd_binaural_struct_set, decoder/impeghd_binaural_renderer.c:244): taps zero-padded to N,
// e^{+j} transform, packed: [Re X0, Re X(N/2), Re X1, -Im X1, ...].  One CTA per filter; work = 4N + 2048 floats each.
__global__ void __launch_bounds__(kThreads)
    spectrum_kernel(const float* __restrict__ taps, int len, float* __restrict__ spec, float* __restrict__ work, int N,
                    Tab T) {
  const size_t f = blockIdx.x;
  float* re = work + f * (4 * size_t(N) + 2048);
  float* im = re + N;
  float* ws = im + N;
  float* ws2 = ws + 2 * N;
  const float* x = taps + f * len;
  for (int i = threadIdx.x; i < N; i += blockDim.x) re[i] = i < len ? x[i] : 0.0f, im[i] = 0.0f;
  cta_transform<true>(re, im, ws, ws2, N, T);
  float* out = spec + f * N;
  const int B = N / 2;
  for (int j = threadIdx.x; j < B; j += blockDim.x) {
    out[2 * j] = re[j];
    out[2 * j + 1] = j == 0 ? re[B] : fmul(-1.0f, im[j]);
  }
}

struct Params {
  const float* in;
  long long in_stream_stride, in_chan_stride, in_chunk_stride;
  float in_scale, out_scale;
  float* out;
  float* state;
  long long state_floats;
  const int32_t* set;
  const float* h_direct;
  const float* h_diffuse;
  const int32_t* direct_len;
  const int32_t* diffuse_len;
  const float* weight;
  int n_in, n_blocks, N;
  Tab T;
};

// d[k] (k = 2j, 2j+1) += a * b in the packed format (impeghd_binaural_mul_add_perm, :498)
__device__ __forceinline__ void mac_perm(float* d, const float* __restrict__ a, float br, float bi, int j) {
  if (j == 0) {
    d[0] = fadd(d[0], fmul(__ldg(a), br));
    d[1] = fadd(d[1], fmul(__ldg(a + 1), bi));
  } else {
    const float ar = __ldg(a + 2 * j), ai = __ldg(a + 2 * j + 1);
    float v = fadd(d[2 * j], fmul(ar, br));
    d[2 * j] = fsub(v, fmul(ai, bi));
    v = fadd(d[2 * j + 1], fmul(ar, bi));
    d[2 * j + 1] = fadd(v, fmul(ai, br));
  }
}

__global__ void __launch_bounds__(kThreads) block_kernel(Params P) {
  const size_t s = blockIdx.x;
  const int tid = threadIdx.x, nt = blockDim.x;
  const int N = P.N, B = N / 2;
  float* state = P.state + s * P.state_floats;
  float* tail = state;                    // [2][B]
  float* prev_in = state + 2 * size_t(B);  // [n_blocks + 1][N]
  int* widx_p = reinterpret_cast<int*>(prev_in + size_t(P.n_blocks + 1) * N);
  float* work = reinterpret_cast<float*>(widx_p) + 4;  // re [N], im [N], acc [2][N], ws [2N], ws2 [2048]
  float* re = work;
  float* im = re + N;
  float* acc = im + N;
  float* ws = acc + 2 * size_t(N);
  float* ws2 = ws + 2 * size_t(N);
  const int widx = *widx_p;
  const int set = P.set ? P.set[s] : 0;
  const float* hd = P.h_direct + size_t(set) * 2 * P.n_in * N;
  const float* hf = P.h_diffuse + size_t(set) * P.n_blocks * 2 * N;
  const int32_t* dlen = P.direct_len + size_t(set) * 2 * P.n_in;
  const int32_t* flen = P.diffuse_len + size_t(set) * P.n_blocks * 2;
  const float* wgt = P.weight + size_t(set) * P.n_in;
  float* dmx = prev_in + size_t(widx) * N;
  for (int i = tid; i < 2 * N; i += nt) acc[i] = 0.0f;
  for (int inp = 0; inp < P.n_in; ++inp) {
    const float* x = P.in + s * P.in_stream_stride + size_t(inp) * P.in_chan_stride;
    for (int n = tid; n < N; n += nt) {
      re[n] = n < B ? fmul(x[(n >> 10) * P.in_chunk_stride + (n & 1023)], P.in_scale) : 0.0f;
      im[n] = 0.0f;
    }
    cta_transform<true>(re, im, ws, ws2, N, P.T);
    const float w = wgt[inp];
    const int l0 = dlen[inp], l1 = dlen[P.n_in + inp];
    for (int j = tid; j < B; j += nt) {
      const float xr = re[j], xi = j == 0 ? re[B] : fmul(-1.0f, im[j]);
      if (inp == 0)
        dmx[2 * j] = fmul(xr, w), dmx[2 * j + 1] = fmul(xi, w);
      else
        dmx[2 * j] = fadd(dmx[2 * j], fmul(xr, w)), dmx[2 * j + 1] = fadd(dmx[2 * j + 1], fmul(xi, w));
      // bins 0 and N/2 are multiplied whatever the length says (impeghd_binaural_mul_add_perm :505-508)
      if (j == 0 || j < l0) mac_perm(acc, hd + size_t(inp) * N, xr, xi, j);
      if (j == 0 || j < l1) mac_perm(acc + N, hd + size_t(P.n_in + inp) * N, xr, xi, j);
    }
    __syncthreads();
  }
  // Im X[0], Im X[N/2] of the last input's transform leak into the first ear's inverse transform (shared scratch)
  float im0 = im[0], imn = im[B];
  __syncthreads();
  {  // diffuse part: the ring, newest block first
    int k = (widx + P.n_blocks) % (P.n_blocks + 1);
    for (int jb = 0; jb < P.n_blocks; ++jb) {
      const float* pv = prev_in + size_t(k) * N;
      for (int o = 0; o < 2; ++o) {
        const int l = flen[jb * 2 + o];
        const float* h = hf + (size_t(jb) * 2 + o) * N;
        for (int j = tid; j < B && (j == 0 || j < l); j += nt) mac_perm(acc + size_t(o) * N, h, pv[2 * j], pv[2 * j + 1], j);
      }
      __syncthreads();
      k = (k + P.n_blocks) % (P.n_blocks + 1);
    }
  }
  const float scale = __fdiv_rn(1.0f, float(N));
  for (int o = 0; o < 2; ++o) {
    const float* a = acc + size_t(o) * N;
    for (int j = tid; j < B; j += nt) {
      if (j == 0) {
        re[0] = a[0], re[B] = a[1], im[0] = im0, im[B] = imn;
      } else {
        const float yi = fmul(-1.0f, a[2 * j + 1]);
        im[j] = yi, im[N - j] = fmul(-1.0f, yi);
        re[j] = a[2 * j], re[N - j] = a[2 * j];
      }
    }
    cta_transform<false>(re, im, ws, ws2, N, P.T);
    float* y = P.out + (s * 2 + o) * size_t(B);
    for (int j = tid; j < B; j += nt) {
      const float v = fadd(fmul(re[j], scale), tail[o * B + j]);
      y[j] = fmul(v, P.out_scale);
      tail[o * B + j] = fmul(re[B + j], scale);
    }
    im0 = im[0], imn = im[B];  // the second ear sees the first ear's leftovers
    __syncthreads();
  }
  if (tid == 0) *widx_p = (widx + 1) % (P.n_blocks + 1);
}

}  // namespace bingen

// ---- host side, called from binaural.cu ----------------------------------------------------------------------------
size_t binaural_generic_work_floats(int N) { return 6 * size_t(N) + 2048; }

cudaError_t binaural_generic_tables(float** d_w, float** d_mc, float** d_ms) {
  const std::vector<float>& w = rom_table(ROM_FFT_TWIDDLE);
  const std::vector<float>& mc = rom_table(ROM_MIXED_RAD_COS);
  const std::vector<float>& ms = rom_table(ROM_MIXED_RAD_SIN);
  cudaError_t e = cudaMalloc(d_w, w.size() * sizeof(float));
  if (e == cudaSuccess) e = cudaMalloc(d_mc, mc.size() * sizeof(float));
  if (e == cudaSuccess) e = cudaMalloc(d_ms, ms.size() * sizeof(float));
  if (e == cudaSuccess) e = cudaMemcpy(*d_w, w.data(), w.size() * sizeof(float), cudaMemcpyHostToDevice);
  if (e == cudaSuccess) e = cudaMemcpy(*d_mc, mc.data(), mc.size() * sizeof(float), cudaMemcpyHostToDevice);
  if (e == cudaSuccess) e = cudaMemcpy(*d_ms, ms.data(), ms.size() * sizeof(float), cudaMemcpyHostToDevice);
  return e;
}

cudaError_t binaural_generic_spectra(const float* d_taps, int len, int n_filters, float* d_spec, int N, const float* d_w,
                                     const float* d_mc, const float* d_ms) {
  float* work = nullptr;
  cudaError_t e = cudaMalloc(&work, size_t(n_filters) * (4 * size_t(N) + 2048) * sizeof(float));
  if (e != cudaSuccess) return e;
  bingen::Tab T{d_w, d_mc, d_ms};
  bingen::spectrum_kernel<<<n_filters, bingen::kThreads>>>(d_taps, len, d_spec, work, N, T);
  e = cudaGetLastError();
  if (e == cudaSuccess) e = cudaDeviceSynchronize();
  cudaFree(work);
  return e;
}

cudaError_t binaural_generic_block(const float* d_in, long long in_stream_stride, long long in_chan_stride,
                                   long long in_chunk_stride, float in_scale, float out_scale, float* d_out, float* d_state,
                                   long long state_floats, const int32_t* d_set, const float* h_direct,
                                   const float* h_diffuse, const int32_t* dlen, const int32_t* flen, const float* weight,
                                   int n_in, int n_blocks, int N, const float* d_w, const float* d_mc, const float* d_ms,
                                   int n_streams, cudaStream_t st) {
  bingen::Params P{};
  P.in = d_in, P.in_stream_stride = in_stream_stride, P.in_chan_stride = in_chan_stride, P.in_chunk_stride = in_chunk_stride;
  P.in_scale = in_scale, P.out_scale = out_scale, P.out = d_out, P.state = d_state, P.state_floats = state_floats;
  P.set = d_set, P.h_direct = h_direct, P.h_diffuse = h_diffuse, P.direct_len = dlen, P.diffuse_len = flen;
  P.weight = weight, P.n_in = n_in, P.n_blocks = n_blocks, P.N = N;
  P.T = bingen::Tab{d_w, d_mc, d_ms};
  bingen::block_kernel<<<n_streams, bingen::kThreads, 0, st>>>(P);
  return cudaGetLastError();
}

}  // namespace mpeghb200
