// fft_exact.cuh -- CTA-level 1024 x DIM2 point complex transforms with the reference's exact rounding sequence.
//
// Mirrors impeghd_rad2_cplx_ifft / impeghd_rad2_cplx_fft for n = 1024 (decoder/impeghd_ifft.c:297,
// decoder/impeghd_fft.c:297) and the 1024 x {2, 8} decompositions built on them, impeghd_cplx_ifft_8k
// (impeghd_ifft.c:1264) and impeghd_cplx_fft_16k (impeghd_fft.c:1415): DIM2 row transforms over the
// decimated input, inter-stage twiddle, DIM2-point column transform, transposed output.  INV = true is the
// e^{+j} ("ifft") family, INV = false the e^{-j} ("fft") family.  The two differ by conjugate-mirrored
// product forms and by the first pass of the e^{-j} transform forming a - 2b as (a - b) - b.
//
// Work layout: DIM2 rows of 1024 float2 in shared memory, element p of a row at phys(p) = p + p/16, rows
// kRowStride float2 apart (both paddings make every access pattern below bank-conflict free).  A CTA has
// 64 * DIM2 threads; each thread owns 16 points per stage:
//   stage A  passes 0 + 1 of the radix-4 DIT in registers (reads digit-reversed positions r3(g) + 64 t)
//   stage B  passes 2 + 3 (elements base + 16 t)
//   stage C  pass 4, four butterflies per thread
//   columns  16 / DIM2 columns per thread: twiddle + DIM2-point transform in registers
#pragma once
#include "common.h"

namespace mpeghb200 {
namespace fftx {

constexpr int kRowStride = 1090;  // float2 units; 1024 + 63 padding, == 2 mod 16

__device__ __forceinline__ int phys(int p) { return p + (p >> 4); }

// a - 2b and a + 2b in the two forms the reference uses
template <bool TWO_STEP>
__device__ __forceinline__ float m2(float a, float b) {
  return TWO_STEP ? fsub(fsub(a, b), b) : sub2(a, b);
}
template <bool TWO_STEP>
__device__ __forceinline__ float p2(float a, float b) {
  return TWO_STEP ? fadd(fadd(a, b), b) : add2(a, b);
}

// x *= (c -/+ j s): INV (r*c + i*s, -(r*s) + i*c), else (r*c - i*s, r*s + i*c)
template <bool INV>
__device__ __forceinline__ void rot(float2& x, float c, float s) {
  float nr, ni;
  if (INV) {
    nr = fadd(fmul(x.x, c), fmul(x.y, s));
    ni = fadd(-fmul(x.x, s), fmul(x.y, c));
  } else {
    nr = fsub(fmul(x.x, c), fmul(x.y, s));
    ni = fadd(fmul(x.x, s), fmul(x.y, c));
  }
  x = make_float2(nr, ni);
}

// Radix-4 butterfly, "a += b; b = a - 2b" form; outputs in natural order a, b, c, d.
template <bool INV, bool TWO_STEP>
__device__ __forceinline__ void bfly(float2& a, float2& b, float2& c, float2& d) {
  float x0r = a.x, x0i = a.y, x1r = b.x, x1i = b.y, x2r = c.x, x2i = c.y, x3r = d.x, x3i = d.y;
  x0r = fadd(x0r, x2r);
  x0i = fadd(x0i, x2i);
  x2r = m2<TWO_STEP>(x0r, x2r);
  x2i = m2<TWO_STEP>(x0i, x2i);
  x1r = fadd(x1r, x3r);
  x1i = fadd(x1i, x3i);
  x3r = m2<TWO_STEP>(x1r, x3r);
  x3i = m2<TWO_STEP>(x1i, x3i);
  x0r = fadd(x0r, x1r);
  x0i = fadd(x0i, x1i);
  x1r = m2<TWO_STEP>(x0r, x1r);
  x1i = m2<TWO_STEP>(x0i, x1i);
  float dr, di;
  if (INV) {
    x2r = fsub(x2r, x3i);
    x2i = fadd(x2i, x3r);
    dr = p2<TWO_STEP>(x2r, x3i);
    di = m2<TWO_STEP>(x2i, x3r);
  } else {
    x2r = fadd(x2r, x3i);
    x2i = fsub(x2i, x3r);
    dr = m2<TWO_STEP>(x2r, x3i);
    di = p2<TWO_STEP>(x2i, x3r);
  }
  a = make_float2(x0r, x0i);
  b = make_float2(x2r, x2i);
  c = make_float2(x1r, x1i);
  d = make_float2(dr, di);
}

// 8-point column transforms (decoder/impeghd_ifft.c:1118 with its swapped bins 2 and 6 kept,
// decoder/impeghd_fft.c:1304)
template <bool INV>
__device__ __forceinline__ void col8(float (&xr)[8], float (&xi)[8]) {
  const float s = 0.7071067811f;
  float ar = fadd(xr[0], xr[4]), ai = fadd(xi[0], xi[4]), br = fsub(xr[0], xr[4]), bi = fsub(xi[0], xi[4]);
  float cr = fadd(xr[2], xr[6]), ci = fadd(xi[2], xi[6]), dr = fsub(xr[2], xr[6]), di = fsub(xi[2], xi[6]);
  float n00 = fadd(ar, cr), n01 = fadd(ai, ci), n20 = fsub(ar, cr), n21 = fsub(ai, ci);
  float n10, n11, n30, n31;
  if (INV) {
    n10 = fsub(br, di), n11 = fadd(bi, dr), n30 = fadd(br, di), n31 = fsub(bi, dr);
  } else {
    n10 = fadd(br, di), n11 = fsub(bi, dr), n30 = fsub(br, di), n31 = fadd(bi, dr);
  }
  ar = fadd(xr[1], xr[5]), ai = fadd(xi[1], xi[5]), br = fsub(xr[1], xr[5]), bi = fsub(xi[1], xi[5]);
  cr = fadd(xr[3], xr[7]), ci = fadd(xi[3], xi[7]), dr = fsub(xr[3], xr[7]), di = fsub(xi[3], xi[7]);
  float e0 = fadd(ar, cr), e1 = fadd(ai, ci), e4 = fsub(ar, cr), e5 = fsub(ai, ci);
  float e2, e3, e6, e7;
  if (INV) {
    e2 = fsub(br, di), e3 = fadd(bi, dr), e6 = fadd(br, di), e7 = fsub(bi, dr);
    const float t0 = fmul(fsub(e2, e3), s), t1 = fmul(fadd(e3, e2), s);
    e2 = t0, e3 = t1;
    const float t2 = fmul(fadd(e6, e7), -s), t3 = fmul(fsub(e6, e7), s);
    e6 = t2, e7 = t3;
  } else {
    e2 = fadd(br, di), e3 = fsub(bi, dr), e6 = fsub(br, di), e7 = fadd(bi, dr);
    const float t0 = fmul(fadd(e2, e3), s), t1 = fmul(fsub(e3, e2), s);
    e2 = t0, e3 = t1;
    const float t2 = fmul(fsub(e7, e6), s), t3 = fmul(fadd(e6, e7), -s);
    e6 = t2, e7 = t3;
    const float t4 = e4;
    e4 = e5;
    e5 = -t4;
  }
  xr[0] = fadd(n00, e0), xi[0] = fadd(n01, e1);
  xr[1] = fadd(n10, e2), xi[1] = fadd(n11, e3);
  xr[3] = fadd(n30, e6), xi[3] = fadd(n31, e7);
  xr[4] = fsub(n00, e0), xi[4] = fsub(n01, e1);
  xr[5] = fsub(n10, e2), xi[5] = fsub(n11, e3);
  xr[7] = fsub(n30, e6), xi[7] = fsub(n31, e7);
  if (INV) {
    xr[2] = fsub(n21, e4), xi[2] = fadd(n20, e5);
    xr[6] = fadd(n21, e4), xi[6] = fsub(n20, e5);
  } else {
    xr[2] = fadd(n20, e4), xi[2] = fadd(n21, e5);
    xr[6] = fsub(n20, e4), xi[6] = fsub(n21, e5);
  }
}

// ---- twiddles of the row transforms, staged in shared memory ------------------------------------------------
// The effective-twiddle table is indexed per lane in stages B and C (16 distinct entries 512 B apart for a warp in
// stage B: a global load costs 16 L1 wavefronts and its latency).  The CTA stages what it needs once, transposed so
// that every warp-wide load is one conflict-free wavefront:
//   a [3][6]      eff[64 m], m = 1..3 (stage A, warp-uniform)
//   b [5][6][16]  entry e for lane class mm = g & 15: e = 0 -> eff[16 mm], e = 1 + a -> eff[4 mm + 64 a]
//   c [6][256]    eff[m], transposed (stage C)
struct RowTw {
  const float* a;
  const float* b;
  const float* c;
};
constexpr int kRowTwFloats = 18 + 480 + 1536;

// Every thread of the CTA calls this; the caller issues the __syncthreads().
__device__ __forceinline__ RowTw fill_row_tw(float* __restrict__ dst, const float4* __restrict__ eff8,
                                             int n_floats = kRowTwFloats) {
  const float* eff = reinterpret_cast<const float*>(eff8);  // [256][8]: c1 s1 c2 s2 c3 s3 0 0
  for (int i = threadIdx.x; i < n_floats; i += blockDim.x) {  // 498: stages A and B only
    int j, f;
    if (i < 18) {
      j = 64 * (1 + i / 6), f = i % 6;
    } else if (i < 498) {
      const int r = i - 18, mm = r & 15, ef = r >> 4, e = ef / 6;
      f = ef - 6 * e;
      j = e == 0 ? 16 * mm : 4 * mm + 64 * (e - 1);
    } else {
      const int r = i - 498;
      j = r & 255, f = r >> 8;
    }
    dst[i] = __ldg(eff + 8 * j + f);
  }
  return RowTw{dst, dst + 18, dst + 498};
}

// b, c, d *= their twiddles (e[0..5] = c1 s1 c2 s2 c3 s3)
template <bool INV>
__device__ __forceinline__ void rot3(float2& b, float2& c, float2& d, const float (&e)[6]) {
  rot<INV>(b, e[0], e[1]);
  rot<INV>(c, e[2], e[3]);
  rot<INV>(d, e[4], e[5]);
}

// ---- stage A: passes 0 and 1 on v[t], t = q + 4 i (butterfly q, input i) -> y[16 g + e] in v[e] -----------
// An index j == 0 into the twiddle table means the reference multiplies nothing (kept: x * 1 + y * 0 would turn a
// -0 into +0).
template <bool INV>
__device__ __forceinline__ void stage_a(float2 (&v)[16], const RowTw tw) {
  float2 y[16];
  float e[6];
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    float2 a = v[q], b = v[q + 4], c = v[q + 8], d = v[q + 12];
    bfly<INV, !INV>(a, b, c, d);
    y[4 * q + 0] = a, y[4 * q + 1] = b, y[4 * q + 2] = c, y[4 * q + 3] = d;
  }
#pragma unroll
  for (int m = 0; m < 4; ++m) {
    float2 a = y[m], b = y[m + 4], c = y[m + 8], d = y[m + 12];
    if (m) {
#pragma unroll
      for (int f = 0; f < 6; ++f) e[f] = tw.a[6 * (m - 1) + f];
      rot3<INV>(b, c, d, e);
    }
    bfly<INV, false>(a, b, c, d);
    v[m] = a, v[m + 4] = b, v[m + 8] = c, v[m + 12] = d;
  }
}

// ---- stage B: passes 2 and 3 on v[t] = y[base + 16 t], t = a + 4 q; mm = base & 15 ------------------------
template <bool INV>
__device__ __forceinline__ void stage_b(float2 (&v)[16], const RowTw tw, int mm) {
  float e[6];
  if (mm) {
#pragma unroll
    for (int f = 0; f < 6; ++f) e[f] = tw.b[f * 16 + mm];
#pragma unroll
    for (int q = 0; q < 4; ++q) rot3<INV>(v[4 * q + 1], v[4 * q + 2], v[4 * q + 3], e);
  }
#pragma unroll
  for (int q = 0; q < 4; ++q) bfly<INV, false>(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
#pragma unroll
  for (int a = 0; a < 4; ++a) {
    if (a || mm) {
#pragma unroll
      for (int f = 0; f < 6; ++f) e[f] = tw.b[((1 + a) * 6 + f) * 16 + mm];
      rot3<INV>(v[a + 4], v[a + 8], v[a + 12], e);
    }
    bfly<INV, false>(v[a], v[a + 4], v[a + 8], v[a + 12]);
  }
}

// ---- warp-level 512-point transform (format converter, domain switcher) ------------------------------------
// 512-point transform of one warp's buffer, natural order in, up to its last pass: what is left is the radix-2 pass
// out[i] = buf[i] + w_i buf[i + 256], out[i + 256] = buf[i] - w_i buf[i + 256] (w_i = rad2[i]), which the two callers
// fuse with what they do to the result (they need half of it, or only its real part).
template <bool INV>
__device__ __forceinline__ void fft512_tail(float2 (&v)[16], float2* __restrict__ buf, int lane, const RowTw eff8) {
  // v[t] = point r3(lane) + 32 t on entry
  stage_a<INV>(v, eff8);
#pragma unroll
  for (int e = 0; e < 16; ++e) buf[phys(16 * lane + e)] = v[e];
  __syncwarp();
  const int base = 256 * (lane >> 4) + (lane & 15);
#pragma unroll
  for (int t = 0; t < 16; ++t) v[t] = buf[phys(base + 16 * t)];
  stage_b<INV>(v, eff8, lane & 15);
#pragma unroll
  for (int t = 0; t < 16; ++t) buf[phys(base + 16 * t)] = v[t];
  __syncwarp();
}

// lane -> its digit-reversed position inside a group of 32 points (the order stage A wants its inputs in)
__device__ __forceinline__ int rev32(int lane) { return 8 * (lane & 3) + 2 * ((lane >> 2) & 3) + (lane >> 4); }

// The row transforms (stages A..C) of all DIM2 rows, in place.  Entry: work holds the rows in natural order,
// preceded by a __syncthreads() issued by the caller.  Exit: rows transformed, __syncthreads() issued.
// REAL_HALF: the input is real and zero beyond position 511 of every row (a zero-padded real block); the caller has
// staged only the real parts, as floats, at rowf[phys(p)] (rowf = the row's memory viewed as float).  The zeros are
// compile-time constants then, and what IEEE arithmetic allows to fold in pass 0 folds.
template <int DIM2, bool INV, bool REAL_HALF = false>
__device__ __forceinline__ void rows_1024(float2* __restrict__ work, const RowTw tw) {
  const int tid = threadIdx.x;
  const int r = tid >> 6, g = tid & 63;
  float2* row = work + r * kRowStride;
  float2 v[16];
  float e[6];
  {
    // digit-reversed read: butterfly b = 4g + q reads r3(g) + 64 q + 256 i
    const int r3 = ((g & 3) << 4) | (g & 12) | (g >> 4);
    if constexpr (REAL_HALF) {
      const float* rowf = reinterpret_cast<const float*>(row);
#pragma unroll
      for (int t = 0; t < 16; ++t)
        v[t] = (t < 8) ? make_float2(rowf[phys(r3 + 64 * t)], 0.0f) : make_float2(0.0f, 0.0f);
    } else {
#pragma unroll
      for (int t = 0; t < 16; ++t) v[t] = row[phys(r3 + 64 * t)];
    }
    __syncthreads();
    stage_a<INV>(v, tw);
#pragma unroll
    for (int t = 0; t < 16; ++t) row[phys(16 * g + t)] = v[t];
    __syncthreads();
  }
  {
    const int mm = g & 15;
    const int base = 256 * (g >> 4) + mm;
#pragma unroll
    for (int t = 0; t < 16; ++t) v[t] = row[phys(base + 16 * t)];
    stage_b<INV>(v, tw, mm);
#pragma unroll
    for (int t = 0; t < 16; ++t) row[phys(base + 16 * t)] = v[t];
    __syncthreads();
  }
  constexpr bool kSameM = (64 * DIM2) % 256 == 0;  // the thread's four pass-4 butterflies share one twiddle
  if (kSameM) {
#pragma unroll
    for (int f = 0; f < 6; ++f) e[f] = tw.c[f * 256 + (tid & 255)];
  }
#pragma unroll
  for (int u = 0; u < 4; ++u) {
    const int w = tid + 64 * DIM2 * u;
    float2* rw = work + (w >> 8) * kRowStride;
    const int m = w & 255;
    float2 a = rw[phys(m)], b = rw[phys(m + 256)], c = rw[phys(m + 512)], d = rw[phys(m + 768)];
    if (m) {
      if (!kSameM) {
#pragma unroll
        for (int f = 0; f < 6; ++f) e[f] = tw.c[f * 256 + m];
      }
      rot3<INV>(b, c, d, e);
    }
    bfly<INV, false>(a, b, c, d);
    rw[phys(m)] = a, rw[phys(m + 256)] = b, rw[phys(m + 512)] = c, rw[phys(m + 768)] = d;
  }
  __syncthreads();
}

// Column i of the transformed rows: inter-stage twiddle (mix[j * 1024 + i] = (cos, sin) of 2 pi i j / N, staged in
// shared memory by the caller) and the DIM2-point transform.  Output k = j * 1024 + i in (xr[j], xi[j]).
template <int DIM2, bool INV>
__device__ __forceinline__ void column(const float2* __restrict__ work, const float2* __restrict__ mix, int i,
                                       float (&xr)[DIM2], float (&xi)[DIM2]) {
#pragma unroll
  for (int j = 0; j < DIM2; ++j) {
    const float2 v = work[j * kRowStride + phys(i)];
    const float2 cs = mix[j * 1024 + i];
    if (INV) {
      xr[j] = fsub(fmul(v.x, cs.x), fmul(v.y, cs.y));
      xi[j] = fadd(fmul(v.x, cs.y), fmul(v.y, cs.x));
    } else {
      xr[j] = fadd(fmul(v.x, cs.x), fmul(v.y, cs.y));
      xi[j] = fsub(fmul(v.y, cs.x), fmul(v.x, cs.y));
    }
  }
  if constexpr (DIM2 == 8) {
    col8<INV>(xr, xi);
  } else {
    static_assert(DIM2 == 2, "column transforms exist for 2 and 8 rows");
    const float a = fadd(xr[0], xr[1]), b = fsub(xr[0], xr[1]), c = fadd(xi[0], xi[1]), d = fsub(xi[0], xi[1]);
    xr[0] = a, xr[1] = b, xi[0] = c, xi[1] = d;
  }
}

}  // namespace fftx
}  // namespace mpeghb200
