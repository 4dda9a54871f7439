// fd_synth.cu -- IMDCT + windowing + overlap-add for FD-coded channels, one warp per channel-frame.
//
// Stands in for ia_core_coder_fd_frm_dec (decoder/ia_core_coder_imdct.c:832) and everything below it:
// scale + pre-twiddle (:87, :366), the N/2-point complex transform impeghd_rad2_cplx_ifft
// (decoder/impeghd_ifft.c:297), post-twiddle (:203), the windowing kernels
// (decoder/ia_core_coder_basic_ops.c:124-508) and the overlap save (:762-815).
//
// Parity: every output sample is produced by the same sequence of separately rounded binary32 operations
// as the reference (no FMA contraction, same butterfly forms, same table entries), so the result is
// bit-identical; only the *schedule* differs.
//
// Schedule (B200): a unit is 4 KB of spectrum in, 4 KB of overlap in/out, 4 KB of PCM out -> HBM bound.
// One warp owns one unit: 512 complex points = 16 per lane.
//   1. lane l loads the 16 coefficient pairs {X[2c], X[2c+1]}, c = l + 32u, as coalesced 8-byte loads; the
//      mirrored operand X[N-1-2c] of the pre-twiddle is the odd half of lane 31-l's pair (one shuffle).
//      Those 16 points are exactly the inputs of the lane's four digit-reversed radix-4 butterflies of
//      pass 0 and, after them, of its four pass-1 butterflies: two passes in registers.
//   2. a padded shared-memory transpose (index i -> i + i/16, conflict free both ways) regroups the data so
//      that passes 2 and 3 again run in registers; the final radix-2 pass pairs lanes l and l^16 by shuffle.
//   3. post-twiddled samples go to shared memory once more so that windowing, overlap-add and the new
//      overlap are computed per float4 with fully coalesced 16-byte global accesses.
// EIGHT_SHORT frames run eight 64-point transforms the same way (4 lanes each).
#include <type_traits>

#include "common.h"

namespace mpeghb200 {
namespace {

constexpr unsigned kFull = 0xffffffffu;

__device__ __forceinline__ int pad16(int i) { return i + (i >> 4); }

// (r, i) *= (c - j s) in the reference's product order: (r*c + i*s, -(r*s) + i*c)
__device__ __forceinline__ void rot(float& r, float& i, float c, float s) {
  const float nr = fadd(fmul(r, c), fmul(i, s));
  const float ni = fadd(-fmul(r, s), fmul(i, c));
  r = nr;
  i = ni;
}

// Radix-4 butterfly in the reference's "a += b; b = a - 2b" form (decoder/impeghd_ifft.c:353-378).
// In: x0..x3. Out (natural order): x0 -> A, x1 -> B, x2 -> C, x3 -> D.
__device__ __forceinline__ void bfly4(float& x0r, float& x0i, float& x1r, float& x1i, float& x2r,
                                      float& x2i, float& x3r, float& x3i) {
  x0r = fadd(x0r, x2r);
  x0i = fadd(x0i, x2i);
  x2r = sub2(x0r, x2r);
  x2i = sub2(x0i, x2i);
  x1r = fadd(x1r, x3r);
  x1i = fadd(x1i, x3i);
  x3r = sub2(x1r, x3r);
  x3i = sub2(x1i, x3i);

  x0r = fadd(x0r, x1r);
  x0i = fadd(x0i, x1i);
  x1r = sub2(x0r, x1r);
  x1i = sub2(x0i, x1i);
  x2r = fsub(x2r, x3i);
  x2i = fadd(x2i, x3r);
  const float dr = add2(x2r, x3i);
  const float di = sub2(x2i, x3r);
  // reorder to natural output order A, B, C, D
  const float cr = x1r, ci = x1i;
  x1r = x2r;
  x1i = x2i;
  x2r = cr;
  x2i = ci;
  x3r = dr;
  x3i = di;
}

// Passes 0 and 1 on the 16 points a lane holds: zr/zi[u], u = t + 4q -> y[16g + e] in place (e = index).
__device__ __forceinline__ void fft16_first_two_passes(float (&r)[16], float (&i)[16], const float* __restrict__ eff) {
  // pass 0: butterfly t takes u = t, t+4, t+8, t+12 and produces y[4t .. 4t+3]
  float ar[16], ai[16];
#pragma unroll
  for (int t = 0; t < 4; ++t) {
    float x0r = r[t], x0i = i[t], x1r = r[t + 4], x1i = i[t + 4], x2r = r[t + 8], x2i = i[t + 8],
          x3r = r[t + 12], x3i = i[t + 12];
    bfly4(x0r, x0i, x1r, x1i, x2r, x2i, x3r, x3i);
    ar[4 * t + 0] = x0r, ai[4 * t + 0] = x0i;
    ar[4 * t + 1] = x1r, ai[4 * t + 1] = x1i;
    ar[4 * t + 2] = x2r, ai[4 * t + 2] = x2i;
    ar[4 * t + 3] = x3r, ai[4 * t + 3] = x3i;
  }
  // pass 1 (span 4, table step 64): butterfly m takes e = m, m+4, m+8, m+12; twiddle index j = 64 m
#pragma unroll
  for (int m = 0; m < 4; ++m) {
    float x0r = ar[m], x0i = ai[m], x1r = ar[m + 4], x1i = ai[m + 4], x2r = ar[m + 8], x2i = ai[m + 8],
          x3r = ar[m + 12], x3i = ai[m + 12];
    if (m > 0) {
      const float* e = eff + 64 * m * 6;
      rot(x1r, x1i, __ldg(e + 0), __ldg(e + 1));
      rot(x2r, x2i, __ldg(e + 2), __ldg(e + 3));
      rot(x3r, x3i, __ldg(e + 4), __ldg(e + 5));
    }
    bfly4(x0r, x0i, x1r, x1i, x2r, x2i, x3r, x3i);
    r[m] = x0r, i[m] = x0i;
    r[m + 4] = x1r, i[m + 4] = x1i;
    r[m + 8] = x2r, i[m + 8] = x2i;
    r[m + 12] = x3r, i[m + 12] = x3i;
  }
}

// One twiddled radix-4 butterfly on registers; j == 0 means "no multiplication at all" in the reference.
__device__ __forceinline__ void bfly4_tw(float& x0r, float& x0i, float& x1r, float& x1i, float& x2r,
                                         float& x2i, float& x3r, float& x3i, const float* __restrict__ e,
                                         bool twiddle) {
  if (twiddle) {
    rot(x1r, x1i, __ldg(e + 0), __ldg(e + 1));
    rot(x2r, x2i, __ldg(e + 2), __ldg(e + 3));
    rot(x3r, x3i, __ldg(e + 4), __ldg(e + 5));
  }
  bfly4(x0r, x0i, x1r, x1i, x2r, x2i, x3r, x3i);
}

struct Side {
  int seq, shape, shape_prev, tkt;
};
__device__ __forceinline__ Side unpack_side(uint32_t s) {
  Side o;
  o.seq = int(s & 7u);
  o.shape = int((s >> 3) & 1u);
  o.shape_prev = int((s >> 4) & 1u);
  o.tkt = int(((s >> 5) & 1u) << 1) | int((s >> 6) & 1u);
  return o;
}

// ---- generic pre-twiddle for transform kernel types 1..3 (decoder/ia_core_coder_imdct.c:98-170) ----------
// xs = scaled spectrum block in shared memory (N floats), nl = N/2, c = output index. Rare path.
__device__ __noinline__ float2 pretwiddle_generic(const float* xs, int N, int tkt, int c, const DeviceTables& T) {
  const int nl = N >> 1;
  float zr, zi;
  if (tkt == 3) {
    const float2 cs = __ldg(((nl == 512) ? T.tw512 : T.tw64) + c);
    const float a = xs[2 * c];               // even entries untouched by the reorder
    const float b = -xs[N - 1 - 2 * c];      // odd entry 2c+1 after the swap-negate reorder
    zr = fsub(fmul(b, cs.x), fmul(a, cs.y));
    zi = fadd(fmul(a, cs.x), fmul(b, cs.y));
    return make_float2(zr, zi);
  }
  // DCT-II / DST-II.  X'[n] = X[N-1-n] for type 1.
  auto X = [&](int n) { return (tkt == 1) ? xs[N - 1 - n] : xs[n]; };
  const float* C = T.dct_cos;
  const float* S = T.dct_sin;
  if (c == 0) {
    const float s4 = 0.70710677f;  // (float)sin(pi/4)
    zr = fadd(X(0), fmul(X(nl), s4));
    zi = fsub(X(0), fmul(X(nl), s4));
    return make_float2(zr, zi);
  }
  const int i = (c <= nl / 2) ? c : nl - c;
  const float re1 = fmul(fadd(fmul(X(i), __ldg(C + i - 1)), fmul(X(N - i), __ldg(S + i - 1))), 0.5f);
  const float re2 = fmul(fadd(fmul(X(nl - i), __ldg(S + nl + i - 1)), fmul(X(nl + i), __ldg(C + nl + i - 1))), 0.5f);
  const float im1 = fmul(fsub(fmul(X(i), __ldg(S + i - 1)), fmul(X(N - i), __ldg(C + i - 1))), 0.5f);
  const float im2 = fmul(fsub(fmul(X(nl - i), __ldg(C + nl + i - 1)), fmul(X(nl + i), __ldg(S + nl + i - 1))), 0.5f);
  const float c4 = __ldg(C + 4 * i - 1), s4 = __ldg(S + 4 * i - 1);
  const float tmp_i = fsub(im1, im2);
  const float ii = fsub(fmul(fsub(re1, re2), c4), fmul(fadd(im1, im2), s4));
  const float tmp_r = fadd(re1, re2);
  const float rr = fadd(fmul(fadd(im1, im2), c4), fmul(fsub(re1, re2), s4));
  if (2 * c < nl) {
    zi = fadd(ii, tmp_i);
    zr = fsub(tmp_r, rr);
  } else if (2 * c > nl) {
    zi = fsub(ii, tmp_i);
    zr = fadd(tmp_r, rr);
  } else {
    // c == nl/2: the reference overwrites the element it is about to read (:140-149)
    zi = fadd(fsub(ii, tmp_i), tmp_i);
    zr = fsub(tmp_r, fsub(tmp_r, fadd(tmp_r, rr)));
  }
  return make_float2(zr, zi);
}

// Post-twiddle of transform output y[k] = (r, q) into the block's time-aliased samples xs[0..N-1]
// (decoder/ia_core_coder_imdct.c:214-273), all four kernel types. `first_odd` selects which of the two
// stores goes first (bank-conflict avoidance only).
__device__ __forceinline__ void posttwiddle_store(float* xs, int N, int tkt, int k, float r, float q, float2 cs,
                                                  bool first_odd) {
  const int nl = N >> 1;
  if (tkt == 1 || tkt == 2) {
    const bool lo = k < nl / 2;
    const int kk = lo ? k : nl - 1 - k;
    const int nr = lo ? 4 * kk : 4 * kk + 3;
    const int nq = lo ? 4 * kk + 2 : 4 * kk + 1;
    xs[nr] = (tkt == 1 && (nr & 1)) ? -r : r;
    xs[nq] = (tkt == 1 && (nq & 1)) ? -q : q;
    return;
  }
  const float xa = -fsub(fmul(r, cs.x), fmul(q, cs.y));
  float xb = -fadd(fmul(q, cs.x), fmul(r, cs.y));
  if (tkt == 3) xb = -xb;
  if (first_odd) {
    xs[N - 1 - 2 * k] = xb;
    xs[2 * k] = xa;
  } else {
    xs[2 * k] = xa;
    xs[N - 1 - 2 * k] = xb;
  }
}

__device__ __forceinline__ float4 ld4(const float* p) { return *reinterpret_cast<const float4*>(p); }
__device__ __forceinline__ float4 ldg4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }
__device__ __forceinline__ float4 ldcs4(const float* p) { return __ldcs(reinterpret_cast<const float4*>(p)); }
__device__ __forceinline__ void st4(float* p, float4 v) { *reinterpret_cast<float4*>(p) = v; }
__device__ __forceinline__ void stcs4(float* p, float4 v) { __stcs(reinterpret_cast<float4*>(p), v); }
__device__ __forceinline__ float4 rev4(float4 v) { return make_float4(v.w, v.z, v.y, v.x); }
__device__ __forceinline__ float4 neg4(float4 v) { return make_float4(-v.x, -v.y, -v.z, -v.w); }

// out = s*wf + o*wr per component
__device__ __forceinline__ float4 win4(float4 s, float4 wf, float4 o, float4 wr) {
  return make_float4(mul_add_mul(s.x, wf.x, o.x, wr.x), mul_add_mul(s.y, wf.y, o.y, wr.y),
                     mul_add_mul(s.z, wf.z, o.z, wr.z), mul_add_mul(s.w, wf.w, o.w, wr.w));
}

// ---- long block: windowing_long1 / windowing_long3 + overlap save, gather form --------------------------
// xs[0..1023] = IMDCT output. For position i: s(i) = xs[512+i] (i < 512), -/+ xs[1535-i] (i >= 512).
__device__ __forceinline__ void window_long(const float* xs, float* __restrict__ ov, float* __restrict__ out,
                                            const Side sd, const DeviceTables& T, int lane) {
  const bool start_like = (sd.seq == 0 || sd.seq == 1);  // ONLY_LONG, LONG_START
  const bool pos_tail = start_like ? (sd.tkt == 2 || sd.tkt == 3) : (sd.tkt == 3);
  const bool hi_pos = (sd.tkt == 1 || sd.tkt == 2);  // sign of overlap[512 + i] = +-x[i]
  const bool lo_pos = (sd.tkt == 2 || sd.tkt == 3);  // sign of overlap[511 - i]
  const float* wl = T.win_long[sd.shape_prev];
  const float* wsh = T.win_short[sd.shape_prev];
#pragma unroll
  for (int it = 0; it < 8; ++it) {
    const int i4 = 4 * (lane + 32 * it);
    float4 s;
    if (i4 < 512) {
      s = ld4(xs + 512 + i4);
    } else {
      s = rev4(ld4(xs + 1532 - i4));
      if (!pos_tail) s = neg4(s);
    }
    float4 o;
    if (start_like) {
      const float4 ovv = ldcs4(ov + i4);
      o = win4(s, ldg4(wl + i4), ovv, rev4(ldg4(wl + 1020 - i4)));
    } else if (i4 < 448) {
      o = ldcs4(ov + i4);
    } else if (i4 < 576) {
      const float4 ovv = ldcs4(ov + i4);
      o = win4(s, ldg4(wsh + i4 - 448), ovv, rev4(ldg4(wsh + 572 - i4)));
    } else {
      o = s;
    }
    stcs4(out + i4, o);
    // new overlap at the same positions this lane just consumed: no cross-lane hazard
    float4 n;
    if (i4 >= 512) {
      n = ld4(xs + i4 - 512);
      if (!hi_pos) n = neg4(n);
    } else {
      n = rev4(ld4(xs + 508 - i4));
      if (!lo_pos) n = neg4(n);
    }
    st4(ov + i4, n);
  }
}

// ---- eight short blocks: windowing_short2/3/4, gather form ---------------------------------------------
// xs[128k + n] = block k's IMDCT output.  Work positions P in [0, 2048): [0,1024) -> out, [1024,2048) ->
// new overlap.  See oracle/oracle_fd.c short_blocks() for the derivation of the accumulation order.
// A lane's four consecutive work positions always fall on the same window phase t0 = (4 lane + 64) & 127
// (all region boundaries 448, 576, 1600 are 64 mod 128 and a row of 32 lanes spans exactly 128 positions), so
// the window entries a lane ever needs are 8 per shape: w[t0..t0+3] and w[127-t0-3..127-t0].
struct ShortWin {
  float f[4], r[4];  // f[e] = w[t0 + e], r[e] = w[127 - t0 - e]
};
__device__ __forceinline__ ShortWin load_short_win(const float* __restrict__ w, int t0) {
  const float4 a = ldg4(w + t0), b = ldg4(w + 124 - t0);
  ShortWin s;
  s.f[0] = a.x, s.f[1] = a.y, s.f[2] = a.z, s.f[3] = a.w;
  s.r[0] = b.w, s.r[1] = b.z, s.r[2] = b.y, s.r[3] = b.x;
  return s;
}

// Iterations it0, it0 + it_step, ... of the 16: several warps can share one unit (it_step a divisor of 8, so that
// the overlap position an iteration writes was read by the same warp eight iterations earlier).
__device__ void window_short(const float* xs, float* __restrict__ ov, float* __restrict__ out, const Side sd,
                             const DeviceTables& T, int lane, int it0 = 0, int it_step = 1) {
  const int tkt = sd.tkt;
  const int t0 = (4 * lane + 64) & 127;
  const ShortWin wc = load_short_win(T.win_short[sd.shape], t0);
  const ShortWin wp = load_short_win(T.win_short[sd.shape_prev], t0);
  const bool lo = t0 < 64;  // t0 .. t0+3 lie on one side of 64
#pragma unroll 1
  for (int it = it0; it < 16; it += it_step) {
    const int P4 = 4 * (lane + 32 * it);
    float v[4];
    if (P4 < 448) {
      const float4 o = ldcs4(ov + P4);
      v[0] = o.x, v[1] = o.y, v[2] = o.z, v[3] = o.w;
    } else if (P4 < 576) {
      // `lo` (and below the block number k) differs between the two halves of a warp: everything that depends on
      // them is a select on an index, a sign or a value, never a branch -- as branches they made the warp run every
      // element twice, with a reconvergence each, and this loop was more than half of a short-block unit
      const float4 o = ldcs4(ov + P4);
      const float oo[4] = {o.x, o.y, o.z, o.w};
      const bool neg = !lo && tkt != 3;
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int t = t0 + e;
        float a = xs[lo ? 64 + t : 191 - t];
        a = neg ? -a : a;
        v[e] = mul_add_mul(a, wp.f[e], oo[e], wp.r[e]);
      }
    } else if (P4 < 1600) {
      const int k = (P4 - 576) >> 7;
      const float* xk = xs + 128 * k;
      const bool last = k == 7;
      const bool quirk = k == 0 && tkt == 3 && lo;  // reference quirk, basic_ops.c:404: w + 0, no 0 + x
      const bool negf = lo ? (tkt != 3) : true;     // falling half of block k
      const bool negr = !lo && tkt != 3;            // rising half of block k + 1
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int t = t0 + e;
        float f = xk[lo ? 63 - t : t - 64];
        f = negf ? -f : f;
        const float w = quirk ? fadd(wc.r[e], 0.0f) : wc.r[e];
        const float m = fmul(f, w);
        float acc = quirk ? m : fadd(0.0f, last ? f : m);
        float r = xk[last ? 0 : (lo ? 128 + 64 + t : 128 + 191 - t)];
        r = negr ? -r : r;
        const float acc2 = fadd(acc, fmul(r, wc.f[e]));
        v[e] = last ? acc : acc2;
      }
    } else {
      v[0] = v[1] = v[2] = v[3] = 0.0f;
    }
    const float4 o4 = make_float4(v[0], v[1], v[2], v[3]);
    if (P4 < 1024) {
      stcs4(out + P4, o4);
      // old overlap beyond 576 is discarded by the reference (memset in windowing_short2)
    }
    // new overlap position P4 - 1024 is the position this same lane read eight iterations ago: the
    // in-place update has no cross-lane hazard
    if (P4 >= 1024) st4(ov + P4 - 1024, o4);
  }
}

// General path: every window sequence and transform kernel type (EIGHT_SHORT frames and kernel switching
// are rare; ONLY_LONG-family frames with kernel type 0 take fd_unit_long_fast instead).
// window_later: for EIGHT_SHORT units stop after the post-twiddle (time-aliased samples in sm); the caller runs
// window_short itself (slow_units shares it between the CTA's warps).
__device__ __noinline__ void fd_unit_generic(const float* __restrict__ X, float* __restrict__ ov, float* __restrict__ o,
                                             const Side sd, float* sm, const DeviceTables& T, int lane,
                                             bool window_later = false) {
  float2* smc = reinterpret_cast<float2*>(sm);
  const bool is_short = (sd.seq == 2);
  const float norm = is_short ? 0.0078125f : 0.0009765625f;  // 2 / (2N)
  // The tables of this rarely taken path are cold in L1, and its table loads sit one by one behind shuffles and
  // dependent arithmetic (the 16 pre-twiddle loads alone: 16 x ~250 cycles of L2 latency, a fifth of a short-block
  // unit).  Ask for the lines now, while the coefficients are on their way; costs no register.
  if (is_short) {
    const char* t64 = reinterpret_cast<const char*>(T.tw64);           // 512 B
    const char* eff = reinterpret_cast<const char*>(T.eff);            // 6 KB
    const char* w0 = reinterpret_cast<const char*>(T.win_short[sd.shape]);
    const char* w1 = reinterpret_cast<const char*>(T.win_short[sd.shape_prev]);
    if (lane < 4) asm volatile("prefetch.global.L1 [%0];" ::"l"(t64 + 128 * lane));
    asm volatile("prefetch.global.L1 [%0];" ::"l"(eff + 128 * lane));
    if (lane < 16) asm volatile("prefetch.global.L1 [%0];" ::"l"(eff + 4096 + 128 * lane));
    if (lane < 4) asm volatile("prefetch.global.L1 [%0];" ::"l"(w0 + 128 * lane));
    if (lane >= 4 && lane < 8) asm volatile("prefetch.global.L1 [%0];" ::"l"(w1 + 128 * (lane - 4)));
  }

  float zr[16], zi[16];
  if (sd.tkt == 0) {
    // c = lane + 32u (long) or, per short block T8 = lane>>2, c = g + 4u with g = lane&3
    float2 eo[16];
    const float2* X2 = reinterpret_cast<const float2*>(X);
#pragma unroll
    for (int u = 0; u < 16; ++u) {
      const int idx = is_short ? (64 * (lane >> 2) + (lane & 3) + 4 * u) : (lane + 32 * u);
      eo[u] = __ldcs(X2 + idx);
    }
#pragma unroll
    for (int u = 0; u < 16; ++u) {
      eo[u].x = fmul(eo[u].x, norm);
      eo[u].y = fmul(eo[u].y, norm);
    }
    const int mirror = is_short ? 3 : 31;
    // Short blocks: the 64 twiddles are parked in the warp's (still unused) shared buffer with one coalesced load
    // that overlaps the coefficient loads.  Read from global inside the loop, each of the 16 loads sat behind the
    // shuffle of the previous iteration and cost its own L2 latency (~4 k cycles, a fifth of a short-block unit);
    // holding all 16 in registers instead cost the fast path of this kernel 2 % (register allocation).
    if (is_short) {
      smc[lane] = __ldg(T.tw64 + lane);
      smc[lane + 32] = __ldg(T.tw64 + lane + 32);
      __syncwarp();
    }
#pragma unroll
    for (int u = 0; u < 16; ++u) {
      const int c = is_short ? ((lane & 3) + 4 * u) : (lane + 32 * u);
      const float2 cs = is_short ? smc[c] : __ldg(T.tw512 + c);
      const float a = eo[u].x;
      const float b = __shfl_xor_sync(kFull, eo[15 - u].y, mirror);
      zr[u] = fsub(-fmul(a, cs.x), fmul(b, cs.y));
      zi[u] = fsub(fmul(b, cs.x), fmul(a, cs.y));
    }
    __syncwarp();  // the buffer is reused by the transpose below
  } else {
    // stage the scaled spectrum in shared memory, then evaluate the type-specific pre-twiddle per point
#pragma unroll
    for (int it = 0; it < 8; ++it) {
      const int i4 = 4 * (lane + 32 * it);
      float4 v = ldcs4(X + i4);
      v.x = fmul(v.x, norm), v.y = fmul(v.y, norm), v.z = fmul(v.z, norm), v.w = fmul(v.w, norm);
      st4(sm + i4, v);
    }
    __syncwarp();
#pragma unroll
    for (int u = 0; u < 16; ++u) {
      const float2 z = is_short ? pretwiddle_generic(sm + 128 * (lane >> 2), 128, sd.tkt, (lane & 3) + 4 * u, T)
                                : pretwiddle_generic(sm, 1024, sd.tkt, lane + 32 * u, T);
      zr[u] = z.x, zi[u] = z.y;
    }
    __syncwarp();
  }

  // passes 0 + 1 in registers: lane now holds y[16 g + e] (per transform)
  fft16_first_two_passes(zr, zi, T.eff);

  // transpose through shared memory
  {
    int base;  // index of y[16 g] in the warp's 512-point space
    if (is_short) {
      base = 64 * (lane >> 2) + 16 * (lane & 3);
    } else {
      const int g = ((lane >> 3) & 3) | (((lane >> 1) & 3) << 2) | ((lane & 1) << 4);
      base = 16 * g;
    }
#pragma unroll
    for (int e = 0; e < 16; ++e) smc[pad16(base + e)] = make_float2(zr[e], zi[e]);
  }
  __syncwarp();

  if (!is_short) {
    const int H = lane >> 4, m16 = lane & 15;
#pragma unroll
    for (int v = 0; v < 16; ++v) {
      const float2 t = smc[pad16(256 * H + m16 + 16 * v)];
      zr[v] = t.x, zi[v] = t.y;
    }
    __syncwarp();
    // pass 2 (span 16, table step 16): butterflies t' on v = 4t'..4t'+3, twiddle index j = 16 m16
    {
      const float* e = T.eff + 16 * m16 * 6;
#pragma unroll
      for (int t = 0; t < 4; ++t)
        bfly4_tw(zr[4 * t], zi[4 * t], zr[4 * t + 1], zi[4 * t + 1], zr[4 * t + 2], zi[4 * t + 2], zr[4 * t + 3],
                 zi[4 * t + 3], e, m16 != 0);
    }
    // pass 3 (span 64, table step 4): butterflies r on v = r, r+4, r+8, r+12, j = 4 (m16 + 16 r)
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      const float* e = T.eff + 4 * (m16 + 16 * r) * 6;
      bfly4_tw(zr[r], zi[r], zr[r + 4], zi[r + 4], zr[r + 8], zi[r + 8], zr[r + 12], zi[r + 12], e,
               (m16 + 16 * r) != 0);
    }
    // final radix-2 pass (decoder/impeghd_ifft.c:779-837) between lanes l and l^16, then post-twiddle
#pragma unroll
    for (int v = 0; v < 16; ++v) {
      const int p = m16 + 16 * v;
      const float pr = __shfl_xor_sync(kFull, zr[v], 16);
      const float pi = __shfl_xor_sync(kFull, zi[v], 16);
      const float x0r = H ? pr : zr[v], x0i = H ? pi : zi[v];
      float x1r = H ? zr[v] : pr, x1i = H ? zi[v] : pi;
      const float2 w = __ldg(T.rad2_512 + p);
      rot(x1r, x1i, w.x, w.y);
      const float yr = H ? fsub(x0r, x1r) : fadd(x0r, x1r);
      const float yi = H ? fsub(x0i, x1i) : fadd(x0i, x1i);
      const int k = 256 * H + p;
      posttwiddle_store(sm, 1024, sd.tkt, k, yr, yi, __ldg(T.tw512 + k), H != 0);
    }
    __syncwarp();
    window_long(sm, ov, o, sd, T, lane);
  } else {
    // 64-point transforms: last pass (span 16, table step 16) butterfly m = g' + 4r on y[m + 16 q]
    const int T8 = lane >> 2, gq = lane & 3;
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const float2 t = smc[pad16(64 * T8 + gq + 4 * r + 16 * q)];
        zr[4 * r + q] = t.x, zi[4 * r + q] = t.y;
      }
    __syncwarp();
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      const int m = gq + 4 * r;
      const float* e = T.eff + 16 * m * 6;
      bfly4_tw(zr[4 * r], zi[4 * r], zr[4 * r + 1], zi[4 * r + 1], zr[4 * r + 2], zi[4 * r + 2], zr[4 * r + 3],
               zi[4 * r + 3], e, m != 0);
    }
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int k = gq + 4 * r + 16 * q;
        posttwiddle_store(sm + 128 * T8, 128, sd.tkt, k, zr[4 * r + q], zi[4 * r + q], __ldg(T.tw64 + k), false);
      }
    __syncwarp();
    if (!window_later) window_short(sm, ov, o, sd, T, lane);
  }
}


// ---- fast path: long block, transform kernel type 0, B units per warp in three phases ------------------
// The arithmetic is that of fd_unit_generic; what changes is the schedule.  Profiling the one-unit-per-warp
// kernel showed the L1/shared data pipe (not HBM, not issue) closest to saturation, and 40 % of its global
// wavefronts were twiddle/window table reads.  So a warp now owns B consecutive units and walks them phase
// by phase with the phase's table entries held in REGISTERS (loaded once per warp instead of once per unit):
//   A  coefficients -> scale, pre-twiddle, radix-4 passes 0-1 -> transposed into the unit's shared buffer
//      (the next unit's 16 coalesced 8-byte loads are in flight while the current one computes; the
//      overlap state of all B units is requested into L2 with one bulk prefetch each)
//   B  passes 2-3, final radix-2 pass across lanes l / l^16, post-twiddle -> time-aliased samples in the
//      same shared buffer
//   C  windowing + overlap-add + overlap save in mirrored form: a lane handles the float4 at i and the one
//      at 1020-i, which share window entries and shared-memory operands (half the table and LDS traffic)
// With B = 3 and 14 resident warps per SM the 6144 units of the benchmark shape are one wave.
// Units that are EIGHT_SHORT or use another transform kernel fall through to fd_unit_generic afterwards.
constexpr int kBufFloats = 1092;  // 545 padded complex values (see tpos), reused as 1024 real samples

// padded position of complex point 16 g + e: 16 lanes of a half-warp hit 16 distinct 8-byte bank pairs both
// for the phase-A stores (fixed e, lane -> g) and the phase-B loads (fixed v, consecutive m16)
__device__ __forceinline__ int tpos_g(int g) { return 17 * g + 2 * (g >> 4); }

__device__ __forceinline__ void rot_sel(float& r, float& i, float c, float s, bool on) {
  float nr = r, ni = i;
  rot(nr, ni, c, s);
  r = on ? nr : r;  // j == 0: the reference does not multiply at all (keeps signed zeros intact)
  i = on ? ni : i;
}

struct TabA {
  float4 pre[8];  // (c, s) of tw512[lane + 32u] for u = 2p, 2p+1
};
struct TabB {
  float4 e2a, e2b;      // pass 2: eff[16 m16]
  float4 e3a[4], e3b[4];  // pass 3: eff[4 (m16 + 16 r)]
  float4 rad[4];        // radix-2: rad2_512[m16 + 16 v], v = 2p, 2p+1 < 8 (v >= 8 reuses them rotated)
  float4 post[8];       // post-twiddle tw512[256 H + m16 + 16 v]
};
struct TabC {
  float4 f[4], r[4];  // window[i4 .. i4+3], window[1020-i4 .. 1023-i4], i4 = 4 (lane + 32 it)
};

__device__ __forceinline__ void load_tab_a(TabA& t, const DeviceTables& T, int lane) {
#pragma unroll
  for (int p = 0; p < 8; ++p) t.pre[p] = __ldg(T.pre4 + p * 32 + lane);
}
__device__ __forceinline__ void load_tab_b(TabB& t, const DeviceTables& T, int lane) {
  const int m16 = lane & 15;
  t.e2a = __ldg(T.eff8 + 2 * 16 * m16);
  t.e2b = __ldg(T.eff8 + 2 * 16 * m16 + 1);
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    const int j = 4 * (m16 + 16 * r);
    t.e3a[r] = __ldg(T.eff8 + 2 * j);
    t.e3b[r] = __ldg(T.eff8 + 2 * j + 1);
  }
#pragma unroll
  for (int p = 0; p < 4; ++p) t.rad[p] = __ldg(T.rad2p4 + p * 32 + lane);
#pragma unroll
  for (int p = 0; p < 8; ++p) t.post[p] = __ldg(T.post4 + p * 32 + lane);
}
__device__ __forceinline__ void load_tab_c(TabC& t, const float* __restrict__ w, int lane) {
#pragma unroll
  for (int it = 0; it < 4; ++it) {
    const int i4 = 4 * (lane + 32 * it);
    t.f[it] = ldg4(w + i4);
    t.r[it] = ldg4(w + 1020 - i4);
  }
}

__device__ __forceinline__ void load_coef(float2 (&eo)[16], const float* __restrict__ X, int lane) {
  const float2* X2 = reinterpret_cast<const float2*>(X);
#pragma unroll
  for (int u = 0; u < 16; ++u) eo[u] = __ldcs(X2 + lane + 32 * u);
}

// Phase A (decoder/ia_core_coder_imdct.c:171-182,371-376; decoder/impeghd_ifft.c:326-449)
__device__ __forceinline__ void phase_a(const float2 (&ein)[16], const TabA& ta, float* sm, const DeviceTables& T,
                                        int lane) {
  const float norm = 0.0009765625f;  // 2 / 2048
  float zr[16], zi[16];
#pragma unroll
  for (int p = 0; p < 8; ++p) {
    const float4 cs = ta.pre[p];
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int u = 2 * p + h;
      const float c = h ? cs.z : cs.x, s = h ? cs.w : cs.y;
      const float a = fmul(ein[u].x, norm);
      const float b = __shfl_xor_sync(kFull, fmul(ein[15 - u].y, norm), 31);
      zr[u] = fsub(-fmul(a, c), fmul(b, s));
      zi[u] = fsub(fmul(b, c), fmul(a, s));
    }
  }
  float ar[16], ai[16];
#pragma unroll
  for (int t = 0; t < 4; ++t) {
    float x0r = zr[t], x0i = zi[t], x1r = zr[t + 4], x1i = zi[t + 4], x2r = zr[t + 8], x2i = zi[t + 8],
          x3r = zr[t + 12], x3i = zi[t + 12];
    bfly4(x0r, x0i, x1r, x1i, x2r, x2i, x3r, x3i);
    ar[4 * t + 0] = x0r, ai[4 * t + 0] = x0i;
    ar[4 * t + 1] = x1r, ai[4 * t + 1] = x1i;
    ar[4 * t + 2] = x2r, ai[4 * t + 2] = x2i;
    ar[4 * t + 3] = x3r, ai[4 * t + 3] = x3i;
  }
  const int g = ((lane >> 3) & 3) | (((lane >> 1) & 3) << 2) | ((lane & 1) << 4);
  float2* dst = reinterpret_cast<float2*>(sm) + tpos_g(g);
#pragma unroll
  for (int m = 0; m < 4; ++m) {
    float x0r = ar[m], x0i = ai[m], x1r = ar[m + 4], x1i = ai[m + 4], x2r = ar[m + 8], x2i = ai[m + 8],
          x3r = ar[m + 12], x3i = ai[m + 12];
    if (m > 0) {  // pass-1 twiddles are warp-uniform: kernel-parameter constant bank
      rot(x1r, x1i, T.pass1[m - 1][0], T.pass1[m - 1][1]);
      rot(x2r, x2i, T.pass1[m - 1][2], T.pass1[m - 1][3]);
      rot(x3r, x3i, T.pass1[m - 1][4], T.pass1[m - 1][5]);
    }
    bfly4(x0r, x0i, x1r, x1i, x2r, x2i, x3r, x3i);
    dst[m] = make_float2(x0r, x0i);
    dst[m + 4] = make_float2(x1r, x1i);
    dst[m + 8] = make_float2(x2r, x2i);
    dst[m + 12] = make_float2(x3r, x3i);
  }
}

// Phase B (decoder/impeghd_ifft.c:450-837; decoder/ia_core_coder_imdct.c:259-273)
__device__ __forceinline__ void phase_b(const TabB& tb, float* sm, int lane) {
  const int H = lane >> 4, m16 = lane & 15;
  float zr[16], zi[16];
  {
    const float2* src = reinterpret_cast<const float2*>(sm) + 274 * H + m16;  // tpos_g(16 H + v) + m16
#pragma unroll
    for (int v = 0; v < 16; ++v) {
      const float2 t = src[17 * v];
      zr[v] = t.x, zi[v] = t.y;
    }
  }
  __syncwarp();  // everyone has its points: the buffer may now receive the post-twiddled samples
  const bool on = (m16 != 0);
#pragma unroll
  for (int t = 0; t < 4; ++t) {
    rot_sel(zr[4 * t + 1], zi[4 * t + 1], tb.e2a.x, tb.e2a.y, on);
    rot_sel(zr[4 * t + 2], zi[4 * t + 2], tb.e2a.z, tb.e2a.w, on);
    rot_sel(zr[4 * t + 3], zi[4 * t + 3], tb.e2b.x, tb.e2b.y, on);
    bfly4(zr[4 * t], zi[4 * t], zr[4 * t + 1], zi[4 * t + 1], zr[4 * t + 2], zi[4 * t + 2], zr[4 * t + 3],
          zi[4 * t + 3]);
  }
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    if (r == 0) {
      rot_sel(zr[r + 4], zi[r + 4], tb.e3a[r].x, tb.e3a[r].y, on);
      rot_sel(zr[r + 8], zi[r + 8], tb.e3a[r].z, tb.e3a[r].w, on);
      rot_sel(zr[r + 12], zi[r + 12], tb.e3b[r].x, tb.e3b[r].y, on);
    } else {
      rot(zr[r + 4], zi[r + 4], tb.e3a[r].x, tb.e3a[r].y);
      rot(zr[r + 8], zi[r + 8], tb.e3a[r].z, tb.e3a[r].w);
      rot(zr[r + 12], zi[r + 12], tb.e3b[r].x, tb.e3b[r].y);
    }
    bfly4(zr[r], zi[r], zr[r + 4], zi[r + 4], zr[r + 8], zi[r + 8], zr[r + 12], zi[r + 12]);
  }
  // final radix-2 pass between lanes l and l^16: the upper lane rotates its own value (x1' = x1 * W) and
  // sends it, the lower lane sends x0; then y = x0 + x1' (lower) or x0 - x1' (upper).  Post-twiddle follows.
  const unsigned sgn = H ? 0x80000000u : 0u;
#pragma unroll
  for (int v = 0; v < 16; ++v) {
    // rad2_512[p], p = m16 + 16 v: for p >= 128 the entry is (ws, -wc) of p - 128 (ctx.cu radix2_twiddles_512)
    const float4 w = tb.rad[(v & 7) >> 1];
    const float w0 = (v & 1) ? w.z : w.x, w1 = (v & 1) ? w.w : w.y;
    const float wc = (v < 8) ? w0 : w1, ws = (v < 8) ? w1 : -w0;
    float tr = zr[v], ti = zi[v];
    rot(tr, ti, wc, ws);
    tr = H ? tr : zr[v];
    ti = H ? ti : zi[v];
    const float rr = __shfl_xor_sync(kFull, tr, 16);
    const float ri = __shfl_xor_sync(kFull, ti, 16);
    const float yr = fadd(rr, __uint_as_float(__float_as_uint(tr) ^ sgn));
    const float yi = fadd(ri, __uint_as_float(__float_as_uint(ti) ^ sgn));
    const float4 cs = tb.post[v >> 1];
    const float c = (v & 1) ? cs.z : cs.x, s = (v & 1) ? cs.w : cs.y;
    const float xa = -fsub(fmul(yr, c), fmul(yi, s));
    const float xb = -fadd(fmul(yi, c), fmul(yr, s));
    const int k = 256 * H + m16 + 16 * v;
    // lower lanes store the even sample first, upper lanes the odd one: 32 distinct banks per store
    sm[H ? (1023 - 2 * k) : (2 * k)] = H ? xb : xa;
    sm[H ? (2 * k) : (1023 - 2 * k)] = H ? xa : xb;
  }
}

__device__ __forceinline__ void load_overlap(float4 (&ovv)[8], const float* __restrict__ ov, int lane) {
#pragma unroll
  for (int it = 0; it < 4; ++it) {
    const int i4 = 4 * (lane + 32 * it);
    ovv[it] = ldcs4(ov + i4);
    ovv[4 + it] = ldcs4(ov + 1020 - i4);
  }
}

// Phase C (decoder/ia_core_coder_basic_ops.c:124-153,230-295; decoder/ia_core_coder_imdct.c:807-813).
// Position block i4 < 512 and its mirror j4 = 1020 - i4 >= 512:
//   s(i4 + e) = xs[512 + i4 + e]            s(j4 + e) = -xs[515 + i4 - e]      (same float4 P, reversed)
//   new_ov(i4 + e) = -xs[511 - i4 - e]      new_ov(j4 + e) = -xs[508 - i4 + e]  (same float4 Q)
__device__ __forceinline__ void phase_c(const float4 (&ovv)[8], const TabC& tc, const float* sm,
                                        float* __restrict__ ov, float* __restrict__ o, int lane) {
#pragma unroll
  for (int it = 0; it < 4; ++it) {
    const int i4 = 4 * (lane + 32 * it), j4 = 1020 - i4;
    const float4 P = ld4(sm + 512 + i4), Q = ld4(sm + 508 - i4);
    const float4 r_lo = win4(P, tc.f[it], ovv[it], rev4(tc.r[it]));
    const float4 r_hi = win4(neg4(rev4(P)), tc.r[it], ovv[4 + it], rev4(tc.f[it]));
    stcs4(o + i4, r_lo);
    stcs4(o + j4, r_hi);
    st4(ov + i4, neg4(rev4(Q)));
    st4(ov + j4, neg4(Q));
  }
}

// LONG_STOP / STOP_START (decoder/ia_core_coder_basic_ops.c:230-295): flat-zero left part, 128-sample window
// slope around the centre.  Rare, kept out of line so its table loads are never speculated into phase_c.
__device__ __noinline__ void phase_c_stop(const float* sm, float* __restrict__ ov, float* __restrict__ o,
                                          const float* __restrict__ wsh, int lane) {
#pragma unroll 1
  for (int it = 0; it < 4; ++it) {
    const int i4 = 4 * (lane + 32 * it), j4 = 1020 - i4;
    const float4 P = ld4(sm + 512 + i4), Q = ld4(sm + 508 - i4);
    const float4 s_hi = neg4(rev4(P));
    float4 r_lo, r_hi;
    if (i4 < 448) {
      r_lo = ldcs4(ov + i4);
      r_hi = s_hi;  // j4 >= 576
    } else {
      // 448 <= i4 < 512 and 512 <= j4 < 576: both inside the slope
      r_lo = win4(P, ldg4(wsh + i4 - 448), ldcs4(ov + i4), rev4(ldg4(wsh + 572 - i4)));
      r_hi = win4(s_hi, ldg4(wsh + j4 - 448), ldcs4(ov + j4), rev4(ldg4(wsh + 572 - j4)));
    }
    stcs4(o + i4, r_lo);
    stcs4(o + j4, r_hi);
    st4(ov + i4, neg4(rev4(Q)));
    st4(ov + j4, neg4(Q));
  }
}

template <int B, int WARPS, int MIN_BLOCKS>
__global__ void __launch_bounds__(WARPS * 32, MIN_BLOCKS)
    fd_synth_kernel(const float* __restrict__ coef, const uint32_t* __restrict__ side, float* __restrict__ overlap,
                    float* __restrict__ out, int n_units, const __grid_constant__ DeviceTables T) {
  __shared__ __align__(16) float smem[WARPS][B][kBufFloats];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int unit0 = (blockIdx.x * WARPS + warp) * B;
  if (unit0 >= n_units) return;
  const int nb = min(B, n_units - unit0);

  // side words; request the overlap state of every unit of the batch into L2 (consumed in phase C)
  Side sd[B];
  bool fast[B];
  bool any_fast = false, any_slow = false;
#pragma unroll
  for (int b = 0; b < B; ++b) {
    const bool valid = b < nb;
    sd[b] = unpack_side(valid ? __ldg(side + unit0 + b) : 0u);
    fast[b] = valid && sd[b].seq != 2 && sd[b].tkt == 0;
    any_fast |= fast[b];
    any_slow |= valid && !fast[b];
  }
  // Everything this warp will read from HBM is requested up front: unit 0's coefficients as ordinary loads
  // (phase A below), the other units' coefficients and the whole batch's overlap state as bulk L2
  // prefetches (units of a batch are contiguous), so later loads are L2 hits and HBM streams while we compute.
  float2 eo[2][16];
  load_coef(eo[0], coef + size_t(unit0) * 1024, lane);
  if (lane == 0 && nb > 1) {
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(coef + size_t(unit0 + 1) * 1024),
                 "r"((nb - 1) * 4096)
                 : "memory");
  }
  if (lane == 1) {
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(overlap + size_t(unit0) * 1024), "r"(nb * 4096)
                 : "memory");
  }

  if (any_fast) {
    // ---- phase A ----
    {
      TabA ta;
      load_tab_a(ta, T, lane);
#pragma unroll
      for (int b = 0; b < B; ++b) {
        if (b + 1 < B && b + 1 < nb) load_coef(eo[(b + 1) & 1], coef + size_t(unit0 + b + 1) * 1024, lane);
        if (fast[b]) phase_a(eo[b & 1], ta, smem[warp][b], T, lane);
      }
    }
    __syncwarp();
    // ---- phase B ----
    {
      TabB tb;
      load_tab_b(tb, T, lane);
#pragma unroll
      for (int b = 0; b < B; ++b)
        if (fast[b]) phase_b(tb, smem[warp][b], lane);
    }
    __syncwarp();
    // ---- phase C: window entries are reloaded only when the previous-window shape changes ----
    {
      TabC tc;
      int cur_shape = -1;
      float4 ovv[2][8];
#pragma unroll
      for (int b = 0; b < B; ++b) {
        if (b == 0 && fast[0]) load_overlap(ovv[0], overlap + size_t(unit0) * 1024, lane);
        // the next unit's overlap (an L2 hit after the prefetch) is requested before this unit's arithmetic
        if (b + 1 < B && fast[(b + 1) % B])
          load_overlap(ovv[(b + 1) & 1], overlap + size_t(unit0 + b + 1) * 1024, lane);
        if (!fast[b]) continue;
        if (sd[b].shape_prev != cur_shape) {
          cur_shape = sd[b].shape_prev;
          load_tab_c(tc, T.win_long[cur_shape], lane);
        }
        float* ov = overlap + size_t(unit0 + b) * 1024;
        float* o = out + size_t(unit0 + b) * 1024;
        if (sd[b].seq == 0 || sd[b].seq == 1) {
          phase_c(ovv[b & 1], tc, smem[warp][b], ov, o, lane);
        } else {
          phase_c_stop(smem[warp][b], ov, o, T.win_short[cur_shape], lane);
        }
      }
    }
  }
  if (any_slow) {
#pragma unroll 1
    for (int b = 0; b < nb; ++b) {
      const Side s = unpack_side(__ldg(side + unit0 + b));
      if (s.seq != 2 && s.tkt == 0) continue;
      const size_t off = size_t(unit0 + b) * 1024;
      __syncwarp();
      fd_unit_generic(coef + off, overlap + off, out + off, s, smem[warp][b], T, lane);
    }
  }
}

// ---- warp-specialised persistent variant -----------------------------------------------------------
// The three phases become three ROLES that keep their tables in registers for the whole kernel and hand units
// to each other through a ring of shared-memory slots:
//   warp 0   A  loads coefficients (next unit's loads in flight while the current one computes), phase_a
//   warp 1,2 B  phase_b on alternate units (B is the longest phase)
//   warp 3   C  overlap loads, phase_c / the generic path, all global stores
// Hand-over uses named barriers in the producer/consumer form (bar.arrive by the producer warp, bar.sync by
// the consumer warp, 64 participants), one barrier per slot and edge: A->B, B->C and C->A ("slot free").
// CTAs are persistent: CTA c processes units c, c + gridDim.x, ...
constexpr int kSlots = 4;
constexpr int kSlowCap = 256;

// Programmatic dependent launch: consecutive frames are consecutive launches of this kernel on one stream.  Every
// CTA lets the next launch be scheduled right away (its CTAs start as this launch's CTAs retire and SMs free up)
// and loads its role's constant tables before it waits for the previous launch to be complete, so the prologue of
// frame f + 1 runs under the tail of frame f.
__device__ __forceinline__ void grid_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void grid_dependency_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

__device__ __forceinline__ void bar_sync64(int id) { asm volatile("bar.sync %0, 64;" ::"r"(id) : "memory"); }
__device__ __forceinline__ void bar_arrive64(int id) {
  __threadfence_block();  // bar.arrive itself orders nothing: make this warp's shared-memory writes visible first
  asm volatile("bar.arrive %0, 64;" ::"r"(id) : "memory");
}

__device__ __noinline__ void slow_units(const float* __restrict__ coef, const uint32_t* __restrict__ side,
                                        float* __restrict__ overlap, float* __restrict__ out, const int* list, int count,
                                        float* slots, const DeviceTables& T) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // What bounds the mixed-sequence step is the latency of the last short-block unit on its warp, and a CTA rarely has
  // more than one or two such units: the transforms of up to four units run one per warp, then the windowing of
  // each of them (more than a third of a unit, and parallel over its 16 float4 rows) is shared by all four warps.
  for (int base = 0; base < count; base += 4) {
    const int k = base + warp;
    if (k < count) {
      const size_t u = size_t(list[k]);
      const Side sd = unpack_side(__ldg(side + u));
      fd_unit_generic(coef + u * 1024, overlap + u * 1024, out + u * 1024, sd, slots + warp * kBufFloats, T, lane,
                      /*window_later=*/true);
    }
    __syncthreads();
    for (int j = base; j < count && j < base + 4; ++j) {
      const size_t u = size_t(list[j]);
      const Side sd = unpack_side(__ldg(side + u));
      if (sd.seq == 2)
        window_short(slots + (j - base) * kBufFloats, overlap + u * 1024, out + u * 1024, sd, T, lane, warp, 4);
    }
    __syncthreads();  // the slots are rewritten by the next batch
  }
}

#ifndef MPEGHB200_FD_CTAS_PER_SM
#define MPEGHB200_FD_CTAS_PER_SM 3
#endif
template <bool DEFER>
__global__ void __launch_bounds__(128, MPEGHB200_FD_CTAS_PER_SM)
    fd_synth_ws_kernel(const float* __restrict__ coef, const uint32_t* __restrict__ side, float* __restrict__ overlap,
                       float* __restrict__ out, int n_units, const __grid_constant__ DeviceTables T) {
  __shared__ __align__(16) float slots[kSlots][kBufFloats];
  __shared__ int fast_flag[kSlots];
  // Units that need the generic path (EIGHT_SHORT, transform-kernel switching) cost ~5x a fast unit and would stall
  // the A/B/C pipeline, so role C only notes them down; after the CTA's fast pass its four warps do them side by
  // side, one unit per warp, each in its own slot.  (List full: the unit is done in line, as before.)
  __shared__ int slow_list[kSlowCap];
  __shared__ int slow_count;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  grid_launch_dependents();
  if (threadIdx.x == 96) slow_count = 0;  // lane 0 of the role-C warp, its only writer until the final barrier
  const int first = blockIdx.x, stride = gridDim.x;
  const int n_local = first < n_units ? (n_units - first + stride - 1) / stride : 0;
  constexpr int AB = 1, BC = 1 + kSlots, FREE = 1 + 2 * kSlots;  // barrier ids (0 is __syncthreads)

  if (warp == 0) {
    // ---------------- role A ----------------
    TabA ta;
    load_tab_a(ta, T, lane);
    grid_dependency_wait();  // tables are constants; everything below reads what the previous launch wrote
    float2 eo[2][16];
    uint32_t sw[2];
    sw[0] = __ldg(side + first);
    load_coef(eo[0], coef + size_t(first) * 1024, lane);
    auto body = [&](auto cur_c, int i) {
      constexpr int cur = decltype(cur_c)::value;
      const int slot = i % kSlots;
      if (i + 1 < n_local) {
        const size_t un = size_t(first) + size_t(i + 1) * stride;
        sw[cur ^ 1] = __ldg(side + un);
        load_coef(eo[cur ^ 1], coef + un * 1024, lane);
        if (lane == 0)
          asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(overlap + un * 1024), "r"(4096) : "memory");
      }
      if (i >= kSlots) bar_sync64(FREE + slot);
      const Side sd = unpack_side(sw[cur]);
      const bool fast = sd.seq != 2 && sd.tkt == 0;
      if (lane == 0) fast_flag[slot] = fast;
      if (fast) phase_a(eo[cur], ta, slots[slot], T, lane);
      bar_arrive64(AB + slot);
    };
    if (lane == 0)
      asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(overlap + size_t(first) * 1024), "r"(4096)
                   : "memory");
    for (int i = 0; i < n_local; i += 2) {
      body(std::integral_constant<int, 0>{}, i);
      if (i + 1 < n_local) body(std::integral_constant<int, 1>{}, i + 1);
    }
  } else if (warp <= 2) {
    // ---------------- role B ----------------
    TabB tb;
    load_tab_b(tb, T, lane);
    grid_dependency_wait();  // tables are constants; everything below reads what the previous launch wrote
    for (int i = warp - 1; i < n_local; i += 2) {
      const int slot = i % kSlots;
      bar_sync64(AB + slot);
      if (fast_flag[slot]) phase_b(tb, slots[slot], lane);
      bar_arrive64(BC + slot);
    }
  } else {
    // ---------------- role C ----------------
    TabC tc0, tc1;
    load_tab_c(tc0, T.win_long[0], lane);
    load_tab_c(tc1, T.win_long[1], lane);
    grid_dependency_wait();  // tables are constants; everything below reads what the previous launch wrote
    for (int i = 0; i < n_local; ++i) {
      const int slot = i % kSlots;
      const size_t u = size_t(first) + size_t(i) * stride;
      const Side sd = unpack_side(__ldg(side + u));
      float* ov = overlap + u * 1024;
      float* o = out + u * 1024;
      const bool fast = sd.seq != 2 && sd.tkt == 0;
      const bool start_like = (sd.seq == 0 || sd.seq == 1);
      float4 ovv[8];
      if (fast && start_like) load_overlap(ovv, ov, lane);  // in flight while we wait for phase B
      bar_sync64(BC + slot);
      if (fast) {
        if (start_like) {
          if (sd.shape_prev)
            phase_c(ovv, tc1, slots[slot], ov, o, lane);
          else
            phase_c(ovv, tc0, slots[slot], ov, o, lane);
        } else {
          phase_c_stop(slots[slot], ov, o, T.win_short[sd.shape_prev], lane);
        }
      } else if (DEFER && slow_count < kSlowCap) {  // only this warp writes slow_count
        if (lane == 0) slow_list[slow_count] = int(u), slow_count = slow_count + 1;
        __syncwarp();
      } else {
        fd_unit_generic(coef + u * 1024, ov, o, sd, slots[slot], T, lane);
      }
      if (i + kSlots < n_local) bar_arrive64(FREE + slot);
    }
  }
  // ---------------- all roles: the generic-path units of this CTA, one per warp ----------------
  if (DEFER) {
    __syncthreads();
    if (slow_count) slow_units(coef, side, overlap, out, slow_list, slow_count, &slots[0][0], T);
  }
}


// ---- F consecutive frames of every unit in one launch ------------------------------------------------
// A decoder that serves files (or any feed with F frames of a stream at hand) can hand them over together:
// coef / side / out are [F][n_units], the overlap state is one buffer per unit.  A CTA walks its units frame by frame,
// so the unit's overlap never leaves the SM between two ONLY_LONG-family frames: role C's lanes write exactly the
// positions they read (i4 and 1020 - i4), so the new overlap is simply kept in the registers that held the old one and
// only the unit's last frame (or a frame followed by a LONG_STOP / generic-path frame) stores it.  Per channel-frame
// that removes up to 8 KB of the 16 KB of traffic, and the launch's ramp and tail are paid once per F frames.
// A unit that meets a generic-path frame (EIGHT_SHORT, other transform kernels) leaves the pipeline from that frame
// on: role C notes (unit, frame) and one warp finishes the unit's remaining frames in order after the fast pass.
template <bool STORE>
__device__ __forceinline__ void phase_c_carry(float4 (&ovv)[8], const TabC& tc, const float* sm, float* __restrict__ ov,
                                              float* __restrict__ o, int lane) {
#pragma unroll
  for (int it = 0; it < 4; ++it) {
    const int i4 = 4 * (lane + 32 * it), j4 = 1020 - i4;
    const float4 P = ld4(sm + 512 + i4), Q = ld4(sm + 508 - i4);
    const float4 r_lo = win4(P, tc.f[it], ovv[it], rev4(tc.r[it]));
    const float4 r_hi = win4(neg4(rev4(P)), tc.r[it], ovv[4 + it], rev4(tc.f[it]));
    stcs4(o + i4, r_lo);
    stcs4(o + j4, r_hi);
    ovv[it] = neg4(rev4(Q));
    ovv[4 + it] = neg4(Q);
    if (STORE) {
      st4(ov + i4, ovv[it]);
      st4(ov + j4, ovv[4 + it]);
    }
  }
}

__device__ __noinline__ void slow_units_frames(const float* __restrict__ coef, const uint32_t* __restrict__ side,
                                               float* __restrict__ overlap, float* __restrict__ out, const int* list,
                                               int count, float* slots, const DeviceTables& T, int n_units,
                                               int n_frames) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int k = warp; k < count; k += 4) {
    const int e = list[k];
    const size_t u = size_t(e) / size_t(n_frames);
    for (int f = e - int(u) * n_frames; f < n_frames; ++f) {
      const size_t q = size_t(f) * n_units + u;
      fd_unit_generic(coef + q * 1024, overlap + u * 1024, out + q * 1024, unpack_side(__ldg(side + q)),
                      slots + warp * kBufFloats, T, lane);
      __syncwarp();  // the next frame's lanes read overlap positions other lanes of this warp just wrote
    }
  }
}

template <bool DEFER>
__global__ void __launch_bounds__(128, MPEGHB200_FD_CTAS_PER_SM)
    fd_synth_ws_frames_kernel(const float* __restrict__ coef, const uint32_t* __restrict__ side,
                              float* __restrict__ overlap, float* __restrict__ out, int n_units, int n_frames,
                              const __grid_constant__ DeviceTables T) {
  __shared__ __align__(16) float slots[kSlots][kBufFloats];
  __shared__ int fast_flag[kSlots];
  __shared__ int slow_list[kSlowCap];  // unit * n_frames + first generic-path frame
  __shared__ int slow_count;
  __shared__ __align__(128) float coefbuf[3][1024];  // role A's coefficient prefetch (bulk copies), three rows
  __shared__ __align__(8) unsigned long long rowbar[3];  // their mbarriers
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  grid_launch_dependents();
  if (threadIdx.x == 96) slow_count = 0;
  const int first = blockIdx.x, stride = gridDim.x;
  const int n_local = first < n_units ? (n_units - first + stride - 1) / stride : 0;
  const int n_items = n_local * n_frames;  // item i = frame i % F of the CTA's unit number i / F
  constexpr int AB = 1, BC = 1 + kSlots, FREE = 1 + 2 * kSlots;
  auto is_fast = [](const Side& sd) { return sd.seq != 2 && sd.tkt == 0; };

  if (warp == 0) {
    // ---------------- role A ----------------
    TabA ta;
    load_tab_a(ta, T, lane);
    grid_dependency_wait();
    // ONE copy of phase A in the loop (role A heads the pipeline, and 15 % of the kernel's stall samples were
    // instruction-cache misses of the three roles' unrolled code; two copies existed only to give the two register
    // sets of the coefficient prefetch compile-time indices).  The next item's 4 KB of coefficients now arrive by
    // bulk copy (TMA) in one of three shared-memory rows of this warp and are read into ONE register set right before
    // phase A.
    float2 eo[16];
    uint32_t sw_cur = 0, sw_nxt = 0;
    int un = first, fn = 0;  // (unit, frame) of the item whose loads are issued next
    bool left = false;       // the current unit has left the pipeline
    auto issue = [&](int buf) {
      const size_t q = size_t(fn) * n_units + un;
      sw_nxt = __ldg(side + q);
      const float* src = coef + q * 1024;
      if (lane == 0) {
        // one bulk copy (TMA) of the unit's 4 KB row, completion counted in bytes on the row's mbarrier
        const unsigned d = static_cast<unsigned>(__cvta_generic_to_shared(&coefbuf[buf][0]));
        const unsigned b = static_cast<unsigned>(__cvta_generic_to_shared(&rowbar[buf]));
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b), "r"(4096) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(d),
                     "l"(src), "r"(4096), "r"(b)
                     : "memory");
      }
      if (fn == 0 && lane == 0)
        asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(overlap + size_t(un) * 1024), "r"(4096)
                     : "memory");
      if (++fn == n_frames) fn = 0, un += stride;
    };
    // three rows: the copies of the next TWO items are outstanding while phase A works on the current one (with one
    // item ahead a quarter of role A's stall samples were waits for the row)
    if (lane == 0) {
#pragma unroll
      for (int r = 0; r < 3; ++r)
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(static_cast<unsigned>(__cvta_generic_to_shared(&rowbar[r]))), "r"(1) : "memory");
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
    if (n_items > 0) issue(0);
    sw_cur = sw_nxt;
    if (n_items > 1) issue(1);
    int f = 0, buf = 0;
    unsigned par = 0;  // phase parity of the row barriers: flips every third item
#pragma unroll 1
    for (int i = 0; i < n_items; ++i) {
      const int slot = i % kSlots;
      {  // item i's row has landed: its mbarrier completes phase (i / 3) & 1
        const unsigned b = static_cast<unsigned>(__cvta_generic_to_shared(&rowbar[buf]));
        const unsigned parity = par;
        asm volatile(
            "{\n\t"
            ".reg .pred p;\n\t"
            "ROW_WAIT_%=:\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
            "@p bra ROW_DONE_%=;\n\t"
            "bra ROW_WAIT_%=;\n\t"
            "ROW_DONE_%=:\n\t"
            "}\n" ::"r"(b),
            "r"(parity)
            : "memory");
      }
      __syncwarp();  // the previous iteration's reads of the row that is refilled next are done in every lane
      const uint32_t sw_i1 = sw_nxt;                                  // side word of item i + 1
      if (i + 2 < n_items) issue(buf == 0 ? 2 : buf - 1);             // row (i + 2) % 3: read into registers in iteration i - 1
      if (i >= kSlots) bar_sync64(FREE + slot);
      if (f == 0) left = false;
      const bool fast = !left && is_fast(unpack_side(sw_cur));
      left = DEFER && !fast;  // without deferral a generic-path frame is done in line and the unit stays in the pipeline
      if (lane == 0) fast_flag[slot] = fast;
      if (fast) {
        const float2* c2 = reinterpret_cast<const float2*>(&coefbuf[buf][0]);
#pragma unroll
        for (int u = 0; u < 16; ++u) eo[u] = c2[lane + 32 * u];
        phase_a(eo, ta, slots[slot], T, lane);
      }
      bar_arrive64(AB + slot);
      if (++f == n_frames) f = 0;
      sw_cur = sw_i1;
      if (buf == 2) buf = 0, par ^= 1u; else ++buf;
    }
  } else if (warp <= 2) {
    // ---------------- role B ----------------
    TabB tb;
    load_tab_b(tb, T, lane);
    grid_dependency_wait();
    for (int i = warp - 1; i < n_items; i += 2) {
      const int slot = i % kSlots;
      bar_sync64(AB + slot);
      if (fast_flag[slot]) phase_b(tb, slots[slot], lane);
      bar_arrive64(BC + slot);
    }
  } else {
    // ---------------- role C ----------------
    TabC tc0, tc1;
    load_tab_c(tc0, T.win_long[0], lane);
    load_tab_c(tc1, T.win_long[1], lane);
    grid_dependency_wait();
    int u = first, f = 0;
    int mode = 0;       // 0: in the pipeline, 1: handed to the slow pass, 2: generic path in line (list full)
    bool have = false;  // ovv holds the unit's current overlap (the previous frame did not store it)
    float4 ovv[8];
    for (int i = 0; i < n_items; ++i) {
      const int slot = i % kSlots;
      const size_t q = size_t(f) * n_units + u;
      const Side sd = unpack_side(__ldg(side + q));
      float* ov = overlap + size_t(u) * 1024;
      float* o = out + q * 1024;
      if (f == 0) mode = 0, have = false;
      const bool fast = mode == 0 && is_fast(sd);
      const bool start_like = (sd.seq == 0 || sd.seq == 1);
      bool keep = false;  // the next frame takes the overlap from the registers
      if (fast && start_like && f + 1 < n_frames) {
        const Side nx = unpack_side(__ldg(side + q + n_units));
        keep = is_fast(nx) && (nx.seq == 0 || nx.seq == 1);
      }
      if (fast && start_like && !have) load_overlap(ovv, ov, lane);
      bar_sync64(BC + slot);
      if (fast) {
        if (start_like) {
          if (keep) {
            if (sd.shape_prev)
              phase_c_carry<false>(ovv, tc1, slots[slot], ov, o, lane);
            else
              phase_c_carry<false>(ovv, tc0, slots[slot], ov, o, lane);
          } else {
            if (sd.shape_prev)
              phase_c_carry<true>(ovv, tc1, slots[slot], ov, o, lane);
            else
              phase_c_carry<true>(ovv, tc0, slots[slot], ov, o, lane);
          }
        } else {
          phase_c_stop(slots[slot], ov, o, T.win_short[sd.shape_prev], lane);
          __syncwarp();  // its overlap stores are read back by other lanes' loads of the next frame
        }
      } else {
        if (!DEFER) {
          fd_unit_generic(coef + q * 1024, ov, o, sd, slots[slot], T, lane);
          __syncwarp();
        } else if (mode == 0) {
          if (slow_count < kSlowCap) {  // only this warp writes slow_count
            if (lane == 0) slow_list[slow_count] = u * n_frames + f, slow_count = slow_count + 1;
            __syncwarp();
            mode = 1;
          } else {
            mode = 2;
          }
        }
        if (DEFER && mode == 2) {
          fd_unit_generic(coef + q * 1024, ov, o, sd, slots[slot], T, lane);
          __syncwarp();
        }
      }
      have = keep;
      if (i + kSlots < n_items) bar_arrive64(FREE + slot);
      if (++f == n_frames) f = 0, u += stride;
    }
  }
  __syncthreads();
  if (slow_count) slow_units_frames(coef, side, overlap, out, slow_list, slow_count, &slots[0][0], T, n_units, n_frames);
}

}  // namespace

template <int B, int WARPS, int MIN_BLOCKS>
static cudaError_t launch_variant(const float* d_coef, const uint32_t* d_side, float* d_overlap, float* d_out,
                                  int n_units, const DeviceTables& T, cudaStream_t stream) {
  auto kern = fd_synth_kernel<B, WARPS, MIN_BLOCKS>;
  cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
  const int per_cta = B * WARPS;
  const int grid = (n_units + per_cta - 1) / per_cta;
  kern<<<grid, WARPS * 32, 0, stream>>>(d_coef, d_side, d_overlap, d_out, n_units, T);
  return cudaGetLastError();
}

mpeghb200_status fd_synth_launch(mpeghb200_ctx* ctx, const float* d_coef, const uint32_t* d_side, float* d_overlap,
                                 float* d_out, int n_units, cudaStream_t stream) {
  if (n_units == 0) return MPEGHB200_OK;
  // Launch shape: units per warp, warps per CTA, CTAs per SM the register budget is sized for.  Variant 0
  // is the default; the others exist for tuning sweeps (MPEGHB200_FD_VARIANT, read at context creation).
  cudaError_t e;
  switch (ctx->fd_variant) {
    case 0: e = launch_variant<3, 1, 14>(d_coef, d_side, d_overlap, d_out, n_units, ctx->tables, stream); break;
    case 1: e = launch_variant<3, 2, 7>(d_coef, d_side, d_overlap, d_out, n_units, ctx->tables, stream); break;
    case 2: e = launch_variant<4, 1, 12>(d_coef, d_side, d_overlap, d_out, n_units, ctx->tables, stream); break;
    case 3: e = launch_variant<2, 1, 16>(d_coef, d_side, d_overlap, d_out, n_units, ctx->tables, stream); break;
    case 4: e = launch_variant<2, 2, 10>(d_coef, d_side, d_overlap, d_out, n_units, ctx->tables, stream); break;
    case 5: e = launch_variant<1, 4, 8>(d_coef, d_side, d_overlap, d_out, n_units, ctx->tables, stream); break;
    case 6: e = launch_variant<6, 1, 7>(d_coef, d_side, d_overlap, d_out, n_units, ctx->tables, stream); break;
    default:
    case 7:
    case 8: {
      const int ctas = ctx->sm_count * MPEGHB200_FD_CTAS_PER_SM;
      cudaLaunchConfig_t cfg{};
      cfg.gridDim = dim3(n_units < ctas ? n_units : ctas), cfg.blockDim = dim3(128), cfg.stream = stream;
      cudaLaunchAttribute attr[1];
      attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
      attr[0].val.programmaticStreamSerializationAllowed = 1;
      cfg.attrs = attr, cfg.numAttrs = ctx->fd_variant == 9 ? 0 : 1;  // variant 9: plain launches, for comparison
      if (ctx->fd_variant == 8)
        e = cudaLaunchKernelEx(&cfg, fd_synth_ws_kernel<false>, d_coef, d_side, d_overlap, d_out, n_units, ctx->tables);
      else
        e = cudaLaunchKernelEx(&cfg, fd_synth_ws_kernel<true>, d_coef, d_side, d_overlap, d_out, n_units, ctx->tables);
      if (e == cudaSuccess) e = cudaGetLastError();
      break;
    }
  }
  ctx->launches.fetch_add(1);
  MB_CUDA(ctx, e);
  return MPEGHB200_OK;
}


mpeghb200_status fd_synth_frames_launch(mpeghb200_ctx* ctx, const float* d_coef, const uint32_t* d_side,
                                        float* d_overlap, float* d_out, int n_units, int n_frames, cudaStream_t stream) {
  if (n_units == 0 || n_frames == 0) return MPEGHB200_OK;
  if (n_frames == 1) return fd_synth_launch(ctx, d_coef, d_side, d_overlap, d_out, n_units, stream);
  const int ctas = ctx->sm_count * MPEGHB200_FD_CTAS_PER_SM;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(n_units < ctas ? n_units : ctas), cfg.blockDim = dim3(128), cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr, cfg.numAttrs = 1;
  // generic-path frames (EIGHT_SHORT, other transform kernels) in line: the unit's later ONLY_LONG-family frames stay on
  // the fast path.  MPEGHB200_FD_VARIANT=10 selects the deferring form (a unit leaves the pipeline at its first such frame).
  cudaError_t e = ctx->fd_variant == 10
                      ? cudaLaunchKernelEx(&cfg, fd_synth_ws_frames_kernel<true>, d_coef, d_side, d_overlap, d_out, n_units,
                                           n_frames, ctx->tables)
                      : cudaLaunchKernelEx(&cfg, fd_synth_ws_frames_kernel<false>, d_coef, d_side, d_overlap, d_out,
                                           n_units, n_frames, ctx->tables);
  if (e == cudaSuccess) e = cudaGetLastError();
  ctx->launches.fetch_add(1);
  MB_CUDA(ctx, e);
  return MPEGHB200_OK;
}

}  // namespace mpeghb200
