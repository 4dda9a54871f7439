// tail.cu -- end of the chain: loudness gain, look-ahead peak limiter, PCM packer.
//
// Stands in for impegh_dec_ln_pl_process (decoder/ia_core_coder_decode_main.c:1496), impegh_dec_peak_limiter
// (:1429), impeghd_peak_limiter_process (decoder/impeghd_peak_limiter.c:190) and ia_core_coder_samples_sat
// (decoder/ia_core_coder_decode_main.c:209).
//
// Limiter.  The reference walks the frame sample by sample on an interleaved copy.  Only one part of that is
// inherently serial: the two-variable gain smoother (gain_modified, pre_smoothed_gain).  Everything around
// it is data parallel and is done by the whole CTA (one CTA per stream):
//   1. m[i]   = max_c |x[c][i]|                       (channel maximum, coalesced float4 reads per channel)
//   2. M[i]   = max(m[i-A..i])                        (A = attack_time_samples = 240 @48 kHz; the reference's
//                                                      max_idx bookkeeping maintains exactly this sliding
//                                                      maximum, see oracle/oracle_tail.c) by window doubling
//   3. g0[i]  = thr >= M ? 1 : thr / M                (IEEE division)
//   4. g[i]   = smoother(g0[i])                       one thread, 1024 steps; skipped when the frame cannot
//                                                      move the gain (all g0 == 1 and the state sits at 1)
//   5. y[c][i] = clip(x[c][i-A] * g[i])               (delay line = last A samples of the previous frame)
// State per stream, time order (oldest first): {gain_modified, pre_smoothed_gain, -, -, max_hist[A],
// delay[n_ch][A]}.  Bytes per stream-frame: read n_ch*4096, write n_ch*4096 (+ state 2*(1+n_ch)*A*4).
//
// PCM packer.  planar float -> interleaved little-endian 16/24/32-bit with the reference's clamp-then-truncate
// and its start-up delay trimming; staged through shared memory so both sides are coalesced.
#include <cmath>
#include <cstring>

#include "common.h"

namespace mpeghb200 {
namespace {

constexpr int kMaxAttack = 240;
constexpr int kFrame = 1024;

struct LimParams {
  int n_ch, attack, limiter_on, stride;  // stride = floats of state per stream
  float thr, ac, rc;
  int fast_release;  // smoother: blocks of eight unit gains take the three-operation release chain (bit-identical)
};

// Steps 1 - 4 for the CTA's stream: leaves the per-sample gains in g[] (< 0 = pass through) and advances the
// smoother / history part of the state.  m, m2: [kMaxAttack + kFrame + 16] each, g: [kFrame], all in shared memory.
// NT threads per stream (128 or 256): a thread owns kFrame / 4 / NT float4 positions of every channel.
template <int NT>
__device__ __forceinline__ void limiter_gains(const float* __restrict__ x, float* __restrict__ st, const LimParams& P,
                                              float lg, bool has_lg, float* m, float* m2, float* g) {
  constexpr int kLimThreads = NT;
  constexpr int kVec = kFrame / 4 / NT;
  const int tid = threadIdx.x;
  const int A = P.attack;
  float* st_hist = st + 4;
  const float psg0 = st[1];
  const bool bypass = (!P.limiter_on) && (psg0 >= 1.0f);

  if (!bypass) {
    // 1. channel maxima (fmax semantics of the reference: NaN operands are dropped)
    float4 acc[kVec];
#pragma unroll
    for (int h = 0; h < kVec; ++h) acc[h] = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 12
    for (int c = 0; c < P.n_ch; ++c) {
#pragma unroll
      for (int h = 0; h < kVec; ++h) {
        float4 v = __ldg(reinterpret_cast<const float4*>(x + size_t(c) * kFrame) + tid + kLimThreads * h);
        if (has_lg) v = make_float4(fmul(v.x, lg), fmul(v.y, lg), fmul(v.z, lg), fmul(v.w, lg));
        acc[h].x = fmaxf(acc[h].x, fabsf(v.x));
        acc[h].y = fmaxf(acc[h].y, fabsf(v.y));
        acc[h].z = fmaxf(acc[h].z, fabsf(v.z));
        acc[h].w = fmaxf(acc[h].w, fabsf(v.w));
      }
    }
    for (int i = tid; i < A; i += kLimThreads) m[i] = st_hist[i];
#pragma unroll
    for (int h = 0; h < kVec; ++h) {
      const int i = 4 * (tid + kLimThreads * h);
      m[A + i] = acc[h].x, m[A + i + 1] = acc[h].y, m[A + i + 2] = acc[h].z, m[A + i + 3] = acc[h].w;
    }
    for (int i = tid; i < 16; i += kLimThreads) m[A + kFrame + i] = 0.f, m2[A + kFrame + i] = 0.f;
    __syncthreads();
    // new history = channel maxima of the last A samples of this frame
    for (int i = tid; i < A; i += kLimThreads) st_hist[i] = m[kFrame + i];
    // 2. sliding maximum over A+1 values ending at sample i: W[i] = max(m[i .. i+A]) in history-first indexing.
    //    Window doubling: after pass k, buf[i] = max(m[i .. i+2^k-1]).  A+1 is covered by two windows of length
    //    L = 2^floor(log2(A+1)).
    const int W = A + 1;
    int L = 1;
    float* src = m;
    float* dst = m2;
    const int total = A + kFrame;
    while (2 * L <= W) {
      for (int i = tid; i < total; i += kLimThreads) {
        const int j = i + L;
        dst[i] = (j < total) ? fmaxf(src[i], src[j]) : src[i];
      }
      __syncthreads();
      float* t = src;
      src = dst;
      dst = t;
      L *= 2;
    }
    // 3. raw gain
    int all_one = 1;
    for (int i = tid; i < kFrame; i += kLimThreads) {
      const float mx = fmaxf(src[i], src[i + W - L]);
      const float g0 = (P.thr >= mx) ? 1.0f : __fdiv_rn(P.thr, mx);
      g[i] = g0;
      all_one &= (g0 == 1.0f);
    }
    const int quiet = __syncthreads_and(all_one);
    // 4. the serial smoother (decoder/impeghd_peak_limiter.c:248-270)
    if (!(quiet && psg0 == 1.0f)) {
      if (tid == 0) {
        // branch-free, eight gains at a time through registers: the chain psg -> gm -> psg is the only thing
        // that has to wait (about ten dependent operations per sample); loads, stores and selects ride along
        float gm = st[0], psg = psg0;
        const float rc = P.rc, ac = P.ac;
        for (int i0 = 0; i0 < kFrame; i0 += 8) {
          const float4 a = *reinterpret_cast<const float4*>(g + i0);
          const float4 b = *reinterpret_cast<const float4*>(g + i0 + 4);
          float gv[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
          // Eight unit gains with the state at or below 1: every sample takes gain >= pre_smoothed_gain, i.e.
          // gain_modified = 1, release coefficient, psg = 1 + rc * (psg - 1) -- and psg <= 1 is kept by each step
          // (psg - 1 <= 0, so is its product with rc >= 0, and 1 plus that rounds to at most 1).  The test on the
          // gains does not depend on the running state, so only the three-operation chain is waited for.
          const bool ones = P.fast_release && fminf(fminf(fminf(a.x, a.y), fminf(a.z, a.w)), fminf(fminf(b.x, b.y), fminf(b.z, b.w))) == 1.0f &&
                            fmaxf(fmaxf(fmaxf(a.x, a.y), fmaxf(a.z, a.w)), fmaxf(fmaxf(b.x, b.y), fmaxf(b.z, b.w))) == 1.0f;
          if (ones && psg <= 1.0f && rc >= 0.0f) {
#pragma unroll
            for (int e = 0; e < 8; ++e) {
              psg = fadd(1.0f, fmul(rc, fsub(psg, 1.0f)));
              gv[e] = psg;
            }
            gm = 1.0f;
          } else {
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            const float gain = gv[e];
            const float cand = fmul(fsub(gain, fmul(0.1f, psg)), 1.11111111f);
            gm = (gain < psg) ? fminf(gm, cand) : gain;
            const bool up = gm >= psg;
            const float r = fadd(gm, fmul(up ? rc : ac, fsub(psg, gm)));
            psg = up ? r : fmaxf(r, gain);
            gv[e] = psg;
          }
          }
          *reinterpret_cast<float4*>(g + i0) = make_float4(gv[0], gv[1], gv[2], gv[3]);
          *reinterpret_cast<float4*>(g + i0 + 4) = make_float4(gv[4], gv[5], gv[6], gv[7]);
        }
        st[0] = gm;
        st[1] = psg;
      }
    } else if (tid == 0) {
      st[0] = 1.0f;  // gain_modified = gain = 1 on every sample; pre_smoothed_gain stays 1
    }
    __syncthreads();
  } else {
    for (int i = tid; i < kFrame; i += kLimThreads) g[i] = -1.0f;
    __syncthreads();
  }

}

template <int NT>
__global__ void __launch_bounds__(NT)
    limiter_kernel(const float* __restrict__ in, float* __restrict__ out, float* __restrict__ state,
                   const float* __restrict__ loud_gain, const LimParams P) {
  constexpr int kLimThreads = NT;
  constexpr int kVec = kFrame / 4 / NT;
  pdl_prologue();
  __shared__ __align__(16) float m[kMaxAttack + kFrame + 16];   // channel maxima incl. history / sliding max
  __shared__ __align__(16) float m2[kMaxAttack + kFrame + 16];  // ping-pong partner
  __shared__ __align__(16) float g[kFrame];                     // gains; < 0 = pass through
  const int tid = threadIdx.x;
  const int A = P.attack;
  const size_t sidx = blockIdx.x;
  const float* x = in + sidx * size_t(P.n_ch) * kFrame;
  float* y = out + sidx * size_t(P.n_ch) * kFrame;
  float* st = state + sidx * size_t(P.stride);
  float* st_delay = st + 4 + A;
  const float lg = loud_gain ? __ldg(loud_gain + sidx) : 1.0f;
  const bool has_lg = loud_gain != nullptr;
  limiter_gains<NT>(x, st, P, lg, has_lg, m, m2, g);

  // 5. delayed output: sample i of the frame leaves as input sample i - A (this frame for i >= A, the delay line
  //    before), times the gain, clipped.  Straight from global memory (the frame is L2-resident since step 1, A is
  //    a multiple of 4 so every float4 stays aligned): no staging, no barriers, loads of several channels in flight.
  const float thr = P.thr;
  float4 gv[kVec];
#pragma unroll
  for (int h = 0; h < kVec; ++h) gv[h] = *reinterpret_cast<const float4*>(g + 4 * (tid + kLimThreads * h));
#pragma unroll 8
  for (int c = 0; c < P.n_ch; ++c) {
    float* dl = st_delay + size_t(c) * A;
    const float* xc = x + size_t(c) * kFrame;
    float4 xv[kVec];
#pragma unroll
    for (int h = 0; h < kVec; ++h) {
      const int i = 4 * (tid + kLimThreads * h);
      if (i >= A) {
        xv[h] = __ldg(reinterpret_cast<const float4*>(xc + i - A));
        if (has_lg) xv[h] = make_float4(fmul(xv[h].x, lg), fmul(xv[h].y, lg), fmul(xv[h].z, lg), fmul(xv[h].w, lg));
      } else {
        xv[h] = *reinterpret_cast<const float4*>(dl + i);  // already carries the loudness gain
      }
    }
    // the last A samples of the frame are the next frame's delay line (same thread that just read dl + 4 tid)
    if (4 * tid < A) {
      float4 t = __ldg(reinterpret_cast<const float4*>(xc + kFrame - A + 4 * tid));
      if (has_lg) t = make_float4(fmul(t.x, lg), fmul(t.y, lg), fmul(t.z, lg), fmul(t.w, lg));
      *reinterpret_cast<float4*>(dl + 4 * tid) = t;
    }
#pragma unroll
    for (int h = 0; h < kVec; ++h) {
      const int i = 4 * (tid + kLimThreads * h);
      float r[4] = {xv[h].x, xv[h].y, xv[h].z, xv[h].w};
      const float gg[4] = {gv[h].x, gv[h].y, gv[h].z, gv[h].w};
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        if (gg[e] >= 0.0f) {
          float t = fmul(r[e], gg[e]);
          if (thr < t)
            t = thr;
          else if (t < -thr)
            t = -thr;
          r[e] = t;
        }
      }
      __stcs(reinterpret_cast<float4*>(y + size_t(c) * kFrame + i), make_float4(r[0], r[1], r[2], r[3]));
    }
  }
}

// The same limiter for launches of few streams (at most four CTAs per SM): with so few warps on an SM the output
// phase above is bound by the latency of its channel loads, which the compiler keeps one behind the other (32
// registers).  Here the loads of CB channels are issued before the first one is used, and the delay line is renewed
// after the output loop (behind a barrier: other threads still read the old one) so that no store sits between them.
template <int CB>
__global__ void __launch_bounds__(256, 4)
    limiter_deep_kernel(const float* __restrict__ in, float* __restrict__ out, float* __restrict__ state,
                        const float* __restrict__ loud_gain, const LimParams P) {
  constexpr int NT = 256;
  pdl_prologue();
  __shared__ __align__(16) float m[kMaxAttack + kFrame + 16];
  __shared__ __align__(16) float m2[kMaxAttack + kFrame + 16];
  __shared__ __align__(16) float g[kFrame];
  const int tid = threadIdx.x;
  const int A = P.attack;
  const size_t sidx = blockIdx.x;
  const float* x = in + sidx * size_t(P.n_ch) * kFrame;
  float* y = out + sidx * size_t(P.n_ch) * kFrame;
  float* st = state + sidx * size_t(P.stride);
  float* st_delay = st + 4 + A;
  const float lg = loud_gain ? __ldg(loud_gain + sidx) : 1.0f;
  const bool has_lg = loud_gain != nullptr;
  limiter_gains<NT>(x, st, P, lg, has_lg, m, m2, g);
  const float thr = P.thr;
  const int i = 4 * tid;
  const float4 gq = *reinterpret_cast<const float4*>(g + i);
  const float gg[4] = {gq.x, gq.y, gq.z, gq.w};
  for (int c0 = 0; c0 < P.n_ch; c0 += CB) {
    float4 xv[CB];
#pragma unroll
    for (int q = 0; q < CB; ++q) {
      const int c = c0 + q;
      if (c < P.n_ch) {
        if (i >= A)
          xv[q] = __ldg(reinterpret_cast<const float4*>(x + size_t(c) * kFrame + i - A));
        else
          xv[q] = *reinterpret_cast<const float4*>(st_delay + size_t(c) * A + i);  // already carries the loudness gain
      }
    }
#pragma unroll
    for (int q = 0; q < CB; ++q) {
      const int c = c0 + q;
      if (c < P.n_ch) {
        float4 v = xv[q];
        if (has_lg && i >= A) v = make_float4(fmul(v.x, lg), fmul(v.y, lg), fmul(v.z, lg), fmul(v.w, lg));
        float r[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          if (gg[e] >= 0.0f) {
            float t = fmul(r[e], gg[e]);
            if (thr < t)
              t = thr;
            else if (t < -thr)
              t = -thr;
            r[e] = t;
          }
        }
        __stcs(reinterpret_cast<float4*>(y + size_t(c) * kFrame + i), make_float4(r[0], r[1], r[2], r[3]));
      }
    }
  }
  __syncthreads();  // every read of the old delay line is done
  const int a4 = A / 4;
  for (int k = tid; k < P.n_ch * a4; k += NT) {
    const int c = k / a4, j = 4 * (k - c * a4);
    float4 t = __ldg(reinterpret_cast<const float4*>(x + size_t(c) * kFrame + kFrame - A + j));
    if (has_lg) t = make_float4(fmul(t.x, lg), fmul(t.y, lg), fmul(t.z, lg), fmul(t.w, lg));
    *reinterpret_cast<float4*>(st_delay + size_t(c) * A + j) = t;
  }
}

__global__ void limiter_state_init_kernel(float* state, int stride, int n_streams) {
  const size_t n = size_t(stride) * n_streams;
  for (size_t i = blockIdx.x * size_t(blockDim.x) + threadIdx.x; i < n; i += size_t(gridDim.x) * blockDim.x) {
    const int k = int(i % stride);
    state[i] = (k < 2) ? 1.0f : 0.0f;
  }
}

// ---- PCM packer ------------------------------------------------------------------------------------
constexpr int kPackSamples = 256;  // samples per CTA
struct PackParams {
  int n_ch, bits;
  int n_samples, ch_stride;  // samples per channel in this call (a multiple of kPackSamples); floats between channels
  unsigned char pos[64];  // planar channel feeding interleave slot ch (the reference's channel_pos[])
};

__device__ __forceinline__ int sat_cast(float v) {
  // x86 cvttss2si semantics for out-of-range / NaN: 0x80000000
  if (!(v < 2147483648.0f) || v < -2147483648.0f) return int(0x80000000u);
  return __float2int_rz(v);
}

__global__ void __launch_bounds__(kPackSamples)
    pcm_pack_kernel(const float* __restrict__ in, unsigned char* __restrict__ out, int* __restrict__ delay,
                    int* __restrict__ out_bytes, const __grid_constant__ PackParams P) {
  pdl_prologue();
  extern __shared__ __align__(16) unsigned char tile[];  // [kPackSamples][n_ch][bytes]
  const int tid = threadIdx.x;
  const int chunk = blockIdx.x, sidx = blockIdx.y;
  const int bps = P.bits >> 3;
  const int frame_bytes = P.n_samples * P.n_ch * bps;
  const int d = delay ? delay[sidx] : 0;
  const int s = chunk * kPackSamples + tid;
  // row offsets of the interleave slots, once per CTA (indexing P.pos[] per load made the compiler copy the
  // table to local memory and redo 64-bit address arithmetic for every value: ncu, 12 instructions per load)
  __shared__ unsigned row_off[64];
  if (tid < P.n_ch) row_off[tid] = unsigned(P.pos[tid]) * unsigned(P.ch_stride);
  __syncthreads();
  const float* xs = in + size_t(sidx) * P.n_ch * P.ch_stride + s;
  constexpr int kBatch = 8;
  for (int ch0 = 0; ch0 < P.n_ch; ch0 += kBatch) {
    float vin[kBatch];  // eight channel loads in flight before the first conversion
#pragma unroll
    for (int q = 0; q < kBatch; ++q) vin[q] = (ch0 + q < P.n_ch) ? __ldcs(xs + row_off[ch0 + q]) : 0.0f;
#pragma unroll
    for (int q = 0; q < kBatch; ++q) {
      const int ch = ch0 + q;
      if (ch >= P.n_ch) break;
      float v = vin[q];
      int w;
      if (P.bits == 24) {
        v = fmul(v, 256.0f);
        if (v < -8388608.0f)
          v = -8388608.0f;
        else if (8388607.0f < v)
          v = 8388607.0f;
        w = sat_cast(v);
      } else if (P.bits == 16) {
        if (v < -32768.0f)
          v = -32768.0f;
        else if (32767.0f < v)
          v = 32767.0f;
        w = sat_cast(v);
      } else {
        v = fmul(v, 65536.0f);
        if (v < -2147483647.0f) v = -2147483647.0f;
        // reference build quirk (see oracle/oracle_tail.c): above 2^31 -> INT32_MAX, exactly 2^31 -> INT32_MIN
        w = (2147483647.0f < v) ? 0x7fffffff : sat_cast(v);
      }
      unsigned char* t = tile + (size_t(tid) * P.n_ch + ch) * bps;
      t[0] = w & 0xff;
      t[1] = (w >> 8) & 0xff;
      if (bps > 2) t[2] = (w >> 16) & 0xff;
      if (bps > 3) t[3] = (w >> 24) & 0xff;
    }
  }
  __syncthreads();
  // tile -> global; sample s lands at byte (s - d) * n_ch * bps (start-up delay trimming)
  const int tile_bytes = kPackSamples * P.n_ch * bps;
  const long long dst0 = (long long)(chunk * kPackSamples - d) * P.n_ch * bps;
  unsigned char* o = out + size_t(sidx) * frame_bytes;
  if (dst0 >= 0 && (dst0 & 15) == 0 && (tile_bytes & 15) == 0) {
    const uint4* t4 = reinterpret_cast<const uint4*>(tile);
    uint4* o4 = reinterpret_cast<uint4*>(o + dst0);
    for (int i = tid; i < tile_bytes / 16; i += kPackSamples) __stcs(o4 + i, t4[i]);
  } else {
    for (int i = tid; i < tile_bytes; i += kPackSamples) {
      const long long p = dst0 + i;
      if (p >= 0) o[p] = tile[i];
    }
  }
  if (chunk == 0 && tid == 0) {
    if (out_bytes) out_bytes[sidx] = (d <= P.n_samples) ? (P.n_samples - d) * P.n_ch * bps : 0;
  }
}

// The same for channel counts that are a multiple of four (22.2, 7.1.4, 5.1.2 ...), sample size at compile time: four
// channels of a sample are converted in registers and leave as 2 / 3 / 4 whole 32-bit words (the generic kernel stores
// byte by byte and spends 46 instructions per value, ncu: issue-active 67 %, LSU wavefronts 66 %, DRAM 33 %).
template <int BITS>
__device__ __forceinline__ unsigned pcm_convert_t(float v) {
  if (BITS == 24) {
    v = fmul(v, 256.0f);
    v = v < -8388608.0f ? -8388608.0f : (8388607.0f < v ? 8388607.0f : v);
    return static_cast<unsigned>(sat_cast(v));
  }
  if (BITS == 16) {
    v = v < -32768.0f ? -32768.0f : (32767.0f < v ? 32767.0f : v);
    return static_cast<unsigned>(sat_cast(v));
  }
  v = fmul(v, 65536.0f);
  if (v < -2147483647.0f) v = -2147483647.0f;
  return (2147483647.0f < v) ? 0x7fffffffu : static_cast<unsigned>(sat_cast(v));
}

template <int BITS>
__global__ void __launch_bounds__(kPackSamples)
    pcm_pack4_kernel(const float* __restrict__ in, unsigned char* __restrict__ out, int* __restrict__ delay,
                     int* __restrict__ out_bytes, const __grid_constant__ PackParams P) {
  pdl_prologue();
  extern __shared__ __align__(16) unsigned char tile[];  // [kPackSamples][n_ch][bytes]
  constexpr int BPS = BITS / 8;
  const int tid = threadIdx.x;
  const int chunk = blockIdx.x, sidx = blockIdx.y;
  const int rec = P.n_ch * BPS;
  const int frame_bytes = P.n_samples * rec;
  const int d = delay ? delay[sidx] : 0;
  const int s = chunk * kPackSamples + tid;
  __shared__ unsigned row_off[64];  // as in pcm_pack_kernel
  if (tid < P.n_ch) row_off[tid] = unsigned(P.pos[tid]) * unsigned(P.ch_stride);
  __syncthreads();
  const float* xs = in + size_t(sidx) * P.n_ch * P.ch_stride + s;
  unsigned* tw = reinterpret_cast<unsigned*>(tile + size_t(tid) * rec);
  constexpr int kB = 24;  // channel loads in flight before the first conversion (n_ch % 4 == 0: whole groups of four)
  for (int ch0 = 0; ch0 < P.n_ch; ch0 += kB) {
    float vin[kB];
#pragma unroll
    for (int q = 0; q < kB; ++q) vin[q] = (ch0 + q < P.n_ch) ? __ldcs(xs + row_off[ch0 + q]) : 0.0f;
#pragma unroll
    for (int g = 0; g < kB / 4; ++g) {
      if (ch0 + 4 * g >= P.n_ch) break;
      const unsigned w0 = pcm_convert_t<BITS>(vin[4 * g]), w1 = pcm_convert_t<BITS>(vin[4 * g + 1]),
                     w2 = pcm_convert_t<BITS>(vin[4 * g + 2]), w3 = pcm_convert_t<BITS>(vin[4 * g + 3]);
      if (BITS == 24) {
        tw[0] = (w0 & 0xffffffu) | (w1 << 24);
        tw[1] = ((w1 >> 8) & 0xffffu) | (w2 << 16);
        tw[2] = ((w2 >> 16) & 0xffu) | (w3 << 8);
      } else if (BITS == 16) {
        tw[0] = (w0 & 0xffffu) | (w1 << 16);
        tw[1] = (w2 & 0xffffu) | (w3 << 16);
      } else {
        tw[0] = w0, tw[1] = w1, tw[2] = w2, tw[3] = w3;
      }
      tw += BPS;
    }
  }
  __syncthreads();
  const int tile_bytes = kPackSamples * rec;
  const long long dst0 = (long long)(chunk * kPackSamples - d) * rec;
  unsigned char* o = out + size_t(sidx) * frame_bytes;
  if (dst0 >= 0 && (dst0 & 15) == 0 && (tile_bytes & 15) == 0) {
    const uint4* t4 = reinterpret_cast<const uint4*>(tile);
    uint4* o4 = reinterpret_cast<uint4*>(o + dst0);
    for (int i = tid; i < tile_bytes / 16; i += kPackSamples) __stcs(o4 + i, t4[i]);
  } else {
    for (int i = tid; i < tile_bytes; i += kPackSamples) {
      const long long p = dst0 + i;
      if (p >= 0) o[p] = tile[i];
    }
  }
  if (chunk == 0 && tid == 0) {
    if (out_bytes) out_bytes[sidx] = (d <= P.n_samples) ? (P.n_samples - d) * rec : 0;
  }
}

// ---- limiter + PCM packer as a gains kernel and an apply-and-pack kernel -------------------------------
// The chain's tail without the planar fp32 round trip between limiter and packer, and without tying the data-parallel
// part of the limiter (delayed samples x gain, clip) to the one CTA per stream that the serial smoother needs:
//   limiter_gains_kernel       one CTA per stream: steps 1 - 4 (channel maxima, sliding maximum, raw gain, smoother),
//                              writes the 1024 gains, hands the OLD delay line to the apply kernel and stores the new
//   limiter_apply_pack_kernel  four CTAs per stream (256 samples each): delayed sample x gain, clip, PCM conversion,
//                              byte tile, coalesced store -- the packer's structure with the limiter's step 5 in front
// Same arithmetic as limiter_kernel + pcm_pack_kernel, node for node.
template <int NT>
__global__ void __launch_bounds__(NT)
    limiter_gains_kernel(const float* __restrict__ in, float* __restrict__ state, const float* __restrict__ loud_gain,
                         float* __restrict__ gains, float* __restrict__ dl_old, const LimParams P) {
  pdl_prologue();
  __shared__ __align__(16) float m[kMaxAttack + kFrame + 16];
  __shared__ __align__(16) float m2[kMaxAttack + kFrame + 16];
  __shared__ __align__(16) float g[kFrame];
  const int tid = threadIdx.x;
  const int A = P.attack;
  const size_t sidx = blockIdx.x;
  const float* x = in + sidx * size_t(P.n_ch) * kFrame;
  float* st = state + sidx * size_t(P.stride);
  float* st_delay = st + 4 + A;
  const float lg = loud_gain ? __ldg(loud_gain + sidx) : 1.0f;
  const bool has_lg = loud_gain != nullptr;
  constexpr int kLimThreads = NT;
  limiter_gains<NT>(x, st, P, lg, has_lg, m, m2, g);
  float4* go = reinterpret_cast<float4*>(gains + sidx * kFrame);
  for (int i = tid; i < kFrame / 4; i += kLimThreads) go[i] = reinterpret_cast<const float4*>(g)[i];
  // delay line: the old one goes to the apply kernel, the last A samples of this frame (with the loudness gain)
  // become the new one; A is a multiple of 4
  float* dlo = dl_old + sidx * size_t(P.n_ch) * A;
  const int a4 = A / 4;
  for (int i = tid; i < P.n_ch * a4; i += kLimThreads) {
    const int c = i / a4, j = 4 * (i - c * a4);
    float4* d = reinterpret_cast<float4*>(st_delay + size_t(c) * A + j);
    *reinterpret_cast<float4*>(dlo + size_t(c) * A + j) = *d;
    float4 t = __ldg(reinterpret_cast<const float4*>(x + size_t(c) * kFrame + kFrame - A + j));
    if (has_lg) t = make_float4(fmul(t.x, lg), fmul(t.y, lg), fmul(t.z, lg), fmul(t.w, lg));
    *d = t;
  }
}

__device__ __forceinline__ int pcm_convert(float v, int bits) {
  if (bits == 24) {
    v = fmul(v, 256.0f);
    if (v < -8388608.0f)
      v = -8388608.0f;
    else if (8388607.0f < v)
      v = 8388607.0f;
    return sat_cast(v);
  }
  if (bits == 16) {
    if (v < -32768.0f)
      v = -32768.0f;
    else if (32767.0f < v)
      v = 32767.0f;
    return sat_cast(v);
  }
  v = fmul(v, 65536.0f);
  if (v < -2147483647.0f) v = -2147483647.0f;
  return (2147483647.0f < v) ? 0x7fffffff : sat_cast(v);
}

__global__ void __launch_bounds__(kPackSamples)
    limiter_apply_pack_kernel(const float* __restrict__ in, const float* __restrict__ gains, const float* __restrict__ dl_old,
                              const float* __restrict__ loud_gain, unsigned char* __restrict__ out,
                              int* __restrict__ delay, int* __restrict__ out_bytes, int attack, float thr,
                              const __grid_constant__ PackParams P) {
  pdl_prologue();
  extern __shared__ __align__(16) unsigned char tile[];  // [kPackSamples][n_ch][bytes]
  const int tid = threadIdx.x;
  const int chunk = blockIdx.x, sidx = blockIdx.y;
  const int bps = P.bits >> 3;
  const int frame_bytes = kFrame * P.n_ch * bps;
  const float* x = in + size_t(sidx) * P.n_ch * kFrame;
  const float* dlo = dl_old + size_t(sidx) * P.n_ch * attack;
  const int d = delay ? delay[sidx] : 0;
  const int s = chunk * kPackSamples + tid;
  const float gi = __ldg(gains + size_t(sidx) * kFrame + s);
  const float lg = loud_gain ? __ldg(loud_gain + sidx) : 1.0f;
  const bool has_lg = loud_gain != nullptr;
  const int rec = P.n_ch * bps;
  __shared__ unsigned char slot_ch[64];  // planar channel of every interleave slot (P.pos[] indexed per load went through local memory)
  if (tid < P.n_ch) slot_ch[tid] = P.pos[tid];
  __syncthreads();
  const bool words = (rec & 3) == 0;  // the sample's record is assembled in registers, one 32-bit store per 4 bytes
  unsigned long long sh = 0;
  int have = 0, wpos = 0;
  unsigned* tw = reinterpret_cast<unsigned*>(tile + size_t(tid) * rec);
  constexpr int kBatch = 8;
  for (int ch0 = 0; ch0 < P.n_ch; ch0 += kBatch) {
    float vin[kBatch];
#pragma unroll
    for (int q = 0; q < kBatch; ++q) {
      if (ch0 + q < P.n_ch) {
        const int c = slot_ch[ch0 + q];
        if (s >= attack) {
          const float t = __ldcs(x + size_t(c) * kFrame + s - attack);
          vin[q] = has_lg ? fmul(t, lg) : t;
        } else {
          vin[q] = __ldcs(dlo + size_t(c) * attack + s);  // already carries the loudness gain
        }
      } else {
        vin[q] = 0.0f;
      }
    }
#pragma unroll
    for (int q = 0; q < kBatch; ++q) {
      const int ch = ch0 + q;
      if (ch >= P.n_ch) break;
      float v = vin[q];
      if (gi >= 0.0f) {
        v = fmul(v, gi);
        if (thr < v)
          v = thr;
        else if (v < -thr)
          v = -thr;
      }
      const unsigned w = static_cast<unsigned>(pcm_convert(v, P.bits));
      if (words) {
        const unsigned long long bytes = bps == 4 ? w : (w & ((1u << (8 * bps)) - 1u));
        sh |= bytes << (8 * have);
        have += bps;
        if (have >= 4) {
          tw[wpos++] = static_cast<unsigned>(sh);
          sh >>= 32;
          have -= 4;
        }
      } else {
        unsigned char* t = tile + size_t(tid) * rec + ch * bps;
        t[0] = w & 0xff;
        t[1] = (w >> 8) & 0xff;
        if (bps > 2) t[2] = (w >> 16) & 0xff;
        if (bps > 3) t[3] = (w >> 24) & 0xff;
      }
    }
  }
  __syncthreads();
  const int tile_bytes = kPackSamples * P.n_ch * bps;
  const long long dst0 = (long long)(chunk * kPackSamples - d) * P.n_ch * bps;
  unsigned char* o = out + size_t(sidx) * frame_bytes;
  if (dst0 >= 0 && (dst0 & 15) == 0 && (tile_bytes & 15) == 0) {
    const uint4* t4 = reinterpret_cast<const uint4*>(tile);
    uint4* o4 = reinterpret_cast<uint4*>(o + dst0);
    for (int i = tid; i < tile_bytes / 16; i += kPackSamples) __stcs(o4 + i, t4[i]);
  } else {
    for (int i = tid; i < tile_bytes; i += kPackSamples) {
      const long long p = dst0 + i;
      if (p >= 0) o[p] = tile[i];
    }
  }
  if (chunk == 0 && tid == 0) {
    if (out_bytes) out_bytes[sidx] = (d <= kFrame) ? (kFrame - d) * P.n_ch * bps : 0;
  }
}

// the delay state is advanced by a separate tiny kernel so that every chunk CTA of a stream sees the old value
__global__ void pcm_delay_advance_kernel(int* delay, int n_streams, int n_samples) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n_streams) {
    const int d = delay[i];
    delay[i] = (d <= n_samples) ? 0 : d - n_samples;
  }
}

}  // namespace
}  // namespace mpeghb200

using namespace mpeghb200;

extern "C" {

mpeghb200_status mpeghb200_limiter_params_init(mpeghb200_limiter_params* p, int n_ch, int sample_rate,
                                               int limiter_on) {
  if (!p || n_ch <= 0 || n_ch > 64 || sample_rate <= 0) return MPEGHB200_ERR_BAD_ARG;
  // decoder/impeghd_peak_limiter.c:152-177; the constants come from the host's double-precision pow(),
  // as in the reference (GPU libm is not bit-compatible)
  const int attack = int(5.0f * sample_rate / 1000);
  if (attack < 4 || attack > kMaxAttack || (attack & 3)) return MPEGHB200_ERR_UNSUPPORTED;
  p->n_ch = n_ch;
  p->attack_samples = attack;
  p->limiter_on = limiter_on;
  p->threshold = 0.89125094f * 32768.0f;
  p->attack_const = float(std::pow(0.1, 1.0 / (attack + 1)));
  p->release_const = float(std::pow(0.1, 1.0 / (50.0f * sample_rate / 1000 + 1)));
  return MPEGHB200_OK;
}

size_t mpeghb200_limiter_state_floats(const mpeghb200_limiter_params* p) {
  return p ? size_t(4 + p->attack_samples * (1 + p->n_ch)) : 0;
}

mpeghb200_status mpeghb200_limiter_state_init_dev(mpeghb200_ctx* ctx, const mpeghb200_limiter_params* p,
                                                  float* d_state, int n_streams, void* cuda_stream) {
  if (!ctx || !p || n_streams < 0 || (n_streams > 0 && !d_state)) return MPEGHB200_ERR_BAD_ARG;
  if (n_streams == 0) return MPEGHB200_OK;
  MB_CUDA(ctx, cudaSetDevice(ctx->device));
  limiter_state_init_kernel<<<148, 256, 0, static_cast<cudaStream_t>(cuda_stream)>>>(
      d_state, int(mpeghb200_limiter_state_floats(p)), n_streams);
  ctx->launches.fetch_add(1);
  MB_CUDA(ctx, cudaGetLastError());
  return MPEGHB200_OK;
}

mpeghb200_status mpeghb200_limiter_dev(mpeghb200_ctx* ctx, const mpeghb200_limiter_params* p, const float* d_in,
                                       float* d_out, float* d_state, const float* d_loudness_gain, int n_streams,
                                       void* cuda_stream) {
  if (!ctx || !p) return MPEGHB200_ERR_BAD_ARG;
  if (n_streams < 0 || (n_streams > 0 && (!d_in || !d_out || !d_state)))
    return fail(ctx, MPEGHB200_ERR_BAD_ARG, "limiter_dev: null buffer or negative stream count");
  if (d_in == d_out) return fail(ctx, MPEGHB200_ERR_BAD_ARG, "limiter_dev: in-place operation is not supported");
  if (n_streams == 0) return MPEGHB200_OK;
  MB_CUDA(ctx, cudaSetDevice(ctx->device));
  LimParams P;
  P.n_ch = p->n_ch;
  P.attack = p->attack_samples;
  P.limiter_on = p->limiter_on;
  P.stride = int(mpeghb200_limiter_state_floats(p));
  P.thr = p->threshold;
  P.ac = p->attack_const;
  P.rc = p->release_const;
  P.fast_release = (ctx->lim_variant & 2) != 0;
  if ((ctx->lim_variant & 4) && n_streams <= 4 * ctx->sm_count)
    MB_CUDA(ctx, launch_pdl(limiter_deep_kernel<8>, dim3(n_streams), dim3(256), 0, static_cast<cudaStream_t>(cuda_stream),
                            d_in, d_out, d_state, d_loudness_gain, P));
  else if (ctx->lim_variant & 1)
    MB_CUDA(ctx, launch_pdl(limiter_kernel<256>, dim3(n_streams), dim3(256), 0, static_cast<cudaStream_t>(cuda_stream),
                            d_in, d_out, d_state, d_loudness_gain, P));
  else
    MB_CUDA(ctx, launch_pdl(limiter_kernel<128>, dim3(n_streams), dim3(128), 0, static_cast<cudaStream_t>(cuda_stream),
                            d_in, d_out, d_state, d_loudness_gain, P));
  ctx->launches.fetch_add(1);
  MB_CUDA(ctx, cudaGetLastError());
  return MPEGHB200_OK;
}

size_t mpeghb200_limiter_pack_scratch_floats(const mpeghb200_limiter_params* p) {
  return p ? size_t(kFrame) + size_t(p->n_ch) * p->attack_samples : 0;  // gains + the old delay line, per stream
}

mpeghb200_status mpeghb200_limiter_pack_dev(mpeghb200_ctx* ctx, const mpeghb200_limiter_params* p, const float* d_in,
                                            float* d_state, float* d_scratch, const float* d_loudness_gain,
                                            int pcm_bits, const int32_t* out_ch_map, int32_t* d_delay, uint8_t* d_pcm,
                                            int32_t* d_pcm_bytes, int n_streams, void* cuda_stream) {
  if (!ctx || !p) return MPEGHB200_ERR_BAD_ARG;
  if (n_streams < 0 || (pcm_bits != 16 && pcm_bits != 24 && pcm_bits != 32) ||
      (n_streams > 0 && (!d_in || !d_state || !d_scratch || !d_pcm)))
    return fail(ctx, MPEGHB200_ERR_BAD_ARG, "limiter_pack_dev: null buffer, bad sample size or negative stream count");
  if (n_streams == 0) return MPEGHB200_OK;
  MB_CUDA(ctx, cudaSetDevice(ctx->device));
  cudaStream_t st = static_cast<cudaStream_t>(cuda_stream);
  LimParams P;
  P.n_ch = p->n_ch;
  P.attack = p->attack_samples;
  P.limiter_on = p->limiter_on;
  P.stride = int(mpeghb200_limiter_state_floats(p));
  P.thr = p->threshold;
  P.ac = p->attack_const;
  P.rc = p->release_const;
  P.fast_release = (ctx->lim_variant & 2) != 0;
  PackParams K;
  K.n_ch = p->n_ch;
  K.bits = pcm_bits;
  K.n_samples = kFrame;
  K.ch_stride = kFrame;
  for (int ch = 0; ch < 64; ++ch) K.pos[ch] = static_cast<unsigned char>(ch < p->n_ch ? ch : 0);
  if (out_ch_map) {
    for (int ch = 0; ch < p->n_ch; ++ch) {
      const int m = out_ch_map[ch] == -1 ? ch : out_ch_map[ch];
      if (m < 0 || m >= p->n_ch) return fail(ctx, MPEGHB200_ERR_BAD_ARG, "limiter_pack_dev: out_ch_map entry out of range");
      K.pos[m] = static_cast<unsigned char>(ch);
    }
  }
  float* d_gains = d_scratch;
  float* d_dl_old = d_scratch + size_t(n_streams) * kFrame;
  if (ctx->lim_variant & 1)
    MB_CUDA(ctx, launch_pdl(limiter_gains_kernel<256>, dim3(n_streams), dim3(256), 0, st, d_in, d_state, d_loudness_gain,
                            d_gains, d_dl_old, P));
  else
    MB_CUDA(ctx, launch_pdl(limiter_gains_kernel<128>, dim3(n_streams), dim3(128), 0, st, d_in, d_state, d_loudness_gain,
                            d_gains, d_dl_old, P));
  ctx->launches.fetch_add(1);
  const size_t smem = size_t(kPackSamples) * p->n_ch * (pcm_bits >> 3);
  if (smem > 48 * 1024)
    MB_CUDA(ctx, cudaFuncSetAttribute(limiter_apply_pack_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
  MB_CUDA(ctx, launch_pdl(limiter_apply_pack_kernel, dim3(kFrame / kPackSamples, n_streams), dim3(kPackSamples), smem, st,
                          d_in, (const float*)d_gains, (const float*)d_dl_old, d_loudness_gain, d_pcm, d_delay, d_pcm_bytes,
                          p->attack_samples, p->threshold, K));
  ctx->launches.fetch_add(1);
  MB_CUDA(ctx, cudaGetLastError());
  if (d_delay) {
    pcm_delay_advance_kernel<<<(n_streams + 255) / 256, 256, 0, st>>>(d_delay, n_streams, kFrame);
    ctx->launches.fetch_add(1);
    MB_CUDA(ctx, cudaGetLastError());
  }
  return MPEGHB200_OK;
}

mpeghb200_status mpeghb200_pcm_pack_dev(mpeghb200_ctx* ctx, const float* d_in, int n_ch, int pcm_bits,
                                        const int32_t* out_ch_map, int32_t* d_delay, uint8_t* d_out,
                                        int32_t* d_out_bytes, int n_streams, void* cuda_stream) {
  return mpeghb200_pcm_pack_block_dev(ctx, d_in, n_ch, kFrame, kFrame, pcm_bits, out_ch_map, d_delay, d_out, d_out_bytes,
                                      n_streams, cuda_stream);
}

mpeghb200_status mpeghb200_pcm_pack_block_dev(mpeghb200_ctx* ctx, const float* d_in, int n_ch, int n_samples,
                                              int ch_stride, int pcm_bits, const int32_t* out_ch_map,
                                              int32_t* d_delay, uint8_t* d_out, int32_t* d_out_bytes, int n_streams,
                                              void* cuda_stream) {
  if (!ctx) return MPEGHB200_ERR_BAD_ARG;
  if (n_samples <= 0 || n_samples % kPackSamples || ch_stride < n_samples)
    return fail(ctx, MPEGHB200_ERR_BAD_ARG, "pcm_pack_block_dev: n_samples must be a positive multiple of 256, ch_stride >= n_samples");
  if (n_ch <= 0 || n_ch > 64 || (pcm_bits != 16 && pcm_bits != 24 && pcm_bits != 32) || n_streams < 0 ||
      (n_streams > 0 && (!d_in || !d_out)))
    return fail(ctx, MPEGHB200_ERR_BAD_ARG, "pcm_pack_dev: bad channel count, sample size or buffer");
  if (n_streams == 0) return MPEGHB200_OK;
  MB_CUDA(ctx, cudaSetDevice(ctx->device));
  PackParams P;
  P.n_ch = n_ch;
  P.bits = pcm_bits;
  P.n_samples = n_samples;
  P.ch_stride = ch_stride;
  // decoder/ia_core_coder_decode_main.c:231-236: channel_pos[out_ch_map[ch]] = ch, -1 meaning identity
  for (int ch = 0; ch < 64; ++ch) P.pos[ch] = static_cast<unsigned char>(ch < n_ch ? ch : 0);
  if (out_ch_map) {
    for (int ch = 0; ch < n_ch; ++ch) {
      const int m = out_ch_map[ch] == -1 ? ch : out_ch_map[ch];
      if (m < 0 || m >= n_ch) return fail(ctx, MPEGHB200_ERR_BAD_ARG, "pcm_pack_dev: out_ch_map entry out of range");
      P.pos[m] = static_cast<unsigned char>(ch);
    }
  }
  cudaStream_t st = static_cast<cudaStream_t>(cuda_stream);
  const size_t smem = size_t(kPackSamples) * n_ch * (pcm_bits >> 3);
  // per device / context attribute: set on every call (a process may hold contexts on several GPUs)
  dim3 grid(n_samples / kPackSamples, n_streams);
  auto launch = [&](void (*k)(const float*, unsigned char*, int*, int*, PackParams)) -> mpeghb200_status {
    // per device / context attribute: set on every call (a process may hold contexts on several GPUs)
    if (smem > 48 * 1024) MB_CUDA(ctx, cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
    MB_CUDA(ctx, launch_pdl(k, grid, dim3(kPackSamples), smem, st, d_in, d_out, d_delay, d_out_bytes, P));
    return MPEGHB200_OK;
  };
  mpeghb200_status ls;
  if (n_ch % 4 == 0 && !ctx->pack_generic)
    ls = pcm_bits == 24 ? launch(pcm_pack4_kernel<24>) : pcm_bits == 16 ? launch(pcm_pack4_kernel<16>) : launch(pcm_pack4_kernel<32>);
  else
    ls = launch(pcm_pack_kernel);
  if (ls != MPEGHB200_OK) return ls;
  ctx->launches.fetch_add(1);
  MB_CUDA(ctx, cudaGetLastError());
  if (d_delay) {
    pcm_delay_advance_kernel<<<(n_streams + 255) / 256, 256, 0, st>>>(d_delay, n_streams, n_samples);
    ctx->launches.fetch_add(1);
    MB_CUDA(ctx, cudaGetLastError());
  }
  return MPEGHB200_OK;
}

}  // extern "C"
