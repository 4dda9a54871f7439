// igf.cu -- intelligent gap filling of FD channel-frames, mono path, in place on the spectra.
//
// Stands in for ia_core_coder_igf_mono (decoder/ia_core_coder_igf_dec.c:2479) -- source-tile whitening (:913),
// band energies of the coded spectrum and of the mapped source (:1083; tile mapping :726-880), group
// normalisation (:1002), energy-matching gain and fill of the zero bins (:1321) with random-sign noise for
// whitening level 2 (:882) -- followed by ia_core_coder_igf_tnf (:2297): temporal noise flattening of the filled
// range (detect :503, reflection / prediction coefficients :330-416, filter :418).  Long blocks (igf_win_type 0)
// and eight-short blocks (type 1); TCX window types 2/3 belong to the LPD path, which stays on the host.
//
// One CTA of 256 threads per channel-frame, everything staged in shared memory (coded spectrum 4 KB, four source
// tiles 16 KB, mapped source 4 KB, TNF buffers 8 KB).  What the reference does bin after bin splits into
//   * per-bin work (whitening envelope, tile / source-bin mapping, fill, TNF FIR): 4 bins per thread,
//   * strictly ordered float sums (band energies, TNF norms and autocorrelations): one thread per sum, in the
//     reference's order, so every partial sum rounds as it does there,
//   * the noise generator: the reference draws one LCG step per zero bin of a level-2 tile in bin order; a block
//     scan gives each such bin its draw index k and the thread jumps the LCG by k + 1 steps (affine powers),
//   * scalar chains (gains with double sqrt, Levinson recursion of order 8): one thread.
// Bytes per channel-frame: spectrum 4096 r + 4096 w, tiles 16384 r, side 548 r; long blocks with TNF enabled
// also write the 8 KB of carried TNF buffers (igf_tnf_spec_float / igf_tnf_mask persist in the reference handle
// and are re-applied by a frame that signals TNF on top of igf_all_zero).
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>

#include "common.h"
#include "bands.h"

namespace mpeghb200 {
namespace {

constexpr int kThreads = 256;
constexpr int kFrame = 1024;
constexpr float kNoise = 2097152.0f;

struct IgfRom {
  float whiten[2][64];  // [0]: 2^((42-n)/2) as the reference ROM prints it, [1]: 2^21 * 2^(n/2)
  float acf_win[8];
  uint32_t lcg_a[11], lcg_c[11];  // x -> a x + c composed 2^b times
};

__device__ __forceinline__ uint32_t lcg_jump(uint32_t seed, uint32_t steps, const IgfRom& R) {
#pragma unroll
  for (int b = 0; b < 11; ++b)
    if (steps & (1u << b)) seed = seed * R.lcg_a[b] + R.lcg_c[b];
  return seed;
}

struct Grid {
  int n_tiles, tile[4];
  int tstart[4];  // first bin of tile t
  int toff[4];    // source offset of tile t for this frame's tile numbers: source bin = target bin - toff
  int bgn, end;
};

__device__ __forceinline__ int tile_of(int tb, const Grid& G) {
  int t = 0;
#pragma unroll
  for (int q = 0; q < 4; ++q) t += (q < G.n_tiles && tb >= G.tstart[q] + G.tile[q]);
  return t;  // == n_tiles when the bin lies behind the last tile, like the reference's loop
}

// ia_core_coder_igf_get_sb (:839) with the per-tile offset precomputed (source_offset below)
__device__ __forceinline__ int source_bin(int igf_min, int tb, int ix, const Grid& G, const int* toff) {
  const int src = G.bgn - igf_min;
  int sb = tb;
  if (src > 0) {
    // select instead of indexing: G and toff stay in registers (tb >= bgn == tstart[0] for every caller's bin)
    const int off = ix == 0 ? toff[0] : ix == 1 ? toff[1] : ix == 2 ? toff[2] : toff[3];
    if (ix < G.n_tiles && tb >= G.bgn) sb = tb - off;
    if (sb < igf_min) sb = igf_min + tb % src;
  }
  return sb;
}

__device__ __forceinline__ int source_offset(int tile_len, int sbs, int bgn, int tile_num) {
  const float sel = float(int8_t(tile_num));
  int offset = int(fsub(fadd(fmul(fadd(2.5f, fmul(-0.5f, sel)), float(tile_len)), float(sbs)), float(bgn)));
  offset += offset % 2;
  return offset;
}

// Tile widths of the grid (ia_core_coder_igf_get_num_tiles :760) by one thread; b2s maps a bin to its band, so the
// reference's "first band edge at or above lim" search is one lookup: j = band(lim - 1) + 1.
__device__ void compute_grid(Grid& G, const mpeghb200_igf_side& sd, const int16_t* __restrict__ tbl,
                             const uint8_t* __restrict__ b2s) {
  auto edge = [&](int i) { return i ? int(__ldg(tbl + i - 1)) : 0; };
  const int s0 = sd.sfb_start, s1 = sd.sfb_stop;
  const int bgn = edge(s0), end = edge(s1);
  int range = end - bgn < 8 ? 8 : end - bgn, mem = s0, sbs = bgn, n = 0;
  int tile[4] = {range >> 2, range >> 2, range >> 2, range >> 2};
  if (bgn > sd.igf_min && bgn < end) {
    for (int i = 0; i < 4; ++i) {
      const int lim = min(sbs + tile[i], end);
      int j = lim > 0 ? int(__ldg(b2s + lim - 1)) + 1 : 1;
      if (j < 1) j = 1;
      if (!sd.use_high_res) {
        if (i == 3)
          j = s1;
        else if ((((j - s0) & 1) == 1) && ((j - mem) > 1))
          --j;
        mem = j;
      }
      tile[i] = max(2, min(edge(j) - sbs, end - sbs));
      ++n;
      sbs += tile[i];
      if (sbs == end) break;
    }
  }
  int st = bgn;
  for (int i = 0; i < 4; ++i) {
    G.tile[i] = i < n ? tile[i] : 0;
    G.tstart[i] = st;
    st += G.tile[i];
  }
  G.n_tiles = n, G.bgn = bgn, G.end = end;
}

__global__ void __launch_bounds__(kThreads, 6)
    igf_kernel(float* __restrict__ coef_g, const float* __restrict__ tiles_g,
               const mpeghb200_igf_side* __restrict__ sides, float* __restrict__ tnf_state,
               uint32_t* __restrict__ seed_out, const int16_t* __restrict__ sfb_long, int n_long,
               const int16_t* __restrict__ sfb_short, int n_short, const uint8_t* __restrict__ b2s_long,
               const uint8_t* __restrict__ b2s_short, const int tnf_only, const IgfRom R) {
  pdl_prologue();
  __shared__ mpeghb200_igf_side sd;
  __shared__ __align__(16) float coef[kFrame];
  __shared__ __align__(16) float tiles[4 * kFrame];
  __shared__ __align__(16) float val[kFrame];   // mapped source value of every bin of the IGF range (noise incl.)
  __shared__ __align__(16) float tnf[kFrame];   // igf_tnf_spec_float
  __shared__ __align__(16) float tbuf[kFrame];  // unfiltered copy for the FIR / igf_tnf_mask staging
  __shared__ float gn[128];                     // gains, [16 * group + sfb] (long: [sfb])
  __shared__ float esum[2][128];                // group-normalised band energies: coded, source
  __shared__ int16_t swb[72];
  __shared__ Grid G;
  __shared__ int toff[4];
  __shared__ uint32_t warp_cnt[kThreads / 32];
  __shared__ float norms[3], acfp[3][8], pred[17];
  __shared__ int tnf_order;

  const int tid = threadIdx.x;
  {
    const uint32_t* src = reinterpret_cast<const uint32_t*>(sides + blockIdx.x);
    uint32_t* dst = reinterpret_cast<uint32_t*>(&sd);
    for (int i = tid; i < int(sizeof(sd) / 4); i += kThreads) dst[i] = src[i];
  }
  __syncthreads();
  const bool is_short = sd.win_type == 1;
  const int bins = is_short ? 128 : kFrame;
  const int n_sfb = is_short ? n_short : n_long;
  const bool noise_ok = !is_short && sd.use_whitening;
  const bool tnf_on = sd.use_tnf && !is_short;
  const bool do_tnf = tnf_on && sd.is_tnf == 1;
  // tnf_only: the TNF stage on its own (after the joint-stereo fill and the stereo tools): same as the reference's
  // igf_tnf on a frame whose fill left its buffers behind
  const bool all_zero = sd.all_zero != 0 || tnf_only;
  if (all_zero && !do_tnf) {
    if (tid == 0 && !tnf_only) seed_out[blockIdx.x] = sd.seed;
    return;
  }
  float* cg = coef_g + size_t(sd.unit) * kFrame;
  float* stg = tnf_state ? tnf_state + size_t(sd.unit) * 2 * kFrame : nullptr;

  // ---- stage: band edges, spectrum, tiles -----------------------------------------------------------
  for (int i = tid; i <= n_sfb; i += kThreads) swb[i] = i ? (is_short ? sfb_short : sfb_long)[i - 1] : 0;
  const float4 rc = reinterpret_cast<const float4*>(cg)[tid];
  float4 rt[4];
  if (!all_zero) {
    const float4* tg = reinterpret_cast<const float4*>(tiles_g + size_t(blockIdx.x) * 4 * kFrame);
#pragma unroll
    for (int t = 0; t < 4; ++t) rt[t] = tg[t * 256 + tid];
  }
  if (tid == 0) {  // while the loads are in flight
    compute_grid(G, sd, is_short ? sfb_short : sfb_long, is_short ? b2s_short : b2s_long);
    for (int i = 0; i < 4; ++i) toff[i] = source_offset(G.tile[i], G.tstart[i], G.bgn, sd.tile_num[i]);
  }
  reinterpret_cast<float4*>(coef)[tid] = rc;
  if (!all_zero) {
#pragma unroll
    for (int t = 0; t < 4; ++t) reinterpret_cast<float4*>(tiles)[t * 256 + tid] = rt[t];
  }
  __syncthreads();
  const int bgn = G.bgn, end = G.end;

  if (!all_zero) {
    // ---- whitening of the level-0 source tiles (long blocks) -----------------------------------------
    if (noise_ok) {
      const int lo = sd.igf_min, stop = bgn;
      for (int t = 0; t < G.n_tiles; ++t) {
        if (sd.whitening_level[t] != 0) continue;  // uniform
        const float* x = tiles + t * kFrame;
        float out[4];
        bool has[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const int i = 4 * tid + k;
          has[k] = i >= lo && i < stop;
          out[k] = 0.0f;
          if (has[k]) {
            const int ie = i < stop - 3 ? i : stop - 4;  // the last three bins reuse the gain of bin stop - 4
            float fac = 0.0f;
            if (ie >= lo) {
              float env = 1e-3f;
#pragma unroll
              for (int j = -3; j <= 3; ++j) env = fadd(env, fmul(x[ie + j], x[ie + j]));
              // the reference searches its power-of-two tables: floor(log2 env) above 1, ceil(-log2 env) below
              const int ex = int((__float_as_uint(env) >> 23) & 0xff) - 127;
              fac = env >= 1.0f ? R.whiten[0][ex & 63] : R.whiten[1][(-ex) & 63];
            }
            out[k] = fmul(x[i], fac);
          }
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < 4; ++k)
          if (has[k]) tiles[t * kFrame + 4 * tid + k] = out[k];
        __syncthreads();
      }
    }

    // ---- map every bin of the IGF range to its source value; number the noise draws ---------------------
    uint32_t my_noise = 0;
    bool noise[4];
    float v[4];
    const Grid Gr = G;  // registers
    int toff_r[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) toff_r[q] = toff[q];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int i = 4 * tid + k;
      const int w = is_short ? i >> 7 : 0;
      const int b = i - w * bins;
      noise[k] = false;
      v[k] = 0.0f;
      if (b >= bgn && b < end) {
        const int ix = tile_of(b, Gr);
        const int sb = source_bin(sd.igf_min, b, ix, Gr, toff_r);
        v[k] = tiles[ix * kFrame + w * bins + sb];
        noise[k] = noise_ok && coef[i] == 0.0f && sd.whitening_level[ix] == 2;
        my_noise += noise[k];
      }
    }
    // exclusive block scan of the draw counts (bin order == the reference's band-by-band order)
    uint32_t incl = my_noise;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const uint32_t o = __shfl_up_sync(0xffffffffu, incl, d);
      if ((tid & 31) >= d) incl += o;
    }
    if ((tid & 31) == 31) warp_cnt[tid >> 5] = incl;
    __syncthreads();
    uint32_t base = 0, total = 0;
#pragma unroll
    for (int wv = 0; wv < kThreads / 32; ++wv) {
      if (wv < (tid >> 5)) base += warp_cnt[wv];
      total += warp_cnt[wv];
    }
    uint32_t kdraw = base + incl - my_noise;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      if (noise[k]) {
        const uint32_t s = lcg_jump(sd.seed, ++kdraw, R);
        v[k] = fmul((s & 0x10000u) ? -1.0f : 1.0f, kNoise);
      }
      val[4 * tid + k] = v[k];
    }
    if (tid == 0) seed_out[blockIdx.x] = lcg_jump(sd.seed, total, R);  // the energy pass drew the same count
    __syncthreads();

    // ---- band energies per (group, band), windows of a group in order (:1083, :1002) ---------------------
    // The squares are formed by everybody first (tnf / tbuf are free until the fill); what stays serial is the
    // pure addition chain of every band, one thread per (sum, group, band).  Adding the +0 of a coded bin to the
    // source-energy chain leaves it unchanged, so that chain runs over all bins of the band, branch-free.
    const int step = (sd.use_high_res || is_short) ? 1 : 2;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int i = 4 * tid + k;
      const float cb = coef[i];
      tnf[i] = fmul(cb, cb);
      tbuf[i] = cb == 0.0f ? fmul(val[i], val[i]) : 0.0f;
    }
    __syncthreads();
    {
      const int which = tid >> 7, r = tid & 127;  // 0: coded energy, 1: source energy on the zero bins
      const int g = is_short ? r >> 4 : 0;
      const int sfb = is_short ? r & 15 : r;
      const bool live = g < sd.num_groups && sfb >= sd.sfb_start && sfb < sd.sfb_stop &&
                        ((sfb - sd.sfb_start) % step) == 0 && (is_short || r < 64);
      if (live) {
        int w0 = 0;
        for (int q = 0; q < g; ++q) w0 += sd.group_len[q];
        const int hi = min(sfb + step, int(sd.sfb_stop));
        const float* arr = which ? tbuf : tnf;
        float acc = 0.0f;
        for (int q = 0; q < sd.group_len[g]; ++q) {
          const float* a = arr + (w0 + q) * bins;
          float e = 0.0f;
          int b = swb[sfb];
          const int b1 = swb[hi];
          for (; (b & 3) == 0 && b + 4 <= b1; b += 4) {  // band edges are multiples of 4 in every standard table
            const float4 x = *reinterpret_cast<const float4*>(a + b);
            e = fadd(fadd(fadd(fadd(e, x.x), x.y), x.z), x.w);
          }
          for (; b < b1; ++b) e = fadd(e, a[b]);
          acc = fadd(acc, e);
        }
        esum[which][(is_short ? 16 * g : 0) + sfb] = __fdiv_rn(acc, float(sd.group_len[g]));
      }
    }
    __syncthreads();
    if (tid < 128) {
      const int g = is_short ? tid >> 4 : 0;
      const int sfb = is_short ? tid & 15 : tid;
      const bool live = g < sd.num_groups && sfb >= sd.sfb_start && sfb < sd.sfb_stop &&
                        ((sfb - sd.sfb_start) % step) == 0 && (is_short || tid < 64);
      if (live) {
        const int slot = (is_short ? 16 * g : 0) + sfb;
        const int hi = min(sfb + step, int(sd.sfb_stop));
        const float s_acc = esum[0][slot], p_acc = esum[1][slot];
        const int width = swb[hi] - swb[sfb];
        const float lev = sd.levels[slot];
        const float mn = fsub(fmul(fmul(lev, lev), float(width)), s_acc);
        float gain = 0.0f;
        if (0.0f < mn && 0.0f < p_acc) {
          gain = float(__dsqrt_rn(__ddiv_rn(double(mn), double(p_acc))));
          gain = gain < 10.f ? gain : 10.f;
        }
        gn[slot] = gain;
        if (step == 2 && sfb + 1 < sd.sfb_stop) gn[slot + 1] = gain;  // both bands of a low-resolution pair
      }
    }
    __syncthreads();

    // ---- fill the zero bins (:1321) -------------------------------------------------------------------
    const uint8_t* b2s = is_short ? b2s_short : b2s_long;
    // group of each window (short blocks)
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int i = 4 * tid + k;
      const int w = is_short ? i >> 7 : 0;
      const int b = i - w * bins;
      float spec = 0.0f, mask = 1.0f;
      if (b >= bgn && b < end) {
        const int sfb_of = __ldg(b2s + b);
        int g = 0;
        if (is_short) {
          int acc = sd.group_len[0];
          while (w >= acc) acc += sd.group_len[++g];
        }
        const float gv = fmul(gn[(is_short ? 16 * g : 0) + sfb_of], val[i]);
        spec = gv;
        if (coef[i] == 0.0f) {
          coef[i] = gv;
          mask = 0.0f;
        }
      }
      tnf[i] = spec;
      tbuf[i] = mask;
    }
    __syncthreads();
    if (tnf_on && stg && !do_tnf) {  // leave the buffers behind for a later all-zero frame
      reinterpret_cast<float4*>(stg)[tid] = reinterpret_cast<const float4*>(tnf)[tid];
      reinterpret_cast<float4*>(stg + kFrame)[tid] = reinterpret_cast<const float4*>(tbuf)[tid];
    }
  } else {
    // igf_all_zero with TNF signalled: the reference flattens and applies what its buffers still hold
    if (tid == 0 && !tnf_only) seed_out[blockIdx.x] = sd.seed;
    reinterpret_cast<float4*>(tnf)[tid] = reinterpret_cast<const float4*>(stg)[tid];
    reinterpret_cast<float4*>(tbuf)[tid] = reinterpret_cast<const float4*>(stg + kFrame)[tid];
    __syncthreads();
  }

  // ---- temporal noise flattening (:503-584, :2297) ------------------------------------------------------
  if (do_tnf) {
    const int range = end - bgn;
    float mask4[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) mask4[k] = tbuf[4 * tid + k];
    if (stg) reinterpret_cast<float4*>(stg + kFrame)[tid] = make_float4(mask4[0], mask4[1], mask4[2], mask4[3]);
    __syncthreads();
    if (tid < 3) {
      const int a = bgn + range * tid / 3, b = bgn + range * (tid + 1) / 3;
      float acc = 0.0f;
      for (int j = a; j < b; ++j) acc = fadd(acc, fmul(tnf[j], tnf[j]));
      norms[tid] = acc;
    } else if (tid >= 32 && tid < 56) {
      const int i = (tid - 32) >> 3, lag = ((tid - 32) & 7) + 1;
      const int a = bgn + range * i / 3, b = bgn + range * (i + 1) / 3, n = b - a;
      float acc = 0.0f;
      if (n - lag) acc = fmul(tnf[a], tnf[a + lag]);
      for (int j = 1; j < n - lag; ++j) acc = fadd(acc, fmul(tnf[a + j], tnf[a + j + lag]));
      acfp[i][lag - 1] = acc;
    }
    __syncthreads();
    if (tid == 0) {
      float acf[9], par[16], mem[32];
      int i;
      for (i = 0; i < 9; ++i) acf[i] = 0.0f;
      for (i = 0; i < 3 && norms[i] > 0.0000000037f; ++i) {
        const float fac = __fdiv_rn(1.0f, norms[i]);
        for (int lag = 1; lag <= 8; ++lag)
          acf[lag] = fadd(acf[lag], fmul(fac, fmul(R.acf_win[lag - 1], acfp[i][lag - 1])));
      }
      int order = 0;
      for (int q = 0; q < 17; ++q) pred[q] = 0.0f;
      if (i == 3) {
        order = 8;
        acf[0] = 3.0f;
        const int lo = min(8, range >> 2);
        for (int q = 0; q < 16; ++q) par[q] = 0.0f;
        for (int q = 0; q < 32; ++q) mem[q] = 0.0f;
        float* pm = mem + lo;
        for (int q = 0; q < lo; ++q) mem[q] = acf[q];
        for (int q = 0; q < lo; ++q) pm[q] = acf[q + 1];
        for (int q = 0; q < lo; ++q) {
          float t = 0.0f;
          if (mem[0] >= 0.0000152f) t = __fdiv_rn(-pm[q], mem[0]);
          t = fminf(0.999f, fmaxf(-0.999f, t));
          par[q] = t;
          for (int j = q; j < lo; ++j) {
            const float t2 = fadd(pm[j], fmul(t, mem[j - q]));
            mem[j - q] = fadd(mem[j - q], fmul(t, pm[j]));
            pm[j] = t2;
          }
        }
        pred[0] = 1.0f;
        pred[1] = par[0];
        for (int q = 1; q < lo; ++q) {
          int j;
          for (j = 0; j < (q >> 1); ++j) {
            const float t = pred[j + 1];
            pred[j + 1] = fadd(pred[j + 1], fmul(par[q], pred[q - j]));
            pred[q - j] = fadd(pred[q - j], fmul(par[q], t));
          }
          if (q & 1) pred[j + 1] = fadd(pred[j + 1], fmul(par[q], pred[j + 1]));
          pred[q + 1] = par[q];
        }
      }
      tnf_order = order;
    }
    __syncthreads();
    // FIR over the unfiltered values; each bin adds its taps in the reference's order (i = 1 .. order - 1)
    float y[4], pr[8];
#pragma unroll
    for (int t = 1; t < 8; ++t) pr[t] = pred[t];
    const int n_taps = tnf_order;  // 0 or 8
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int i = 4 * tid + k;
      y[k] = tnf[i];
      if (i >= bgn && i < end && n_taps) {
        const int j = i - bgn;
#pragma unroll
        for (int t = 1; t < 8; ++t)
          if (j >= t) y[k] = fadd(y[k], fmul(pr[t], tnf[i - t]));
      }
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int i = 4 * tid + k;
      tnf[i] = y[k];
      if (i >= bgn && i < end && mask4[k] == 0.0f) coef[i] = y[k];
    }
    if (stg) reinterpret_cast<float4*>(stg)[tid] = make_float4(y[0], y[1], y[2], y[3]);
    __syncthreads();
  }
  reinterpret_cast<float4*>(cg)[tid] = reinterpret_cast<const float4*>(coef)[tid];
}

// ---- joint-stereo fill of a channel pair (ia_core_coder_igf_apply_stereo :1661, calc :1169, proc :1431) -------
// Same building blocks as the mono kernel, both channels side by side in one CTA: the source value of every bin of
// the range is determined once per channel (tile mapping; for whitening level 2 a noise draw on EVERY bin of that
// tile, zero or not), mixed to mid/side on the masked bands, and then feeds the energies, the gains (float
// quotient, double square root here) and the fill.  The TNF buffers of both channels are always left behind: the
// stereo tools act on the filled spectra and on those buffers before mpeghb200_igf_tnf_dev flattens them.
__global__ void __launch_bounds__(kThreads)
    igf_stereo_kernel(float* __restrict__ coef_g, const float* __restrict__ tiles_g,
                      const mpeghb200_igf_side* __restrict__ sides, const uint8_t* __restrict__ masks,
                      float* __restrict__ tnf_state, uint32_t* __restrict__ seed_out,
                      const int16_t* __restrict__ sfb_long, int n_long, const int16_t* __restrict__ sfb_short,
                      int n_short, const uint8_t* __restrict__ b2s_long, const uint8_t* __restrict__ b2s_short,
                      const IgfRom R) {
  pdl_prologue();
  extern __shared__ __align__(16) float dsm[];
  float* coef = dsm;                   // [2][1024]
  float* tiles = coef + 2 * kFrame;    // [2][4][1024]
  float* val = tiles + 8 * kFrame;     // [2][1024]
  float* tnf = val + 2 * kFrame;       // [2][1024] spectrum
  float* tmask = tnf + 2 * kFrame;     // [2][1024] 1 = coded, 0 = filled
  __shared__ mpeghb200_igf_side sd2[2];
  __shared__ uint8_t mask[128];
  __shared__ float gn[2][128];
  __shared__ int16_t swb[72];
  __shared__ Grid G;
  __shared__ int toff2[2][4];
  __shared__ uint32_t warp_cnt[2][kThreads / 32];
  const int tid = threadIdx.x;
  {
    const uint32_t* src = reinterpret_cast<const uint32_t*>(sides + 2 * blockIdx.x);
    uint32_t* dst = reinterpret_cast<uint32_t*>(sd2);
    for (int i = tid; i < int(2 * sizeof(mpeghb200_igf_side) / 4); i += kThreads) dst[i] = src[i];
    if (tid < 32) reinterpret_cast<uint32_t*>(mask)[tid] = reinterpret_cast<const uint32_t*>(masks + 128 * blockIdx.x)[tid];
  }
  __syncthreads();
  const mpeghb200_igf_side& sl = sd2[0];
  if (sd2[0].all_zero && sd2[1].all_zero) {
    if (tid < 2) seed_out[2 * blockIdx.x + tid] = sd2[tid].seed;
    return;
  }
  const bool is_short = sl.win_type == 1;
  const int bins = is_short ? 128 : kFrame;
  const int n_sfb = is_short ? n_short : n_long;
  const bool noise_ok = !is_short && sl.use_whitening;
  const bool tnf_on = sl.use_tnf && !is_short;
  float* cg[2] = {coef_g + size_t(sd2[0].unit) * kFrame, coef_g + size_t(sd2[1].unit) * kFrame};

  for (int i = tid; i <= n_sfb; i += kThreads) swb[i] = i ? (is_short ? sfb_short : sfb_long)[i - 1] : 0;
#pragma unroll
  for (int ch = 0; ch < 2; ++ch) {
    reinterpret_cast<float4*>(coef + ch * kFrame)[tid] = reinterpret_cast<const float4*>(cg[ch])[tid];
    const float4* tg = reinterpret_cast<const float4*>(tiles_g + (size_t(2 * blockIdx.x) + ch) * 4 * kFrame);
#pragma unroll
    for (int t = 0; t < 4; ++t) reinterpret_cast<float4*>(tiles + ch * 4 * kFrame)[t * 256 + tid] = tg[t * 256 + tid];
  }
  __syncthreads();
  if (tid == 0) {
    compute_grid(G, sl, is_short ? sfb_short : sfb_long, is_short ? b2s_short : b2s_long);
    for (int ch = 0; ch < 2; ++ch)
      for (int i = 0; i < 4; ++i) toff2[ch][i] = source_offset(G.tile[i], G.tstart[i], G.bgn, sd2[ch].tile_num[i]);
  }
  __syncthreads();
  const int bgn = G.bgn, end = G.end;

  // ---- whitening of the level-0 source tiles of both channels (long blocks) --------------------------------
  if (noise_ok) {
    const int lo = sl.igf_min, stop = bgn;
    for (int ct = 0; ct < 2 * G.n_tiles; ++ct) {
      const int ch = ct / G.n_tiles, t = ct - ch * G.n_tiles;
      if (sd2[ch].whitening_level[t] != 0) continue;  // uniform
      float* x = tiles + (ch * 4 + t) * kFrame;
      float out[4];
      bool has[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int i = 4 * tid + k;
        has[k] = i >= lo && i < stop;
        out[k] = 0.0f;
        if (has[k]) {
          const int ie = i < stop - 3 ? i : stop - 4;
          float fac = 0.0f;
          if (ie >= lo) {
            float env = 1e-3f;
#pragma unroll
            for (int j = -3; j <= 3; ++j) env = fadd(env, fmul(x[ie + j], x[ie + j]));
            const int ex = int((__float_as_uint(env) >> 23) & 0xff) - 127;
            fac = env >= 1.0f ? R.whiten[0][ex & 63] : R.whiten[1][(-ex) & 63];
          }
          out[k] = fmul(x[i], fac);
        }
      }
      __syncthreads();
#pragma unroll
      for (int k = 0; k < 4; ++k)
        if (has[k]) x[4 * tid + k] = out[k];
      __syncthreads();
    }
  }

  // ---- source values of both channels, noise draws numbered in bin order, mid/side on the masked bands -------
  const int step = (sl.use_high_res || is_short) ? 1 : 2;
  const uint8_t* b2s = is_short ? b2s_short : b2s_long;
  int sfb4[4], grp4 = 0;
  {
    const int w = is_short ? (4 * tid) >> 7 : 0;
    if (is_short) {
      int acc = sl.group_len[0];
      while (w >= acc) acc += sl.group_len[++grp4];
    }
  }
  float v[2][4];
  bool in_range[4];
  uint32_t my_noise[2] = {0, 0};
  bool noise[2][4];
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int i = 4 * tid + k;
    const int w = is_short ? i >> 7 : 0;
    const int b = i - w * bins;
    in_range[k] = b >= bgn && b < end;
    sfb4[k] = 0;
    v[0][k] = v[1][k] = 0.0f;
    noise[0][k] = noise[1][k] = false;
    if (in_range[k]) {
      // the reference walks the range in steps of one or two bands and indexes mask, level and gain by the first
      // band of the step
      const int sfb_of = __ldg(b2s + b);
      sfb4[k] = sl.sfb_start + ((sfb_of - sl.sfb_start) / step) * step;
      const int ix = tile_of(b, G);
#pragma unroll
      for (int ch = 0; ch < 2; ++ch) {
        const int sb = source_bin(sl.igf_min, b, ix, G, toff2[ch]);
        v[ch][k] = tiles[(ch * 4 + ix) * kFrame + w * bins + sb];
        noise[ch][k] = noise_ok && sd2[ch].whitening_level[ix] == 2;
        my_noise[ch] += noise[ch][k];
      }
    }
  }
  uint32_t incl[2] = {my_noise[0], my_noise[1]};
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const uint32_t o0 = __shfl_up_sync(0xffffffffu, incl[0], d), o1 = __shfl_up_sync(0xffffffffu, incl[1], d);
    if ((tid & 31) >= d) incl[0] += o0, incl[1] += o1;
  }
  if ((tid & 31) == 31) warp_cnt[0][tid >> 5] = incl[0], warp_cnt[1][tid >> 5] = incl[1];
  __syncthreads();
#pragma unroll
  for (int ch = 0; ch < 2; ++ch) {
    uint32_t base = 0, total = 0;
#pragma unroll
    for (int wv = 0; wv < kThreads / 32; ++wv) {
      if (wv < (tid >> 5)) base += warp_cnt[ch][wv];
      total += warp_cnt[ch][wv];
    }
    uint32_t kdraw = base + incl[ch] - my_noise[ch];
#pragma unroll
    for (int k = 0; k < 4; ++k)
      if (noise[ch][k]) {
        const uint32_t s = lcg_jump(sd2[ch].seed, ++kdraw, R);
        v[ch][k] = fmul(kNoise, (s & 0x10000u) ? -1.0f : 1.0f);
      }
    if (tid == 0) seed_out[2 * blockIdx.x + ch] = lcg_jump(sd2[ch].seed, total, R);
  }
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    if (in_range[k] && mask[(is_short ? 16 * grp4 : 0) + sfb4[k]]) {
      const float t = v[0][k];
      v[0][k] = fmul(0.5f, fadd(t, v[1][k]));
      v[1][k] = fmul(0.5f, fsub(t, v[1][k]));
    }
    val[4 * tid + k] = v[0][k];
    val[kFrame + 4 * tid + k] = v[1][k];
  }
  __syncthreads();

  // ---- band energies and gains per (channel, group, band) ----------------------------------------------
  {
    const int ch = tid >> 7, r = tid & 127;
    const int g = is_short ? r >> 4 : 0;
    const int sfb = is_short ? r & 15 : r;
    const bool live = g < sl.num_groups && sfb >= sl.sfb_start && sfb < sl.sfb_stop &&
                      ((sfb - sl.sfb_start) % step) == 0 && (is_short || r < 64);
    if (live) {
      int w0 = 0;
      for (int q = 0; q < g; ++q) w0 += sl.group_len[q];
      const int hi = min(sfb + step, int(sl.sfb_stop));
      float s_acc = 0.0f, p_acc = 0.0f;
      for (int q = 0; q < sl.group_len[g]; ++q) {
        const float* c = coef + ch * kFrame + (w0 + q) * bins;
        const float* sv = val + ch * kFrame + (w0 + q) * bins;
        float e = 0.0f, pe = 0.0f;
        for (int b = swb[sfb]; b < swb[hi]; ++b) {
          const float cb = c[b];
          e = fadd(e, fmul(cb, cb));
          if (cb == 0.0f) pe = fadd(pe, fmul(sv[b], sv[b]));
        }
        s_acc = fadd(s_acc, e);
        p_acc = fadd(p_acc, pe);
      }
      if (sl.group_len[g] != 1) {
        const float gl = float(sl.group_len[g]);
        s_acc = __fdiv_rn(s_acc, gl);
        p_acc = __fdiv_rn(p_acc, gl);
      }
      const int width = swb[hi] - swb[sfb];
      const float lev = sd2[ch].levels[(is_short ? 16 * g : 0) + sfb];
      const float mn = fsub(fmul(fmul(lev, lev), float(width)), s_acc);
      float gain = 0.0f;
      if (0.0f < mn && 0.0f < p_acc) {
        gain = float(__dsqrt_rn(double(__fdiv_rn(mn, p_acc))));
        gain = gain < 10.f ? gain : 10.f;
      }
      const int slot = (is_short ? 16 * g : 0) + sfb;
      gn[ch][slot] = gain;
      if (step == 2 && sfb + 1 < sl.sfb_stop) gn[ch][slot + 1] = gain;
    }
  }
  __syncthreads();

  // ---- fill; TNF buffers with the either-channel-coded rule on the masked bands (:1560-1580) ---------------
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int i = 4 * tid + k;
    float spec[2] = {0.0f, 0.0f}, mk[2] = {1.0f, 1.0f};
    if (in_range[k]) {
      const int slot = (is_short ? 16 * grp4 : 0) + sfb4[k];
#pragma unroll
      for (int ch = 0; ch < 2; ++ch) {
        spec[ch] = fmul(gn[ch][slot], val[ch * kFrame + i]);
        if (coef[ch * kFrame + i] == 0.0f) {
          coef[ch * kFrame + i] = spec[ch];
          mk[ch] = 0.0f;
        }
      }
      if (mask[slot]) mk[0] = mk[1] = (mk[0] != 0.0f || mk[1] != 0.0f) ? 1.0f : 0.0f;
    }
#pragma unroll
    for (int ch = 0; ch < 2; ++ch) tnf[ch * kFrame + i] = spec[ch], tmask[ch * kFrame + i] = mk[ch];
  }
  __syncthreads();
#pragma unroll
  for (int ch = 0; ch < 2; ++ch) {
    reinterpret_cast<float4*>(cg[ch])[tid] = reinterpret_cast<const float4*>(coef + ch * kFrame)[tid];
    if (tnf_on) {
      float* stg = tnf_state + size_t(sd2[ch].unit) * 2 * kFrame;
      reinterpret_cast<float4*>(stg)[tid] = reinterpret_cast<const float4*>(tnf + ch * kFrame)[tid];
      reinterpret_cast<float4*>(stg + kFrame)[tid] = reinterpret_cast<const float4*>(tmask + ch * kFrame)[tid];
    }
  }
}

IgfRom make_rom() {
  IgfRom R{};
  for (int n = 0; n < 64; ++n) {
    char buf[64];
    std::snprintf(buf, sizeof(buf), "%.7f", std::pow(2.0, (42 - n) / 2.0));  // the ROM prints seven decimals
    R.whiten[0][n] = std::strtof(buf, nullptr);
    R.whiten[1][n] = float(2097152.0 * std::pow(2.0, n / 2.0));
  }
  R.whiten[0][41] = 1.4142135f;  // truncated, not rounded, in decoder/ia_core_coder_rom.c:1296
  const float win[8] = {0.997803f, 0.991211f, 0.980225f, 0.964844f, 0.945068f, 0.920898f, 0.892334f, 0.859375f};
  std::memcpy(R.acf_win, win, sizeof(win));
  uint32_t a = 69069u, c = 5u;
  for (int b = 0; b < 11; ++b) {
    R.lcg_a[b] = a, R.lcg_c[b] = c;
    c = a * c + c;
    a = a * a;
  }
  return R;
}

}  // namespace
}  // namespace mpeghb200

using namespace mpeghb200;

static_assert(sizeof(mpeghb200_igf_side) == 548, "mpeghb200_igf_side layout");

extern "C" {

void mpeghb200_igf_whitening_rom(float* out128) {
  if (!out128) return;
  const IgfRom R = make_rom();
  std::memcpy(out128, R.whiten, sizeof(R.whiten));
}

mpeghb200_status mpeghb200_igf_dev(mpeghb200_ctx* ctx, const mpeghb200_bands* bands, float* d_coef,
                                   const float* d_tiles, const mpeghb200_igf_side* d_side, float* d_tnf_state,
                                   uint32_t* d_seed_out, int n_units, void* cuda_stream) {
  if (!ctx || !bands) return MPEGHB200_ERR_BAD_ARG;
  if (n_units < 0 || (n_units > 0 && (!d_coef || !d_tiles || !d_side || !d_tnf_state || !d_seed_out)))
    return fail(ctx, MPEGHB200_ERR_BAD_ARG, "igf_dev: null buffer or negative unit count");
  if (n_units == 0) return MPEGHB200_OK;
  MB_CUDA(ctx, cudaSetDevice(ctx->device));
  static const IgfRom R = make_rom();
  MB_CUDA(ctx, launch_pdl(igf_kernel, dim3(n_units), dim3(kThreads), 0, static_cast<cudaStream_t>(cuda_stream), d_coef,
                          d_tiles, d_side, d_tnf_state, d_seed_out, bands->d_sfb_long, bands->n_long, bands->d_sfb_short,
                          bands->n_short, bands->d_bin2sfb_long, bands->d_bin2sfb_short, 0, R));
  ctx->launches.fetch_add(1);
  MB_CUDA(ctx, cudaGetLastError());
  return MPEGHB200_OK;
}

mpeghb200_status mpeghb200_igf_tnf_dev(mpeghb200_ctx* ctx, const mpeghb200_bands* bands, float* d_coef,
                                       const mpeghb200_igf_side* d_side, float* d_tnf_state, int n_units,
                                       void* cuda_stream) {
  if (!ctx || !bands) return MPEGHB200_ERR_BAD_ARG;
  if (n_units < 0 || (n_units > 0 && (!d_coef || !d_side || !d_tnf_state)))
    return fail(ctx, MPEGHB200_ERR_BAD_ARG, "igf_tnf_dev: null buffer or negative unit count");
  if (n_units == 0) return MPEGHB200_OK;
  MB_CUDA(ctx, cudaSetDevice(ctx->device));
  static const IgfRom R = make_rom();
  MB_CUDA(ctx, launch_pdl(igf_kernel, dim3(n_units), dim3(kThreads), 0, static_cast<cudaStream_t>(cuda_stream), d_coef,
                          static_cast<const float*>(nullptr), d_side, d_tnf_state, static_cast<uint32_t*>(nullptr),
                          bands->d_sfb_long, bands->n_long, bands->d_sfb_short, bands->n_short, bands->d_bin2sfb_long,
                          bands->d_bin2sfb_short, 1, R));
  ctx->launches.fetch_add(1);
  MB_CUDA(ctx, cudaGetLastError());
  return MPEGHB200_OK;
}

mpeghb200_status mpeghb200_igf_stereo_dev(mpeghb200_ctx* ctx, const mpeghb200_bands* bands, float* d_coef,
                                          const float* d_tiles, const mpeghb200_igf_side* d_side,
                                          const uint8_t* d_mask, float* d_tnf_state, uint32_t* d_seed_out,
                                          int n_pairs, void* cuda_stream) {
  if (!ctx || !bands) return MPEGHB200_ERR_BAD_ARG;
  if (n_pairs < 0 || (n_pairs > 0 && (!d_coef || !d_tiles || !d_side || !d_mask || !d_tnf_state || !d_seed_out)))
    return fail(ctx, MPEGHB200_ERR_BAD_ARG, "igf_stereo_dev: null buffer or negative pair count");
  if (n_pairs == 0) return MPEGHB200_OK;
  MB_CUDA(ctx, cudaSetDevice(ctx->device));
  static const IgfRom R = make_rom();
  constexpr size_t smem = size_t(16) * kFrame * sizeof(float);
  // per device / context attribute: set on every call (a process may hold contexts on several GPUs)
  MB_CUDA(ctx, cudaFuncSetAttribute(igf_stereo_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)));
  MB_CUDA(ctx, launch_pdl(igf_stereo_kernel, dim3(n_pairs), dim3(kThreads), smem, static_cast<cudaStream_t>(cuda_stream),
                          d_coef, d_tiles, d_side, d_mask, d_tnf_state, d_seed_out, bands->d_sfb_long, bands->n_long,
                          bands->d_sfb_short, bands->n_short, bands->d_bin2sfb_long, bands->d_bin2sfb_short, R));
  ctx->launches.fetch_add(1);
  MB_CUDA(ctx, cudaGetLastError());
  return MPEGHB200_OK;
}

}  // extern "C"
