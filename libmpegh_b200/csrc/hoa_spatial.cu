// hoa_spatial.cu -- HOA spatial decoding: transport channels -> HOA coefficient signals.
//
// Stands in for the signal half of impeghd_hoa_spatial_process (decoder/impeghd_hoa_spatial_decoder.c:663),
// i.e. everything after impeghd_hoa_spatial_decode_frame_side_info (which stays host code):
//   impeghd_hoa_inv_dyn_correction_process             decoder/impeghd_hoa_inverse_dyn_correction.c:75
//   impeghd_hoa_ambience_synthesis_process             decoder/impeghd_hoa_amb_syn.c:76
//   impeghd_hoa_ch_reassignment_process                decoder/impeghd_hoa_ch_reassignment.c:71
//   impeghd_hoa_vector_based_predom_sound_syn_process  decoder/impeghd_hoa_vector_based_predom_sound_syn.c:76
//   impeghd_hoa_dir_based_pre_dom_sound_syn_process    decoder/impeghd_hoa_dir_based_pre_dom_sound_syn.c:304
//     with ..._compute_hoa_repr_of_dir_sigs :145, ..._smoothed_hoa_representation_of_spat_pred_sigs :362,
//     ..._compute_hoa_representation_of_spat_pred :523
//
// The reference runs these as ~10 passes over [49][1024] frame buffers.  Here every output sample
// out[c][t] is one expression of the T transport samples x[.][t] (gather form, same products, sums and order
// as the reference, so bit-identical); nothing but small side information is carried between frames.
// Two launches per frame:
//   plan kernel    one small CTA per stream: resolves the frame's side info against the carried state into a
//                  flat "plan" (which signals contribute to what, with which cross-fade window, the gain
//                  correction factors), then advances the state exactly like the reference's bookkeeping.
//   render kernel  (stream, 256-sample chunk) CTAs: stage the T corrected transport samples in shared memory,
//                  then each thread evaluates all n_coef outputs of its sample; control flow is uniform per CTA.
// HBM traffic per stream-frame: T * 4 KB in, n_coef * 4 KB out, ~6 KB side info, <= 12 KB plan.
#include <cstddef>
#include <cstring>
#include <new>

#include "common.h"

namespace mpeghb200 {
namespace {

constexpr int kMaxCh = MPEGHB200_HOASP_MAX_CH;
constexpr int kMaxCoef = MPEGHB200_HOASP_MAX_COEF;
constexpr int kMaxPred = MPEGHB200_HOASP_MAX_PRED;
constexpr int kFrame = 1024;
constexpr int kChunk = 256;
using Side = mpeghb200_hoa_spatial_side;

struct Cfg {
  int n_coef, n_transport, n_addnl, low_order, vec_start, coded_v_vec_length, spat_interp_time, pred_dir_sigs,
      num_bits_per_sf;
  const float* order_matrix;  // [low][low]
  const float* fine;          // [900][49]
  const float* coarse;        // [n_coef][n_coef]
  const float* wins;          // [4][1024]: in_vec, out_vec, in_dir, out_dir
  const float* dyn_pow;       // [7][1024]
  const float* dyn_win;       // [1024]
  const float* pow2;          // [128]
};

// carried between frames (the reference's ia_spatial_dec_str members that outlive a frame)
struct State {
  float prev_gain[kMaxCh];
  int prev_ps_set[kMaxCh];
  int num_prev_vec, num_prev_dir, num_prev_ps, pad;
  int prev_vec_index[kMaxCh];
  int old_ps_set[kMaxCh];
  int last_dir_ch[kMaxCh + 1], last_dir_id[kMaxCh + 1];
  int last_pred_type[kMaxCoef];
  int last_pred_idx[kMaxPred][kMaxCoef];
  float last_pred_fac[kMaxPred][kMaxCoef];
  float prev_vec[kMaxCh][kMaxCoef];
};

enum Win : int { kWinNone = 0, kWinInVec = 1, kWinOutVec = 2, kWinInDir = 3, kWinOutDir = 4 };

constexpr int kVecPitch = (kMaxCoef + 3) / 4 * 4;
struct Plan {
  float dyn_a[kMaxCh];
  int dyn_mode[kMaxCh];  // 0: a*x, 1: (x*a)*pow_table[arg][t], 2: pow(win[t], arg) * (a*x)
  int dyn_arg[kMaxCh];
  int clear[kMaxCh];
  int reassign[kMaxCoef];
  int n_vec_e, n_dir_e, n_pred[2];  // pred [0] = last parameters, [1] = current
  int pred_flag[2], pad[2];
  int in_non[kMaxCoef], in_en[kMaxCoef], in_dis[kMaxCoef];
  int vec_ch[2 * kMaxCh], vec_ns[2 * kMaxCh], vec_win[2 * kMaxCh];
  int dir_ch[2 * kMaxCh], dir_grid[2 * kMaxCh], dir_win[2 * kMaxCh];
  int pred_g[2][kMaxCoef], pred_n[2][kMaxCoef];
  int pred_idx[2][kMaxCoef][kMaxPred];
  float pred_fac[2][kMaxCoef][kMaxPred];
  // rows indexed by the coefficient itself (entries below vec_start unused), 16-byte pitch: the render kernel reads
  // them as float4 (every thread reads the same row: one LDS.128 instead of four LDS.32 on the LSU pipe that bounds it)
  alignas(16) float vec_v[2 * kMaxCh][kVecPitch];
};

struct Slot {
  State st;
  Plan plan;
};

__device__ __forceinline__ float pow2f(const Cfg& C, int e) { return e >= 0 ? __ldg(C.pow2 + e) : __ldg(C.pow2 + 64 - e); }

__device__ __forceinline__ int sig_set(const Side& sd, int i) {
  // impeghd_hoa_pre_dom_sound_syn.c:83-91: direction channels first, then overwritten from slot 0 by vectors
  if (i < sd.n_vec) return sd.vec_index[i];
  if (i < sd.n_dir) return sd.dir_ch[i];
  return 0;
}

__global__ void __launch_bounds__(128)
    hoasp_plan_kernel(const Side* __restrict__ sides, Slot* __restrict__ slots, const __grid_constant__ Cfg C) {
  const int tid = threadIdx.x;
  const Side& sd = sides[blockIdx.x];
  State& st = slots[blockIdx.x].st;
  Plan& pl = slots[blockIdx.x].plan;
  __shared__ int s_vec_ok[2 * kMaxCh], s_vec_ns[2 * kMaxCh], s_vec_win[2 * kMaxCh], s_vec_ch[2 * kMaxCh];
  __shared__ int s_dir_ok[2 * kMaxCh], s_dir_win[2 * kMaxCh], s_dir_ch[2 * kMaxCh], s_dir_grid[2 * kMaxCh];
  __shared__ int s_vec_slot[2 * kMaxCh];
  __shared__ float s_fac[kMaxPred][kMaxCoef];
  float new_gain = 0.0f;

  if (tid < kMaxCh) {
    // ---- gain correction factors + next gain (inverse_dyn_correction.c:84-135) and the clearing mask
    const int ch = tid;
    if (ch < C.n_transport) {
      const int e = sd.exponents[ch];
      int offset = (e < 0) ? (-e + 5) : (e - 1);
      if (offset > 6) offset = 6;
      float last = st.prev_gain[ch];
      if (sd.prev_gain_reset[ch] > 0.0f) last = sd.prev_gain_reset[ch];
      const float reci = __double2float_rn(__ddiv_rn(1.0, double(last)));
      int mode = 0, arg = 0;
      float a = reci;
      if (e != 0) {
        if (sd.is_exception[ch]) {
          a = fmul(pow2f(C, e), reci);
          last = fmul(pow2f(C, -e), last);
        } else if (e > 6 || e < -6) {
          mode = 2, arg = e;
          last = fmul(last, __double2float_rn(pow(double(__ldg(C.dyn_win + kFrame - 1)), double(e))));
        } else {
          mode = 1, arg = offset;
          last = __fdiv_rn(last, __ldg(C.dyn_pow + offset * kFrame + kFrame - 1));
        }
      }
      new_gain = last;
      pl.dyn_a[ch] = a, pl.dyn_mode[ch] = mode, pl.dyn_arg[ch] = arg;
      bool keep = ch >= C.n_addnl;
      for (int i = 0; i < sd.n_vec; ++i) keep |= (sd.vec_index[i] - 1 == ch);
      for (int i = 0; i < sd.n_dir; ++i) keep |= (sd.dir_ch[i] - 1 == ch);
      pl.clear[ch] = keep ? 0 : 1;
    }
  } else if (tid < 2 * kMaxCh) {
    // ---- previous frame's vectors, faded out (vector_based...c:94-137)
    const int i = tid - kMaxCh;
    int ok = 0, ns = kFrame, win = kWinOutDir, ch = 0;
    if (i < st.num_prev_vec) {
      const int vi = st.prev_vec_index[i];
      int it = 0;
      for (; it < kMaxCh; ++it)
        if (sig_set(sd, it) == vi) break;
      if (it != kMaxCh) {
        ok = 1, ch = vi - 1;
        int mp = 0;
        for (; mp < kMaxCh; ++mp)
          if (sd.vec_index[mp] == vi) break;
        if (mp != kMaxCh) ns = C.spat_interp_time, win = kWinOutVec;
      }
    }
    s_vec_ok[i] = ok, s_vec_ns[i] = ns, s_vec_win[i] = win, s_vec_ch[i] = ch;
  } else if (tid < 3 * kMaxCh) {
    // ---- this frame's vectors, faded in when the channel was predominant before (:139-166)
    const int i = tid - 2 * kMaxCh;
    int ok = 0, win = kWinNone, ch = 0;
    if (i < sd.n_vec) {
      const int vi = sd.vec_index[i];
      ok = 1, ch = vi - 1;
      for (int it = 0; it < kMaxCh; ++it)
        if (st.prev_ps_set[it] == vi) {
          win = kWinInVec;
          break;
        }
    }
    s_vec_ok[kMaxCh + i] = ok, s_vec_ns[kMaxCh + i] = kFrame, s_vec_win[kMaxCh + i] = win, s_vec_ch[kMaxCh + i] = ch;
  } else if (tid < 4 * kMaxCh) {
    // ---- previous frame's direction signals (dir_based...c:169-224); the "not found" test compares with 49
    const int i = tid - 3 * kMaxCh;
    int ok = 0, win = kWinNone, ch = 0, grid = 0;
    if (i < st.num_prev_dir && st.last_dir_id[i] != 0) {
      grid = st.last_dir_id[i];
      const int dch = st.last_dir_ch[i];
      ok = 1, ch = dch - 1;
      int di = 0;
      for (; di < sd.n_dir; ++di)
        if (sd.dir_ch[di] == dch) break;
      if (di == kMaxCoef) {
        for (int q = 0; q < kMaxCh; ++q)
          if (sig_set(sd, q) == dch) {
            win = kWinOutVec;
            break;
          }
      } else if (sd.dir_id[di] != 0) {
        win = kWinOutDir;
      }
    }
    s_dir_ok[i] = ok, s_dir_win[i] = win, s_dir_ch[i] = ch, s_dir_grid[i] = grid;
  } else if (tid < 5 * kMaxCh) {
    // ---- this frame's direction signals (:226-279)
    const int i = tid - 4 * kMaxCh;
    int ok = 0, win = kWinNone, ch = 0, grid = 0;
    if (i < sd.n_dir && sd.dir_id[i] != 0) {
      grid = sd.dir_id[i];
      const int dch = sd.dir_ch[i];
      ok = 1, ch = dch - 1;
      int di = 0;
      for (; di < st.num_prev_dir; ++di)
        if (dch == st.last_dir_ch[di]) break;
      if (di == kMaxCoef) {
        for (int q = 0; q < st.num_prev_ps; ++q)
          if (dch == st.old_ps_set[q]) {
            win = kWinInDir;
            break;
          }
      } else if (st.last_dir_id[di] != 0) {
        win = kWinInDir;
      }
    }
    s_dir_ok[kMaxCh + i] = ok, s_dir_win[kMaxCh + i] = win, s_dir_ch[kMaxCh + i] = ch, s_dir_grid[kMaxCh + i] = grid;
  }
  // ---- per coefficient: reassignment source, membership in the three coefficient sets, dequantised factors
  if (tid < kMaxCoef) {
    const int c = tid;
    int src = -1;
    for (int ch = 0; ch < C.n_transport; ++ch)
      if (sd.amb_hoa_assign[ch] - 1 == c && sd.amb_hoa_assign[ch] != 0) src = ch;
    pl.reassign[c] = src;
    int in_non = 0, in_en = 0, in_dis = 0;
    for (int i = 0; i < kMaxCoef; ++i) in_non |= (sd.non_en_dis[i] == c + 1);
    for (int i = 0; i < sd.n_enable; ++i) in_en += (sd.enable[i] - 1 == c);
    for (int i = 0; i < sd.n_disable; ++i) in_dis += (sd.disable[i] - 1 == c);
    pl.in_non[c] = in_non, pl.in_en[c] = in_en, pl.in_dis[c] = in_dis;
    if (c < C.n_coef) {
      const float scale = pow2f(C, 1 - C.num_bits_per_sf);
      for (int r = 0; r < C.pred_dir_sigs; ++r)
        s_fac[r][c] = __double2float_rn(double(scale) * (double(float(sd.pred_quant[r][c])) + 0.5));
    }
  }
  __syncthreads();

  // ---- compaction (order preserved) and the prediction lists ---------------------------------------------
  if (tid == 0) {
    int n = 0;
    for (int i = 0; i < 2 * kMaxCh; ++i)
      if (s_vec_ok[i]) {
        pl.vec_ch[n] = s_vec_ch[i], pl.vec_ns[n] = s_vec_ns[i], pl.vec_win[n] = s_vec_win[i];
        s_vec_slot[n] = i;
        ++n;
      }
    pl.n_vec_e = n;
    for (int i = n; i < 2 * kMaxCh; ++i) s_vec_slot[i] = -1;
  } else if (tid == 32) {
    int n = 0;
    for (int i = 0; i < 2 * kMaxCh; ++i)
      if (s_dir_ok[i]) {
        pl.dir_ch[n] = s_dir_ch[i], pl.dir_grid[n] = s_dir_grid[i], pl.dir_win[n] = s_dir_win[i];
        ++n;
      }
    pl.n_dir_e = n;
  } else if (tid == 64 || tid == 96) {
    const int k = (tid == 96) ? 1 : 0;  // 0 = last frame's parameters, 1 = this frame's
    int n = 0;
    if (C.n_addnl != 0) {
      for (int g = 0; g < C.n_coef; ++g) {
        const int type = k ? sd.pred_type[g] : st.last_pred_type[g];
        if (type == 0) continue;
        int nv = 0;
        while (nv < C.pred_dir_sigs && (k ? sd.pred_idx[nv][g] : st.last_pred_idx[nv][g]) != 0) ++nv;
        pl.pred_g[k][n] = g, pl.pred_n[k][n] = nv;
        for (int r = 0; r < nv; ++r) {
          pl.pred_idx[k][n][r] = k ? sd.pred_idx[r][g] : st.last_pred_idx[r][g];
          pl.pred_fac[k][n][r] = k ? s_fac[r][g] : st.last_pred_fac[r][g];
        }
        ++n;
      }
    }
    pl.n_pred[k] = n;
    pl.pred_flag[k] = n > 0;
  }
  __syncthreads();
  {
    const int nel = C.n_coef - C.vec_start;
    for (int idx = tid; idx < 2 * kMaxCh * kMaxCoef; idx += blockDim.x) {
      const int n = idx / kMaxCoef, e = idx - n * kMaxCoef;
      const int src = s_vec_slot[n];
      if (src < 0 || e >= nel) continue;
      pl.vec_v[n][e + C.vec_start] = (src < kMaxCh) ? st.prev_vec[src][e] : sd.vec[src - kMaxCh][e];
    }
  }
  __syncthreads();

  // ---- state for the next frame (vector_based...c:168-224, dir_based...c:281-288, :496-510) ----------------
  if (tid < kMaxCh) {
    if (tid < C.n_transport && !(sd.exponents[tid] == 0 && !(sd.prev_gain_reset[tid] > 0.0f)))
      st.prev_gain[tid] = (sd.exponents[tid] == 0) ? sd.prev_gain_reset[tid] : new_gain;
    st.prev_ps_set[tid] = sig_set(sd, tid);
    if (tid < sd.n_vec) {
      st.prev_vec_index[tid] = sd.vec_index[tid];
      st.old_ps_set[tid] = sig_set(sd, tid);
    }
    if (tid < sd.n_dir) {
      st.last_dir_ch[tid] = sd.dir_ch[tid];
      st.last_dir_id[tid] = sd.dir_id[tid];
    }
    if (tid == 0) {
      st.num_prev_vec = sd.n_vec;
      st.num_prev_ps = sd.n_vec;
      st.num_prev_dir = sd.n_dir;
    }
  }
  for (int idx = tid; idx < sd.n_vec * kMaxCoef; idx += blockDim.x) {
    const int i = idx / kMaxCoef, e = idx - i * kMaxCoef;
    float v = sd.vec[i][e];
    if (C.coded_v_vec_length == 1)
      for (int k = 0; k < sd.n_enable; ++k)
        if (sd.enable[k] - 1 - C.vec_start == e) v = 0.0f;
    st.prev_vec[i][e] = v;
  }
  if (tid < C.n_coef) {
    st.last_pred_type[tid] = sd.pred_type[tid];
    for (int r = 0; r < C.pred_dir_sigs; ++r) {
      st.last_pred_fac[r][tid] = s_fac[r][tid];
      st.last_pred_idx[r][tid] = sd.pred_idx[r][tid];
    }
  }
}

__global__ void hoasp_state_init_kernel(Slot* slots, int n) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= n) return;
  State& st = slots[s].st;
  for (int i = 0; i < kMaxCh; ++i) st.prev_gain[i] = 1.0f, st.prev_ps_set[i] = -1;
}

__device__ __forceinline__ float win_mul(float v, int win, const float (&w)[4]) {
  switch (win) {
    case kWinInVec: return fmul(v, w[0]);
    case kWinOutVec: return fmul(v, w[1]);
    case kWinInDir: return fmul(v, w[2]);
    case kWinOutDir: return fmul(v, w[3]);
    default: return v;
  }
}

constexpr int kPredCache = 6;

// |exponent| > 6: the reference's per-sample double pow (kept out of line: rare, and it would bloat the kernel)
__device__ __noinline__ float dyn_pow_path(float win, int e, float v) {
  return fmul(__double2float_rn(pow(double(win), double(e))), v);
}

// NC = (order + 1)^2 at compile time: the per-coefficient accumulators of a sample live in registers and the
// contribution lists are walked entry-outermost (the order of additions per coefficient stays the reference's).
template <int NC>
__global__ void __launch_bounds__(kChunk, (NC <= 25) ? 3 : 2)
    hoasp_render_kernel(const float* __restrict__ in, float* __restrict__ out, const Slot* __restrict__ slots,
                        const __grid_constant__ Cfg C) {
  __shared__ float xs[kMaxCh][kChunk];        // corrected transport samples (before clearing)
  __shared__ __align__(16) Plan pl;           // this stream's plan: every thread walks it, so keep it on chip
  constexpr int NC4 = (NC + 3) / 4;
  __shared__ __align__(16) float fine_rows[2 * kMaxCh][NC4 * 4]; // mode-matrix rows of the direction entries
  const int tid = threadIdx.x;
  {
    const Plan& gp = slots[blockIdx.x].plan;
    const uint32_t* src = reinterpret_cast<const uint32_t*>(&gp);
    uint32_t* dst = reinterpret_cast<uint32_t*>(&pl);
    constexpr int kHead = int(offsetof(Plan, pred_g) / 4), kPredEnd = int(offsetof(Plan, vec_v) / 4);
    for (int i = tid; i < kHead; i += kChunk) dst[i] = src[i];
    if (gp.n_pred[0] | gp.n_pred[1])
      for (int i = kHead + tid; i < kPredEnd; i += kChunk) dst[i] = src[i];
    const int nv = gp.n_vec_e * kVecPitch;
    for (int i = tid; i < nv; i += kChunk) dst[kPredEnd + i] = src[kPredEnd + i];
    const int nd = gp.n_dir_e;
    for (int i = tid; i < nd * NC; i += kChunk) {
      const int n = i / NC, c = i - n * NC;
      fine_rows[n][c] = __ldg(C.fine + (gp.dir_grid[n] - 1) * kMaxCoef + c);
    }
  }
  const int t = blockIdx.y * kChunk + tid;
  const float* x = in + size_t(blockIdx.x) * C.n_transport * kFrame;
  float* y = out + size_t(blockIdx.x) * NC * kFrame;
  const float w[4] = {__ldg(C.wins + t), __ldg(C.wins + kFrame + t), __ldg(C.wins + 2 * kFrame + t),
                      __ldg(C.wins + 3 * kFrame + t)};
  {
    // all transport samples of this thread's column are requested before the first one is used
    float xv[kMaxCh];
#pragma unroll
    for (int ch = 0; ch < kMaxCh; ++ch) xv[ch] = (ch < C.n_transport) ? __ldcs(x + size_t(ch) * kFrame + t) : 0.0f;
    __syncthreads();  // plan staged
#pragma unroll
    for (int ch = 0; ch < kMaxCh; ++ch) {
      if (ch < C.n_transport) {
        const float v = xv[ch];
        const float a = pl.dyn_a[ch];
        const int mode = pl.dyn_mode[ch], arg = pl.dyn_arg[ch];
        float r;
        if (mode == 0) {
          r = fmul(a, v);
        } else if (mode == 1) {
          r = fmul(fmul(v, a), __ldg(C.dyn_pow + arg * kFrame + t));
        } else {
          r = dyn_pow_path(__ldg(C.dyn_win + t), arg, fmul(a, v));
        }
        xs[ch][tid] = r;
      }
    }
  }
  // own column only: no barrier needed.  xc(ch) = sample after the channel clearing
  auto xc = [&](int ch) { return pl.clear[ch] ? 0.0f : xs[ch][tid]; };
  auto wsel = [&](int win) { return win == kWinInVec ? w[0] : win == kWinOutVec ? w[1] : win == kWinInDir ? w[2] : w[3]; };

  // ---- vector-based part: v[c] += (V[n][c] * win) * x ---------------------------------------------------
  float v[NC];
#pragma unroll
  for (int c = 0; c < NC; ++c) v[c] = 0.0f;
  const int vs = C.vec_start;
  for (int n = 0; n < pl.n_vec_e; ++n) {
    if (t >= pl.vec_ns[n]) continue;
    const float xv = xc(pl.vec_ch[n]);
    const int win = pl.vec_win[n];
    const float4* V4 = reinterpret_cast<const float4*>(pl.vec_v[n]);
    if (win == kWinNone) {
#pragma unroll
      for (int q = 0; q < NC4; ++q) {
        const float4 a = V4[q];
        const float V[4] = {a.x, a.y, a.z, a.w};
#pragma unroll
        for (int e = 0; e < 4; ++e)
          if (4 * q + e < NC && 4 * q + e >= vs) v[4 * q + e] = fadd(v[4 * q + e], fmul(V[e], xv));
      }
    } else {
      const float wv = wsel(win);
#pragma unroll
      for (int q = 0; q < NC4; ++q) {
        const float4 a = V4[q];
        const float V[4] = {a.x, a.y, a.z, a.w};
#pragma unroll
        for (int e = 0; e < 4; ++e)
          if (4 * q + e < NC && 4 * q + e >= vs) v[4 * q + e] = fadd(v[4 * q + e], fmul(fmul(V[e], wv), xv));
      }
    }
  }
  // ---- direction signals: d[c] += (x * M[grid][c]) * win ------------------------------------------------------
  float d[NC];
#pragma unroll
  for (int c = 0; c < NC; ++c) d[c] = 0.0f;
  for (int n = 0; n < pl.n_dir_e; ++n) {
    const float xv = xc(pl.dir_ch[n]);
    const int win = pl.dir_win[n];
    const float4* F4 = reinterpret_cast<const float4*>(fine_rows[n]);
    if (win == kWinNone) {
#pragma unroll
      for (int q = 0; q < NC4; ++q) {
        const float4 a = F4[q];
        const float F[4] = {a.x, a.y, a.z, a.w};
#pragma unroll
        for (int e = 0; e < 4; ++e)
          if (4 * q + e < NC) d[4 * q + e] = fadd(d[4 * q + e], fmul(xv, F[e]));
      }
    } else {
      const float wv = wsel(win);
#pragma unroll
      for (int q = 0; q < NC4; ++q) {
        const float4 a = F4[q];
        const float F[4] = {a.x, a.y, a.z, a.w};
#pragma unroll
        for (int e = 0; e < 4; ++e)
          if (4 * q + e < NC) d[4 * q + e] = fadd(d[4 * q + e], fmul(fmul(xv, F[e]), wv));
      }
    }
  }

  // predicted grid signals of the first few active grid points are kept, the rest recomputed per coefficient
  auto pred_tmp = [&](int k, int n) {
    float a = 0.0f;
    const int nv = pl.pred_n[k][n];
    for (int s = 0; s < nv; ++s) a = fadd(a, fmul(pl.pred_fac[k][n][s], xc(pl.pred_idx[k][n][s] - 1)));
    return fmul(a, k ? w[2] : w[3]);
  };
  const bool any_pred = (pl.n_pred[0] | pl.n_pred[1]) != 0;
  float tcache[2][kPredCache];
#pragma unroll
  for (int k = 0; k < 2; ++k)
#pragma unroll
    for (int n = 0; n < kPredCache; ++n) tcache[k][n] = (any_pred && n < pl.n_pred[k]) ? pred_tmp(k, n) : 0.0f;

  // Ambience + spatial prediction per coefficient in a ROLLED loop (its body is big: unrolled 25 times it thrashed
  // the instruction cache -- ncu: most stall samples of the first version were "no instruction"); the sum is parked
  // in a shared-memory column this thread owns (a round trip through the output element itself cost an L2 latency
  // per coefficient) and picked up again by the small unrolled loop below, which needs the register arrays v[], d[].
  extern __shared__ float ap_s[];  // [NC][kChunk]: the parked sums, one column per thread
  // Frames without spatial prediction (the usual case; the flag is the same for the whole CTA) do not need that loop:
  // for a coefficient above the ambient order the sum is one reassigned transport sample, which the unrolled loop below
  // fetches itself, and the few ambient-order coefficients are a small matrix product whose inputs are parked once in
  // the (then unused) top rows of ap_s.  ncu of the rolled loop with the bench's side info: 38 instructions per
  // coefficient and sample for one load and one addition, 49 % of the kernel's instructions.  Same operations in the
  // same order, so bit-identical; frames with prediction (or an ambient order that covers more than half of the
  // coefficients) take the general loop.
  const int low = C.low_order;
  const bool direct = !any_pred && 2 * low <= NC;
  if (direct) {
    for (int m = 0; m < low; ++m) {
      const int src = pl.reassign[m];
      ap_s[(NC - 1 - m) * kChunk + tid] = src < 0 ? 0.0f : xs[src][tid];
    }
    for (int c = 0; c < low; ++c) {
      float amb = 0.0f;
      for (int m = 0; m < low; ++m)
        amb = fadd(amb, fmul(__ldg(C.order_matrix + c * low + m), ap_s[(NC - 1 - m) * kChunk + tid]));
      ap_s[c * kChunk + tid] = fadd(amb, 0.0f);
    }
  }
#pragma unroll 1
  for (int c = direct ? NC : 0; c < NC; ++c) {
    const int in_en = pl.in_en[c], in_dis = pl.in_dis[c];
    float amb;
    if (c >= C.low_order) {
      const int src = pl.reassign[c];
      amb = (src < 0) ? 0.0f : xs[src][tid];
    } else {
      amb = 0.0f;
      for (int m = 0; m < C.low_order; ++m) {
        const int src = pl.reassign[m];
        amb = fadd(amb, fmul(__ldg(C.order_matrix + c * C.low_order + m), src < 0 ? 0.0f : xs[src][tid]));
      }
    }
    float pred = 0.0f;
    if (any_pred) {
      float pk[2];
#pragma unroll
      for (int k = 0; k < 2; ++k) {
        float p = 0.0f;
        const int np = pl.n_pred[k];
        for (int n = 0; n < np; ++n) {
          float tv;
          if (n < kPredCache) {
            tv = 0.0f;
#pragma unroll
            for (int q = 0; q < kPredCache; ++q)
              if (q == n) tv = tcache[k][q];
          } else {
            tv = pred_tmp(k, n);
          }
          p = fadd(p, fmul(tv, __ldg(C.coarse + c * NC + pl.pred_g[k][n])));
        }
        pk[k] = p;
      }
      pred = pl.in_non[c] ? 0.0f : fadd(pk[1], pk[0]);
      if (pl.pred_flag[1])
        for (int i = 0; i < in_en; ++i) pred = fmul(pred, w[3]);
      if (pl.pred_flag[0])
        for (int i = 0; i < in_dis; ++i) pred = fmul(pred, w[2]);
    }
    ap_s[c * kChunk + tid] = fadd(amb, pred);
  }
#pragma unroll
  for (int c = 0; c < NC; ++c) {
    float vv = v[c];
    if (C.coded_v_vec_length == 1) {
      const int in_en = pl.in_en[c], in_dis = pl.in_dis[c];
      for (int i = 0; i < in_en; ++i) vv = fmul(vv, w[3]);
      for (int i = 0; i < in_dis; ++i) vv = fmul(vv, w[2]);
    }
    float ap;
    if (direct && c >= low) {
      const int src = pl.reassign[c];
      ap = fadd(src < 0 ? 0.0f : xs[src][tid], 0.0f);
    } else {
      ap = ap_s[c * kChunk + tid];
    }
    __stcs(y + size_t(c) * kFrame + t, fadd(vv, fadd(ap, d[c])));
  }
}

template <int NC>
cudaError_t launch_render(const float* in, float* out, const Slot* slots, const Cfg& cfg, int n_streams,
                          cudaStream_t st) {
  const int smem = NC * kChunk * int(sizeof(float));
  cudaError_t e = cudaFuncSetAttribute(hoasp_render_kernel<NC>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  if (e != cudaSuccess) return e;
  hoasp_render_kernel<NC><<<dim3(n_streams, kFrame / kChunk), kChunk, smem, st>>>(in, out, slots, cfg);
  return cudaGetLastError();
}

}  // namespace
}  // namespace mpeghb200

using namespace mpeghb200;

struct mpeghb200_hoa_spatial {
  mpeghb200_ctx* ctx = nullptr;
  Cfg cfg{};
  float* d_tables = nullptr;
};

extern "C" {

mpeghb200_status mpeghb200_hoa_spatial_create(mpeghb200_ctx* ctx, const mpeghb200_hoa_spatial_desc* d,
                                              mpeghb200_hoa_spatial** out) {
  if (!ctx || !out || !d) return MPEGHB200_ERR_BAD_ARG;
  *out = nullptr;
  const int nc = (d->order + 1) * (d->order + 1);
  if (d->order < 1 || d->order > 6 || d->n_transport < 1 || d->n_transport > kMaxCh || d->n_addnl_coders < 0 ||
      d->n_addnl_coders > d->n_transport || d->low_order_mat_sz < 1 || d->low_order_mat_sz > nc ||
      d->vec_start < 0 || d->vec_start >= nc || d->pred_dir_sigs < 0 || d->pred_dir_sigs > kMaxPred ||
      d->spat_interp_time < 0 || d->spat_interp_time > kFrame || d->num_bits_per_sf < 1 || d->num_bits_per_sf > 32 ||
      !d->order_matrix || !d->fine_grid_mode_matrix || !d->coarse_grid_mode_matrix || !d->fade_windows ||
      !d->dyn_win_pow_table || !d->dyn_win_func || !d->pow2_table)
    return fail(ctx, MPEGHB200_ERR_BAD_ARG, "hoa_spatial_create: parameter out of range or null table");
  if (d->use_phase_shift_decorr)
    return fail(ctx, MPEGHB200_ERR_UNSUPPORTED,
                "hoa_spatial_create: use_phase_shift_decorr (the reference leaves the low-order ambience "
                "untouched in that mode, decoder/impeghd_hoa_amb_syn.c:124)");
  MB_CUDA(ctx, cudaSetDevice(ctx->device));
  auto* h = new (std::nothrow) mpeghb200_hoa_spatial();
  if (!h) return MPEGHB200_ERR_NO_MEM;
  h->ctx = ctx;
  const int low = d->low_order_mat_sz;
  const size_t n_order = size_t(low) * low, n_fine = 900 * size_t(kMaxCoef), n_coarse = size_t(nc) * nc;
  const size_t total = n_order + n_fine + n_coarse + 4 * kFrame + 7 * kFrame + kFrame + 128;
  cudaError_t e = cudaMalloc(&h->d_tables, total * sizeof(float));
  if (e != cudaSuccess) {
    delete h;
    return cuda_fail(ctx, e, "hoa_spatial_create");
  }
  float* p = h->d_tables;
  auto put = [&](const float* src, size_t n) {
    const float* at = p;
    if (e == cudaSuccess) e = cudaMemcpy(p, src, n * sizeof(float), cudaMemcpyHostToDevice);
    p += n;
    return at;
  };
  Cfg& c = h->cfg;
  c.n_coef = nc, c.n_transport = d->n_transport, c.n_addnl = d->n_addnl_coders, c.low_order = low;
  c.vec_start = d->vec_start, c.coded_v_vec_length = d->coded_v_vec_length, c.spat_interp_time = d->spat_interp_time;
  c.pred_dir_sigs = d->pred_dir_sigs, c.num_bits_per_sf = d->num_bits_per_sf;
  c.order_matrix = put(d->order_matrix, n_order);
  c.fine = put(d->fine_grid_mode_matrix, n_fine);
  c.coarse = put(d->coarse_grid_mode_matrix, n_coarse);
  c.wins = put(d->fade_windows, 4 * kFrame);
  c.dyn_pow = put(d->dyn_win_pow_table, 7 * kFrame);
  c.dyn_win = put(d->dyn_win_func, kFrame);
  c.pow2 = put(d->pow2_table, 128);
  if (e != cudaSuccess) {
    cudaFree(h->d_tables);
    delete h;
    return cuda_fail(ctx, e, "hoa_spatial_create upload");
  }
  *out = h;
  return MPEGHB200_OK;
}

void mpeghb200_hoa_spatial_destroy(mpeghb200_hoa_spatial* h) {
  if (!h) return;
  cudaSetDevice(h->ctx->device);
  cudaFree(h->d_tables);
  delete h;
}

size_t mpeghb200_hoa_spatial_state_bytes(const mpeghb200_hoa_spatial* h) { return h ? sizeof(Slot) : 0; }
int mpeghb200_hoa_spatial_num_coeffs(const mpeghb200_hoa_spatial* h) { return h ? h->cfg.n_coef : 0; }
int mpeghb200_hoa_spatial_num_transport(const mpeghb200_hoa_spatial* h) { return h ? h->cfg.n_transport : 0; }

mpeghb200_status mpeghb200_hoa_spatial_state_init_dev(mpeghb200_ctx* ctx, const mpeghb200_hoa_spatial* h,
                                                      void* d_state, int n_streams, void* cuda_stream) {
  if (!ctx || !h || (n_streams > 0 && !d_state) || n_streams < 0) return MPEGHB200_ERR_BAD_ARG;
  if (n_streams == 0) return MPEGHB200_OK;
  MB_CUDA(ctx, cudaSetDevice(ctx->device));
  cudaStream_t st = static_cast<cudaStream_t>(cuda_stream);
  MB_CUDA(ctx, cudaMemsetAsync(d_state, 0, sizeof(Slot) * size_t(n_streams), st));
  hoasp_state_init_kernel<<<(n_streams + 127) / 128, 128, 0, st>>>(static_cast<Slot*>(d_state), n_streams);
  ctx->launches.fetch_add(1);
  MB_CUDA(ctx, cudaGetLastError());
  return MPEGHB200_OK;
}

mpeghb200_status mpeghb200_hoa_spatial_dev(mpeghb200_ctx* ctx, const mpeghb200_hoa_spatial* h,
                                           const mpeghb200_hoa_spatial_side* d_side, const float* d_in, float* d_out,
                                           void* d_state, int n_streams, void* cuda_stream) {
  if (!ctx || !h) return MPEGHB200_ERR_BAD_ARG;
  if (n_streams < 0 || (n_streams > 0 && (!d_side || !d_in || !d_out || !d_state)))
    return fail(ctx, MPEGHB200_ERR_BAD_ARG, "hoa_spatial_dev: null buffer or negative stream count");
  if (n_streams == 0) return MPEGHB200_OK;
  MB_CUDA(ctx, cudaSetDevice(ctx->device));
  cudaStream_t st = static_cast<cudaStream_t>(cuda_stream);
  Slot* slots = static_cast<Slot*>(d_state);
  hoasp_plan_kernel<<<n_streams, 128, 0, st>>>(d_side, slots, h->cfg);
  cudaError_t e;
  switch (h->cfg.n_coef) {
    case 4: e = launch_render<4>(d_in, d_out, slots, h->cfg, n_streams, st); break;
    case 9: e = launch_render<9>(d_in, d_out, slots, h->cfg, n_streams, st); break;
    case 16: e = launch_render<16>(d_in, d_out, slots, h->cfg, n_streams, st); break;
    case 25: e = launch_render<25>(d_in, d_out, slots, h->cfg, n_streams, st); break;
    case 36: e = launch_render<36>(d_in, d_out, slots, h->cfg, n_streams, st); break;
    default: e = launch_render<49>(d_in, d_out, slots, h->cfg, n_streams, st); break;
  }
  ctx->launches.fetch_add(2);
  MB_CUDA(ctx, e);
  return MPEGHB200_OK;
}

}  // extern "C"
