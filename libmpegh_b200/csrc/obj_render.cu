// obj_render.cu -- object renderer: VBAP gains (triangle search, MDAP spread, imaginary-speaker downmix,
// power normalisation) and the ramped mix of object signals into the loudspeaker channels.
//
// Stands in for impeghd_obj_renderer_dec (decoder/impeghd_obj_ren_dec.c:1319), impegh_obj_ren_vbap_process
// (decoder/impeghd_obj_ren_vbap.c:485), impeghd_calc_vbap (:437), impeghd_calc_one_source_point (:156),
// impeghd_spreading_calc_mdap_vectors (:336), impeghd_calc_spread_gains (:113), impeghd_nrm_values (:82).
// The layout-dependent tables (triangulation, inverse matrices, imaginary-speaker downmix) are produced once
// per layout by the reference's host init (impeghd_obj_renderer_dec_init, :1057) and handed over as a
// descriptor; the per-object trigonometry (double-precision libm sin/cos/tan, not bit-reproducible on the GPU)
// is done by the host helper mpeghb200_obj_side_from_polar.
//
// Two kernels:
//   obj_gains_kernel  one thread per (stream, object): final gains -> ramp start/step per output channel,
//                     state update (previous gains, first-frame flag).  Tiny.
//   obj_mix_kernel    out[ch][j] = sum_obj in[obj][j] * scale_{obj,ch}(j), scale_{j+1} = scale_j + step.
//                     The running scale is a chain of separately rounded additions (a closed form drifts by
//                     an ulp and flips PCM LSBs).  The chain is evaluated once per (object, channel) to get the
//                     scale at the 16 chunk boundaries; after that the 16 chunks are independent and a thread
//                     walks its 64 samples with the 32 per-object scales in registers.  Object sums follow the
//                     reference's order (object 0 first, starting from the zeroed output buffer).
// Per stream-frame (32 objects -> 7.1.4): reads 32*4096 B, writes 12*4096 B; 1.08 MFLOP of unfused mul/add:
// at the HBM roofline that is ~36 TFLOP/s of FP32 issue, so this stage is FP-issue bound, not HBM bound.
#include <cmath>
#include <cstring>
#include <new>

#include "common.h"

struct mpeghb200_obj_layout {
  mpeghb200_ctx* ctx = nullptr;
  mpeghb200_obj_layout_desc host{};
  void* d_desc = nullptr;  // device copy of DevLayout
  int nonlfe_to_spk[MPEGHB200_OBJ_MAX_LS] = {};
};

namespace mpeghb200 {
namespace {

constexpr int kMaxLs = MPEGHB200_OBJ_MAX_LS;  // 44
constexpr int kMaxObj = 32;
constexpr int kFrame = 1024;
constexpr float kDeg2Rad = float(M_PI / 180.0f);
constexpr float kMinVecCtrDist = 0.01f;

struct DevLayout {
  int n_spk, n_non_lfe, n_imag, n_tri, height_present;
  float unit_gain;  // (float)(1 / sqrt((float)tot)), host double
  int tri[kMaxLs][3];
  float inv[kMaxLs][9];
  float dmx[kMaxLs][kMaxLs];
  int nonlfe_to_spk[kMaxLs];
  int lfe_spk[kMaxLs];
  int n_lfe;
};

__device__ void one_source_point(const DevLayout& L, float vx, float vy, float vz, float* ls) {
  int winner = 0, hi_lo = 3;
  float highest_least = -100000.f, g0 = 0.f, g1 = 0.f, g2 = 0.f;
  for (int i = 0; i < L.n_tri; ++i) {
    float gains[3], min_gain = 10000000.f;
    int neg = 3;
#pragma unroll
    for (int j = 2; j >= 0; --j) {
      const float* r = &L.inv[i][j * 3];
      float gj = fmul(vx, __ldg(r));
      gj = fadd(gj, fmul(vy, __ldg(r + 1)));
      gj = fadd(gj, fmul(vz, __ldg(r + 2)));
      gains[j] = gj;
      if (gj >= -kMinVecCtrDist) neg--;
      if (gj < min_gain) min_gain = gj;
    }
    if ((neg < hi_lo) || (neg <= hi_lo && min_gain > highest_least)) {
      g0 = gains[0], g1 = gains[1], g2 = gains[2];
      highest_least = min_gain;
      hi_lo = neg;
      winner = i;
    }
  }
  if (g2 < 0.f) g2 = 0.f;  // covers both the -0.01 test (:204-210) and the < 0 clamps (:224-235)
  if (g1 < 0.f) g1 = 0.f;
  if (g0 < 0.f) g0 = 0.f;
  const float power = __fsqrt_rn(fadd(fadd(fmul(g0, g0), fmul(g1, g1)), fmul(g2, g2)));
  if (power != 0.f) {
    g0 = __fdiv_rn(g0, power);
    g1 = __fdiv_rn(g1, power);
    g2 = __fdiv_rn(g2, power);
  }
  const int i2 = __ldg(&L.tri[winner][2]), i1 = __ldg(&L.tri[winner][1]), i0 = __ldg(&L.tri[winner][0]);
  ls[i2] = fadd(g2, ls[i2]);
  ls[i1] = fadd(g1, ls[i1]);
  ls[i0] = fadd(g0, ls[i0]);
}

struct V3 {
  float x, y, z;
};
__device__ __forceinline__ V3 sva(V3 a, float ga, V3 b, float gb) {
  V3 r;
  r.z = fadd(fmul(ga, a.z), fmul(gb, b.z));
  r.y = fadd(fmul(ga, a.y), fmul(gb, b.y));
  r.x = fadd(fmul(ga, a.x), fmul(gb, b.x));
  return r;
}

// scratch per (stream, object): [n_non_lfe][2] = {ramp start, ramp step}
__global__ void __launch_bounds__(128)
    obj_gains_kernel(const DevLayout* __restrict__ Lp, const mpeghb200_obj_side* __restrict__ side,
                     float* __restrict__ state, float* __restrict__ scratch, int n_obj, int n_streams) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= n_obj * n_streams) return;
  const DevLayout& L = *Lp;
  const mpeghb200_obj_side s = side[idx];
  const int NL = L.n_non_lfe, tot = NL + L.n_imag;
  float ls[kMaxLs];
  for (int i = 0; i < tot; ++i) ls[i] = 0.f;
  if (s.spread_w <= 0.0f && s.spread_h <= 0.0f) {
    one_source_point(L, s.vec_c[0], s.vec_c[1], s.vec_c[2], ls);
  } else {
    V3 p0{s.p0[0], s.p0[1], s.p0[2]}, v{s.v[0], s.v[1], s.v[2]}, u;
    u.z = fsub(fmul(v.x, p0.y), fmul(v.y, p0.x));
    u.y = fsub(fmul(v.z, p0.x), fmul(v.x, p0.z));
    u.x = fsub(fmul(v.y, p0.z), fmul(v.z, p0.y));
    if (!L.height_present) v.x = v.y = v.z = 0.0f;
    V3 m[19];
    m[0] = V3{s.vec_c[0], s.vec_c[1], s.vec_c[2]};
    m[1] = u;
    m[2] = sva(u, 0.75f, p0, 0.25f);
    m[3] = sva(u, 0.375f, p0, 0.625f);
    m[4] = V3{-u.x, -u.y, -u.z};
    m[5] = sva(u, -0.75f, p0, 0.25f);
    m[6] = sva(u, -0.375f, p0, 0.625f);
    m[7] = sva(sva(u, 0.5f, v, 0.866f), 1.0f, p0, 0.333f);
    m[8] = sva(m[7], 0.5f, p0, 0.5f);
    m[9] = sva(m[7], 0.25f, p0, 0.75f);
    m[10] = sva(sva(u, -0.5f, v, 0.866f), 1.0f, p0, 0.333f);
    m[11] = sva(m[10], 0.5f, p0, 0.5f);
    m[12] = sva(m[10], 0.25f, p0, 0.75f);
    m[13] = sva(sva(u, -0.5f, v, -0.866f), 1.0f, p0, 0.333f);
    m[14] = sva(m[13], 0.5f, p0, 0.5f);
    m[15] = sva(m[13], 0.25f, p0, 0.75f);
    m[16] = sva(sva(u, 0.5f, v, -0.866f), 1.0f, p0, 0.333f);
    m[17] = sva(m[16], 0.5f, p0, 0.5f);
    m[18] = sva(m[16], 0.25f, p0, 0.75f);
    const float px = __fdiv_rn(p0.x, s.tan_alpha), py = __fdiv_rn(p0.y, s.tan_alpha), pz = __fdiv_rn(p0.z, s.tan_alpha);
    for (int i = 0; i < 19; ++i) {
      float x = fadd(m[i].x, px), y = fadd(m[i].y, py), z = fadd(m[i].z, pz);
      const float vl = __fsqrt_rn(fadd(fmul(x, x), fadd(fmul(y, y), fmul(z, z))));
      if (vl != 0.0f) {
        const float inv = __fdiv_rn(1.0f, vl);
        x = fmul(x, inv), y = fmul(y, inv), z = fmul(z, inv);
      }
      one_source_point(L, x, y, z, ls);
    }
    // impeghd_calc_spread_gains (:113) with the raw width; the double-precision expression is IEEE on both sides
    float alpha = float(double(fmul(s.spread_w, kDeg2Rad)) / (M_PI / 2.0) - 1.0);
    float norm = 0.0f;
    for (int i = 0; i < tot; ++i) norm = fadd(norm, fmul(ls[i], ls[i]));
    norm = __fsqrt_rn(norm);
    if (alpha > 1.0f)
      alpha = 1.0f;
    else if (alpha < 0.0f)
      alpha = 0.0f;
    const float oma = fsub(1.f, alpha), au = fmul(alpha, L.unit_gain);
    for (int i = 0; i < tot; ++i) ls[i] = fadd(__fdiv_rn(fmul(oma, ls[i]), norm), au);
  }
  // impeghd_nrm_values (:82)
  float len = 0.0f;
  for (int i = 0; i < tot; ++i) len = fadd(len, fmul(ls[i], ls[i]));
  len = __fdiv_rn(1.0f, __fsqrt_rn(len));
  if (len != 0.0f)
    for (int i = 0; i < tot; ++i) ls[i] = fmul(ls[i], len);
  // imaginary-speaker downmix + power normalisation (:563-586), then the ramp of impeghd_obj_renderer_dec
  float fg[24];
  float gain_sum = 0.0f;
  for (int i = 0; i < NL; ++i) {
    float value = 0.f;
    for (int j = 0; j < tot; ++j) value = fadd(value, fmul(__ldg(&L.dmx[i][j]), ls[j]));
    fg[i] = value;
    gain_sum = fadd(gain_sum, fmul(value, value));
  }
  gain_sum = __fsqrt_rn(gain_sum);
  float* st = state + size_t(idx) * (NL + 1);
  const bool first = st[0] != 0.0f;
  float* sc = scratch + size_t(idx) * NL * 2;
  for (int i = 0; i < NL; ++i) {
    float f = fg[i];
    if (gain_sum != 0.0f) f = __fdiv_rn(fmul(f, s.gain), gain_sum);
    const float init = first ? f : st[1 + i];
    const float step = __fdiv_rn(fsub(f, init), 1024.0f);
    sc[2 * i] = fadd(init, step);
    sc[2 * i + 1] = step;
    st[1 + i] = f;
  }
  st[0] = 0.0f;
}

// One CTA per stream, thread = (output channel, 64-sample chunk).
//   phase A  ramp values at the 16 chunk boundaries for every (object, channel): a chain of 64 c separately
//            rounded additions each, computed once per frame into shared memory (threads stride over pairs)
//   phase B  the mix: 8 samples at a time, the 32 per-object scales of this thread's channel in registers,
//            object signals read straight from global memory (each 32-byte piece is shared by the channel
//            threads of the chunk and served by L1)
constexpr int kChunks = 16;
constexpr int kCLen = kFrame / kChunks;  // 64

__global__ void __launch_bounds__(384)
    obj_mix_kernel(const DevLayout* __restrict__ Lp, const float* __restrict__ in, float* __restrict__ out,
                   const float* __restrict__ scratch, int n_obj) {
  extern __shared__ float bound[];  // [n_obj * NL][kChunks + 1]: chunk-start scales, last = step
  const DevLayout& L = *Lp;
  const int NL = L.n_non_lfe;
  const int s = blockIdx.x;
  const int n_pairs = n_obj * NL;
  const float* sc = scratch + size_t(s) * n_pairs * 2;
  for (int p = threadIdx.x; p < n_pairs; p += blockDim.x) {
    const float2 t = __ldg(reinterpret_cast<const float2*>(sc) + p);
    float v = t.x;
    float* b = bound + p * (kChunks + 1);
    b[kChunks] = t.y;
    for (int c = 0; c < kChunks; ++c) {
      b[c] = v;
#pragma unroll 8
      for (int k = 0; k < kCLen; ++k) v = fadd(v, t.y);
    }
  }
  __syncthreads();
  const int ci = threadIdx.x % NL, chunk = threadIdx.x / NL;
  if (chunk < kChunks) {
    float scale[kMaxObj], step[kMaxObj];
#pragma unroll
    for (int o = 0; o < kMaxObj; ++o) {
      if (o < n_obj) {
        const float* b = bound + (o * NL + ci) * (kChunks + 1);
        scale[o] = b[chunk];
        step[o] = b[kChunks];
      } else {
        scale[o] = 0.f, step[o] = 0.f;
      }
    }
    const float* x = in + size_t(s) * n_obj * kFrame + chunk * kCLen;
    float* y = out + (size_t(s) * L.n_spk + L.nonlfe_to_spk[ci]) * kFrame + chunk * kCLen;
    for (int j = 0; j < kCLen; j += 8) {
      float acc[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) acc[e] = 0.0f;
#pragma unroll
      for (int o = 0; o < kMaxObj; ++o) {
        if (o < n_obj) {
          const float4 a = __ldg(reinterpret_cast<const float4*>(x + size_t(o) * kFrame + j));
          const float4 b = __ldg(reinterpret_cast<const float4*>(x + size_t(o) * kFrame + j + 4));
          const float xv[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
          float sc_o = scale[o];
          const float st_o = step[o];
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            acc[e] = fadd(acc[e], fmul(xv[e], sc_o));
            sc_o = fadd(sc_o, st_o);
          }
          scale[o] = sc_o;
        }
      }
      __stcs(reinterpret_cast<float4*>(y + j), make_float4(acc[0], acc[1], acc[2], acc[3]));
      __stcs(reinterpret_cast<float4*>(y + j + 4), make_float4(acc[4], acc[5], acc[6], acc[7]));
    }
  }
  // LFE rows stay zero (decode_main.c:1374 memset)
  for (int l = 0; l < L.n_lfe; ++l) {
    float4* z = reinterpret_cast<float4*>(out + (size_t(s) * L.n_spk + L.lfe_spk[l]) * kFrame);
    for (int j = threadIdx.x; j < kFrame / 4; j += blockDim.x) z[j] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
}

__global__ void obj_state_init_kernel(float* state, const int* first, int n_obj, int nl1, size_t total) {
  for (size_t i = blockIdx.x * size_t(blockDim.x) + threadIdx.x; i < total; i += size_t(gridDim.x) * blockDim.x) {
    const int k = int(i % nl1);
    const int o = int((i / nl1) % n_obj);
    state[i] = (k == 0) ? (first ? float(first[o] != 0) : 1.0f) : 0.0f;
  }
}

void pol2cart(float phi, float theta, float radius, float* c) {
  // decoder/impeghd_3d_vec_basic_ops.c:86 (host, glibc double sin/cos like the reference)
  const float sin_phi = float(std::sin(double(phi * kDeg2Rad))), cos_phi = float(std::cos(double(phi * kDeg2Rad)));
  const float sin_theta = float(std::sin(double(theta * kDeg2Rad))), cos_theta = float(std::cos(double(theta * kDeg2Rad)));
  c[2] = radius * sin_phi;
  c[1] = (radius * sin_theta) * cos_phi;
  c[0] = (radius * cos_theta) * cos_phi;
}

}  // namespace
}  // namespace mpeghb200

using namespace mpeghb200;

extern "C" {

void mpeghb200_obj_side_from_polar(const float* params7, int n, mpeghb200_obj_side* out) {
  for (int o = 0; o < n; ++o) {
    const float* q = params7 + size_t(o) * 7;
    mpeghb200_obj_side& s = out[o];
    std::memset(&s, 0, sizeof(s));
    const float theta = q[0], phi = q[1];
    pol2cart(phi, theta, q[2], s.vec_c);
    s.gain = q[3];
    s.spread_w = q[4];
    s.spread_h = q[5];
    if (!(q[4] <= 0.0f && q[5] <= 0.0f)) {
      // decoder/impeghd_obj_ren_vbap.c:351-389: clamp, tan in double, p0 and the elevation-rotated v
      float alpha = q[4];
      if (alpha > 90.0f)
        alpha = 90.0f;
      else if (alpha < 0.001f)
        alpha = 0.001f;
      s.tan_alpha = float(std::tan(double(alpha * kDeg2Rad)));
      pol2cart(phi, theta, 1.0f, s.p0);
      const float phi2 = (phi < 0) ? phi + 90 : phi - 90;
      pol2cart(phi2, theta, 1.0f, s.v);
    }
  }
}

mpeghb200_status mpeghb200_obj_layout_create(mpeghb200_ctx* ctx, const mpeghb200_obj_layout_desc* d,
                                             mpeghb200_obj_layout** out) {
  if (!ctx || !d || !out) return MPEGHB200_ERR_BAD_ARG;
  *out = nullptr;
  const int tot = d->n_non_lfe + d->n_imag;
  if (d->n_spk <= 0 || d->n_spk > kMaxLs || d->n_non_lfe <= 0 || d->n_non_lfe > 24 || d->n_non_lfe > d->n_spk ||
      d->n_imag < 0 || tot > kMaxLs || d->n_tri <= 0 || d->n_tri > kMaxLs)
    return fail(ctx, MPEGHB200_ERR_BAD_ARG, "obj_layout_create: counts out of range");
  MB_CUDA(ctx, cudaSetDevice(ctx->device));
  auto* L = new (std::nothrow) mpeghb200_obj_layout();
  if (!L) return MPEGHB200_ERR_NO_MEM;
  L->ctx = ctx;
  L->host = *d;
  DevLayout h{};
  h.n_spk = d->n_spk, h.n_non_lfe = d->n_non_lfe, h.n_imag = d->n_imag, h.n_tri = d->n_tri;
  h.height_present = d->height_present;
  h.unit_gain = float(1.f / std::sqrt(double(float(tot))));
  std::memcpy(h.tri, d->tri, sizeof(h.tri));
  std::memcpy(h.inv, d->inv, sizeof(h.inv));
  std::memcpy(h.dmx, d->dmx, sizeof(h.dmx));
  int nl = 0, nlfe = 0;
  for (int ch = 0; ch < d->n_spk; ++ch) {
    if (d->lfe[ch])
      h.lfe_spk[nlfe++] = ch;
    else
      h.nonlfe_to_spk[nl++] = ch;
  }
  h.n_lfe = nlfe;
  for (int t = 0; t < d->n_tri; ++t)
    for (int k = 0; k < 3; ++k)
      if (d->tri[t][k] < 0 || d->tri[t][k] >= tot) nl = -1;
  if (nl != d->n_non_lfe) {
    delete L;
    return fail(ctx, MPEGHB200_ERR_BAD_ARG, "obj_layout_create: lfe flags / triangle indices inconsistent with counts");
  }
  std::memcpy(L->nonlfe_to_spk, h.nonlfe_to_spk, sizeof(h.nonlfe_to_spk));
  cudaError_t e = cudaMalloc(&L->d_desc, sizeof(DevLayout));
  if (e == cudaSuccess) e = cudaMemcpy(L->d_desc, &h, sizeof(DevLayout), cudaMemcpyHostToDevice);
  if (e != cudaSuccess) {
    cudaFree(L->d_desc);
    delete L;
    return cuda_fail(ctx, e, "obj_layout_create upload");
  }
  *out = L;
  return MPEGHB200_OK;
}

void mpeghb200_obj_layout_destroy(mpeghb200_obj_layout* L) {
  if (!L) return;
  cudaSetDevice(L->ctx->device);
  cudaFree(L->d_desc);
  delete L;
}

size_t mpeghb200_obj_state_floats(const mpeghb200_obj_layout* L, int n_obj) {
  return L ? size_t(n_obj) * (L->host.n_non_lfe + 1) : 0;
}
size_t mpeghb200_obj_scratch_floats(const mpeghb200_obj_layout* L, int n_obj) {
  return L ? size_t(n_obj) * L->host.n_non_lfe * 2 : 0;
}

mpeghb200_status mpeghb200_obj_state_init_dev(mpeghb200_ctx* ctx, const mpeghb200_obj_layout* L, int n_obj,
                                              const int32_t* d_first_flags, float* d_state, int n_streams,
                                              void* cuda_stream) {
  if (!ctx || !L || n_obj <= 0 || n_obj > kMaxObj || n_streams < 0 || (n_streams > 0 && !d_state))
    return MPEGHB200_ERR_BAD_ARG;
  if (n_streams == 0) return MPEGHB200_OK;
  MB_CUDA(ctx, cudaSetDevice(ctx->device));
  const size_t total = mpeghb200_obj_state_floats(L, n_obj) * n_streams;
  obj_state_init_kernel<<<148, 256, 0, static_cast<cudaStream_t>(cuda_stream)>>>(d_state, d_first_flags, n_obj,
                                                                                L->host.n_non_lfe + 1, total);
  ctx->launches.fetch_add(1);
  MB_CUDA(ctx, cudaGetLastError());
  return MPEGHB200_OK;
}

mpeghb200_status mpeghb200_obj_render_dev(mpeghb200_ctx* ctx, const mpeghb200_obj_layout* L, int n_obj,
                                          const mpeghb200_obj_side* d_side, const float* d_in, float* d_out,
                                          float* d_state, float* d_scratch, int n_streams, void* cuda_stream) {
  if (!ctx || !L) return MPEGHB200_ERR_BAD_ARG;
  if (n_obj <= 0 || n_obj > kMaxObj) return fail(ctx, MPEGHB200_ERR_UNSUPPORTED, "obj_render_dev: 1..32 objects");
  if (n_streams < 0 || (n_streams > 0 && (!d_side || !d_in || !d_out || !d_state || !d_scratch)))
    return fail(ctx, MPEGHB200_ERR_BAD_ARG, "obj_render_dev: null buffer or negative stream count");
  if (n_streams == 0) return MPEGHB200_OK;
  MB_CUDA(ctx, cudaSetDevice(ctx->device));
  cudaStream_t st = static_cast<cudaStream_t>(cuda_stream);
  const DevLayout* dl = static_cast<const DevLayout*>(L->d_desc);
  const int n_pairs = n_obj * n_streams;
  obj_gains_kernel<<<(n_pairs + 127) / 128, 128, 0, st>>>(dl, d_side, d_state, d_scratch, n_obj, n_streams);
  ctx->launches.fetch_add(1);
  MB_CUDA(ctx, cudaGetLastError());
  const int NL = L->host.n_non_lfe;
  const size_t smem = size_t(n_obj) * NL * (kChunks + 1) * sizeof(float);
  static size_t configured_smem = 48 * 1024;
  if (smem > configured_smem) {
    MB_CUDA(ctx, cudaFuncSetAttribute(obj_mix_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)));
    configured_smem = smem;
  }
  obj_mix_kernel<<<n_streams, NL * kChunks, smem, st>>>(dl, d_in, d_out, d_scratch, n_obj);
  ctx->launches.fetch_add(1);
  MB_CUDA(ctx, cudaGetLastError());
  return MPEGHB200_OK;
}

}  // extern "C"
