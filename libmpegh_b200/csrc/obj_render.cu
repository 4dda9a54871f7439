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

// One source direction -> gains of the best triangle, accumulated into ls[] (impeghd_calc_one_source_point :156).
// The reference scans the triangles in order and keeps a candidate when it has fewer negative gains, or as many
// and a larger smallest gain: a strict lexicographic improvement, so the survivor is the lexicographic best with
// the lowest index among exact ties -- which a warp can find with one triangle per lane and a shuffle reduction.
// Called by all 32 lanes of a warp with identical arguments; ls[] lives in shared memory.
__device__ void one_source_point(const DevLayout& L, float vx, float vy, float vz, float* ls, int lane) {
  int winner = 0x7fffffff, hi_lo = 3;
  float highest_least = -100000.f, g0 = 0.f, g1 = 0.f, g2 = 0.f;
  for (int i = lane; i < L.n_tri; i += 32) {
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
#pragma unroll
  for (int d = 16; d >= 1; d >>= 1) {
    const int o_neg = __shfl_xor_sync(0xffffffffu, hi_lo, d), o_win = __shfl_xor_sync(0xffffffffu, winner, d);
    const float o_least = __shfl_xor_sync(0xffffffffu, highest_least, d);
    const float o0 = __shfl_xor_sync(0xffffffffu, g0, d), o1 = __shfl_xor_sync(0xffffffffu, g1, d),
                o2 = __shfl_xor_sync(0xffffffffu, g2, d);
    const bool take = (o_neg < hi_lo) || (o_neg == hi_lo && (o_least > highest_least ||
                                                              (o_least == highest_least && o_win < winner)));
    if (take) hi_lo = o_neg, highest_least = o_least, winner = o_win, g0 = o0, g1 = o1, g2 = o2;
  }
  if (winner == 0x7fffffff) winner = 0;  // nothing beat the initial state: the reference keeps triangle 0, gains 0
  if (g2 < 0.f) g2 = 0.f;  // covers both the -0.01 test (:204-210) and the < 0 clamps (:224-235)
  if (g1 < 0.f) g1 = 0.f;
  if (g0 < 0.f) g0 = 0.f;
  const float power = __fsqrt_rn(fadd(fadd(fmul(g0, g0), fmul(g1, g1)), fmul(g2, g2)));
  if (power != 0.f) {
    g0 = __fdiv_rn(g0, power);
    g1 = __fdiv_rn(g1, power);
    g2 = __fdiv_rn(g2, power);
  }
  if (lane == 0) {
    const int i2 = __ldg(&L.tri[winner][2]), i1 = __ldg(&L.tri[winner][1]), i0 = __ldg(&L.tri[winner][0]);
    ls[i2] = fadd(g2, ls[i2]);
    ls[i1] = fadd(g1, ls[i1]);
    ls[i0] = fadd(g0, ls[i0]);
  }
  __syncwarp();
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
// One warp per (stream, object).  The triangle search runs one triangle per lane; the loudspeaker gain vector
// ls[] sits in shared memory; the ordered sums of the reference (spread norm, vector norm, power sum) are walked
// in index order, the imaginary-speaker downmix has one output channel per lane.
constexpr int kGainWarps = 4;

__global__ void __launch_bounds__(32 * kGainWarps)
    obj_gains_kernel(const DevLayout* __restrict__ Lp, const mpeghb200_obj_side* __restrict__ side,
                     float* __restrict__ state, float* __restrict__ scratch, int n_obj, int n_streams, float frame_len,
                     const int32_t* __restrict__ meta_set) {
  pdl_prologue();
  __shared__ float ls_all[kGainWarps][kMaxLs];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const int idx = blockIdx.x * kGainWarps + wib;
  if (idx >= n_obj * n_streams) return;
  const DevLayout& L = *Lp;
  const mpeghb200_obj_side s = side[idx];
  const int NL = L.n_non_lfe, tot = NL + L.n_imag;
  float* ls = ls_all[wib];
  for (int i = lane; i < tot; i += 32) ls[i] = 0.f;
  __syncwarp();
  if (s.spread_w <= 0.0f && s.spread_h <= 0.0f) {
    one_source_point(L, s.vec_c[0], s.vec_c[1], s.vec_c[2], ls, lane);
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
      one_source_point(L, x, y, z, ls, lane);
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
    __syncwarp();
    for (int i = lane; i < tot; i += 32) ls[i] = fadd(__fdiv_rn(fmul(oma, ls[i]), norm), au);
    __syncwarp();
  }
  // impeghd_nrm_values (:82)
  float len = 0.0f;
  for (int i = 0; i < tot; ++i) len = fadd(len, fmul(ls[i], ls[i]));
  len = __fdiv_rn(1.0f, __fsqrt_rn(len));
  __syncwarp();
  if (len != 0.0f)
    for (int i = lane; i < tot; i += 32) ls[i] = fmul(ls[i], len);
  __syncwarp();
  // imaginary-speaker downmix + power normalisation (:563-586), then the ramp of impeghd_obj_renderer_dec;
  // output channel = lane
  float value = 0.f;
  if (lane < NL)
    for (int j = 0; j < tot; ++j) value = fadd(value, fmul(__ldg(&L.dmx[lane][j]), ls[j]));
  const float sq = fmul(value, value);
  float gain_sum = 0.0f;
  for (int i = 0; i < NL; ++i) gain_sum = fadd(gain_sum, __shfl_sync(0xffffffffu, sq, i));
  gain_sum = __fsqrt_rn(gain_sum);
  float* st = state + size_t(idx) * (NL + 1);
  // The first-frame rule (start the ramp at the final gains) only fires for calls rendering with metadata set 0:
  // the reference tests first_frame_flag[sub_frame_offset + obj] (decoder/impeghd_obj_ren_vbap.c:497), an entry it
  // never clears for sets > 0 and whose gains land in a row the mixer does not read; the object's own flag is
  // cleared by every call all the same (decoder/impeghd_obj_ren_dec.c:1367).
  const bool first = st[0] != 0.0f && (meta_set == nullptr || meta_set[idx / n_obj] == 0);
  __syncwarp();
  float* sc = scratch + size_t(idx) * NL * 2;
  if (lane < NL) {
    float f = value;
    if (gain_sum != 0.0f) f = __fdiv_rn(fmul(f, s.gain), gain_sum);
    const float init = first ? f : st[1 + lane];
    const float step = __fdiv_rn(fsub(f, init), frame_len);  // OAM frame length: 1024, or a sub-frame of 512 / 256
    reinterpret_cast<float2*>(sc)[lane] = make_float2(fadd(init, step), step);
    st[1 + lane] = f;
  }
  if (lane == 0) st[0] = 0.0f;
}

// Two CTAs per stream (each takes half of the non-LFE output channels), thread = (output channel, 64-sample chunk).
//   phase A  ramp values at the 16 chunk boundaries for every (object, channel) of this CTA: a chain of 64 c
//            separately rounded additions each, computed once per frame into shared memory (two chains per thread
//            at a time)
//   phase B  the mix in 16 steps of 4 samples per chunk, the per-object scales and steps of this thread's channel
//            in registers.  The object signals of a step (16 chunks x NOBJ objects x 16 B) are staged in shared
//            memory with cp.async, double buffered, so the FP pipe never waits on an L2 / HBM round trip: the
//            step-0 copy overlaps phase A, the copy of step t + 1 overlaps the arithmetic of step t.
// NOBJ (8, 16, 24, 32) is the compile-time object count the loops are unrolled for; a stream with fewer objects
// is padded with zero signals and zero scales, which is exact: the accumulator starts at +0, so adding a (+-0)
// product never changes it.
constexpr int kChunks = 16;
constexpr int kCLen = kFrame / kChunks;  // 64
constexpr int kStep = 4;                 // samples per chunk and step
constexpr int kSteps = kCLen / kStep;    // 16
constexpr int kMixMaxThreads = 12 * kChunks;  // up to 12 channels per CTA (24 non-LFE loudspeakers)

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc) {
  const unsigned d = static_cast<unsigned>(__cvta_generic_to_shared(smem_dst));
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(d), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;\n" ::"n"(N) : "memory");
}

template <int NOBJ>
struct MixSmem {
  static constexpr int kChunkStride = NOBJ * kStep + 4;  // floats; +16 B skews the chunks across banks
  static constexpr int kStageFloats = 2 * kChunks * kChunkStride;
  static size_t bytes(int n_ch_cta) { return (size_t(kStageFloats) + size_t(NOBJ) * n_ch_cta * (kChunks + 1)) * 4; }
};

template <int NOBJ>
__global__ void __launch_bounds__(kMixMaxThreads)
    obj_mix_kernel(const DevLayout* __restrict__ Lp, const float* __restrict__ in, float* __restrict__ out,
                   const float* __restrict__ scratch, int n_obj, int n_chunks, int sample_offset) {
  pdl_prologue();
  extern __shared__ __align__(16) float smem_f[];
  using SM = MixSmem<NOBJ>;
  const DevLayout& L = *Lp;
  const int NL = L.n_non_lfe;
  const int s = blockIdx.y;
  const int half = (NL + 1) / 2;
  const int c_first = blockIdx.x * half;                       // first channel of this CTA
  const int n_c = blockIdx.x == 0 ? half : NL - half;          // its channel count
  const int n_pairs = NOBJ * n_c;                              // (object, channel) chains, padded objects included
  float* stage = smem_f;                                       // [2][kChunks][kChunkStride]: x[chunk][obj][4]
  float* bound = smem_f + SM::kStageFloats;                    // [n_pairs][kChunks + 1]: chunk-start scales, step
  const float* xs = in + size_t(s) * n_obj * kFrame + sample_offset;  // n_chunks * 64 samples from sample_offset on

  auto issue = [&](int t) {
    float* dst = stage + (t & 1) * kChunks * SM::kChunkStride;
    for (int p = threadIdx.x; p < n_chunks * NOBJ; p += blockDim.x) {
      const int chunk = p / NOBJ, o = p % NOBJ;
      if (o < n_obj)
        cp_async16(dst + chunk * SM::kChunkStride + o * kStep, xs + size_t(o) * kFrame + chunk * kCLen + t * kStep);
    }
    cp_async_commit();
  };
  if (n_obj < NOBJ)  // padded objects: zero signal in both buffers, never overwritten
    for (int p = threadIdx.x; p < 2 * kChunks * (NOBJ - n_obj) * kStep; p += blockDim.x) {
      const int e = p % kStep, q = p / kStep, o = n_obj + q % (NOBJ - n_obj), cb = q / (NOBJ - n_obj);
      stage[cb * SM::kChunkStride + o * kStep + e] = 0.0f;
    }
  issue(0);

  const float* sc = scratch + size_t(s) * n_obj * NL * 2;
  auto ramp = [&](int p) {  // chain p = o * n_c + c  ->  (start, step) of (object o, channel c_first + c)
    const int o = p / n_c, c = p - o * n_c;
    return o < n_obj ? __ldg(reinterpret_cast<const float2*>(sc) + o * NL + c_first + c) : make_float2(0.f, 0.f);
  };
  for (int p = threadIdx.x; p < n_pairs; p += 2 * blockDim.x) {
    const int q = p + blockDim.x;
    const bool two = q < n_pairs;
    const float2 t0 = ramp(p);
    const float2 t1 = two ? ramp(q) : make_float2(0.f, 0.f);
    float v0 = t0.x, v1 = t1.x;
    float* b0 = bound + p * (kChunks + 1);
    float* b1 = bound + (two ? q : p) * (kChunks + 1);
    for (int c = 0; c < n_chunks; ++c) {
      b0[c] = v0;
      if (two) b1[c] = v1;
      if (c + 1 < n_chunks) {
#pragma unroll 8
        for (int k = 0; k < kCLen; ++k) v0 = fadd(v0, t0.y), v1 = fadd(v1, t1.y);
      }
    }
    b0[kChunks] = t0.y;
    if (two) b1[kChunks] = t1.y;
  }
  __syncthreads();
  const int ci = threadIdx.x % n_c, chunk = threadIdx.x / n_c;
  const bool active = chunk < n_chunks;
  float scale[NOBJ], step[NOBJ];
#pragma unroll
  for (int o = 0; o < NOBJ; ++o) {
    const float* b = bound + (o * n_c + ci) * (kChunks + 1);
    scale[o] = active ? b[chunk] : 0.f;
    step[o] = active ? b[kChunks] : 0.f;
  }
  float* y = out + (size_t(s) * L.n_spk + L.nonlfe_to_spk[c_first + ci]) * kFrame + sample_offset +
             (active ? chunk : 0) * kCLen;
  for (int t = 0; t < kSteps; ++t) {
    if (t + 1 < kSteps) {
      issue(t + 1);
      cp_async_wait<1>();
    } else {
      cp_async_wait<0>();
    }
    __syncthreads();
    if (active) {
      const float* x = stage + (t & 1) * kChunks * SM::kChunkStride + chunk * SM::kChunkStride;
      float acc[kStep];
#pragma unroll
      for (int e = 0; e < kStep; ++e) acc[e] = 0.0f;
#pragma unroll
      for (int o = 0; o < NOBJ; ++o) {
        const float4 a = *reinterpret_cast<const float4*>(x + o * kStep);
        const float xv[4] = {a.x, a.y, a.z, a.w};
        float sc_o = scale[o];
        const float st_o = step[o];
#pragma unroll
        for (int e = 0; e < kStep; ++e) {
          acc[e] = fadd(acc[e], fmul(xv[e], sc_o));
          sc_o = fadd(sc_o, st_o);
        }
        scale[o] = sc_o;
      }
      __stcs(reinterpret_cast<float4*>(y + t * kStep), make_float4(acc[0], acc[1], acc[2], acc[3]));
    }
    __syncthreads();  // the buffer of step t is overwritten by the copy issued in iteration t + 1
  }
  // LFE rows stay zero (decode_main.c:1374 memset)
  if (blockIdx.x == 0)
    for (int l = 0; l < L.n_lfe; ++l) {
      float4* z = reinterpret_cast<float4*>(out + (size_t(s) * L.n_spk + L.lfe_spk[l]) * kFrame + sample_offset);
      for (int j = threadIdx.x; j < n_chunks * kCLen / 4; j += blockDim.x) z[j] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
}

template <int NOBJ>
cudaError_t launch_mix(const DevLayout* dl, int NL, const float* d_in, float* d_out, const float* d_scratch, int n_obj,
                       int n_streams, int n_chunks, int sample_offset, cudaStream_t st) {
  const int half = (NL + 1) / 2;
  const size_t smem = MixSmem<NOBJ>::bytes(half);
  if (smem > 48 * 1024) {  // per device / context attribute: set on every call
    cudaError_t e = cudaFuncSetAttribute(obj_mix_kernel<NOBJ>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem));
    if (e != cudaSuccess) return e;
  }
  return launch_pdl(obj_mix_kernel<NOBJ>, dim3(NL > 1 ? 2 : 1, n_streams), dim3(half * kChunks), smem, st, dl, d_in, d_out,
                    d_scratch, n_obj, n_chunks, sample_offset);
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

int mpeghb200_obj_layout_num_speakers(const mpeghb200_obj_layout* L) { return L ? L->host.n_spk : 0; }

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
  return mpeghb200_obj_render_sub_dev(ctx, L, n_obj, d_side, d_in, d_out, d_state, d_scratch, n_streams, kFrame, 0,
                                      nullptr, cuda_stream);
}

mpeghb200_status mpeghb200_obj_render_sub_dev(mpeghb200_ctx* ctx, const mpeghb200_obj_layout* L, int n_obj,
                                              const mpeghb200_obj_side* d_side, const float* d_in, float* d_out,
                                              float* d_state, float* d_scratch, int n_streams, int frame_len,
                                              int sample_offset, const int32_t* d_metadata_set, void* cuda_stream) {
  if (!ctx || !L) return MPEGHB200_ERR_BAD_ARG;
  if ((frame_len != 256 && frame_len != 512 && frame_len != 1024) || sample_offset < 0 || sample_offset % frame_len ||
      sample_offset + frame_len > kFrame)
    return fail(ctx, MPEGHB200_ERR_BAD_ARG, "obj_render_sub_dev: OAM frame length 256 / 512 / 1024 at a multiple of itself");
  if (n_obj <= 0 || n_obj > kMaxObj) return fail(ctx, MPEGHB200_ERR_UNSUPPORTED, "obj_render_dev: 1..32 objects");
  if (n_streams < 0 || (n_streams > 0 && (!d_side || !d_in || !d_out || !d_state || !d_scratch)))
    return fail(ctx, MPEGHB200_ERR_BAD_ARG, "obj_render_dev: null buffer or negative stream count");
  if (n_streams == 0) return MPEGHB200_OK;
  MB_CUDA(ctx, cudaSetDevice(ctx->device));
  cudaStream_t st = static_cast<cudaStream_t>(cuda_stream);
  const DevLayout* dl = static_cast<const DevLayout*>(L->d_desc);
  const int n_pairs = n_obj * n_streams;
  MB_CUDA(ctx, launch_pdl(obj_gains_kernel, dim3((n_pairs + kGainWarps - 1) / kGainWarps), dim3(32 * kGainWarps), 0, st, dl,
                          d_side, d_state, d_scratch, n_obj, n_streams, float(frame_len), d_metadata_set));
  ctx->launches.fetch_add(1);
  MB_CUDA(ctx, cudaGetLastError());
  const int NL = L->host.n_non_lfe;
  if (NL > 2 * (kMixMaxThreads / kChunks))
    return fail(ctx, MPEGHB200_ERR_UNSUPPORTED, "obj_render_dev: more than 24 non-LFE loudspeakers");
  cudaError_t e;
  if (n_obj <= 8)
    e = launch_mix<8>(dl, NL, d_in, d_out, d_scratch, n_obj, n_streams, frame_len / kCLen, sample_offset, st);
  else if (n_obj <= 16)
    e = launch_mix<16>(dl, NL, d_in, d_out, d_scratch, n_obj, n_streams, frame_len / kCLen, sample_offset, st);
  else if (n_obj <= 24)
    e = launch_mix<24>(dl, NL, d_in, d_out, d_scratch, n_obj, n_streams, frame_len / kCLen, sample_offset, st);
  else
    e = launch_mix<32>(dl, NL, d_in, d_out, d_scratch, n_obj, n_streams, frame_len / kCLen, sample_offset, st);
  MB_CUDA(ctx, e);
  ctx->launches.fetch_add(1);
  MB_CUDA(ctx, cudaGetLastError());
  return MPEGHB200_OK;
}

}  // extern "C"
