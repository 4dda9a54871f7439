// format_conv.cu -- format converter: STFT-domain active downmix with ERB-band energy equalisation.
//
// Stands in for impeghd_format_converter_process (decoder/ia_core_coder_process.c:641) and
// impeghd_format_conv_active_dmx_stft_process (:127).  Per frame and stream, 4 slots of 256 samples:
//   1. every input channel: sine-windowed [previous 256 | current 256] block -> 512-point e^{+j} transform
//   2. Y[out] = sum_in D[out][in][band] * X[in] over the inputs whose D[out][in][0] != 0, in input order;
//      target energy (sum of |D X|^2) and realised energy (|Y|^2) per ERB band, one-pole smoothing against the
//      previous slot, EQ = sqrt(target / (1e-35 + realised)) clipped to [-10 dB, +8 dB], Y *= EQ
//   3. every output channel: Hermitian extension -> 512-point e^{-j} transform -> /512 -> windowed overlap-add.
// The downmix tensor D comes from the reference's host init (impeghd_format_conv_init_data or its VBAP
// fallback, decoder/ia_core_coder_decode_main.c:519) and is handed over.
//
// One CTA (12 warps) per stream walks the four slots; spectra never leave shared memory.  A warp owns one
// 512-point transform (16 points per lane, same exact radix-4 DAG as fft_exact.cuh plus the final radix-2 pass).
// The energy sums are strictly sequential float chains in the reference; they are cut where the reference's
// data flow allows it: one chain per (output, ERB band), 58 x n_out chains running side by side, while the
// spectrum accumulation itself runs per (output, bin).  Rounding order within every chain is the reference's.
// HBM traffic per stream-frame: (n_in + n_out) * 4 KB of audio + (n_in + n_out) * 1 KB + n_out * 464 B of state.
#include <cstring>
#include <new>
#include <vector>

#include "fft_exact.cuh"

namespace mpeghb200 {
namespace {

using namespace fftx;

constexpr int kMaxFcCh = 24;
constexpr int kErb = 58;
constexpr int kWarps = 12;
constexpr int kBufStride = 544;  // float2: 512 + 512/16 padding
constexpr int kYStride = 514;    // floats per output spectrum row: the chains of a warp walk rows 514 apart conflict free
static_assert(kMaxFcCh * kYStride <= 2 * kWarps * kBufStride, "the squares of a batch must fit the transform buffers");
constexpr float kAlpha = 0.0435f;
constexpr float kBeta = (1 - kAlpha);
constexpr float kEps = 1e-35f;
constexpr float kEqMax = 2.51188636f;
constexpr float kEqMin = 0.316227764f;

__constant__ int c_erb_stop[kErb] = {1,  2,  3,  4,  5,  6,  7,   8,   9,   10,  11,  12,  13,  14,  15,
                                     16, 17, 18, 19, 20, 21, 22,  23,  24,  25,  26,  27,  28,  29,  30,
                                     31, 32, 33, 36, 39, 43, 47,  51,  55,  60,  65,  70,  77,  84,  91,
                                     98, 106, 115, 125, 136, 148, 160, 174, 189, 205, 222, 241, 256};
__constant__ unsigned char c_erb_of_bin[256];

struct FcParams {
  const float* in;   // [stream][n_in][1024]
  float* out;        // [stream][n_out][1024]
  float* state;      // [stream][state_floats]: in_prev [n_in][256], out_prev [n_out][256], t_prev, r_prev [n_out][58]
  const float* D;    // [n_out][n_in][257]
  const int* active;  // [n_out][1 + n_in]: count, then input indices with D[out][in][0] != 0
  const float* win;  // [256]
  const float4* eff8;
  const float2* rad2;
  int n_in, n_out, state_floats, max_active, ranks;
};

__device__ __forceinline__ float clip_eq(float t, float r) {
  float eq = __fsqrt_rn(__fdiv_rn(t, fadd(kEps, r)));
  return eq > kEqMax ? kEqMax : (eq < kEqMin ? kEqMin : eq);
}

__global__ void __launch_bounds__(kWarps * 32) format_conv_kernel(const __grid_constant__ FcParams P) {
  extern __shared__ float smem[];
  float2* work = reinterpret_cast<float2*>(smem);                       // [kWarps][kBufStride]
  float* X = smem + 2 * kWarps * kBufStride;                            // [n_in][512] packed spectra
  float* Y = X + P.n_in * 512;                                          // [n_out][kYStride]
  float* sq = smem;                                                     // [n_out][kYStride], aliases work (idle in phase 2)
  float* outprev = Y + P.n_out * kYStride;                              // [n_out][256]
  float* eqs = outprev + P.n_out * 256;                                 // [n_out][58]
  unsigned char* erb_s = reinterpret_cast<unsigned char*>(eqs + ((P.n_out * kErb + 1) & ~1));  // [256] band of a bin
  float* twf = eqs + ((P.n_out * kErb + 1) & ~1) + 64;                       // twiddles: stages A, B (498) | rad2 [256] float2
  float2* rad2_s = reinterpret_cast<float2*>(twf + 500);
  const RowTw tw = fill_row_tw(twf, P.eff8, 498);
  for (int i = threadIdx.x; i < 256; i += blockDim.x) rad2_s[i] = __ldg(P.rad2 + i), erb_s[i] = c_erb_of_bin[i];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const size_t s = blockIdx.x;
  const float* x = P.in + s * size_t(P.n_in) * 1024;
  float* yo = P.out + s * size_t(P.n_out) * 1024;
  float* st_in = P.state + s * size_t(P.state_floats);
  float* st_out = st_in + P.n_in * 256;
  float* st_t = st_out + P.n_out * 256;
  float* st_r = st_t + P.n_out * kErb;
  float2* buf = work + warp * kBufStride;

  for (int i = tid; i < P.n_out * 256; i += blockDim.x) outprev[i] = st_out[i];
  // The lane's sample positions are i = r3 + 32 u, r3 = its digit-reversed lane number: the order in which stage A
  // of a transform takes its inputs, so the input blocks go from global memory straight into the butterflies (any
  // permutation inside a group of 32 samples is as coalesced as the identity).
  const int r3 = rev32(lane);
  float wv[8], wr[8];  // window entries of those positions
#pragma unroll
  for (int u = 0; u < 8; ++u) {
    wv[u] = __ldg(P.win + r3 + 32 * u);
    wr[u] = __ldg(P.win + 255 - r3 - 32 * u);
  }
  if (tid < P.n_in)  // the frame's input rows on their way into L2 while the state is being read
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(x + size_t(tid) * 1024), "r"(4096) : "memory");
  if (tid >= 32 && tid < 32 + P.n_in)  // and the previous frame's last input slots
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(st_in + size_t(tid - 32) * 256), "r"(1024)
                 : "memory");
  // per (output, ERB band) chains owned by this thread: smoothed energies carried in registers across slots
  constexpr int kChains = 4;  // ceil(24 * 58 / 384)
  float tprev[kChains], rprev[kChains];
  // Chain j of the CTA is (erb, co) = (j / n_out, j % n_out): band-major, so the 32 chains of a warp have (nearly) the
  // same number of bins.  Output-major order put a 1-bin and a 19-bin band into the same warp, and the warp paid
  // for the longest in every pass over the inputs.
  const int n_chain = P.n_out * kErb;
  int chain_co[kChains], chain_erb[kChains], chain_na[kChains], chain_k[kChains];  // chain_k = first bin | end << 16
#pragma unroll
  for (int q = 0; q < kChains; ++q) {
    const int j = tid + q * blockDim.x;
    chain_erb[q] = (j < n_chain) ? j / P.n_out : -1;
    chain_co[q] = j - chain_erb[q] * P.n_out;
    chain_na[q] = (j < n_chain) ? P.active[chain_co[q] * (1 + P.n_in)] : 0;
    const int erb = chain_erb[q] < 0 ? 0 : chain_erb[q];
    chain_k[q] = (erb <= 1 ? 1 : c_erb_stop[erb - 1]) | ((erb == 0 ? 1 : c_erb_stop[erb]) << 16);
    const int c = chain_co[q] * kErb + chain_erb[q];
    tprev[q] = (j < n_chain) ? st_t[c] : 0.0f;
    rprev[q] = (j < n_chain) ? st_r[c] : 0.0f;
  }
  __syncthreads();

  for (int slot = 0; slot < 4; ++slot) {
    // ---- 1. input transforms ------------------------------------------------------------------------
    for (int ci = warp; ci < P.n_in; ci += kWarps) {
      const float* cur = x + size_t(ci) * 1024 + slot * 256;
      const float* prev = slot ? cur - 256 : st_in + ci * 256;
      float2 v[16];
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        v[u] = make_float2(fmul(prev[r3 + 32 * u], wv[u]), 0.0f);
        v[8 + u] = make_float2(fmul(wr[u], cur[r3 + 32 * u]), 0.0f);
      }
      fft512_tail<true>(v, buf, lane, tw);
      float2* Xc = reinterpret_cast<float2*>(X + ci * 512);
#pragma unroll
      for (int u = 0; u < 8; ++u) {  // last pass: bins 0..255, and Re of bin 256 into the imaginary slot of bin 0
        const int k = lane + 32 * u;
        const float2 x0 = buf[phys(k)];
        float2 t = buf[phys(k + 256)];
        const float2 cs = rad2_s[k];
        rot<true>(t, cs.x, cs.y);
        float2 v = make_float2(fadd(x0.x, t.x), fadd(x0.y, t.y));
        if (k == 0) v.y = fsub(x0.x, t.x);
        Xc[k] = v;
      }
      __syncwarp();
    }
    __syncthreads();

    // ---- 2a/2b. downmix, one batch per rank a: the a-th active input of every output at once ---------------------
    // All threads form the terms D * X of the batch (one (output, bin) item each), add them into Y and leave their
    // squares in `sq`; then the (output, ERB band) chains add the squares in the reference's order (inputs outermost,
    // bins ascending, re before im).  The chains carry only dependent additions: the products, the table loads and
    // the spectrum loads are off their critical path.
    float tcur[kChains];
#pragma unroll
    for (int q = 0; q < kChains; ++q) tcur[q] = 0.0f;
    // max_active >= 1: the first pass also clears Y of outputs without inputs.  A pass covers `ranks` consecutive
    // ranks (as many as the squares buffer holds, 2 for 12 outputs): half the barriers.
    for (int a0 = 0; a0 < P.max_active; a0 += P.ranks) {
      // a warp takes whole outputs: which input, which table row are warp-uniform, the lane's 8 bins k = lane + 32 u
      for (int co = warp; co < P.n_out; co += kWarps) {
        const int* act = P.active + co * (1 + P.n_in);
        const int na = __ldg(act);
        float2* yp = reinterpret_cast<float2*>(Y + co * kYStride) + lane;
        float2 y[8];
        if (a0 && a0 < na) {
#pragma unroll
          for (int u = 0; u < 8; ++u) y[u] = yp[32 * u];
        } else {
#pragma unroll
          for (int u = 0; u < 8; ++u) y[u] = make_float2(0.0f, 0.0f);
        }
        for (int r = 0; r < P.ranks && a0 + r < na; ++r) {
          const int ci = __ldg(act + 1 + a0 + r);
          const float* d = P.D + (size_t(co) * P.n_in + ci) * 257 + lane;
          const float2* xp = reinterpret_cast<const float2*>(X + ci * 512) + lane;
          float2* sp = reinterpret_cast<float2*>(sq + (r * P.n_out + co) * kYStride) + lane;
          float dk[8];
#pragma unroll
          for (int u = 0; u < 8; ++u) dk[u] = __ldg(d + 32 * u);
          const float d_nyq = __ldg(d + 256 - lane);
#pragma unroll
          for (int u = 0; u < 8; ++u) {
            const float2 xv = xp[32 * u];
            const float d1 = (u == 0 && lane == 0) ? d_nyq : dk[u];
            const float tr = fmul(dk[u], xv.x), ti = fmul(d1, xv.y);
            y[u] = make_float2(fadd(tr, y[u].x), fadd(ti, y[u].y));
            sp[32 * u] = make_float2(fmul(tr, tr), fmul(ti, ti));
          }
        }
        if (a0 < na || a0 == 0) {
#pragma unroll
          for (int u = 0; u < 8; ++u) yp[32 * u] = y[u];
        }
      }
      __syncthreads();
#pragma unroll
      for (int q = 0; q < kChains; ++q) {
        if (chain_erb[q] >= 0) {
          const int erb = chain_erb[q];
          float t = tcur[q];
          for (int r = 0; r < P.ranks && a0 + r < chain_na[q]; ++r) {
            const float* sc = sq + (r * P.n_out + chain_co[q]) * kYStride;
            if (erb == 0) t = fadd(t, sc[0]);                 // DC
            if (erb == kErb - 1) t = fadd(t, sc[1]);          // Nyquist, ahead of the band's bins
            const int k0 = (erb == 0) ? 256 : (chain_k[q] & 0xffff), k1 = chain_k[q] >> 16;
#pragma unroll 4
            for (int k = k0; k < k1; ++k) {
              const float2 v = *reinterpret_cast<const float2*>(sc + 2 * k);
              t = fadd(fadd(t, v.x), v.y);
            }
          }
          tcur[q] = t;
        }
      }
      // the squares are next overwritten by the following pass; after the last pass by phase 3, which starts behind
      // the barrier that ends 2c: no barrier needed here then, and 2c joins the last pass's chain phase
      if (a0 + P.ranks < P.max_active) __syncthreads();
    }
    // ---- 2c. realised energy, smoothing, EQ ----------------------------------------------------------------
#pragma unroll
    for (int q = 0; q < kChains; ++q) {
      if (chain_erb[q] >= 0) {
        const int co = chain_co[q], erb = chain_erb[q], c = co * kErb + erb;
        const float* yc = Y + co * kYStride;
        float r = 0.0f;
        if (erb == 0) r = fmul(yc[0], yc[0]);
        if (erb == kErb - 1) r = fmul(yc[1], yc[1]);
        const int k0 = (erb == 0) ? 256 : (chain_k[q] & 0xffff), k1 = chain_k[q] >> 16;
#pragma unroll 4
        for (int k = k0; k < k1; ++k) {
          const float2 v = *reinterpret_cast<const float2*>(yc + 2 * k);
          r = fadd(r, fmul(v.x, v.x));
          r = fadd(r, fmul(v.y, v.y));
        }
        const float tn = fadd(fmul(kAlpha, tcur[q]), fmul(kBeta, tprev[q]));
        const float rn = fadd(fmul(kAlpha, r), fmul(kBeta, rprev[q]));
        tprev[q] = tn, rprev[q] = rn;
        eqs[c] = clip_eq(tn, rn);
      }
    }
    __syncthreads();

    // ---- 3. output transforms + overlap-add ----------------------------------------------------------------
    for (int co = warp; co < P.n_out; co += kWarps) {
      const float2* Yc = reinterpret_cast<const float2*>(Y + co * kYStride);
      const float* eq = eqs + co * kErb;
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const int k = lane + 32 * u;
        const float2 v = Yc[k];
        if (k == 0) {
          buf[phys(0)] = make_float2(fmul(v.x, eq[0]), 0.0f);
          buf[phys(256)] = make_float2(fmul(v.y, eq[kErb - 1]), 0.0f);
        } else {
          const float e = eq[erb_s[k]];
          const float re = fmul(v.x, e), im = fmul(e, v.y);
          buf[phys(k)] = make_float2(re, im);
          buf[phys(512 - k)] = make_float2(re, -im);
        }
      }
      __syncwarp();
      float2 v[16];
#pragma unroll
      for (int t = 0; t < 16; ++t) v[t] = buf[phys(r3 + 32 * t)];
      __syncwarp();
      fft512_tail<false>(v, buf, lane, tw);
      float* o = yo + size_t(co) * 1024 + slot * 256;
      float* op = outprev + co * 256;
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const int i = r3 + 32 * u;
        // last pass, real parts only: Re(w t) = t.x c - t.y s for the e^{-j} family
        const float2 x0 = buf[phys(i)], t = buf[phys(i + 256)], cs = rad2_s[i];
        const float tr = fsub(fmul(t.x, cs.x), fmul(t.y, cs.y));
        const float a = fmul(fadd(x0.x, tr), 0.001953125f);  // == x / 512 bit for bit (power of two)
        __stcs(o + i, fadd(fmul(a, wv[u]), fmul(op[i], wr[u])));
        op[i] = fmul(fsub(x0.x, tr), 0.001953125f);
      }
      __syncwarp();
    }
    __syncthreads();
  }

  // ---- state for the next frame --------------------------------------------------------------------------
  for (int i = tid; i < P.n_in * 256; i += blockDim.x) st_in[i] = x[size_t(i >> 8) * 1024 + 768 + (i & 255)];
  for (int i = tid; i < P.n_out * 256; i += blockDim.x) st_out[i] = outprev[i];
#pragma unroll
  for (int q = 0; q < kChains; ++q) {
    const int c = chain_co[q] * kErb + chain_erb[q];
    if (chain_erb[q] >= 0) st_t[c] = tprev[q], st_r[c] = rprev[q];
  }
}

}  // namespace
}  // namespace mpeghb200

using namespace mpeghb200;

struct mpeghb200_format_conv {
  mpeghb200_ctx* ctx = nullptr;
  int n_in = 0, n_out = 0, max_active = 0;
  float* d_D = nullptr;
  int* d_active = nullptr;
  float* d_win = nullptr;
};

extern "C" {

mpeghb200_status mpeghb200_format_conv_create(mpeghb200_ctx* ctx, int n_in, int n_out, const float* dmx,
                                              mpeghb200_format_conv** out) {
  if (!ctx || !out || !dmx) return MPEGHB200_ERR_BAD_ARG;
  *out = nullptr;
  if (n_in < 1 || n_in > kMaxFcCh || n_out < 1 || n_out > kMaxFcCh)
    return fail(ctx, MPEGHB200_ERR_BAD_ARG, "format_conv_create: 1..24 input and output channels");
  MB_CUDA(ctx, cudaSetDevice(ctx->device));
  auto* h = new (std::nothrow) mpeghb200_format_conv();
  if (!h) return MPEGHB200_ERR_NO_MEM;
  h->ctx = ctx, h->n_in = n_in, h->n_out = n_out;
  std::vector<int> active(size_t(n_out) * (1 + n_in), 0);
  for (int o = 0; o < n_out; ++o) {
    int n = 0;
    for (int i = 0; i < n_in; ++i)
      if (dmx[(size_t(o) * n_in + i) * 257] != 0.0f) active[size_t(o) * (1 + n_in) + 1 + n++] = i;
    active[size_t(o) * (1 + n_in)] = n;
    if (n > h->max_active) h->max_active = n;
  }
  static bool erb_table_done[64] = {};
  cudaError_t e = cudaSuccess;
  if (ctx->device < 64 && !erb_table_done[ctx->device]) {
    static const int stop[kErb] = {1,  2,  3,  4,  5,  6,  7,   8,   9,   10,  11,  12,  13,  14,  15,
                                   16, 17, 18, 19, 20, 21, 22,  23,  24,  25,  26,  27,  28,  29,  30,
                                   31, 32, 33, 36, 39, 43, 47,  51,  55,  60,  65,  70,  77,  84,  91,
                                   98, 106, 115, 125, 136, 148, 160, 174, 189, 205, 222, 241, 256};
    unsigned char erb_of_bin[256] = {};
    for (int erb = 1, k = 1; erb < kErb; ++erb)
      for (; k < stop[erb]; ++k) erb_of_bin[k] = (unsigned char)erb;
    e = cudaMemcpyToSymbol(c_erb_of_bin, erb_of_bin, sizeof(erb_of_bin));
    erb_table_done[ctx->device] = (e == cudaSuccess);
  }
  const std::vector<float>& win = rom_table(ROM_SINE_256);
  if (e == cudaSuccess) e = cudaMalloc(&h->d_D, size_t(n_out) * n_in * 257 * sizeof(float));
  if (e == cudaSuccess) e = cudaMalloc(&h->d_active, active.size() * sizeof(int));
  if (e == cudaSuccess) e = cudaMalloc(&h->d_win, 256 * sizeof(float));
  if (e == cudaSuccess) e = cudaMemcpy(h->d_D, dmx, size_t(n_out) * n_in * 257 * sizeof(float), cudaMemcpyHostToDevice);
  if (e == cudaSuccess) e = cudaMemcpy(h->d_active, active.data(), active.size() * sizeof(int), cudaMemcpyHostToDevice);
  if (e == cudaSuccess) e = cudaMemcpy(h->d_win, win.data(), 256 * sizeof(float), cudaMemcpyHostToDevice);
  if (e != cudaSuccess) {
    cudaFree(h->d_D), cudaFree(h->d_active), cudaFree(h->d_win);
    delete h;
    return cuda_fail(ctx, e, "format_conv_create");
  }
  *out = h;
  return MPEGHB200_OK;
}

void mpeghb200_format_conv_destroy(mpeghb200_format_conv* h) {
  if (!h) return;
  cudaSetDevice(h->ctx->device);
  cudaFree(h->d_D), cudaFree(h->d_active), cudaFree(h->d_win);
  delete h;
}

int mpeghb200_format_conv_num_in(const mpeghb200_format_conv* h) { return h ? h->n_in : 0; }
int mpeghb200_format_conv_num_out(const mpeghb200_format_conv* h) { return h ? h->n_out : 0; }

size_t mpeghb200_format_conv_state_floats(const mpeghb200_format_conv* h) {
  return h ? size_t(h->n_in) * 256 + size_t(h->n_out) * (256 + 2 * kErb) : 0;
}

mpeghb200_status mpeghb200_format_conv_dev(mpeghb200_ctx* ctx, const mpeghb200_format_conv* h, const float* d_in,
                                           float* d_out, float* d_state, int n_streams, void* cuda_stream) {
  if (!ctx || !h) return MPEGHB200_ERR_BAD_ARG;
  if (n_streams < 0 || (n_streams > 0 && (!d_in || !d_out || !d_state)))
    return fail(ctx, MPEGHB200_ERR_BAD_ARG, "format_conv_dev: null buffer or negative stream count");
  if (n_streams == 0) return MPEGHB200_OK;
  MB_CUDA(ctx, cudaSetDevice(ctx->device));
  FcParams P{};
  P.in = d_in, P.out = d_out, P.state = d_state, P.D = h->d_D, P.active = h->d_active, P.win = h->d_win;
  P.eff8 = ctx->tables.eff8, P.rad2 = ctx->tables.rad2_512;
  P.n_in = h->n_in, P.n_out = h->n_out, P.max_active = h->max_active > 0 ? h->max_active : 1;
  P.ranks = (2 * kWarps * kBufStride) / (h->n_out * kYStride);  // >= 1 by the static_assert above
  P.ranks = P.ranks > 4 ? 4 : P.ranks;
  P.state_floats = int(mpeghb200_format_conv_state_floats(h));
  const size_t smem = sizeof(float) * (size_t(2) * kWarps * kBufStride + size_t(h->n_in) * 512 +
                                       size_t(h->n_out) * (kYStride + 256 + kErb) + 2 + 64 + 500 + 512);
  MB_CUDA(ctx, cudaFuncSetAttribute(format_conv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)));
  format_conv_kernel<<<n_streams, kWarps * 32, smem, static_cast<cudaStream_t>(cuda_stream)>>>(P);
  ctx->launches.fetch_add(1);
  MB_CUDA(ctx, cudaGetLastError());
  return MPEGHB200_OK;
}

}  // extern "C"
