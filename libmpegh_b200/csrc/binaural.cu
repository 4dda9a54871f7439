// binaural.cu -- time-domain binaural renderer: FFT overlap-add fast convolution, one CTA per stream-block.
//
// Stands in for impeghd_binaural_struct_process_ld_ola (decoder/impeghd_binaural_renderer.c:543, the body of
// its block loop), impeghd_binaural_mul_add_perm (:498), the packed-spectrum helpers (:112, :466) and the
// filter preparation of impeghd_binaural_struct_set (:244), on top of the long transforms
// impeghd_cplx_ifft_8k / impeghd_cplx_fft_16k (decoder/impeghd_ifft.c:1264, decoder/impeghd_fft.c:1415).
//
// One block = B = N/2 new samples of each of n_in inputs -> B samples for each ear.  Per input the zero-padded
// block goes through the N-point e^{+j} transform; the transform's column stage leaves every thread with the
// bins it owns in registers, and those are multiplied into per-thread register accumulators for both ears
// (direct filters) and into the weighted mono downmix (diffuse path) right there: no spectrum is ever
// written back except the downmix, which is part of the state.  Accumulation order per bin is the
// reference's (inputs 0..n-1, then diffuse blocks) and every product/sum is rounded separately, so the
// result is bit-identical.  Then, per ear, the Hermitian spectrum is rebuilt in shared memory, transformed
// with the e^{-j} family, scaled by 1/N and overlap-added with the tail kept from the previous block.
//
// HBM traffic per stream-block (algorithmic): inputs n_in*B*4, filters 2*n_in*N*4 (when every stream has its
// own set), tail 2*2*B*4, outputs 2*B*4, downmix ring (1 + n_blocks)*N*4.
//
// Reference quirks kept (oracle/oracle_binaural.c has the long version): swapped bins of the 8-point e^{+j}
// column transform; Im Y[0], Im Y[N/2] of the inverse transform are whatever the shared imaginary scratch
// held (last input's forward transform for the left ear, the left ear's inverse transform for the right).
#include <cstring>
#include <new>
#include <vector>

#include "fft_exact.cuh"

namespace mpeghb200 {
namespace {

using namespace fftx;

struct BinParams {
  const float* in;      // input samples
  long long in_stream_stride, in_chan_stride, in_chunk_stride;  // floats; a chunk = 1024 consecutive samples
  float in_scale, out_scale;
  float* out;           // [stream][2][B]
  float* state;         // [stream][state_floats]: tail [2][B], prev_in [n_blocks + 1][N], write_idx (int bits)
  long long state_floats;
  const int32_t* set;   // [stream] filter-set index or nullptr
  const float* h_direct;   // [set][2][n_in][N] perm spectra
  const float* h_diffuse;  // [set][n_blocks][2][N]
  const int32_t* direct_len;   // [set][2][n_in]
  const int32_t* diffuse_len;  // [set][n_blocks][2]
  const float* weight;         // [set][n_in]
  const float4* eff8;
  const float2* mix;    // [DIM2][1024]
  int n_in, n_blocks;
  int in_vec;  // input rows are 16-byte aligned
};

template <int DIM2>
struct Geo {
  static constexpr int NT = 64 * DIM2;       // threads per CTA
  static constexpr int N = 1024 * DIM2;      // transform length
  static constexpr int B = N / 2;            // block length
  static constexpr int HB = DIM2 / 2;        // bins per column kept (k < N/2)
  static constexpr int COLS = 1024 / NT;     // columns per thread
  static constexpr int NB = COLS * HB;       // bins per thread (= 8)
  // shared memory: transform rows | row twiddles (fft_exact.cuh RowTw) | inter-stage twiddles [DIM2][1024]
  static constexpr int kWorkF2 = DIM2 * kRowStride;
  static constexpr int kTwF2 = (kRowTwFloats + 3) / 4 * 2;  // even: what follows stays 16-byte aligned
  static constexpr int kTabF2 = kWorkF2 + kTwF2 + DIM2 * 1024;
  // block kernel only: | filter spectra of the current input, both ears [2][N] | samples of the next input [B]
  static constexpr size_t kSmemTab = size_t(kTabF2) * sizeof(float2);
  static constexpr size_t kSmem = kSmemTab + size_t(2 * N + B) * sizeof(float);
};

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(unsigned(__cvta_generic_to_shared(smem_dst))), "l"(src)
               : "memory");
}
__device__ __forceinline__ void cp_async4(void* smem_dst, const void* src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(unsigned(__cvta_generic_to_shared(smem_dst))), "l"(src)
               : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

// Tables every CTA stages once; returns the row twiddles, *mix_s = the staged inter-stage twiddles.
template <int DIM2>
__device__ __forceinline__ RowTw stage_tables(float2* __restrict__ work, const float4* __restrict__ eff8,
                                              const float2* __restrict__ mix, const float2** mix_s) {
  using G = Geo<DIM2>;
  const RowTw tw = fill_row_tw(reinterpret_cast<float*>(work + G::kWorkF2), eff8);
  float2* m = work + G::kWorkF2 + G::kTwF2;
  for (int i = threadIdx.x; i < DIM2 * 1024; i += G::NT) m[i] = __ldg(mix + i);
  *mix_s = m;
  return tw;  // the first __syncthreads() of forward_real orders these writes before their first use
}

// Streaming 8-byte load that does not allocate in L1: the filter spectra (1.5 MB per stream-block) must not evict
// what the transforms keep re-reading.
__device__ __forceinline__ float2 ld_stream(const float2* p) {
  float2 v;
  asm volatile("ld.global.nc.L1::no_allocate.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "l"(p));
  return v;
}

// Forward transform of one real, zero-padded block.  src(n) returns sample n < n_valid (scaled).  On return
// X[b] = (Re X[k], -Im X[k]) for the thread's bins k = (b % HB) * 1024 + tid + NT * (b / HB); for thread 0
// X[0] = (Re X[0], Re X[N/2]); im0/imn = Im X[0], Im X[N/2] (thread 0 only).
// prefetch() runs once every thread has staged its samples: the place to start asynchronous copies into buffers the
// previous block of work was still reading; they are complete and visible to the CTA on return.
template <int DIM2, typename Src, typename Prefetch>
__device__ __forceinline__ void forward_real(float2* __restrict__ work, const RowTw tw,
                                             const float2* __restrict__ mix, Src src, int n_valid,
                                             float2 (&X)[8], float& im0, float& imn, Prefetch prefetch) {
  using G = Geo<DIM2>;
  const int tid = threadIdx.x;
  // only the first half of every row is non-zero and all of it is real: staged as floats (see rows_1024 REAL_HALF)
  float* workf = reinterpret_cast<float*>(work);
#pragma unroll
  for (int q = 0; q < G::B / G::NT; ++q) {
    const int n = tid + G::NT * q;
    const float x = (n < n_valid) ? src(n) : 0.0f;
    workf[(n % DIM2) * (2 * kRowStride) + phys(n / DIM2)] = x;
  }
  __syncthreads();
  prefetch();
  rows_1024<DIM2, true, true>(work, tw);
#pragma unroll
  for (int c = 0; c < G::COLS; ++c) {
    const int i = tid + G::NT * c;
    float xr[DIM2], xi[DIM2];
    column<DIM2, true>(work, mix, i, xr, xi);
#pragma unroll
    for (int j = 0; j < G::HB; ++j) X[c * G::HB + j] = make_float2(xr[j], -xi[j]);
    if (c == 0 && tid == 0) {
      X[0] = make_float2(xr[0], xr[G::HB]);
      im0 = xi[0];
      imn = xi[G::HB];
    }
  }
  cp_async_wait_all();
  __syncthreads();  // the caller may overwrite work
}

// dst += h * x in the packed format (impeghd_binaural_mul_add_perm): k == 0 is the (DC, Nyquist) pair.
__device__ __forceinline__ void mac_perm(float2& d, float2 h, float2 x, int k, int len) {
  if (k == 0) {
    d.x = fadd(d.x, fmul(h.x, x.x));
    d.y = fadd(d.y, fmul(h.y, x.y));
  } else if (k < len) {
    d.x = fsub(fadd(d.x, fmul(h.x, x.x)), fmul(h.y, x.y));
    d.y = fadd(fadd(d.y, fmul(h.x, x.y)), fmul(h.y, x.x));
  }
}

template <int DIM2>
__global__ void __launch_bounds__(64 * DIM2) binaural_block_kernel(const __grid_constant__ BinParams P) {
  using G = Geo<DIM2>;
  extern __shared__ float2 work[];
  const int tid = threadIdx.x;
  const size_t s = blockIdx.x;
  const int set = P.set ? P.set[s] : 0;
  float* state = P.state + s * P.state_floats;
  float* tail = state;
  float* prev_in = state + 2 * G::B;
  int* widx_p = reinterpret_cast<int*>(state + 2 * G::B + size_t(P.n_blocks + 1) * G::N);
  const int widx = *widx_p;
  const float* hd = P.h_direct + size_t(set) * 2 * P.n_in * G::N;
  const int32_t* dlen = P.direct_len + size_t(set) * 2 * P.n_in;
  const float* wgt = P.weight + size_t(set) * P.n_in;

  const float2* mix_s;
  const RowTw tw = stage_tables<DIM2>(work, P.eff8, P.mix, &mix_s);

  int kbin[G::NB];
#pragma unroll
  for (int b = 0; b < G::NB; ++b) kbin[b] = (b % G::HB) * 1024 + tid + G::NT * (b / G::HB);

  float2 acc[2][G::NB], dmx[G::NB];
#pragma unroll
  for (int b = 0; b < G::NB; ++b) acc[0][b] = acc[1][b] = dmx[b] = make_float2(0.0f, 0.0f);
  float im0 = 0.0f, imn = 0.0f;

  // The filter spectra of an input (both ears, 2 N floats from DRAM when every stream has its own set) and the samples
  // of the next input are copied into shared memory asynchronously while the input is being transformed.
  float* hbuf = reinterpret_cast<float*>(work + G::kTabF2);
  float* xs = hbuf + 2 * G::N;
  auto fetch_input = [&](int inp) {
    const float* x = P.in + s * P.in_stream_stride + inp * P.in_chan_stride;
    if (P.in_vec) {
      for (int n = 4 * tid; n < G::B; n += 4 * G::NT) cp_async16(xs + n, x + (n >> 10) * P.in_chunk_stride + (n & 1023));
    } else {
      for (int n = tid; n < G::B; n += G::NT) cp_async4(xs + n, x + (n >> 10) * P.in_chunk_stride + (n & 1023));
    }
  };
  auto fetch_filters = [&](int inp) {
    const float* h0 = hd + size_t(inp) * G::N;
    const float* h1 = hd + size_t(P.n_in + inp) * G::N;
    for (int n = 4 * tid; n < G::N; n += 4 * G::NT) {
      cp_async16(hbuf + n, h0 + n);
      cp_async16(hbuf + G::N + n, h1 + n);
    }
  };
  fetch_input(0);
  cp_async_wait_all();
  __syncthreads();

  for (int inp = 0; inp < P.n_in; ++inp) {
    float2 X[G::NB];
    forward_real<DIM2>(
        work, tw, mix_s, [&](int n) { return fmul(xs[n], P.in_scale); }, G::B, X, im0, imn, [&]() {
          fetch_filters(inp);
          if (inp + 1 < P.n_in) fetch_input(inp + 1);
        });
    const int l0 = dlen[inp], l1 = dlen[P.n_in + inp];
    const float w = wgt[inp];
#pragma unroll
    for (int b = 0; b < G::NB; ++b) {
      const float2 a0 = reinterpret_cast<const float2*>(hbuf)[kbin[b]];
      const float2 a1 = reinterpret_cast<const float2*>(hbuf + G::N)[kbin[b]];
      mac_perm(acc[0][b], a0, X[b], kbin[b], l0);
      mac_perm(acc[1][b], a1, X[b], kbin[b], l1);
      if (P.n_blocks) {
        if (inp == 0) {
          dmx[b] = make_float2(fmul(X[b].x, w), fmul(X[b].y, w));
        } else {
          dmx[b] = make_float2(fadd(dmx[b].x, fmul(X[b].x, w)), fadd(dmx[b].y, fmul(X[b].y, w)));
        }
      }
    }
  }

  // diffuse path: this block's downmix goes into the ring, older ones are convolved
  if (P.n_blocks) {
    float2* slot = reinterpret_cast<float2*>(prev_in + size_t(widx) * G::N);
#pragma unroll
    for (int b = 0; b < G::NB; ++b) slot[kbin[b]] = dmx[b];
    const float* hf = P.h_diffuse + size_t(set) * P.n_blocks * 2 * G::N;
    const int32_t* flen = P.diffuse_len + size_t(set) * P.n_blocks * 2;
    int k = (widx + P.n_blocks) % (P.n_blocks + 1);
    for (int j = 0; j < P.n_blocks; ++j) {
      const float2* pin = reinterpret_cast<const float2*>(prev_in + size_t(k) * G::N);
#pragma unroll
      for (int b = 0; b < G::NB; ++b) {
        const float2 xin = pin[kbin[b]];
#pragma unroll
        for (int o = 0; o < 2; ++o) {
          const float2 h = ld_stream(reinterpret_cast<const float2*>(hf + size_t(j * 2 + o) * G::N) + kbin[b]);
          mac_perm(acc[o][b], h, xin, kbin[b], flen[j * 2 + o]);
        }
      }
      k = (k + P.n_blocks) % (P.n_blocks + 1);
    }
  }

  const float inv_n = 1.0f / float(G::N);
#pragma unroll
  for (int o = 0; o < 2; ++o) {
    // Hermitian spectrum into the row layout: element n -> row n % DIM2, position n / DIM2
#pragma unroll
    for (int b = 0; b < G::NB; ++b) {
      const int k = kbin[b];
      if (k == 0) {
        work[0] = make_float2(acc[o][b].x, im0);
        work[(G::B % DIM2) * kRowStride + phys(G::B / DIM2)] = make_float2(acc[o][b].y, imn);
      } else {
        const float yi = -acc[o][b].y;
        const int m = G::N - k;
        work[(k % DIM2) * kRowStride + phys(k / DIM2)] = make_float2(acc[o][b].x, yi);
        work[(m % DIM2) * kRowStride + phys(m / DIM2)] = make_float2(acc[o][b].x, -yi);
      }
    }
    __syncthreads();
    rows_1024<DIM2, false>(work, tw);
    float* yo = P.out + (s * 2 + o) * G::B;
    float* tl = tail + o * G::B;
#pragma unroll
    for (int c = 0; c < G::COLS; ++c) {
      const int i = tid + G::NT * c;
      float xr[DIM2], xi[DIM2];
      column<DIM2, false>(work, mix_s, i, xr, xi);
#pragma unroll
      for (int j = 0; j < G::HB; ++j) {
        const int t = j * 1024 + i;
        const float v = fadd(fmul(xr[j], inv_n), tl[t]);
        __stcs(yo + t, fmul(v, P.out_scale));
        tl[t] = fmul(xr[j + G::HB], inv_n);
      }
      if (c == 0 && tid == 0) {
        im0 = xi[0];
        imn = xi[G::HB];
      }
    }
    __syncthreads();
  }
  if (tid == 0) *widx_p = (widx + 1) % (P.n_blocks + 1);
}

// Filter preparation: one CTA per filter, taps[len] -> packed spectrum [N].
template <int DIM2>
__global__ void __launch_bounds__(64 * DIM2)
    binaural_spectrum_kernel(const float* __restrict__ taps, int len, float* __restrict__ spec,
                             const float4* __restrict__ eff8, const float2* __restrict__ mix) {
  using G = Geo<DIM2>;
  extern __shared__ float2 work[];
  const float* t = taps + size_t(blockIdx.x) * len;
  float2 X[G::NB];
  float im0, imn;
  const float2* mix_s;
  const RowTw tw = stage_tables<DIM2>(work, eff8, mix, &mix_s);
  forward_real<DIM2>(work, tw, mix_s, [&](int n) { return t[n]; }, len < G::B ? len : G::B, X, im0, imn, []() {});
  float2* o = reinterpret_cast<float2*>(spec + size_t(blockIdx.x) * G::N);
#pragma unroll
  for (int b = 0; b < G::NB; ++b) o[(b % G::HB) * 1024 + threadIdx.x + G::NT * (b / G::HB)] = X[b];
}

}  // namespace
}  // namespace mpeghb200

using namespace mpeghb200;

struct mpeghb200_binaural {
  mpeghb200_ctx* ctx = nullptr;
  int n_in = 0, len = 0, n_blocks = 0, n_sets = 0, dim2 = 0, N = 0;
  float* d_hdirect = nullptr;
  float* d_hdiffuse = nullptr;
  int32_t* d_dlen = nullptr;
  int32_t* d_flen = nullptr;
  float* d_weight = nullptr;
  float2* d_mix = nullptr;
  // transform lengths without a tuned kernel (N <= 1024, N = 16384): binaural_generic.cu
  bool generic = false;
  float* d_w = nullptr;
  float* d_mc = nullptr;
  float* d_ms = nullptr;
  void release() {
    cudaFree(d_hdirect), cudaFree(d_hdiffuse), cudaFree(d_dlen), cudaFree(d_flen), cudaFree(d_weight), cudaFree(d_mix);
    cudaFree(d_w), cudaFree(d_mc), cudaFree(d_ms);
  }
};

namespace mpeghb200 {
size_t binaural_generic_work_floats(int N);
cudaError_t binaural_generic_tables(float** d_w, float** d_mc, float** d_ms);
cudaError_t binaural_generic_spectra(const float* d_taps, int len, int n_filters, float* d_spec, int N, const float* d_w,
                                     const float* d_mc, const float* d_ms);
cudaError_t binaural_generic_block(const float* d_in, long long in_stream_stride, long long in_chan_stride,
                                   long long in_chunk_stride, float in_scale, float out_scale, float* d_out, float* d_state,
                                   long long state_floats, const int32_t* d_set, const float* h_direct,
                                   const float* h_diffuse, const int32_t* dlen, const int32_t* flen, const float* weight,
                                   int n_in, int n_blocks, int N, const float* d_w, const float* d_mc, const float* d_ms,
                                   int n_streams, cudaStream_t st);
}  // namespace mpeghb200
// which transform lengths exist: 0 = none, 1 = tuned kernels, 2 = generic path
static int binaural_size_class(int N) {
  if (N == 2048 || N == 8192) return 1;
  if ((N >= 16 && N <= 1024) || N == 16384) return 2;
  return 0;
}

template <int DIM2>
static cudaError_t run_spectra(const mpeghb200_binaural* b, const float* d_taps, int n_filters, float* d_spec) {
  using G = Geo<DIM2>;
  cudaError_t e = cudaFuncSetAttribute(binaural_spectrum_kernel<DIM2>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       int(G::kSmemTab));
  if (e != cudaSuccess) return e;
  binaural_spectrum_kernel<DIM2><<<n_filters, G::NT, G::kSmemTab>>>(d_taps, b->len, d_spec, b->ctx->tables.eff8, b->d_mix);
  return cudaGetLastError();
}

extern "C" {

mpeghb200_status mpeghb200_binaural_create(mpeghb200_ctx* ctx, int n_in, int len_direct, int n_diffuse_blocks,
                                           int n_sets, const float* taps_direct, const float* taps_diffuse,
                                           const float* direct_fc, const float* diffuse_fc,
                                           const float* inv_diffuse_weight, mpeghb200_binaural** out) {
  if (!ctx || !out) return MPEGHB200_ERR_BAD_ARG;
  *out = nullptr;
  if (n_in < 1 || n_in > 32 || n_diffuse_blocks < 0 || n_diffuse_blocks > 2 || n_sets < 1 || len_direct < 1 ||
      !taps_direct || !direct_fc || !inv_diffuse_weight || (n_diffuse_blocks && (!taps_diffuse || !diffuse_fc)))
    return fail(ctx, MPEGHB200_ERR_BAD_ARG, "binaural_create: 1..32 inputs, 0..2 diffuse blocks, non-null tables");
  int N = 2;
  while (N < 2 * len_direct) N <<= 1;  // impeghd_binaural_struct_init (decoder/impeghd_binaural_renderer.c:200-205)
  if (!binaural_size_class(N))
    return fail(ctx, MPEGHB200_ERR_UNSUPPORTED,
                N == 4096 ? "binaural_create: direct lengths of 1025 .. 2048 taps give a 4096-point transform, for which the "
                            "reference's output depends on stale scratch memory (impeghd_cplx_ifft_4 reads behind its "
                            "buffer, decoder/impeghd_ifft.c:1225-1228): there is no result to reproduce"
                          : "binaural_create: direct length must be 8 .. 8192 taps");
  MB_CUDA(ctx, cudaSetDevice(ctx->device));
  auto* b = new (std::nothrow) mpeghb200_binaural();
  if (!b) return MPEGHB200_ERR_NO_MEM;
  b->ctx = ctx, b->n_in = n_in, b->len = len_direct, b->n_blocks = n_diffuse_blocks, b->n_sets = n_sets;
  b->N = N, b->dim2 = N / 1024;
  b->generic = binaural_size_class(N) == 2;
  const int B = N / 2;
  const size_t n_dir = size_t(n_sets) * 2 * n_in, n_dif = size_t(n_sets) * n_diffuse_blocks * 2;

  // inter-stage twiddles [j][i] from the [1024][16] table (decoder/impeghd_fft_ifft_rom.c:1122,4401)
  std::vector<float2> mix(b->generic ? 1 : size_t(b->dim2) * 1024);
  if (!b->generic) {
    const std::vector<float>& mc = rom_table(ROM_MIXED_RAD_COS);
    const std::vector<float>& ms = rom_table(ROM_MIXED_RAD_SIN);
    const int fac = 16 / b->dim2;
    for (int j = 0; j < b->dim2; ++j)
      for (int i = 0; i < 1024; ++i)
        mix[size_t(j) * 1024 + i] = make_float2(mc[(i * b->dim2 + j) * fac], ms[(i * b->dim2 + j) * fac]);
  }
  // MAC lengths: (WORD32)(min(fc, 1) * filter_size) (renderer.c:312-330)
  std::vector<int32_t> dlen(n_dir), flen(n_dif ? n_dif : 1);
  for (size_t i = 0; i < n_dir; ++i) dlen[i] = int32_t((direct_fc[i] < 1.0f ? direct_fc[i] : 1.0f) * float(B));
  for (int st = 0; st < n_sets; ++st)
    for (int j = 0; j < n_diffuse_blocks; ++j)
      for (int o = 0; o < 2; ++o) {
        const float fc = diffuse_fc[(size_t(st) * 2 + o) * n_diffuse_blocks + j];  // [set][2][n_blocks]
        flen[(size_t(st) * n_diffuse_blocks + j) * 2 + o] = int32_t((fc < 1.0f ? fc : 1.0f) * float(B));
      }

  float* d_taps = nullptr;
  cudaError_t e = cudaMalloc(&b->d_mix, mix.size() * sizeof(float2));
  if (e == cudaSuccess) e = cudaMemcpy(b->d_mix, mix.data(), mix.size() * sizeof(float2), cudaMemcpyHostToDevice);
  if (e == cudaSuccess && b->generic) e = binaural_generic_tables(&b->d_w, &b->d_mc, &b->d_ms);
  if (e == cudaSuccess) e = cudaMalloc(&b->d_hdirect, n_dir * N * sizeof(float));
  if (e == cudaSuccess) e = cudaMalloc(&b->d_hdiffuse, (n_dif ? n_dif : 1) * N * sizeof(float));
  if (e == cudaSuccess) e = cudaMalloc(&b->d_dlen, dlen.size() * sizeof(int32_t));
  if (e == cudaSuccess) e = cudaMalloc(&b->d_flen, flen.size() * sizeof(int32_t));
  if (e == cudaSuccess) e = cudaMalloc(&b->d_weight, size_t(n_sets) * n_in * sizeof(float));
  if (e == cudaSuccess) e = cudaMemcpy(b->d_dlen, dlen.data(), dlen.size() * sizeof(int32_t), cudaMemcpyHostToDevice);
  if (e == cudaSuccess) e = cudaMemcpy(b->d_flen, flen.data(), flen.size() * sizeof(int32_t), cudaMemcpyHostToDevice);
  if (e == cudaSuccess)
    e = cudaMemcpy(b->d_weight, inv_diffuse_weight, size_t(n_sets) * n_in * sizeof(float), cudaMemcpyHostToDevice);
  const size_t max_f = n_dir > n_dif ? n_dir : n_dif;
  if (e == cudaSuccess) e = cudaMalloc(&d_taps, max_f * len_direct * sizeof(float));
  for (int pass = 0; pass < 2 && e == cudaSuccess; ++pass) {
    const size_t nf = pass ? n_dif : n_dir;
    if (!nf) continue;
    e = cudaMemcpy(d_taps, pass ? taps_diffuse : taps_direct, nf * len_direct * sizeof(float), cudaMemcpyHostToDevice);
    if (e == cudaSuccess && b->generic)
      e = binaural_generic_spectra(d_taps, len_direct, int(nf), pass ? b->d_hdiffuse : b->d_hdirect, N, b->d_w, b->d_mc,
                                   b->d_ms);
    else if (e == cudaSuccess)
      e = (b->dim2 == 8) ? run_spectra<8>(b, d_taps, int(nf), pass ? b->d_hdiffuse : b->d_hdirect)
                         : run_spectra<2>(b, d_taps, int(nf), pass ? b->d_hdiffuse : b->d_hdirect);
    ctx->launches.fetch_add(1);
    if (e == cudaSuccess) e = cudaDeviceSynchronize();
  }
  cudaFree(d_taps);
  if (e != cudaSuccess) {
    b->release();
    delete b;
    return cuda_fail(ctx, e, "binaural_create");
  }
  *out = b;
  return MPEGHB200_OK;
}

mpeghb200_status mpeghb200_binaural_create_from_state(mpeghb200_ctx* ctx, int fft_size, int n_in, int n_diffuse_blocks,
                                                      const float* h_direct, size_t direct_row_stride,
                                                      size_t direct_ear_stride, const float* h_diffuse,
                                                      size_t diffuse_row_stride,
                                                      const int32_t* direct_process_size, int direct_size_stride,
                                                      const int32_t* diffuse_process_size,
                                                      const float* diffuse_weight, mpeghb200_binaural** out) {
  if (!ctx || !out) return MPEGHB200_ERR_BAD_ARG;
  *out = nullptr;
  if (n_in < 1 || n_in > 32 || n_diffuse_blocks < 0 || n_diffuse_blocks > 2 || !h_direct || !direct_process_size ||
      !diffuse_weight || direct_size_stride < n_in || direct_row_stride < size_t(fft_size) ||
      direct_ear_stride < size_t(n_in) * direct_row_stride ||
      (n_diffuse_blocks && (!h_diffuse || !diffuse_process_size || diffuse_row_stride < size_t(fft_size))))
    return fail(ctx, MPEGHB200_ERR_BAD_ARG, "binaural_create_from_state: 1..32 inputs, 0..2 diffuse blocks, non-null tables");
  if (!binaural_size_class(fft_size))
    return fail(ctx, MPEGHB200_ERR_UNSUPPORTED,
                "binaural_create_from_state: transform lengths 16 .. 1024, 2048, 8192 and 16384 (the reference's "
                "4096-point result depends on stale scratch memory)");
  MB_CUDA(ctx, cudaSetDevice(ctx->device));
  auto* b = new (std::nothrow) mpeghb200_binaural();
  if (!b) return MPEGHB200_ERR_NO_MEM;
  const int N = fft_size;
  b->ctx = ctx, b->n_in = n_in, b->len = N / 2, b->n_blocks = n_diffuse_blocks, b->n_sets = 1, b->N = N, b->dim2 = N / 1024;
  b->generic = binaural_size_class(N) == 2;
  std::vector<float2> mix(b->generic ? 1 : size_t(b->dim2) * 1024);
  if (!b->generic) {
    const std::vector<float>& mc = rom_table(ROM_MIXED_RAD_COS);
    const std::vector<float>& ms = rom_table(ROM_MIXED_RAD_SIN);
    const int fac = 16 / b->dim2;
    for (int j = 0; j < b->dim2; ++j)
      for (int i = 0; i < 1024; ++i)
        mix[size_t(j) * 1024 + i] = make_float2(mc[(i * b->dim2 + j) * fac], ms[(i * b->dim2 + j) * fac]);
  }
  const size_t n_dir = size_t(2) * n_in, n_dif = size_t(n_diffuse_blocks) * 2;
  std::vector<int32_t> dlen(n_dir), flen(n_dif ? n_dif : 1);
  for (int o = 0; o < 2; ++o)
    for (int i = 0; i < n_in; ++i) dlen[size_t(o) * n_in + i] = direct_process_size[o * direct_size_stride + i];
  for (size_t i = 0; i < n_dif; ++i) flen[i] = diffuse_process_size[i];  // [block][2]
  cudaError_t e = cudaMalloc(&b->d_mix, mix.size() * sizeof(float2));
  if (e == cudaSuccess) e = cudaMemcpy(b->d_mix, mix.data(), mix.size() * sizeof(float2), cudaMemcpyHostToDevice);
  if (e == cudaSuccess && b->generic) e = binaural_generic_tables(&b->d_w, &b->d_mc, &b->d_ms);
  if (e == cudaSuccess) e = cudaMalloc(&b->d_hdirect, n_dir * N * sizeof(float));
  if (e == cudaSuccess) e = cudaMalloc(&b->d_hdiffuse, (n_dif ? n_dif : 1) * N * sizeof(float));
  if (e == cudaSuccess) e = cudaMalloc(&b->d_dlen, dlen.size() * sizeof(int32_t));
  if (e == cudaSuccess) e = cudaMalloc(&b->d_flen, flen.size() * sizeof(int32_t));
  if (e == cudaSuccess) e = cudaMalloc(&b->d_weight, size_t(n_in) * sizeof(float));
  if (e == cudaSuccess) e = cudaMemcpy(b->d_dlen, dlen.data(), dlen.size() * sizeof(int32_t), cudaMemcpyHostToDevice);
  if (e == cudaSuccess) e = cudaMemcpy(b->d_flen, flen.data(), flen.size() * sizeof(int32_t), cudaMemcpyHostToDevice);
  if (e == cudaSuccess) e = cudaMemcpy(b->d_weight, diffuse_weight, size_t(n_in) * sizeof(float), cudaMemcpyHostToDevice);
  for (int o = 0; o < 2 && e == cudaSuccess; ++o)  // rows of N floats out of rows of row_stride floats
    e = cudaMemcpy2D(b->d_hdirect + size_t(o) * n_in * N, size_t(N) * sizeof(float), h_direct + o * direct_ear_stride,
                     direct_row_stride * sizeof(float), size_t(N) * sizeof(float), size_t(n_in), cudaMemcpyHostToDevice);
  if (e == cudaSuccess && n_dif)
    e = cudaMemcpy2D(b->d_hdiffuse, size_t(N) * sizeof(float), h_diffuse, diffuse_row_stride * sizeof(float),
                     size_t(N) * sizeof(float), n_dif, cudaMemcpyHostToDevice);
  if (e != cudaSuccess) {
    b->release();
    delete b;
    return cuda_fail(ctx, e, "binaural_create_from_state");
  }
  *out = b;
  return MPEGHB200_OK;
}

void mpeghb200_binaural_destroy(mpeghb200_binaural* b) {
  if (!b) return;
  cudaSetDevice(b->ctx->device);
  b->release();
  delete b;
}

int mpeghb200_binaural_block_len(const mpeghb200_binaural* b) { return b ? b->N / 2 : 0; }
int mpeghb200_binaural_num_in(const mpeghb200_binaural* b) { return b ? b->n_in : 0; }

size_t mpeghb200_binaural_state_floats(const mpeghb200_binaural* b) {
  if (!b) return 0;
  // tail [2][B], ring, write index (+ padding); the generic path keeps its per-stream workspace behind them
  return size_t(b->N) + size_t(b->n_blocks + 1) * b->N + 4 + (b->generic ? binaural_generic_work_floats(b->N) : 0);
}

mpeghb200_status mpeghb200_binaural_spectra(mpeghb200_ctx* ctx, const mpeghb200_binaural* b, float* h_direct,
                                            float* h_diffuse) {
  if (!ctx || !b) return MPEGHB200_ERR_BAD_ARG;
  MB_CUDA(ctx, cudaSetDevice(ctx->device));
  if (h_direct)
    MB_CUDA(ctx, cudaMemcpy(h_direct, b->d_hdirect, size_t(b->n_sets) * 2 * b->n_in * b->N * sizeof(float),
                            cudaMemcpyDeviceToHost));
  if (h_diffuse && b->n_blocks)
    MB_CUDA(ctx, cudaMemcpy(h_diffuse, b->d_hdiffuse, size_t(b->n_sets) * b->n_blocks * 2 * b->N * sizeof(float),
                            cudaMemcpyDeviceToHost));
  return MPEGHB200_OK;
}

mpeghb200_status mpeghb200_binaural_block_dev(mpeghb200_ctx* ctx, const mpeghb200_binaural* b, const float* d_in,
                                              int in_layout, float in_scale, float out_scale, float* d_out,
                                              float* d_state, const int32_t* d_filter_set, int n_streams,
                                              void* cuda_stream) {
  if (!ctx || !b) return MPEGHB200_ERR_BAD_ARG;
  if (n_streams < 0 || (in_layout != 0 && in_layout != 1) || (n_streams > 0 && (!d_in || !d_out || !d_state)))
    return fail(ctx, MPEGHB200_ERR_BAD_ARG, "binaural_block_dev: null buffer, bad layout or negative stream count");
  if (n_streams == 0) return MPEGHB200_OK;
  MB_CUDA(ctx, cudaSetDevice(ctx->device));
  BinParams P{};
  const long long B = b->N / 2;
  P.in = d_in;
  if (in_layout == 0) {  // [stream][n_in][B]
    P.in_stream_stride = b->n_in * B, P.in_chan_stride = B, P.in_chunk_stride = 1024;
  } else {  // [B/1024][stream][n_in][1024]: consecutive decoder frames stacked
    P.in_stream_stride = b->n_in * 1024LL, P.in_chan_stride = 1024, P.in_chunk_stride = (long long)n_streams * b->n_in * 1024;
  }
  P.in_scale = in_scale, P.out_scale = out_scale;
  P.in_vec = (reinterpret_cast<uintptr_t>(d_in) & 15) == 0;
  P.out = d_out, P.state = d_state, P.state_floats = (long long)mpeghb200_binaural_state_floats(b);
  P.set = d_filter_set;
  P.h_direct = b->d_hdirect, P.h_diffuse = b->d_hdiffuse, P.direct_len = b->d_dlen, P.diffuse_len = b->d_flen;
  P.weight = b->d_weight, P.eff8 = ctx->tables.eff8, P.mix = b->d_mix;
  P.n_in = b->n_in, P.n_blocks = b->n_blocks;
  cudaStream_t st = static_cast<cudaStream_t>(cuda_stream);
  cudaError_t e;
  if (b->generic) {
    if (in_layout == 1 && (B % 1024))
      return fail(ctx, MPEGHB200_ERR_BAD_ARG, "binaural_block_dev: layout 1 stacks whole 1024-sample frames; this renderer's "
                                              "block is shorter");
    e = binaural_generic_block(d_in, P.in_stream_stride, P.in_chan_stride, P.in_chunk_stride, in_scale, out_scale, d_out,
                               d_state, P.state_floats, d_filter_set, b->d_hdirect, b->d_hdiffuse, b->d_dlen, b->d_flen,
                               b->d_weight, b->n_in, b->n_blocks, b->N, b->d_w, b->d_mc, b->d_ms, n_streams, st);
  } else if (b->dim2 == 8) {
    e = cudaFuncSetAttribute(binaural_block_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(Geo<8>::kSmem));
    if (e == cudaSuccess) binaural_block_kernel<8><<<n_streams, Geo<8>::NT, Geo<8>::kSmem, st>>>(P);
  } else {
    e = cudaFuncSetAttribute(binaural_block_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(Geo<2>::kSmem));
    if (e == cudaSuccess) binaural_block_kernel<2><<<n_streams, Geo<2>::NT, Geo<2>::kSmem, st>>>(P);
  }
  if (e == cudaSuccess) e = cudaGetLastError();
  ctx->launches.fetch_add(1);
  MB_CUDA(ctx, e);
  return MPEGHB200_OK;
}

}  // extern "C"
