// capi_fd.cu -- C ABI of the FD synthesis stage (include/mpeghb200.h): device-pointer entry point and the
// host-buffer session (device-resident overlap state + pinned staging + one stream).
#include <cstring>
#include <new>
#include <utility>
#include <vector>

#include "common.h"

namespace mpeghb200 {
mpeghb200_status fd_synth_launch(mpeghb200_ctx* ctx, const float* d_coef, const uint32_t* d_side, float* d_overlap,
                                 float* d_out, int n_units, cudaStream_t stream);
}
using namespace mpeghb200;

struct mpeghb200_fd_session {
  mpeghb200_ctx* ctx = nullptr;
  int n_streams = 0, n_ch = 0;
  size_t n_units = 0;
  static constexpr int kStreams = 3;
  cudaStream_t stream[kStreams] = {nullptr, nullptr, nullptr};
  float* d_coef = nullptr;
  float* d_out = nullptr;
  float* d_overlap = nullptr;
  uint32_t* d_side = nullptr;
  float* p_coef = nullptr;  // pinned staging, allocated on first use with a buffer that cannot be pinned in place
  float* p_out = nullptr;
  uint32_t* p_side = nullptr;
  cudaEvent_t done[8] = {};
};

// Is p DMA-able as it is?  True for page-locked host memory (mpeghb200_host_alloc, cudaHostAlloc, torch
// pin_memory).  Ordinary malloc'ed buffers are staged through the session's pinned copies, chunk by chunk.
static bool dma_ready(const void* p) {
  cudaPointerAttributes at{};
  if (cudaPointerGetAttributes(&at, p) == cudaSuccess && at.type == cudaMemoryTypeHost) return true;
  cudaGetLastError();
  return false;
}

extern "C" {

mpeghb200_status mpeghb200_fd_synth_dev(mpeghb200_ctx* ctx, const float* d_coef, const uint32_t* d_side,
                                        float* d_overlap, float* d_out, int n_units, void* cuda_stream) {
  if (!ctx) return MPEGHB200_ERR_BAD_ARG;
  if (n_units < 0 || (n_units > 0 && (!d_coef || !d_side || !d_overlap || !d_out)))
    return fail(ctx, MPEGHB200_ERR_BAD_ARG, "fd_synth_dev: null buffer or negative unit count");
  MB_CUDA(ctx, cudaSetDevice(ctx->device));
  return fd_synth_launch(ctx, d_coef, d_side, d_overlap, d_out, n_units, static_cast<cudaStream_t>(cuda_stream));
}

mpeghb200_status mpeghb200_fd_open(mpeghb200_ctx* ctx, int n_streams, int n_ch, mpeghb200_fd_session** out) {
  if (!ctx || !out) return MPEGHB200_ERR_BAD_ARG;
  *out = nullptr;
  if (n_streams <= 0 || n_ch <= 0 || n_ch > 64) return fail(ctx, MPEGHB200_ERR_BAD_ARG, "fd_open: bad shape");
  MB_CUDA(ctx, cudaSetDevice(ctx->device));
  auto* s = new (std::nothrow) mpeghb200_fd_session();
  if (!s) return MPEGHB200_ERR_NO_MEM;
  s->ctx = ctx;
  s->n_streams = n_streams;
  s->n_ch = n_ch;
  s->n_units = size_t(n_streams) * n_ch;
  const size_t fb = s->n_units * 1024 * sizeof(float);
  cudaError_t e = cudaSuccess;
  for (int i = 0; i < mpeghb200_fd_session::kStreams && e == cudaSuccess; ++i)
    e = cudaStreamCreateWithFlags(&s->stream[i], cudaStreamNonBlocking);
  for (int i = 0; i < 8 && e == cudaSuccess; ++i) e = cudaEventCreateWithFlags(&s->done[i], cudaEventDisableTiming);
  if (e == cudaSuccess) e = cudaMalloc(&s->d_coef, fb);
  if (e == cudaSuccess) e = cudaMalloc(&s->d_out, fb);
  if (e == cudaSuccess) e = cudaMalloc(&s->d_overlap, fb);
  if (e == cudaSuccess) e = cudaMalloc(&s->d_side, s->n_units * sizeof(uint32_t));
  if (e == cudaSuccess) e = cudaMemset(s->d_overlap, 0, fb);
  if (e != cudaSuccess) {
    mpeghb200_status st = cuda_fail(ctx, e, "fd_open allocation");
    mpeghb200_fd_close(s);
    return st;
  }
  *out = s;
  return MPEGHB200_OK;
}

mpeghb200_status mpeghb200_fd_execute(mpeghb200_fd_session* s, const float* h_coef, const uint32_t* h_side,
                                      float* h_out) {
  if (!s || !h_coef || !h_side || !h_out) return MPEGHB200_ERR_BAD_ARG;
  mpeghb200_ctx* ctx = s->ctx;
  MB_CUDA(ctx, cudaSetDevice(ctx->device));
  const size_t fb = s->n_units * 1024 * sizeof(float);
  const size_t sb = s->n_units * sizeof(uint32_t);
  // Source/destination of the DMA: the caller's buffers when they are page-locked, otherwise pinned staging
  // copies filled / drained chunk by chunk so that the host memcpy of one chunk overlaps the DMA of another.
  const bool pin_coef = dma_ready(h_coef), pin_side = dma_ready(h_side), pin_out = dma_ready(h_out);
  if (!pin_coef && !s->p_coef) MB_CUDA(ctx, cudaHostAlloc(&s->p_coef, fb, cudaHostAllocDefault));
  if (!pin_side && !s->p_side) MB_CUDA(ctx, cudaHostAlloc(&s->p_side, sb, cudaHostAllocDefault));
  if (!pin_out && !s->p_out) MB_CUDA(ctx, cudaHostAlloc(&s->p_out, fb, cudaHostAllocDefault));
  const float* src_coef = pin_coef ? h_coef : s->p_coef;
  const uint32_t* src_side = pin_side ? h_side : s->p_side;
  float* dst_out = pin_out ? h_out : s->p_out;
  if (!pin_side) std::memcpy(s->p_side, h_side, sb);
  // Chunked over three streams: the upload of chunk i+1, the kernel of chunk i and the download of chunk
  // i-1 overlap (PCIe is full duplex; units are independent, so any split is valid).
  const size_t n = s->n_units;
  const size_t n_chunks = n >= 2048 ? 8 : (n >= 512 ? 4 : 1);
  const size_t per = ((n + n_chunks - 1) / n_chunks + 5) / 6 * 6;  // whole warps for every launch shape in use
  size_t used = 0;
  for (size_t a = 0, i = 0; a < n; a += per, ++i) {
    const size_t cnt = (a + per <= n) ? per : n - a;
    cudaStream_t st = s->stream[i % mpeghb200_fd_session::kStreams];
    if (!pin_coef) std::memcpy(s->p_coef + a * 1024, h_coef + a * 1024, cnt * 4096);
    MB_CUDA(ctx, cudaMemcpyAsync(s->d_coef + a * 1024, src_coef + a * 1024, cnt * 4096, cudaMemcpyHostToDevice, st));
    MB_CUDA(ctx, cudaMemcpyAsync(s->d_side + a, src_side + a, cnt * sizeof(uint32_t), cudaMemcpyHostToDevice, st));
    mpeghb200_status stt = fd_synth_launch(ctx, s->d_coef + a * 1024, s->d_side + a, s->d_overlap + a * 1024,
                                           s->d_out + a * 1024, int(cnt), st);
    if (stt != MPEGHB200_OK) return stt;
    MB_CUDA(ctx, cudaMemcpyAsync(dst_out + a * 1024, s->d_out + a * 1024, cnt * 4096, cudaMemcpyDeviceToHost, st));
    MB_CUDA(ctx, cudaEventRecord(s->done[i], st));
    used = i + 1;
  }
  for (size_t a = 0, i = 0; i < used; a += per, ++i) {
    MB_CUDA(ctx, cudaEventSynchronize(s->done[i]));
    const size_t cnt = (a + per <= n) ? per : n - a;
    if (!pin_out) std::memcpy(h_out + a * 1024, s->p_out + a * 1024, cnt * 4096);
  }
  return MPEGHB200_OK;
}

mpeghb200_status mpeghb200_fd_reset_stream(mpeghb200_fd_session* s, int stream_idx) {
  if (!s || stream_idx < 0 || stream_idx >= s->n_streams) return MPEGHB200_ERR_BAD_ARG;
  mpeghb200_ctx* ctx = s->ctx;
  MB_CUDA(ctx, cudaSetDevice(ctx->device));
  const size_t per_stream = size_t(s->n_ch) * 1024;
  MB_CUDA(ctx, cudaMemsetAsync(s->d_overlap + per_stream * stream_idx, 0, per_stream * sizeof(float), s->stream[0]));
  MB_CUDA(ctx, cudaStreamSynchronize(s->stream[0]));
  return MPEGHB200_OK;
}

void mpeghb200_fd_close(mpeghb200_fd_session* s) {
  if (!s) return;
  cudaSetDevice(s->ctx->device);
  for (auto st : s->stream)
    if (st) cudaStreamDestroy(st);
  for (auto ev : s->done)
    if (ev) cudaEventDestroy(ev);
  cudaFree(s->d_coef);
  cudaFree(s->d_out);
  cudaFree(s->d_overlap);
  cudaFree(s->d_side);
  cudaFreeHost(s->p_coef);
  cudaFreeHost(s->p_out);
  cudaFreeHost(s->p_side);
  delete s;
}

}  // extern "C"
