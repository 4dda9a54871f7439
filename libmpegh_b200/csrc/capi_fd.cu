// capi_fd.cu -- C ABI of the FD synthesis stage (include/mpeghb200.h): device-pointer entry point and the
// host-buffer session (device-resident overlap state + pinned staging + one stream).
#include <cstring>
#include <new>

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
  cudaStream_t stream = nullptr;
  float* d_coef = nullptr;
  float* d_out = nullptr;
  float* d_overlap = nullptr;
  uint32_t* d_side = nullptr;
  float* p_coef = nullptr;  // pinned staging
  float* p_out = nullptr;
  uint32_t* p_side = nullptr;
};

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
  cudaError_t e = cudaStreamCreateWithFlags(&s->stream, cudaStreamNonBlocking);
  if (e == cudaSuccess) e = cudaMalloc(&s->d_coef, fb);
  if (e == cudaSuccess) e = cudaMalloc(&s->d_out, fb);
  if (e == cudaSuccess) e = cudaMalloc(&s->d_overlap, fb);
  if (e == cudaSuccess) e = cudaMalloc(&s->d_side, s->n_units * sizeof(uint32_t));
  if (e == cudaSuccess) e = cudaMemset(s->d_overlap, 0, fb);
  if (e == cudaSuccess) e = cudaHostAlloc(&s->p_coef, fb, cudaHostAllocDefault);
  if (e == cudaSuccess) e = cudaHostAlloc(&s->p_out, fb, cudaHostAllocDefault);
  if (e == cudaSuccess) e = cudaHostAlloc(&s->p_side, s->n_units * sizeof(uint32_t), cudaHostAllocDefault);
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
  // Caller buffers are ordinary pageable memory (the reference's mem tables are malloc'ed by the caller):
  // stage through pinned memory so the DMA engines run at full PCIe rate and the copies are truly async.
  std::memcpy(s->p_coef, h_coef, fb);
  std::memcpy(s->p_side, h_side, sb);
  MB_CUDA(ctx, cudaMemcpyAsync(s->d_coef, s->p_coef, fb, cudaMemcpyHostToDevice, s->stream));
  MB_CUDA(ctx, cudaMemcpyAsync(s->d_side, s->p_side, sb, cudaMemcpyHostToDevice, s->stream));
  mpeghb200_status st =
      fd_synth_launch(ctx, s->d_coef, s->d_side, s->d_overlap, s->d_out, int(s->n_units), s->stream);
  if (st != MPEGHB200_OK) return st;
  MB_CUDA(ctx, cudaMemcpyAsync(s->p_out, s->d_out, fb, cudaMemcpyDeviceToHost, s->stream));
  MB_CUDA(ctx, cudaStreamSynchronize(s->stream));
  std::memcpy(h_out, s->p_out, fb);
  return MPEGHB200_OK;
}

mpeghb200_status mpeghb200_fd_reset_stream(mpeghb200_fd_session* s, int stream_idx) {
  if (!s || stream_idx < 0 || stream_idx >= s->n_streams) return MPEGHB200_ERR_BAD_ARG;
  mpeghb200_ctx* ctx = s->ctx;
  MB_CUDA(ctx, cudaSetDevice(ctx->device));
  const size_t per_stream = size_t(s->n_ch) * 1024;
  MB_CUDA(ctx, cudaMemsetAsync(s->d_overlap + per_stream * stream_idx, 0, per_stream * sizeof(float), s->stream));
  MB_CUDA(ctx, cudaStreamSynchronize(s->stream));
  return MPEGHB200_OK;
}

void mpeghb200_fd_close(mpeghb200_fd_session* s) {
  if (!s) return;
  cudaSetDevice(s->ctx->device);
  if (s->stream) cudaStreamDestroy(s->stream);
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
