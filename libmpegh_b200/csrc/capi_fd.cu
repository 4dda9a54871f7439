// capi_fd.cu -- C ABI of the FD synthesis stage (include/mpeghb200.h): device-pointer entry point and the
// host-buffer session (device-resident overlap state + pinned staging + one stream).
#include <cstring>
#include <new>
#include <utility>
#include <vector>

#include "common.h"

namespace mpeghb200 {
mpeghb200_status fd_synth_frames_launch(mpeghb200_ctx* ctx, const float* d_coef, const uint32_t* d_side,
                                        float* d_overlap, float* d_out, int n_units, int n_frames, cudaStream_t stream);
mpeghb200_status fd_synth_launch(mpeghb200_ctx* ctx, const float* d_coef, const uint32_t* d_side, float* d_overlap,
                                 float* d_out, int n_units, cudaStream_t stream);
}
using namespace mpeghb200;

struct mpeghb200_fd_session {
  mpeghb200_ctx* ctx = nullptr;
  int n_streams = 0, n_ch = 0;
  size_t n_units = 0;
  static constexpr int kDepth = 2;   // frames in flight (mpeghb200_fd_submit / _wait)
  static constexpr int kChunks = 8;  // chunks per frame at most
  cudaStream_t up = nullptr;         // uploads
  cudaStream_t comp = nullptr;       // kernels, in frame order (the overlap state is carried on this stream)
  cudaStream_t down = nullptr;       // downloads
  float* d_overlap = nullptr;
  struct Slot {
    float* d_coef = nullptr;
    float* d_out = nullptr;
    uint32_t* d_side = nullptr;
    float* p_coef = nullptr;  // pinned staging, allocated on first use with a buffer that cannot be pinned in place
    float* p_out = nullptr;
    uint32_t* p_side = nullptr;
    cudaEvent_t uploaded[kChunks] = {};  // chunk's spectra + side info on the device
    cudaEvent_t computed[kChunks] = {};  // chunk's kernel finished
    cudaEvent_t done[kChunks] = {};      // chunk downloaded
    size_t used = 0, per = 0;
    float* h_out = nullptr;  // caller's output buffer when it has to be filled from p_out in wait()
    bool in_flight = false;
  } slot[kDepth];
  uint64_t n_submitted = 0, n_waited = 0;
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

mpeghb200_status mpeghb200_fd_synth_frames_dev(mpeghb200_ctx* ctx, const float* d_coef, const uint32_t* d_side,
                                               float* d_overlap, float* d_out, int n_units, int n_frames,
                                               void* cuda_stream) {
  if (!ctx) return MPEGHB200_ERR_BAD_ARG;
  if (n_units < 0 || n_frames < 0 || (n_units > 0 && n_frames > 0 && (!d_coef || !d_side || !d_overlap || !d_out)))
    return fail(ctx, MPEGHB200_ERR_BAD_ARG, "fd_synth_frames_dev: null buffer or negative count");
  if (n_frames > 0 && n_units > 0x7fffffff / n_frames)
    return fail(ctx, MPEGHB200_ERR_BAD_ARG, "fd_synth_frames_dev: n_units * n_frames does not fit 31 bits");
  // the coefficient rows travel by bulk copy, overlap and output as float4: every buffer on a 16-byte boundary
  if ((reinterpret_cast<uintptr_t>(d_coef) | reinterpret_cast<uintptr_t>(d_overlap) | reinterpret_cast<uintptr_t>(d_out)) & 15)
    return fail(ctx, MPEGHB200_ERR_BAD_ARG, "fd_synth_frames_dev: coef, overlap and out must be 16-byte aligned");
  MB_CUDA(ctx, cudaSetDevice(ctx->device));
  return fd_synth_frames_launch(ctx, d_coef, d_side, d_overlap, d_out, n_units, n_frames,
                                static_cast<cudaStream_t>(cuda_stream));
}

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
  cudaError_t e = cudaStreamCreateWithFlags(&s->up, cudaStreamNonBlocking);
  if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&s->comp, cudaStreamNonBlocking);
  if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&s->down, cudaStreamNonBlocking);
  for (auto& sl : s->slot) {
    for (int i = 0; i < mpeghb200_fd_session::kChunks && e == cudaSuccess; ++i) {
      e = cudaEventCreateWithFlags(&sl.done[i], cudaEventDisableTiming);
      if (e == cudaSuccess) e = cudaEventCreateWithFlags(&sl.computed[i], cudaEventDisableTiming);
      if (e == cudaSuccess) e = cudaEventCreateWithFlags(&sl.uploaded[i], cudaEventDisableTiming);
    }
    if (e == cudaSuccess) e = cudaMalloc(&sl.d_coef, fb);
    if (e == cudaSuccess) e = cudaMalloc(&sl.d_out, fb);
    if (e == cudaSuccess) e = cudaMalloc(&sl.d_side, s->n_units * sizeof(uint32_t));
  }
  if (e == cudaSuccess) e = cudaMalloc(&s->d_overlap, fb);
  if (e == cudaSuccess) e = cudaMemset(s->d_overlap, 0, fb);
  if (e != cudaSuccess) {
    mpeghb200_status st = cuda_fail(ctx, e, "fd_open allocation");
    mpeghb200_fd_close(s);
    return st;
  }
  *out = s;
  return MPEGHB200_OK;
}

// Enqueue one frame: chunk after chunk the upload stream copies spectra + side info in, the compute stream runs
// the chunk's kernel as soon as it has arrived, the download stream copies it out as soon as the kernel has finished.  PCIe is full duplex and the two copy
// directions use different DMA engines; with two frames in flight the download of frame f overlaps the upload of
// frame f + 1, so both directions stay busy.  Frames of one session execute in submission order (the overlap
// state lives on the compute stream); each frame slot has its own device staging, and the upload stream waits for
// the slot's previous download before the kernel overwrites its output buffer.
mpeghb200_status mpeghb200_fd_submit(mpeghb200_fd_session* s, const float* h_coef, const uint32_t* h_side,
                                     float* h_out) {
  if (!s || !h_coef || !h_side || !h_out) return MPEGHB200_ERR_BAD_ARG;
  mpeghb200_ctx* ctx = s->ctx;
  if (s->n_submitted - s->n_waited >= mpeghb200_fd_session::kDepth)
    return fail(ctx, MPEGHB200_ERR_BAD_ARG, "fd_submit: two frames are in flight already, call mpeghb200_fd_wait first");
  MB_CUDA(ctx, cudaSetDevice(ctx->device));
  auto& sl = s->slot[s->n_submitted % mpeghb200_fd_session::kDepth];
  const size_t fb = s->n_units * 1024 * sizeof(float);
  const size_t sb = s->n_units * sizeof(uint32_t);
  // Source/destination of the DMA: the caller's buffers when they are page-locked, otherwise pinned staging
  // copies filled / drained chunk by chunk so that the host memcpy of one chunk overlaps the DMA of another.
  const bool pin_coef = dma_ready(h_coef), pin_side = dma_ready(h_side), pin_out = dma_ready(h_out);
  if (!pin_coef && !sl.p_coef) MB_CUDA(ctx, cudaHostAlloc(&sl.p_coef, fb, cudaHostAllocDefault));
  if (!pin_side && !sl.p_side) MB_CUDA(ctx, cudaHostAlloc(&sl.p_side, sb, cudaHostAllocDefault));
  if (!pin_out && !sl.p_out) MB_CUDA(ctx, cudaHostAlloc(&sl.p_out, fb, cudaHostAllocDefault));
  const float* src_coef = pin_coef ? h_coef : sl.p_coef;
  const uint32_t* src_side = pin_side ? h_side : sl.p_side;
  float* dst_out = pin_out ? h_out : sl.p_out;
  sl.h_out = pin_out ? nullptr : h_out;
  if (!pin_side) std::memcpy(sl.p_side, h_side, sb);
  const size_t n = s->n_units;
  // A lone frame is cut into chunks so that its own upload, kernel and download overlap; when another frame is
  // already in flight the overlap comes from that frame, and fewer, larger copies cost less per-copy overhead.
  const bool overlapped = s->n_submitted > s->n_waited;
  const size_t n_chunks = overlapped ? (n >= 2048 ? 2 : 1) : (n >= 2048 ? 4 : (n >= 512 ? 2 : 1));
  const size_t per = ((n + n_chunks - 1) / n_chunks + 5) / 6 * 6;  // whole warps for every launch shape in use
  // the slot's previous frame has been waited for by the caller (depth check above), so its downloads are complete
  // and its device buffers are free
  // A failure in the middle of the frame must not leave copies in flight that the caller cannot wait for (earlier
  // chunks are enqueued already, including downloads into the caller's h_out): drain the three streams, then report.
  auto enqueue = [&]() -> mpeghb200_status {
    for (size_t a = 0, i = 0; a < n; a += per, ++i) {
      const size_t cnt = (a + per <= n) ? per : n - a;
      if (!pin_coef) std::memcpy(sl.p_coef + a * 1024, h_coef + a * 1024, cnt * 4096);
      MB_CUDA(ctx, cudaMemcpyAsync(sl.d_coef + a * 1024, src_coef + a * 1024, cnt * 4096, cudaMemcpyHostToDevice, s->up));
      MB_CUDA(ctx, cudaMemcpyAsync(sl.d_side + a, src_side + a, cnt * sizeof(uint32_t), cudaMemcpyHostToDevice, s->up));
      MB_CUDA(ctx, cudaEventRecord(sl.uploaded[i], s->up));
      MB_CUDA(ctx, cudaStreamWaitEvent(s->comp, sl.uploaded[i], 0));
      mpeghb200_status stt = fd_synth_launch(ctx, sl.d_coef + a * 1024, sl.d_side + a, s->d_overlap + a * 1024,
                                             sl.d_out + a * 1024, int(cnt), s->comp);
      if (stt != MPEGHB200_OK) return stt;
      MB_CUDA(ctx, cudaEventRecord(sl.computed[i], s->comp));
      MB_CUDA(ctx, cudaStreamWaitEvent(s->down, sl.computed[i], 0));
      MB_CUDA(ctx, cudaMemcpyAsync(dst_out + a * 1024, sl.d_out + a * 1024, cnt * 4096, cudaMemcpyDeviceToHost, s->down));
      MB_CUDA(ctx, cudaEventRecord(sl.done[i], s->down));
      sl.used = i + 1;
    }
    return MPEGHB200_OK;
  };
  sl.used = 0, sl.per = per;
  const mpeghb200_status st = enqueue();
  if (st != MPEGHB200_OK) {
    cudaStreamSynchronize(s->up), cudaStreamSynchronize(s->comp), cudaStreamSynchronize(s->down);
    cudaGetLastError();
    sl.used = 0, sl.in_flight = false;
    return st;  // the frame was not submitted: n_submitted is unchanged, nothing of it is in flight any more
  }
  sl.in_flight = true;
  s->n_submitted++;
  return MPEGHB200_OK;
}

// Wait for the oldest frame in flight; its h_out is complete on return.
mpeghb200_status mpeghb200_fd_wait(mpeghb200_fd_session* s) {
  if (!s) return MPEGHB200_ERR_BAD_ARG;
  mpeghb200_ctx* ctx = s->ctx;
  if (s->n_waited == s->n_submitted) return MPEGHB200_OK;  // nothing in flight
  MB_CUDA(ctx, cudaSetDevice(ctx->device));
  auto& sl = s->slot[s->n_waited % mpeghb200_fd_session::kDepth];
  const size_t n = s->n_units;
  for (size_t a = 0, i = 0; i < sl.used; a += sl.per, ++i) {
    MB_CUDA(ctx, cudaEventSynchronize(sl.done[i]));
    const size_t cnt = (a + sl.per <= n) ? sl.per : n - a;
    if (sl.h_out) std::memcpy(sl.h_out + a * 1024, sl.p_out + a * 1024, cnt * 4096);
  }
  sl.in_flight = false;
  s->n_waited++;
  return MPEGHB200_OK;
}

mpeghb200_status mpeghb200_fd_execute(mpeghb200_fd_session* s, const float* h_coef, const uint32_t* h_side,
                                      float* h_out) {
  if (!s) return MPEGHB200_ERR_BAD_ARG;
  while (s->n_waited < s->n_submitted) {  // frames submitted earlier complete first, in order
    const mpeghb200_status st = mpeghb200_fd_wait(s);
    if (st != MPEGHB200_OK) return st;
  }
  const mpeghb200_status st = mpeghb200_fd_submit(s, h_coef, h_side, h_out);
  return st != MPEGHB200_OK ? st : mpeghb200_fd_wait(s);
}

mpeghb200_status mpeghb200_fd_reset_stream(mpeghb200_fd_session* s, int stream_idx) {
  if (!s || stream_idx < 0 || stream_idx >= s->n_streams) return MPEGHB200_ERR_BAD_ARG;
  mpeghb200_ctx* ctx = s->ctx;
  MB_CUDA(ctx, cudaSetDevice(ctx->device));
  const size_t per_stream = size_t(s->n_ch) * 1024;
  MB_CUDA(ctx, cudaMemsetAsync(s->d_overlap + per_stream * stream_idx, 0, per_stream * sizeof(float), s->comp));
  MB_CUDA(ctx, cudaStreamSynchronize(s->comp));
  return MPEGHB200_OK;
}

void mpeghb200_fd_close(mpeghb200_fd_session* s) {
  if (!s) return;
  cudaSetDevice(s->ctx->device);
  if (s->up) cudaStreamSynchronize(s->up);
  if (s->comp) cudaStreamSynchronize(s->comp);
  if (s->down) cudaStreamSynchronize(s->down);
  if (s->up) cudaStreamDestroy(s->up);
  if (s->comp) cudaStreamDestroy(s->comp);
  if (s->down) cudaStreamDestroy(s->down);
  for (auto& sl : s->slot) {
    for (auto ev : sl.done)
      if (ev) cudaEventDestroy(ev);
    for (auto ev : sl.computed)
      if (ev) cudaEventDestroy(ev);
    for (auto ev : sl.uploaded)
      if (ev) cudaEventDestroy(ev);
    cudaFree(sl.d_coef);
    cudaFree(sl.d_out);
    cudaFree(sl.d_side);
    cudaFreeHost(sl.p_coef);
    cudaFreeHost(sl.p_out);
    cudaFreeHost(sl.p_side);
  }
  cudaFree(s->d_overlap);
  delete s;
}

}  // extern "C"
