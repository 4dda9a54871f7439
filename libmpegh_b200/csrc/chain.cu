// chain.cu -- mpeghb200_chain_*: one BASELINE.json configuration of the post-parse path end to end behind one
// host-buffer call (include/mpeghb200.h, "chains").  The chain owns the per-stream state of every stage, the
// device-resident intermediates, two slots of input / output staging and three streams (upload, compute, download);
// the stages themselves are the library's *_dev entry points, enqueued back to back on the compute stream in the
// order of decoder/ia_core_coder_decode_main.c:2736-2893 (each is a programmatic-dependent launch, common.h).
#include <cstdlib>
#include <cstring>
#include <new>
#include <vector>

#include "common.h"

using namespace mpeghb200;

namespace mpeghb200 {
mpeghb200_status fd_synth_launch(mpeghb200_ctx* ctx, const float* d_coef, const uint32_t* d_side, float* d_overlap,
                                 float* d_out, int n_units, cudaStream_t stream);
}

struct mpeghb200_chain {
  mpeghb200_ctx* ctx = nullptr;
  mpeghb200_chain_desc d{};
  int S = 0;
  int n_core = 0;      // core channels per stream
  int n_mix = 0;       // channels of the mix bus (before the format converter)
  int n_fc_out = 0;    // channels after the format converter (= n_mix without one)
  int n_out = 0;       // channels of the PCM record
  int fpr = 1;         // frames per PCM record (binaural: block_len / 1024)
  int max_frames = 1;  // frames per call the slots are sized for
  int block_len = 1024;
  size_t core_floats = 0, rec_bytes = 0;
  bool has_map = false;
  int32_t ch_map[64] = {};
  mpeghb200_limiter_params lim{};

  cudaStream_t up = nullptr, comp = nullptr, down = nullptr;
  // state
  float* d_overlap = nullptr;
  float* d_obj_state = nullptr;
  float* d_obj_scratch = nullptr;
  void* d_hoa_state = nullptr;
  float* d_nfc_state = nullptr;
  float* d_fc_state = nullptr;
  float* d_bin_state = nullptr;
  float* d_lim_state = nullptr;
  float* d_lim_scratch = nullptr;
  float* d_gain = nullptr;
  int32_t* d_delay = nullptr;
  int32_t* d_bin_set = nullptr;
  // intermediates (one set: the compute stream runs the frames in order)
  float* d_td = nullptr;       // FD synthesis output, core layout
  float* d_objout = nullptr;   // [S][n_mix][1024]
  float* d_hoacoef = nullptr;  // [S][n_coef][1024]
  float* d_hoaout = nullptr;   // [S][n_mix][1024]
  float* d_mixbuf = nullptr;   // [S][n_mix][1024]
  float* d_fcout = nullptr;    // [S][n_fc_out][1024]
  float* d_ring = nullptr;     // [fpr][S][n_in][1024] binaural input block
  float* d_binout = nullptr;   // [S][2][block_len]
  float* d_limout = nullptr;   // [S][n_out][1024]
  int ring_pos = 0;            // frames of the current binaural block already stored
  // The tail as mpeghb200_limiter_dev + mpeghb200_pcm_pack_dev (default) or as limiter gains kernel + apply-and-pack
  // kernel (MPEGHB200_CHAIN_FUSED_TAIL=1).  Measured equal on B200 (config 3: 0.400 vs 0.414 ms per step, config 5:
  // 0.161 vs 0.162): ncu of the gains kernel (profiles/r2_limiter_gains_ncu.json) puts 46 % of its 63 us on the global
  // loads of the channel maxima and 32 % on the barrier behind the serial smoother -- the output phase the split moves
  // out was never the cost.  (A single fused kernel, one CTA per stream, was slower than both.)
  bool fused_tail = false;
  std::vector<void*> owned;

  static constexpr int kDepth = 2;
  struct Slot {
    float* d_core = nullptr;
    uint32_t* d_fd_side = nullptr;
    mpeghb200_obj_side* d_obj_side = nullptr;
    mpeghb200_hoa_spatial_side* d_hoa_side = nullptr;
    uint8_t* d_pcm = nullptr;
    int32_t* d_pcm_bytes = nullptr;
    cudaEvent_t uploaded = nullptr, computed = nullptr, done = nullptr;
  } slot[kDepth];
  uint64_t n_submitted = 0, n_waited = 0;
};

namespace {

template <typename T>
cudaError_t dalloc(mpeghb200_chain* c, T** p, size_t count, bool zero) {
  *p = nullptr;
  if (count == 0) return cudaSuccess;
  void* v = nullptr;
  cudaError_t e = cudaMalloc(&v, count * sizeof(T));
  if (e != cudaSuccess) return e;
  c->owned.push_back(v);
  *p = static_cast<T*>(v);
  return zero ? cudaMemset(v, 0, count * sizeof(T)) : cudaSuccess;
}

__global__ void fill_kernel(float* p, float v, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) p[i] = v;
}
__global__ void fill_i32_kernel(int32_t* p, int32_t v, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) p[i] = v;
}

// One frame on device buffers.  `core` is consumed (the bed region may be summed into in place).
// block_in: with the binaural renderer, non-NULL when this frame's mix is the f-th [S][n_in][1024] slice of a
// caller-provided contiguous block [fpr][S][n_in][1024] starting at block_in (time-domain bed only): nothing is copied.
mpeghb200_status run_frame(mpeghb200_chain* c, float* core, const uint32_t* fd_side, const mpeghb200_obj_side* obj_side,
                           const mpeghb200_hoa_spatial_side* hoa_side, uint8_t* pcm, int32_t* pcm_bytes,
                           const float* block_in, cudaStream_t st) {
  mpeghb200_ctx* ctx = c->ctx;
  const mpeghb200_chain_desc& d = c->d;
  const int S = c->S;
  mpeghb200_status s;
  const bool bed_only = d.n_bed_ch > 0 && d.n_obj == 0 && d.n_hoa_transport == 0;
  // the binaural input block is filled in place when the mix IS the FD output / the bed region and nothing else
  // writes it: no copy into the ring
  float* ring_slot = d.binaural ? c->d_ring + size_t(c->ring_pos) * S * c->n_fc_out * 1024 : nullptr;
  const bool direct_ring = d.binaural && bed_only && !d.format_conv && d.core_is_spectra;
  float* td = core;
  if (d.core_is_spectra) {
    td = direct_ring ? ring_slot : c->d_td;
    if ((s = fd_synth_launch(ctx, core, fd_side, c->d_overlap, td, S * c->n_core, st)) != MPEGHB200_OK) return s;
  }
  float* bed = td;
  float* obj = bed + size_t(S) * d.n_bed_ch * 1024;
  float* hoa = obj + size_t(S) * d.n_obj * 1024;
  const size_t mix_floats = size_t(S) * c->n_mix * 1024;
  float* mix = d.n_bed_ch > 0 ? bed : nullptr;
  if (d.n_obj > 0) {
    if ((s = mpeghb200_obj_render_dev(ctx, d.obj_layout, d.n_obj, obj_side, obj, c->d_objout, c->d_obj_state,
                                      c->d_obj_scratch, S, st)) != MPEGHB200_OK)
      return s;
    if (mix) {  // time_sample_vector[ch] += ptr_out_buf[ch] (decode_main.c:2185-2195)
      if ((s = mpeghb200_bus_add_dev(ctx, mix, c->d_objout, c->d_mixbuf, mix_floats, st)) != MPEGHB200_OK) return s;
      mix = c->d_mixbuf;
    } else {
      mix = c->d_objout;
    }
  }
  if (d.n_hoa_transport > 0) {
    if ((s = mpeghb200_hoa_spatial_dev(ctx, d.hoa_spatial, hoa_side, hoa, c->d_hoacoef, c->d_hoa_state, S, st)) !=
        MPEGHB200_OK)
      return s;
    if ((s = mpeghb200_hoa_render_dev(ctx, d.hoa_renderer, c->d_hoacoef, c->d_hoaout, c->d_nfc_state, S, st)) !=
        MPEGHB200_OK)
      return s;
    if (mix) {  // objects + HOA: tsv = ia_add_flt(hoa, obj) (:2203-2213, :2290-2300)
      if ((s = mpeghb200_bus_add_dev(ctx, c->d_hoaout, mix, c->d_mixbuf, mix_floats, st)) != MPEGHB200_OK) return s;
      mix = c->d_mixbuf;
    } else {
      mix = c->d_hoaout;
    }
  }
  if (d.format_conv) {
    if ((s = mpeghb200_format_conv_dev(ctx, d.format_conv, mix, c->d_fcout, c->d_fc_state, S, st)) != MPEGHB200_OK)
      return s;
    mix = c->d_fcout;
  }
  if (d.binaural) {
    if (mix != ring_slot && !block_in)
      MB_CUDA(ctx, cudaMemcpyAsync(ring_slot, mix, size_t(S) * c->n_fc_out * 4096, cudaMemcpyDeviceToDevice, st));
    if (++c->ring_pos < c->fpr) return MPEGHB200_OK;  // block not complete: no PCM for this frame
    c->ring_pos = 0;
    if (!pcm) return fail(ctx, MPEGHB200_ERR_BAD_ARG, "chain: this frame completes a binaural block, pcm must not be NULL");
    // ia_div_q15_flt on the way in, * 32768 on the way out (decode_main.c:2833, :2851)
    if ((s = mpeghb200_binaural_block_dev(ctx, d.binaural, block_in ? block_in : c->d_ring, 1, 1.0f / 32768.0f, 32768.0f, c->d_binout,
                                          c->d_bin_state, c->d_bin_set, S, st)) != MPEGHB200_OK)
      return s;
    return mpeghb200_pcm_pack_block_dev(ctx, c->d_binout, 2, c->block_len, c->block_len, d.pcm_bits,
                                        c->has_map ? c->ch_map : nullptr, c->d_delay, pcm, pcm_bytes, S, st);
  }
  if (d.limiter && !c->fused_tail) {
    if ((s = mpeghb200_limiter_dev(ctx, &c->lim, mix, c->d_limout, c->d_lim_state, c->d_gain, S, st)) != MPEGHB200_OK)
      return s;
    return mpeghb200_pcm_pack_dev(ctx, c->d_limout, c->n_out, d.pcm_bits, c->has_map ? c->ch_map : nullptr, c->d_delay,
                                  pcm, pcm_bytes, S, st);
  }
  if (d.limiter)  // gains kernel + apply-and-pack kernel: no planar fp32 round trip between limiter and packer
    return mpeghb200_limiter_pack_dev(ctx, &c->lim, mix, c->d_lim_state, c->d_lim_scratch, c->d_gain, d.pcm_bits,
                                      c->has_map ? c->ch_map : nullptr, c->d_delay, pcm, pcm_bytes, S, st);
  return mpeghb200_pcm_pack_dev(ctx, mix, c->n_out, d.pcm_bits, c->has_map ? c->ch_map : nullptr, c->d_delay, pcm,
                                pcm_bytes, S, st);
}

// n_frames consecutive frames laid out as mpeghb200_chain_submit describes, on device buffers.
mpeghb200_status run_frames(mpeghb200_chain* c, float* core, const uint32_t* fd_side, const mpeghb200_obj_side* obj_side,
                            const mpeghb200_hoa_spatial_side* hoa_side, uint8_t* pcm, int32_t* pcm_bytes, int n_frames,
                            cudaStream_t st) {
  const mpeghb200_chain_desc& d = c->d;
  const int S = c->S;
  if (d.binaural && c->ring_pos != 0)
    return fail(c->ctx, MPEGHB200_ERR_BAD_ARG, "chain: a call must start on a binaural block boundary");
  const bool in_place_block = d.binaural && !d.core_is_spectra && !d.format_conv && d.n_obj == 0 && d.n_hoa_transport == 0;
  for (int f = 0; f < n_frames; ++f) {
    const size_t rec = size_t(f / c->fpr);
    const bool completes = (f % c->fpr) == c->fpr - 1;
    const mpeghb200_status s = run_frame(
        c, core + size_t(f) * c->core_floats, d.core_is_spectra ? fd_side + size_t(f) * S * c->n_core : nullptr,
        d.n_obj > 0 ? obj_side + size_t(f) * S * d.n_obj : nullptr, d.n_hoa_transport > 0 ? hoa_side + size_t(f) * S : nullptr,
        completes ? pcm + rec * S * c->rec_bytes : nullptr, completes && pcm_bytes ? pcm_bytes + rec * S : nullptr,
        in_place_block ? core + rec * c->fpr * c->core_floats : nullptr, st);
    if (s != MPEGHB200_OK) return s;
  }
  return MPEGHB200_OK;
}

}  // namespace

extern "C" {

mpeghb200_status mpeghb200_chain_open(mpeghb200_ctx* ctx, const mpeghb200_chain_desc* desc, mpeghb200_chain** out) {
  if (!ctx || !desc || !out) return MPEGHB200_ERR_BAD_ARG;
  *out = nullptr;
  const mpeghb200_chain_desc& d = *desc;
  if (d.n_streams <= 0 || d.n_bed_ch < 0 || d.n_obj < 0 || d.n_hoa_transport < 0 ||
      d.n_bed_ch + d.n_obj + d.n_hoa_transport <= 0 || (d.pcm_bits != 16 && d.pcm_bits != 24 && d.pcm_bits != 32) ||
      d.limiter < 0 || d.limiter > 2 || d.start_delay < 0)
    return fail(ctx, MPEGHB200_ERR_BAD_ARG, "chain_open: bad shape");
  if ((d.n_obj > 0 && !d.obj_layout) || (d.n_hoa_transport > 0 && (!d.hoa_spatial || !d.hoa_renderer)))
    return fail(ctx, MPEGHB200_ERR_BAD_ARG, "chain_open: a signal group without its renderer handle");
  if (d.n_bed_ch > 0 && d.n_hoa_transport > 0)
    return fail(ctx, MPEGHB200_ERR_UNSUPPORTED, "chain_open: bed + HOA scenes are not chained (bed + objects, objects + HOA are)");
  if (d.n_obj > MPEGHB200_OBJ_MAX_OBJECTS) return fail(ctx, MPEGHB200_ERR_BAD_ARG, "chain_open: more than 32 objects");
  // the mix bus: every group renders into the same loudspeaker layout
  int n_mix = d.n_bed_ch;
  if (d.n_obj > 0) {
    const int n = mpeghb200_obj_layout_num_speakers(d.obj_layout);
    if (n_mix && n != n_mix) return fail(ctx, MPEGHB200_ERR_BAD_ARG, "chain_open: object layout and bed differ in channels");
    n_mix = n;
  }
  if (d.n_hoa_transport > 0) {
    if (mpeghb200_hoa_spatial_num_transport(d.hoa_spatial) != d.n_hoa_transport ||
        mpeghb200_hoa_spatial_num_coeffs(d.hoa_spatial) != mpeghb200_hoa_renderer_num_coeffs(d.hoa_renderer))
      return fail(ctx, MPEGHB200_ERR_BAD_ARG, "chain_open: HOA decoder / renderer shapes do not match");
    const int n = mpeghb200_hoa_renderer_num_out(d.hoa_renderer);
    if (n_mix && n != n_mix) return fail(ctx, MPEGHB200_ERR_BAD_ARG, "chain_open: HOA renderer and mix differ in channels");
    n_mix = n;
  }
  int n_fc_out = n_mix;
  if (d.format_conv) {
    if (mpeghb200_format_conv_num_in(d.format_conv) != n_mix)
      return fail(ctx, MPEGHB200_ERR_BAD_ARG, "chain_open: format converter input count differs from the mix");
    n_fc_out = mpeghb200_format_conv_num_out(d.format_conv);
  }
  if (d.binaural && mpeghb200_binaural_num_in(d.binaural) != n_fc_out)
    return fail(ctx, MPEGHB200_ERR_BAD_ARG, "chain_open: binaural renderer input count differs from the mix");
  if (n_mix > 64 || n_fc_out > 64) return fail(ctx, MPEGHB200_ERR_BAD_ARG, "chain_open: more than 64 channels");
  MB_CUDA(ctx, cudaSetDevice(ctx->device));
  auto* c = new (std::nothrow) mpeghb200_chain();
  if (!c) return MPEGHB200_ERR_NO_MEM;
  c->ctx = ctx;
  c->d = d;
  c->S = d.n_streams;
  c->n_core = d.n_bed_ch + d.n_obj + d.n_hoa_transport;
  c->n_mix = n_mix;
  c->n_fc_out = n_fc_out;
  c->n_out = d.binaural ? 2 : n_fc_out;
  c->block_len = d.binaural ? mpeghb200_binaural_block_len(d.binaural) : 1024;
  c->fpr = c->block_len / 1024;
  c->max_frames = d.max_frames_per_call > 0 ? d.max_frames_per_call : c->fpr;
  if (c->max_frames > MPEGHB200_CHAIN_MAX_FRAMES || (c->fpr >= 1 && c->max_frames % c->fpr)) {
    delete c;
    return fail(ctx, MPEGHB200_ERR_BAD_ARG, "chain_open: max_frames_per_call out of range or not whole binaural blocks");
  }
  const size_t F = size_t(c->max_frames);
  c->core_floats = size_t(c->S) * c->n_core * 1024;
  c->rec_bytes = size_t(c->block_len) * c->n_out * (d.pcm_bits / 8);
  if (d.out_ch_map) {
    c->has_map = true;
    std::memcpy(c->ch_map, d.out_ch_map, sizeof(int32_t) * c->n_out);
    c->d.out_ch_map = nullptr;
  }
  const int S = c->S;
  mpeghb200_status st = MPEGHB200_OK;
  cudaError_t e = cudaSuccess;
  auto ok = [&]() { return e == cudaSuccess && st == MPEGHB200_OK; };
  if (d.binaural && (c->block_len < 1024 || c->block_len % 1024 || c->fpr > MPEGHB200_CHAIN_MAX_FRAMES))
    st = fail(ctx, MPEGHB200_ERR_UNSUPPORTED, "chain_open: binaural block length must be 1024..8192 in whole frames");
  if (ok()) e = cudaStreamCreateWithFlags(&c->up, cudaStreamNonBlocking);
  if (ok()) e = cudaStreamCreateWithFlags(&c->comp, cudaStreamNonBlocking);
  if (ok()) e = cudaStreamCreateWithFlags(&c->down, cudaStreamNonBlocking);
  // ---- per-stream state ----
  if (ok() && d.core_is_spectra) e = dalloc(c, &c->d_overlap, c->core_floats, true);
  if (ok() && d.core_is_spectra) e = dalloc(c, &c->d_td, c->core_floats, false);
  if (ok() && d.n_obj > 0) {
    e = dalloc(c, &c->d_obj_state, S * mpeghb200_obj_state_floats(d.obj_layout, d.n_obj), false);
    if (ok()) e = dalloc(c, &c->d_obj_scratch, S * mpeghb200_obj_scratch_floats(d.obj_layout, d.n_obj), false);
    if (ok()) e = dalloc(c, &c->d_objout, size_t(S) * n_mix * 1024, false);
    // first_frame_flag[] as the reference handle starts out: set for its first MAX_NUM_OAM_OBJS = 24 objects only
    // (ia_obj_ren_dec_state_struct is zero-allocated and impeghd_obj_renderer_dec_init marks 24 entries)
    int32_t* d_first = nullptr;
    if (ok()) e = dalloc(c, &d_first, size_t(d.n_obj), false);
    if (ok()) {
      int32_t first[MPEGHB200_OBJ_MAX_OBJECTS];
      for (int o = 0; o < d.n_obj; ++o) first[o] = o < 24 ? 1 : 0;
      e = cudaMemcpy(d_first, first, sizeof(int32_t) * d.n_obj, cudaMemcpyHostToDevice);
    }
    if (ok()) st = mpeghb200_obj_state_init_dev(ctx, d.obj_layout, d.n_obj, d_first, c->d_obj_state, S, c->comp);
  }
  if (ok() && d.n_hoa_transport > 0) {
    uint8_t* p = nullptr;
    e = dalloc(c, &p, S * mpeghb200_hoa_spatial_state_bytes(d.hoa_spatial), false);
    c->d_hoa_state = p;
    if (ok()) st = mpeghb200_hoa_spatial_state_init_dev(ctx, d.hoa_spatial, c->d_hoa_state, S, c->comp);
    if (ok()) e = dalloc(c, &c->d_nfc_state, S * mpeghb200_hoa_nfc_state_floats(d.hoa_renderer), true);
    if (ok()) e = dalloc(c, &c->d_hoacoef, size_t(S) * mpeghb200_hoa_spatial_num_coeffs(d.hoa_spatial) * 1024, false);
    if (ok()) e = dalloc(c, &c->d_hoaout, size_t(S) * n_mix * 1024, false);
  }
  if (ok() && (d.n_bed_ch > 0) + (d.n_obj > 0) + (d.n_hoa_transport > 0) > 1)
    e = dalloc(c, &c->d_mixbuf, size_t(S) * n_mix * 1024, false);
  if (ok() && d.format_conv) {
    e = dalloc(c, &c->d_fc_state, S * mpeghb200_format_conv_state_floats(d.format_conv), true);
    if (ok()) e = dalloc(c, &c->d_fcout, size_t(S) * n_fc_out * 1024, false);
  }
  if (ok() && d.binaural) {
    e = dalloc(c, &c->d_bin_state, S * mpeghb200_binaural_state_floats(d.binaural), true);
    if (ok()) e = dalloc(c, &c->d_ring, size_t(c->fpr) * S * n_fc_out * 1024, false);
    if (ok()) e = dalloc(c, &c->d_binout, size_t(S) * 2 * c->block_len, false);
    if (ok() && d.binaural_set_of_stream) {
      e = dalloc(c, &c->d_bin_set, size_t(S), false);
      if (ok()) e = cudaMemcpy(c->d_bin_set, d.binaural_set_of_stream, size_t(S) * sizeof(int32_t), cudaMemcpyHostToDevice);
      c->d.binaural_set_of_stream = nullptr;
    }
  }
  if (ok() && d.limiter && !d.binaural) {
    st = mpeghb200_limiter_params_init(&c->lim, c->n_out, 48000, d.limiter == 1);
    if (st != MPEGHB200_OK) st = fail(ctx, st, "chain_open: limiter parameters");
    if (ok()) e = dalloc(c, &c->d_lim_state, S * mpeghb200_limiter_state_floats(&c->lim), false);
    if (ok()) st = mpeghb200_limiter_state_init_dev(ctx, &c->lim, c->d_lim_state, S, c->comp);
    if (const char* ev = getenv("MPEGHB200_CHAIN_FUSED_TAIL")) c->fused_tail = atoi(ev) != 0;
    if (ok() && !c->fused_tail) e = dalloc(c, &c->d_limout, size_t(S) * c->n_out * 1024, false);
    if (ok() && c->fused_tail) e = dalloc(c, &c->d_lim_scratch, S * mpeghb200_limiter_pack_scratch_floats(&c->lim), false);
    if (ok() && d.loudness_gain != 1.0f) {
      e = dalloc(c, &c->d_gain, size_t(S), false);
      if (ok()) fill_kernel<<<(S + 255) / 256, 256, 0, c->comp>>>(c->d_gain, d.loudness_gain, S);
    }
  }
  if (ok()) e = dalloc(c, &c->d_delay, size_t(S), false);
  if (ok()) fill_i32_kernel<<<(S + 255) / 256, 256, 0, c->comp>>>(c->d_delay, d.start_delay, S);
  // ---- frame slots ----
  for (auto& sl : c->slot) {
    if (ok()) e = dalloc(c, &sl.d_core, c->core_floats * F, false);
    if (ok() && d.core_is_spectra) e = dalloc(c, &sl.d_fd_side, size_t(S) * c->n_core * F, false);
    if (ok() && d.n_obj > 0) e = dalloc(c, &sl.d_obj_side, size_t(S) * d.n_obj * F, false);
    if (ok() && d.n_hoa_transport > 0) e = dalloc(c, &sl.d_hoa_side, size_t(S) * F, false);
    if (ok()) e = dalloc(c, &sl.d_pcm, size_t(S) * c->rec_bytes * (F / c->fpr), false);
    if (ok()) e = dalloc(c, &sl.d_pcm_bytes, size_t(S) * F, true);
    if (ok()) e = cudaEventCreateWithFlags(&sl.uploaded, cudaEventDisableTiming);
    if (ok()) e = cudaEventCreateWithFlags(&sl.computed, cudaEventDisableTiming);
    if (ok()) e = cudaEventCreateWithFlags(&sl.done, cudaEventDisableTiming);
  }
  if (ok()) e = cudaStreamSynchronize(c->comp);
  if (ok()) e = cudaGetLastError();
  if (!ok()) {
    if (e != cudaSuccess) st = cuda_fail(ctx, e, "chain_open");
    mpeghb200_chain_close(c);
    return st;
  }
  *out = c;
  return MPEGHB200_OK;
}

void mpeghb200_chain_close(mpeghb200_chain* c) {
  if (!c) return;
  cudaSetDevice(c->ctx->device);
  for (cudaStream_t s : {c->up, c->comp, c->down})
    if (s) cudaStreamSynchronize(s);
  for (cudaStream_t s : {c->up, c->comp, c->down})
    if (s) cudaStreamDestroy(s);
  for (auto& sl : c->slot)
    for (cudaEvent_t ev : {sl.uploaded, sl.computed, sl.done})
      if (ev) cudaEventDestroy(ev);
  for (void* p : c->owned) cudaFree(p);
  delete c;
}

int mpeghb200_chain_out_channels(const mpeghb200_chain* c) { return c ? c->n_out : 0; }
size_t mpeghb200_chain_core_floats(const mpeghb200_chain* c) { return c ? c->core_floats : 0; }
size_t mpeghb200_chain_pcm_record_bytes(const mpeghb200_chain* c) { return c ? c->rec_bytes : 0; }
int mpeghb200_chain_frames_per_record(const mpeghb200_chain* c) { return c ? c->fpr : 0; }

mpeghb200_status mpeghb200_chain_run_dev(mpeghb200_chain* c, const mpeghb200_chain_frame* f, int n_frames,
                                         void* cuda_stream) {
  if (!c || !f) return MPEGHB200_ERR_BAD_ARG;
  mpeghb200_ctx* ctx = c->ctx;
  const mpeghb200_chain_desc& d = c->d;
  if (n_frames < 1 || n_frames > MPEGHB200_CHAIN_MAX_FRAMES || n_frames % c->fpr)
    return fail(ctx, MPEGHB200_ERR_BAD_ARG, "chain_run_dev: n_frames out of range or not whole binaural blocks");
  if (!f->core || !f->pcm || (d.core_is_spectra && !f->fd_side) || (d.n_obj > 0 && !f->obj_side) ||
      (d.n_hoa_transport > 0 && !f->hoa_side))
    return fail(ctx, MPEGHB200_ERR_BAD_ARG, "chain_run_dev: a buffer the configuration needs is NULL");
  MB_CUDA(ctx, cudaSetDevice(ctx->device));
  return run_frames(c, const_cast<float*>(f->core), f->fd_side, f->obj_side, f->hoa_side, f->pcm, f->pcm_bytes, n_frames,
                    static_cast<cudaStream_t>(cuda_stream));
}

mpeghb200_status mpeghb200_chain_submit(mpeghb200_chain* c, const mpeghb200_chain_frame* h, int n_frames) {
  if (!c || !h) return MPEGHB200_ERR_BAD_ARG;
  mpeghb200_ctx* ctx = c->ctx;
  const mpeghb200_chain_desc& d = c->d;
  if (n_frames < 1 || n_frames > c->max_frames || n_frames % c->fpr)
    return fail(ctx, MPEGHB200_ERR_BAD_ARG, "chain_submit: n_frames beyond max_frames_per_call or not whole binaural blocks");
  if (!h->core || !h->pcm || (d.core_is_spectra && !h->fd_side) || (d.n_obj > 0 && !h->obj_side) ||
      (d.n_hoa_transport > 0 && !h->hoa_side))
    return fail(ctx, MPEGHB200_ERR_BAD_ARG, "chain_submit: a buffer the configuration needs is NULL");
  if (c->n_submitted - c->n_waited >= mpeghb200_chain::kDepth)
    return fail(ctx, MPEGHB200_ERR_BAD_ARG, "chain_submit: two calls are in flight already, call mpeghb200_chain_wait first");
  MB_CUDA(ctx, cudaSetDevice(ctx->device));
  auto& sl = c->slot[c->n_submitted % mpeghb200_chain::kDepth];
  const int S = c->S;
  const size_t F = size_t(n_frames), n_rec = F / c->fpr;
  auto enqueue = [&]() -> mpeghb200_status {
    // the slot's previous call has been waited for (depth check above): its buffers are free
    MB_CUDA(ctx, cudaMemcpyAsync(sl.d_core, h->core, c->core_floats * F * sizeof(float), cudaMemcpyHostToDevice, c->up));
    if (d.core_is_spectra)
      MB_CUDA(ctx, cudaMemcpyAsync(sl.d_fd_side, h->fd_side, size_t(S) * c->n_core * F * sizeof(uint32_t),
                                   cudaMemcpyHostToDevice, c->up));
    if (d.n_obj > 0)
      MB_CUDA(ctx, cudaMemcpyAsync(sl.d_obj_side, h->obj_side, size_t(S) * d.n_obj * F * sizeof(mpeghb200_obj_side),
                                   cudaMemcpyHostToDevice, c->up));
    if (d.n_hoa_transport > 0)
      MB_CUDA(ctx, cudaMemcpyAsync(sl.d_hoa_side, h->hoa_side, size_t(S) * F * sizeof(mpeghb200_hoa_spatial_side),
                                   cudaMemcpyHostToDevice, c->up));
    MB_CUDA(ctx, cudaEventRecord(sl.uploaded, c->up));
    MB_CUDA(ctx, cudaStreamWaitEvent(c->comp, sl.uploaded, 0));
    {
      const mpeghb200_status s = run_frames(c, sl.d_core, sl.d_fd_side, sl.d_obj_side, sl.d_hoa_side, sl.d_pcm,
                                            sl.d_pcm_bytes, n_frames, c->comp);
      if (s != MPEGHB200_OK) return s;
    }
    MB_CUDA(ctx, cudaEventRecord(sl.computed, c->comp));
    MB_CUDA(ctx, cudaStreamWaitEvent(c->down, sl.computed, 0));
    MB_CUDA(ctx, cudaMemcpyAsync(h->pcm, sl.d_pcm, n_rec * S * c->rec_bytes, cudaMemcpyDeviceToHost, c->down));
    if (h->pcm_bytes)
      MB_CUDA(ctx, cudaMemcpyAsync(h->pcm_bytes, sl.d_pcm_bytes, n_rec * S * sizeof(int32_t), cudaMemcpyDeviceToHost,
                                   c->down));
    MB_CUDA(ctx, cudaEventRecord(sl.done, c->down));
    return MPEGHB200_OK;
  };
  const mpeghb200_status st = enqueue();
  if (st != MPEGHB200_OK) {
    // nothing of a failed call may stay in flight towards the caller's buffers
    cudaStreamSynchronize(c->up), cudaStreamSynchronize(c->comp), cudaStreamSynchronize(c->down);
    cudaGetLastError();
    return st;
  }
  c->n_submitted++;
  return MPEGHB200_OK;
}

mpeghb200_status mpeghb200_chain_wait(mpeghb200_chain* c) {
  if (!c) return MPEGHB200_ERR_BAD_ARG;
  if (c->n_waited == c->n_submitted) return MPEGHB200_OK;
  mpeghb200_ctx* ctx = c->ctx;
  MB_CUDA(ctx, cudaSetDevice(ctx->device));
  auto& sl = c->slot[c->n_waited % mpeghb200_chain::kDepth];
  MB_CUDA(ctx, cudaEventSynchronize(sl.done));
  c->n_waited++;
  return MPEGHB200_OK;
}

mpeghb200_status mpeghb200_chain_execute(mpeghb200_chain* c, const mpeghb200_chain_frame* h, int n_frames) {
  if (!c) return MPEGHB200_ERR_BAD_ARG;
  while (c->n_waited < c->n_submitted) {
    const mpeghb200_status st = mpeghb200_chain_wait(c);
    if (st != MPEGHB200_OK) return st;
  }
  const mpeghb200_status st = mpeghb200_chain_submit(c, h, n_frames);
  return st != MPEGHB200_OK ? st : mpeghb200_chain_wait(c);
}

}  // extern "C"
