/* mpeghb200.h -- C ABI of libmpeghb200.so: the B200 (sm_100a) implementation of libmpegh's post-parse
 * signal path, batched over many independent streams.
 *
 * Boundary.  The reference decodes one frame of one stream per ia_mpegh_dec_execute call
 * (decoder/impeghd_api.h:47-51).  Inside that call the bitstream half (MHAS/USAC parse, arithmetic
 * decoding, scale factors, OAM/HOA side info) ends at decoder/ia_core_coder_process.c:1092 with the
 * dequantised spectra usac_data->coef[ch][1024] and the per-channel side info complete; everything after
 * that line is signal processing on planar FLOAT32 buffers.  This library replaces that second half.
 * Each entry point below names the reference function it stands in for; INTEGRATION.md shows the patch a
 * maintainer applies to the reference so that N decoder handles advance one frame per batched call.
 *
 * Conventions
 *   - plain C types only; every buffer is caller-owned; sizes in elements unless a name ends in _bytes
 *   - *_dev entry points take DEVICE pointers and a CUDA stream (cudaStream_t passed as void*, NULL =
 *     default stream) and return after enqueueing; *_host entry points take HOST pointers, perform the
 *     host<->device copies themselves and return when the result is in the host buffer
 *   - status: 0 = OK; errors follow the reference's IA_ERRORCODE style (decoder/impeghd_error_codes.h):
 *     negative (bit 31 set) = fatal for that call, nothing was written
 *   - audio is planar fp32 at 16-bit scale (+-32768), [stream][channel][1024], as in the reference's
 *     time_sample_vector (decoder/ia_core_coder_main.h:124)
 *   - there is no CPU fallback: with no CUDA device every call returns MPEGHB200_ERR_NO_DEVICE
 */
#ifndef MPEGHB200_H
#define MPEGHB200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MPEGHB200_FRAME_LEN 1024

typedef int32_t mpeghb200_status;
#define MPEGHB200_OK 0
#define MPEGHB200_ERR_NO_DEVICE ((mpeghb200_status)0xFFFF8001)
#define MPEGHB200_ERR_CUDA ((mpeghb200_status)0xFFFF8002)
#define MPEGHB200_ERR_BAD_ARG ((mpeghb200_status)0xFFFF8003)
#define MPEGHB200_ERR_NO_MEM ((mpeghb200_status)0xFFFF8004)
#define MPEGHB200_ERR_UNSUPPORTED ((mpeghb200_status)0xFFFF8005)

typedef struct mpeghb200_ctx mpeghb200_ctx;

/* ---- context -------------------------------------------------------------------------------- */
const char *mpeghb200_version(void);
/* One context per GPU (streams are sharded across GPUs by the caller: no cross-GPU traffic). Uploads
 * the constant tables. */
mpeghb200_status mpeghb200_create(mpeghb200_ctx **out, int device_ordinal);
void mpeghb200_destroy(mpeghb200_ctx *ctx);
const char *mpeghb200_last_error(const mpeghb200_ctx *ctx); /* ctx may be NULL: last create() error */
/* Number of kernel launches issued through this context so far (bench.py's gpu_launches). */
uint64_t mpeghb200_launch_count(const mpeghb200_ctx *ctx);

/* Host copy of a generated constant table (ids: libmpegh_b200/csrc/rom_tables.h). patched = 0 returns
 * the pure-formula table (used by tools/gen_rom_patches.py). No GPU needed. */
const float *mpeghb200_rom_table(int id, int patched, int *n_out);
int mpeghb200_rom_count(void);
const char *mpeghb200_rom_name(int id);

/* Device memory helpers for callers without a CUDA runtime of their own (ctypes, cgo, JNI). */
mpeghb200_status mpeghb200_dev_alloc(mpeghb200_ctx *ctx, size_t bytes, void **dptr);
mpeghb200_status mpeghb200_dev_free(mpeghb200_ctx *ctx, void *dptr);
mpeghb200_status mpeghb200_dev_memset(mpeghb200_ctx *ctx, void *dptr, int value, size_t bytes);
mpeghb200_status mpeghb200_h2d(mpeghb200_ctx *ctx, void *dptr, const void *hptr, size_t bytes);
mpeghb200_status mpeghb200_d2h(mpeghb200_ctx *ctx, void *hptr, const void *dptr, size_t bytes);
mpeghb200_status mpeghb200_sync(mpeghb200_ctx *ctx);
/* Page-locked host memory.  The *_execute entry points DMA straight from/to such buffers (and from any other
 * page-locked memory); ordinary malloc'ed buffers work too but are staged through an internal pinned copy.
 * A decoder handle's coef / OUTPUT mem tables should live here. */
mpeghb200_status mpeghb200_host_alloc(mpeghb200_ctx *ctx, size_t bytes, void **hptr);
mpeghb200_status mpeghb200_host_free(mpeghb200_ctx *ctx, void *hptr);

/* ---- FD synthesis: IMDCT + windowing + overlap-add ------------------------------------------
 * Replaces ia_core_coder_fd_frm_dec (decoder/ia_core_coder_imdct.c:832) for FD-coded channels with FD
 * history and no FAC data (LPD-coded channels are synthesised by the host).
 * One unit = one channel of one stream for one frame.
 *
 * side[unit] packs what the reference keeps in ia_usac_data_struct (decoder/ia_core_coder_main.h:127-200):
 *   bits 0-2 window_sequence[ch]   (0 ONLY_LONG, 1 LONG_START, 2 EIGHT_SHORT, 3 LONG_STOP, 4 STOP_START)
 *   bit  3   window_shape[ch]      (0 sine, 1 KBD)
 *   bit  4   window_shape_prev[ch]
 *   bit  5   prev_aliasing_symmetry[ch]
 *   bit  6   curr_aliasing_symmetry[ch]
 */
#define MPEGHB200_FD_SIDE(seq, shape, shape_prev, prev_sym, curr_sym)                              \
  ((uint32_t)(((seq)&7) | (((shape)&1) << 3) | (((shape_prev)&1) << 4) | (((prev_sym)&1) << 5) |   \
              (((curr_sym)&1) << 6)))

/* coef [n_units][1024] (read only), overlap [n_units][1024] (read + replaced; the per-channel state the
 * reference keeps in overlap_data_ptr), out [n_units][1024]. */
mpeghb200_status mpeghb200_fd_synth_dev(mpeghb200_ctx *ctx, const float *d_coef, const uint32_t *d_side,
                                        float *d_overlap, float *d_out, int n_units, void *cuda_stream);

/* n_frames consecutive frames of every unit in ONE launch, for feeds that have them at hand (file decoding, any
 * look-ahead of the parse half): coef / side / out are [n_frames][n_units][...], overlap [n_units][1024] as above
 * (read before frame 0, replaced after the last frame).  The result is what n_frames calls of mpeghb200_fd_synth_dev
 * give, bit for bit; between two ONLY_LONG-family frames a unit's overlap stays on the SM (ia_core_coder_fd_frm_dec
 * writes and re-reads it through memory, decoder/ia_core_coder_imdct.c:807-813), which removes up to half of the
 * launch's HBM traffic.  n_units * n_frames < 2^31; coef, overlap and out on 16-byte boundaries (the coefficient rows
 * travel by TMA bulk copy), else MPEGHB200_ERR_BAD_ARG. */
mpeghb200_status mpeghb200_fd_synth_frames_dev(mpeghb200_ctx *ctx, const float *d_coef, const uint32_t *d_side,
                                               float *d_overlap, float *d_out, int n_units, int n_frames,
                                               void *cuda_stream);

/* Host-buffer session: device-resident overlap state for n_streams x n_ch channels, pinned staging. */
typedef struct mpeghb200_fd_session mpeghb200_fd_session;
mpeghb200_status mpeghb200_fd_open(mpeghb200_ctx *ctx, int n_streams, int n_ch, mpeghb200_fd_session **out);
/* One frame for every stream: h_coef/h_out [n_streams][n_ch][1024], h_side [n_streams][n_ch]. */
mpeghb200_status mpeghb200_fd_execute(mpeghb200_fd_session *s, const float *h_coef, const uint32_t *h_side,
                                      float *h_out);
/* The same, asynchronous: submit enqueues the frame (uploads, kernels, downloads) and returns; wait blocks until
 * the OLDEST submitted frame is complete (its h_out filled).  At most two frames may be in flight; with two, the
 * download of frame f overlaps the upload of frame f + 1 (PCIe is full duplex).  Frames execute in submission
 * order.  The host buffers of a frame must stay valid and untouched until its wait returns.
 * mpeghb200_fd_execute = drain, submit, wait. */
mpeghb200_status mpeghb200_fd_submit(mpeghb200_fd_session *s, const float *h_coef, const uint32_t *h_side,
                                     float *h_out);
mpeghb200_status mpeghb200_fd_wait(mpeghb200_fd_session *s);
/* Zero the state of one stream (the reference's flush / re-init, decoder/impeghd_api.c:1156). */
mpeghb200_status mpeghb200_fd_reset_stream(mpeghb200_fd_session *s, int stream);
void mpeghb200_fd_close(mpeghb200_fd_session *s);

/* ---- tail of the chain: loudness gain, peak limiter, PCM packer -----------------------------------
 * Limiter: replaces impeghd_peak_limiter_init / _process (decoder/impeghd_peak_limiter.c:146,190) together
 * with the planar<->interleaved transposes of impegh_dec_peak_limiter (decoder/ia_core_coder_decode_main.c:1429)
 * and the loudness-normalisation multiply of impegh_dec_ln_pl_process (:1533-1544).
 * Audio in/out: planar [stream][n_ch][1024] at 16-bit scale; output is delayed by attack_samples (240 @48 kHz)
 * exactly like the reference.  Per-stream state lives in d_state (mpeghb200_limiter_state_floats() floats per
 * stream, initialised by mpeghb200_limiter_state_init_dev; re-initialise a stream's slice on flush). */
typedef struct mpeghb200_limiter_params {
  int32_t n_ch;
  int32_t attack_samples; /* (int)(5 ms * fs / 1000) */
  int32_t limiter_on;     /* 0: pure delay while the smoothed gain is >= 1 (impeghd_peak_limiter.c:207-211) */
  float threshold;        /* 0.89125094f * 32768.0f */
  float attack_const;     /* (float)pow(0.1, 1 / (attack_samples + 1)), host double pow like the reference */
  float release_const;    /* (float)pow(0.1, 1 / (50 ms * fs / 1000 + 1)) */
} mpeghb200_limiter_params;
mpeghb200_status mpeghb200_limiter_params_init(mpeghb200_limiter_params *p, int n_ch, int sample_rate,
                                               int limiter_on);
size_t mpeghb200_limiter_state_floats(const mpeghb200_limiter_params *p);
mpeghb200_status mpeghb200_limiter_state_init_dev(mpeghb200_ctx *ctx, const mpeghb200_limiter_params *p,
                                                  float *d_state, int n_streams, void *cuda_stream);
/* d_loudness_gain: per-stream linear gain (float)pow(10, dB/20) computed by the host, or NULL. d_in != d_out. */
mpeghb200_status mpeghb200_limiter_dev(mpeghb200_ctx *ctx, const mpeghb200_limiter_params *p, const float *d_in,
                                       float *d_out, float *d_state, const float *d_loudness_gain,
                                       int n_streams, void *cuda_stream);

/* PCM packer: replaces ia_core_coder_samples_sat (decoder/ia_core_coder_decode_main.c:209).
 * d_in planar [stream][n_ch][1024]; d_out [stream][1024 * n_ch * pcm_bits/8] interleaved little-endian;
 * out_ch_map: HOST array of n_ch entries as in the reference (-1 = identity) or NULL;
 * d_delay: per-stream start-up delay in samples (the handle's delay_in_samples), read and advanced, or NULL;
 * d_out_bytes: per-stream valid byte count of this frame (the reference's *out_bytes), or NULL. */
mpeghb200_status mpeghb200_pcm_pack_dev(mpeghb200_ctx *ctx, const float *d_in, int n_ch, int pcm_bits,
                                        const int32_t *out_ch_map, int32_t *d_delay, uint8_t *d_out,
                                        int32_t *d_out_bytes, int n_streams, void *cuda_stream);

/* Limiter and packer together (what a chain's tail runs): a gains kernel (one CTA per stream: channel maxima, sliding
 * maximum, the serial smoother) and an apply-and-pack kernel (delayed sample x gain, clip, PCM bytes; four CTAs per
 * stream) -- the limited samples never make the planar fp32 round trip through HBM and the data-parallel half of the
 * limiter is not tied to the smoother's one CTA per stream.  d_scratch [n_streams *
 * mpeghb200_limiter_pack_scratch_floats] (gains + the previous delay line).  Other arguments as the two calls above;
 * results identical to mpeghb200_limiter_dev followed by mpeghb200_pcm_pack_dev. */
size_t mpeghb200_limiter_pack_scratch_floats(const mpeghb200_limiter_params *p);
mpeghb200_status mpeghb200_limiter_pack_dev(mpeghb200_ctx *ctx, const mpeghb200_limiter_params *p, const float *d_in,
                                            float *d_state, float *d_scratch, const float *d_loudness_gain,
                                            int pcm_bits, const int32_t *out_ch_map, int32_t *d_delay, uint8_t *d_pcm,
                                            int32_t *d_pcm_bytes, int n_streams, void *cuda_stream);

/* The same for a block of n_samples per channel (a multiple of 256) whose channels lie ch_stride floats apart:
 * d_in [stream][n_ch][ch_stride].  The binaural renderer hands over fft_size/2 samples per channel at a time
 * (num_samples_processed, decoder/ia_core_coder_decode_main.c:2839-2856) and the reference packs them with one
 * ia_core_coder_samples_sat call.  mpeghb200_pcm_pack_dev = n_samples 1024, ch_stride 1024. */
mpeghb200_status mpeghb200_pcm_pack_block_dev(mpeghb200_ctx *ctx, const float *d_in, int n_ch, int n_samples,
                                              int ch_stride, int pcm_bits, const int32_t *out_ch_map,
                                              int32_t *d_delay, uint8_t *d_out, int32_t *d_out_bytes, int n_streams,
                                              void *cuda_stream);

/* ---- object renderer: VBAP gains + ramped mix ------------------------------------------------------
 * Replaces impeghd_obj_renderer_dec (decoder/impeghd_obj_ren_dec.c:1319) and everything below
 * impegh_obj_ren_vbap_process (decoder/impeghd_obj_ren_vbap.c:485); OAM frame length 1024 here, sub-frames of 512 /
 * 256 samples through mpeghb200_obj_render_sub_dev below.
 *
 * Layout tables: the reference's init (impeghd_obj_renderer_dec_init, impeghd_obj_ren_dec.c:1057) stays host
 * code; after it has run, copy these members of ia_obj_ren_dec_state_struct
 * (decoder/impeghd_obj_ren_dec_struct_def.h:90-120) into the descriptor:
 *   n_spk = num_cicp_speakers, n_non_lfe = num_non_lfe_ls, n_imag = num_imag_ls, n_tri = num_triangles,
 *   height_present = height_spks_present, lfe[ch] = ptr_cicp_ls_geo[ch]->lfe_flag,
 *   tri[t][k] = ls_triangle[t].spk_index[k], inv[t][3*r + {0,1,2}] = ls_inv_mtx[t].row[r].{x,y,z},
 *   cart[i] = non_lfe_ls_str[i].ls_cart_coord.{x,y,z}, dmx[i][j] = downmix_mtx[i][j]. */
#define MPEGHB200_OBJ_MAX_LS 44
#define MPEGHB200_OBJ_MAX_OBJECTS 32
typedef struct mpeghb200_obj_layout_desc {
  int32_t n_spk, n_non_lfe, n_imag, n_tri, height_present;
  int32_t lfe[MPEGHB200_OBJ_MAX_LS];
  int32_t tri[MPEGHB200_OBJ_MAX_LS][3];
  float inv[MPEGHB200_OBJ_MAX_LS][9];
  float cart[MPEGHB200_OBJ_MAX_LS][3];
  float dmx[MPEGHB200_OBJ_MAX_LS][MPEGHB200_OBJ_MAX_LS];
} mpeghb200_obj_layout_desc;
typedef struct mpeghb200_obj_layout mpeghb200_obj_layout;
mpeghb200_status mpeghb200_obj_layout_create(mpeghb200_ctx *ctx, const mpeghb200_obj_layout_desc *desc,
                                             mpeghb200_obj_layout **out);
void mpeghb200_obj_layout_destroy(mpeghb200_obj_layout *layout);
int mpeghb200_obj_layout_num_speakers(const mpeghb200_obj_layout *layout); /* n_spk */

/* Per object and frame: the OAM decoder's descaled metadata after the host-side trigonometry (the reference
 * evaluates sin/cos/tan in double-precision libm, decoder/impeghd_3d_vec_basic_ops.c:86 and
 * impeghd_obj_ren_vbap.c:370; those are not bit-reproducible on a GPU, everything else is). 64 bytes. */
typedef struct mpeghb200_obj_side {
  float vec_c[3];  /* impeghd_pol_2_cart_degree(elevation, azimuth, radius) */
  float gain;      /* gain_descaled */
  float spread_w;  /* spread_width_descaled (raw) */
  float spread_h;  /* spread_height_descaled */
  float tan_alpha; /* (float)tan(clamp(width, 0.001, 90) * DEG_2_RAD); only read when spreading */
  float pad0;
  float p0[3];     /* pol_2_cart(elevation, azimuth, 1) */
  float pad1;
  float v[3];      /* pol_2_cart(elevation -/+ 90, azimuth, 1) */
  float pad2;
} mpeghb200_obj_side;
/* Host helper (no GPU needed): params7[n][7] = {azimuth, elevation, radius, gain, spread width, height, depth}
 * as ia_oam_dec_state_struct keeps them (decoder/impeghd_oam_dec_struct_def.h:86-92). */
void mpeghb200_obj_side_from_polar(const float *params7, int n, mpeghb200_obj_side *out);

/* State per stream: previous gains + first-frame flag of every object (mpeghb200_obj_state_floats floats per
 * stream).  d_first_flags: DEVICE int32[n_obj] copy of the reference's first_frame_flag[] or NULL (all 1). */
size_t mpeghb200_obj_state_floats(const mpeghb200_obj_layout *layout, int n_obj);
size_t mpeghb200_obj_scratch_floats(const mpeghb200_obj_layout *layout, int n_obj);
mpeghb200_status mpeghb200_obj_state_init_dev(mpeghb200_ctx *ctx, const mpeghb200_obj_layout *layout, int n_obj,
                                              const int32_t *d_first_flags, float *d_state, int n_streams,
                                              void *cuda_stream);
/* d_side [n_streams][n_obj], d_in [n_streams][n_obj][1024], d_out [n_streams][n_spk][1024] (overwritten; LFE
 * rows zero, like the memset at decoder/ia_core_coder_decode_main.c:1374), d_scratch
 * [n_streams][mpeghb200_obj_scratch_floats]. */
mpeghb200_status mpeghb200_obj_render_dev(mpeghb200_ctx *ctx, const mpeghb200_obj_layout *layout, int n_obj,
                                          const mpeghb200_obj_side *d_side, const float *d_in, float *d_out,
                                          float *d_state, float *d_scratch, int n_streams, void *cuda_stream);

/* OAM sub-frames (decoder/ia_core_coder_decode_main.c:1383-1410): when the OAM frame length is 512 or 256 the
 * reference renders a core frame in 2 or 4 calls of impeghd_obj_renderer_dec, each over frame_len samples starting at
 * sample_offset (a multiple of frame_len) of the same 1024-sample rows, with that sub-frame's metadata (the host
 * passes the matching side records; sub_frame_number bookkeeping stays host) and the gain ramp spread over frame_len
 * samples.  Only [sample_offset, sample_offset + frame_len) of d_out is written.  d_metadata_set [n_streams]
 * (device, or NULL = all 0) is sub_frame_number - 1 of each stream's call: the reference applies its first-frame rule
 * (ramp starts at the final gains) only to calls rendering with set 0, because it tests first_frame_flag[set *
 * num_objects + obj] (decoder/impeghd_obj_ren_vbap.c:497); an object first rendered with another set ramps up from
 * zero gains instead.  mpeghb200_obj_render_dev = frame_len 1024, sample_offset 0, set 0. */
mpeghb200_status mpeghb200_obj_render_sub_dev(mpeghb200_ctx *ctx, const mpeghb200_obj_layout *layout, int n_obj,
                                              const mpeghb200_obj_side *d_side, const float *d_in, float *d_out,
                                              float *d_state, float *d_scratch, int n_streams, int frame_len,
                                              int sample_offset, const int32_t *d_metadata_set, void *cuda_stream);

/* ---- HOA loudspeaker rendering: NFC filtering + render matrix + clip ---------------------------------
 * Replaces impeghd_hoa_ren_block_process (decoder/impeghd_hoa_renderer.c:181) incl. the near-field
 * compensation pre-filter (impeghd_hoa_ren_nfc_filter_apply, decoder/impeghd_hoa_nfc_filtering.c:201).
 * The reference's host init stays in charge of the design; after impeghd_hoa_ren_renderer_init / _nfc_init:
 *   matrix [n_out][n_coef]  = ia_render_hoa_str.v_render_mtrx[mtrx_selected].sm_mtrx.mtrx (row-major),
 *   nfc_coeffs [n_coef][17] = per ia_render_hoa_nfc_filtering_str: {global_gain, num_iir_2_filts,
 *                             num_iir_1_filts, (a1, a2, b1, b2) of str_iir_2[0..2], (a1, b1) of str_iir_1},
 *                             or NULL when use_nfc == 0,
 *   clip                    = is_clipping_protect_enabled (1 at the reference's only call site,
 *                             decoder/impeghd_hoa_decoder.c:223).
 * d_in [stream][n_coef][1024] (not modified), d_out [stream][n_out][1024], d_nfc_state
 * [stream][mpeghb200_hoa_nfc_state_floats] zero-initialised by the caller (cudaMemset / dev_memset). */
typedef struct mpeghb200_hoa_renderer mpeghb200_hoa_renderer;
mpeghb200_status mpeghb200_hoa_renderer_create(mpeghb200_ctx *ctx, const float *matrix, int n_out, int n_coef,
                                               const float *nfc_coeffs, int clip, mpeghb200_hoa_renderer **out);
void mpeghb200_hoa_renderer_destroy(mpeghb200_hoa_renderer *r);
size_t mpeghb200_hoa_nfc_state_floats(const mpeghb200_hoa_renderer *r);
int mpeghb200_hoa_renderer_num_out(const mpeghb200_hoa_renderer *r);
int mpeghb200_hoa_renderer_num_coeffs(const mpeghb200_hoa_renderer *r);
/* Opt-in throughput modes for the one dense contraction of the path (out = M x, impeghd_hoa_renderer.c:231-253),
 * which apply the render matrix on the tensor cores instead of the reference's left-to-right fp32 sum:
 *   1  legacy mma.sync TF32 with the two-term 3xTF32 operand split (any order, NFC on or off);
 *   2  tcgen05.mma kind::tf32, accumulators in tensor memory, EXACT three-term TF32 split of both operands (six
 *      products per k step): NFC-off renderers with at most 32 coefficients (order <= 4) and 32 outputs.
 * Results are then no longer bit-identical but within BASELINE.json's float-stage tolerance, max |err| <= 2^-20 of
 * full scale (2^-5 at the 16-bit scale of the path); mode 2 also holds the final 24-bit PCM within +-1 LSB for
 * programme-level signals (a few ulp of fp32 accumulation-order difference is all that is left).  The NFC cascade
 * and the clip are unchanged.  0 (default) = the bit-exact fp32 kernels. */
mpeghb200_status mpeghb200_hoa_renderer_set_tensor_mode(mpeghb200_hoa_renderer *r, int on);
mpeghb200_status mpeghb200_hoa_render_dev(mpeghb200_ctx *ctx, const mpeghb200_hoa_renderer *r, const float *d_in,
                                          float *d_out, float *d_nfc_state, int n_streams, void *cuda_stream);

/* ---- binaural renderer: FFT overlap-add fast convolution -----------------------------------------------
 * Replaces impeghd_binaural_struct_process_ld_ola (decoder/impeghd_binaural_renderer.c:543) and the filter
 * preparation of impeghd_binaural_renderer_init / impeghd_binaural_struct_set (:346, :244), including the long
 * transforms impeghd_cplx_ifft_8k / impeghd_cplx_fft_16k (decoder/impeghd_ifft.c:1264, impeghd_fft.c:1415).
 *
 * create: the members of ia_td_binaural_ren_param_str (decoder/impeghd_binaural.h:76-87) after the BRIR
 * parameterisation, for n_sets independent filter sets (1 = every stream shares one BRIR set):
 *   taps_direct [set][2][n_in][len_direct], taps_diffuse [set][n_diffuse_blocks][2][len_direct],
 *   direct_fc [set][2][n_in], diffuse_fc [set][2][n_diffuse_blocks], inv_diffuse_weight [set][n_in].
 * The transform length N is the smallest power of two >= 2 * len_direct (impeghd_binaural_struct_init, :200-205); a
 * block is B = N/2 samples per input.  N = 2048 and N = 8192 (direct filters of 513 .. 1024 and 2049 .. 4096 taps) run on
 * tuned kernels; N = 16 .. 1024 (8 .. 512 taps) and N = 16384 (4097 .. 8192 taps) on a generic one that mirrors the
 * reference's transforms pass for pass -- for N = 16384 including the missing column stage of impeghd_cplx_ifft_8k
 * (decoder/impeghd_ifft.c:1318-1359), i.e. it reproduces what the reference computes, which is not a convolution.
 * N = 4096 (1025 .. 2048 taps) is refused with MPEGHB200_ERR_UNSUPPORTED: impeghd_cplx_ifft_4 (:1212) reads two floats
 * behind the transform's buffer, so the reference's result depends on what an earlier tool left in the shared scratch
 * (tests/test_oracle_binaural.py::test_ifft_4096_depends_on_stale_scratch) and there is nothing to reproduce.
 *
 * block_dev: one block for every stream.  d_in layout 0 = [stream][n_in][B]; layout 1 = [B/1024][stream][n_in][1024]
 * (B/1024 consecutive decoder frames stacked, what the reference accumulates in input_mem_buf).  Inputs are
 * multiplied by in_scale (1/32768: ia_div_q15_flt at decoder/ia_core_coder_decode_main.c:2833) and outputs by
 * out_scale (32768, :2851); pass 1.0f for the renderer's own +-1 domain.  d_out [stream][2][B].
 * d_state [stream][mpeghb200_binaural_state_floats], zero-initialised by the caller, holds the overlap tail,
 * the diffuse downmix ring and its write index.  d_filter_set [stream] selects the set per stream (NULL: set 0).
 * With 2 diffuse blocks the reference indexes prev_in_buf[2][..] out of bounds; this library keeps a 3-slot
 * ring instead (0 and 1 diffuse blocks are bit-identical to the reference). */
typedef struct mpeghb200_binaural mpeghb200_binaural;
mpeghb200_status mpeghb200_binaural_create(mpeghb200_ctx *ctx, int n_in, int len_direct, int n_diffuse_blocks,
                                           int n_sets, const float *taps_direct, const float *taps_diffuse,
                                           const float *direct_fc, const float *diffuse_fc,
                                           const float *inv_diffuse_weight, mpeghb200_binaural **out);
/* The same renderer from what the reference's own init left in ia_binaural_ren_pers_mem_str
 * (decoder/impeghd_binaural_renderer.h:52-76): fft_size, the packed filter spectra taps_direct[2][n_in][..] (rows of
 * direct_row_stride floats, the first fft_size of each are used, the two ears direct_ear_stride floats apart; in
 * direct_process_size[2][..] the ears are direct_size_stride entries apart), taps_diffuse[block][2][..], the multiply lengths direct_process_size /
 * diffuse_process_size[block][2] and diffuse_weight[n_in].  One filter set.  This is what the drop-in uses: nothing is
 * re-designed, the filters are the reference's bit for bit.  fft_size: every length mpeghb200_binaural_create takes. */
mpeghb200_status mpeghb200_binaural_create_from_state(mpeghb200_ctx *ctx, int fft_size, int n_in, int n_diffuse_blocks,
                                                      const float *h_direct, size_t direct_row_stride,
                                                      size_t direct_ear_stride, const float *h_diffuse,
                                                      size_t diffuse_row_stride,
                                                      const int32_t *direct_process_size, int direct_size_stride,
                                                      const int32_t *diffuse_process_size,
                                                      const float *diffuse_weight, mpeghb200_binaural **out);
void mpeghb200_binaural_destroy(mpeghb200_binaural *b);
int mpeghb200_binaural_block_len(const mpeghb200_binaural *b);
int mpeghb200_binaural_num_in(const mpeghb200_binaural *b);
size_t mpeghb200_binaural_state_floats(const mpeghb200_binaural *b);
/* Host copies of the prepared filter spectra in the reference's packed format (taps_direct / taps_diffuse of
 * ia_binaural_ren_pers_mem_str): h_direct [set][2][n_in][N], h_diffuse [set][n_diffuse_blocks][2][N]; either
 * may be NULL. */
mpeghb200_status mpeghb200_binaural_spectra(mpeghb200_ctx *ctx, const mpeghb200_binaural *b, float *h_direct,
                                            float *h_diffuse);
mpeghb200_status mpeghb200_binaural_block_dev(mpeghb200_ctx *ctx, const mpeghb200_binaural *b, const float *d_in,
                                              int in_layout, float in_scale, float out_scale, float *d_out,
                                              float *d_state, const int32_t *d_filter_set, int n_streams,
                                              void *cuda_stream);

/* ---- HOA spatial decoding: transport channels -> HOA coefficient signals ----------------------------------
 * Replaces the signal half of impeghd_hoa_spatial_process (decoder/impeghd_hoa_spatial_decoder.c:663): inverse
 * dynamic correction (impeghd_hoa_inverse_dyn_correction.c:75), ambience synthesis (impeghd_hoa_amb_syn.c:76),
 * channel reassignment (impeghd_hoa_ch_reassignment.c:71), vector-based and direction-based predominant sound
 * synthesis incl. spatial prediction (impeghd_hoa_vector_based_predom_sound_syn.c:76,
 * impeghd_hoa_dir_based_pre_dom_sound_syn.c:145-605).  The bitstream side, impeghd_hoa_frame and
 * impeghd_hoa_spatial_decode_frame_side_info (:291), stays host code; what it leaves in
 * ia_hoa_dec_frame_param_str (decoder/impeghd_hoa_frame_params.h:38-62) is handed over per frame as one
 * mpeghb200_hoa_spatial_side per stream.  Copy the arrays whole: the reference itself reads slots beyond the
 * current counts (stale entries of vectors[].index and of active_and_grid_dir_indices[]).
 *
 * desc: after impeghd_hoa_spatial_init (:154) copy from ia_spatial_dec_str / ia_hoa_config_struct
 *   order, n_transport = num_transport_ch, n_addnl_coders = num_addnl_coders,
 *   low_order_mat_sz = amb_syn_handle.low_order_mat_sz, vec_start, coded_v_vec_length,
 *   spat_interp_time = spat_interpolation_time, pred_dir_sigs, num_bits_per_sf, use_phase_shift_decorr (must be 0),
 *   order_matrix            = amb_syn_handle.order_matrix                      [low][low]
 *   fine_grid_mode_matrix   = transpose_mod_mtx_grid_pts                       [900][49]
 *   coarse_grid_mode_matrix = ...dir_based_pre_dom_sound_syn_handle.mode_mt_all_coarse_grid_points [n_coef][n_coef]
 *   fade_windows            = fade_in_win_for_vec_sigs, fade_out_win_for_vec_sigs, fade_in_win_for_dir_sigs,
 *                             fade_out_win_for_dir_sigs                        [4][1024]
 *   dyn_win_pow_table       = ia_hoa_inv_dyn_win_pow_table                     [7][1024]
 *   dyn_win_func            = inv_dyn_corr_win_func                            [1024]
 *   pow2_table              = impeghd_hoa_pow_2(0..63), impeghd_hoa_pow_2(-0..-63)  [128] (the reference's
 *                             negative-power table is a decimal literal, exact only down to 2^-17)
 * Results are bit-identical to the reference except for gain-correction exponents beyond +-6, where the
 * reference evaluates double-precision pow() per sample (device pow: <= 1 ulp of the float result). */
#define MPEGHB200_HOASP_MAX_CH 16
#define MPEGHB200_HOASP_MAX_COEF 49
#define MPEGHB200_HOASP_MAX_PRED 4
typedef struct mpeghb200_hoa_spatial_side {
  int32_t n_vec, n_dir, n_enable, n_disable; /* num_vec/dir_predom_sounds, num_enable/disable_coeff */
  int32_t exponents[MPEGHB200_HOASP_MAX_CH];
  int32_t is_exception[MPEGHB200_HOASP_MAX_CH];
  float prev_gain_reset[MPEGHB200_HOASP_MAX_CH]; /* hoa_independency_flag: pow_2(gain_corr_prev_amp_exp), else 0 */
  int32_t amb_hoa_assign[MPEGHB200_HOASP_MAX_CH];
  int32_t vec_index[MPEGHB200_HOASP_MAX_CH];     /* vectors[].index, all 16 slots */
  int32_t dir_ch[MPEGHB200_HOASP_MAX_CH + 1];    /* active_and_grid_dir_indices[0..16].ch_idx */
  int32_t dir_id[MPEGHB200_HOASP_MAX_CH + 1];
  int32_t pad0[2];
  int32_t enable[MPEGHB200_HOASP_MAX_COEF];      /* amb_coeff_indices_to_enable (sorted, -1 padded) */
  int32_t disable[MPEGHB200_HOASP_MAX_COEF];
  int32_t non_en_dis[MPEGHB200_HOASP_MAX_COEF];  /* non_en_dis_able_act_hoa_coeff_indices */
  int32_t pred_type[MPEGHB200_HOASP_MAX_COEF];
  int32_t pred_idx[MPEGHB200_HOASP_MAX_PRED][MPEGHB200_HOASP_MAX_COEF];   /* prediction_indices_mat[r * n_coef + g] */
  int32_t pred_quant[MPEGHB200_HOASP_MAX_PRED][MPEGHB200_HOASP_MAX_COEF]; /* quant_pred_fact_mat */
  float vec[MPEGHB200_HOASP_MAX_CH][MPEGHB200_HOASP_MAX_COEF];            /* vectors[].sig_id */
} mpeghb200_hoa_spatial_side;
typedef struct mpeghb200_hoa_spatial_desc {
  int32_t order, n_transport, n_addnl_coders, low_order_mat_sz, vec_start, coded_v_vec_length, spat_interp_time,
      pred_dir_sigs, num_bits_per_sf, use_phase_shift_decorr;
  const float *order_matrix, *fine_grid_mode_matrix, *coarse_grid_mode_matrix, *fade_windows, *dyn_win_pow_table,
      *dyn_win_func, *pow2_table;
} mpeghb200_hoa_spatial_desc;
typedef struct mpeghb200_hoa_spatial mpeghb200_hoa_spatial;
mpeghb200_status mpeghb200_hoa_spatial_create(mpeghb200_ctx *ctx, const mpeghb200_hoa_spatial_desc *desc,
                                              mpeghb200_hoa_spatial **out);
void mpeghb200_hoa_spatial_destroy(mpeghb200_hoa_spatial *h);
int mpeghb200_hoa_spatial_num_coeffs(const mpeghb200_hoa_spatial *h);
int mpeghb200_hoa_spatial_num_transport(const mpeghb200_hoa_spatial *h);
/* Per-stream device state (what ia_spatial_dec_str carries between frames + per-frame scratch), in bytes. */
size_t mpeghb200_hoa_spatial_state_bytes(const mpeghb200_hoa_spatial *h);
mpeghb200_status mpeghb200_hoa_spatial_state_init_dev(mpeghb200_ctx *ctx, const mpeghb200_hoa_spatial *h,
                                                      void *d_state, int n_streams, void *cuda_stream);
/* d_side [stream] (device), d_in [stream][n_transport][1024] (perceptually decoded transport signals),
 * d_out [stream][n_coef][1024]. */
mpeghb200_status mpeghb200_hoa_spatial_dev(mpeghb200_ctx *ctx, const mpeghb200_hoa_spatial *h,
                                           const mpeghb200_hoa_spatial_side *d_side, const float *d_in, float *d_out,
                                           void *d_state, int n_streams, void *cuda_stream);

/* ---- format converter: STFT-domain active downmix ------------------------------------------------------
 * Replaces impeghd_format_converter_process (decoder/ia_core_coder_process.c:641) and
 * impeghd_format_conv_active_dmx_stft_process (:127).  The frequency-dependent downmix tensor is designed by
 * the reference's host init (impeghd_format_conv_init, decoder/ia_core_coder_decode_main.c:519) and handed over:
 *   dmx [n_out][n_in][257] = fc_params->active_dmx_stft->downmix_mat[out][in][band].
 * d_in [stream][n_in][1024] -> d_out [stream][n_out][1024] (the reference converts in place in
 * time_sample_vector; here input and output are separate buffers).  d_state
 * [stream][mpeghb200_format_conv_state_floats], zero-initialised by the caller: STFT input/output memories
 * (stft_input_buf, stft_output_buf) and the smoothed ERB-band energies (target/realized_energy_prev). */
typedef struct mpeghb200_format_conv mpeghb200_format_conv;
mpeghb200_status mpeghb200_format_conv_create(mpeghb200_ctx *ctx, int n_in, int n_out, const float *dmx,
                                              mpeghb200_format_conv **out);
void mpeghb200_format_conv_destroy(mpeghb200_format_conv *h);
size_t mpeghb200_format_conv_state_floats(const mpeghb200_format_conv *h);
int mpeghb200_format_conv_num_in(const mpeghb200_format_conv *h);
int mpeghb200_format_conv_num_out(const mpeghb200_format_conv *h);
mpeghb200_status mpeghb200_format_conv_dev(mpeghb200_ctx *ctx, const mpeghb200_format_conv *h, const float *d_in,
                                           float *d_out, float *d_state, int n_streams, void *cuda_stream);

/* ---- TNS: all-pole filtering of spectral ranges, in place, before the FD synthesis ------------------------
 * Replaces the filtering of ia_core_coder_tns_apply (decoder/ia_core_coder_tns.c:204) /
 * impeghd_tns_decode_coeff_ar_filter (:83).  The host resolves each transmitted filter the way :240-297 does
 * (band clamps against tns_max_bands / nbands / IGF, sfb table lookup) and decodes its coefficients with
 * mpeghb200_tns_lpc_from_coef (double-precision sin like the reference, no GPU needed), then hands over one
 * mpeghb200_tns_filter per filter with order > 0 and size > 0.  Filters of the same window of the same channel
 * form a chain and are applied in array order; chain k = filters [chain_first[k], chain_first[k+1]). */
typedef struct mpeghb200_tns_filter {
  int32_t unit;  /* channel-frame index into d_coef [n_units][1024] */
  int32_t start; /* first bin of the range inside the unit (window offset 128 * w included) */
  int32_t size;  /* number of bins */
  int32_t inc;   /* +1: upwards from start; -1: downwards from start + size - 1 (direction != 0) */
  int32_t order; /* 1..15 */
  int32_t pad[3];
  float lpc[16]; /* lpc[j] = the reference's lpc[j + 1] */
} mpeghb200_tns_filter;
mpeghb200_status mpeghb200_tns_lpc_from_coef(int order, int coef_res, const int16_t *coef, float *lpc16);
/* Host half of ia_core_coder_tns_apply for one channel-frame (no GPU needed): the transmitted TNS data as the
 * parser leaves it in ia_tns_frame_info_struct (decoder/ia_core_coder_tns_usac.h:50-71) -> filter records, one
 * chain per window with an effective filter.  Same clamps as the reference (:240-284: tns_max_bands, max_sfb, the
 * IGF start / stop band via ia_core_coder_igf_nbands :147); a band index beyond the table returns
 * MPEGHB200_ERR_BAD_ARG where the reference returns IA_MPEGH_DEC_EXE_FATAL_INVALID_TNS_SFB. */
typedef struct mpeghb200_tns_channel {
  int32_t unit;    /* channel-frame index the records will carry */
  int32_t is_long; /* 1: one window of 1024 bins, 0: eight windows of 128 */
  int32_t max_sfb; /* the nbands argument of ia_core_coder_tns_apply */
  int32_t igf_active, igf_after_tns_synth, igf_sfb_start, igf_sfb_stop; /* igf_grid of the block type */
  int32_t n_filt[8];   /* per window */
  int32_t coef_res[8]; /* 3 or 4, per window */
  struct {
    int32_t order, start_band, stop_band, direction;
    int16_t coef[16];
  } filt[8][3];
} mpeghb200_tns_channel;
mpeghb200_status mpeghb200_tns_resolve(const mpeghb200_tns_channel *ch, const int16_t *sfb_top, int sfb_per_sbk,
                                       int tns_max_band, mpeghb200_tns_filter *filters_out, int filters_cap,
                                       int32_t *chain_len_out, int chains_cap, int *n_filters_out,
                                       int *n_chains_out);
mpeghb200_status mpeghb200_tns_dev(mpeghb200_ctx *ctx, float *d_coef, const mpeghb200_tns_filter *d_filters,
                                   const int32_t *d_chain_first, int n_chains, void *cuda_stream);

/* ---- joint-stereo tools of a channel-pair element, on the spectra, before TNS / FD synthesis ---------------
 * Replaces impeghd_ms_stereo (decoder/ia_core_coder_ext_ch_ele.c:1427), ia_core_coder_cplx_pred_upmixing (:561)
 * incl. ia_core_coder_estimate_dmx_im (:495), the previous-frame downmix ia_core_coder_cplx_prev_mdct_dmx (:703)
 * and the spectrum save ia_core_coder_usac_cplx_save_prev (:93).  The host keeps the bitstream side: mask and
 * alpha_q decoding (ia_core_coder_read_ms_mask :262, ia_core_coder_cplx_pred_data :128) and expands window
 * groups to windows.  Order of the launches for one frame, as ia_core_coder_data_process (:1695) runs them:
 *   tns_on_lr == 0:  mpeghb200_tns_dev, mpeghb200_stereo_dev      tns_on_lr == 1:  stereo, then TNS
 *   then mpeghb200_stereo_save_dev, then mpeghb200_fd_synth_dev.
 * The same side record applied to other unit pairs (left/right of the IGF source tiles) repeats the tool there.
 * Pair state: mpeghb200_stereo_state_floats() floats per slot (the two saved spectra), zero-initialised. */
typedef struct mpeghb200_stereo_pair {
  int32_t left, right; /* unit indices of the two spectra in d_coef [n_units][1024] */
  int32_t slot;        /* pair-state slot */
  uint8_t tool;        /* ms_mask_present: 0 none, 1 (or 2 with a full mask) M/S by band mask, 3 complex prediction */
  uint8_t is_short;    /* EIGHT_SHORT: 8 windows of 128 bins */
  uint8_t window_sequence, window_shape, window_shape_prev; /* select the MDST filter pair */
  uint8_t pred_dir, complex_coef, use_prev_frame;
  uint8_t first_band, last_band; /* M/S band range; band range of real-only prediction (0, max_sfb_ste) */
  uint8_t prev_dmx; /* 1: previous-frame downmix from the saved spectra; 0: zeros (independency flag, LPD history) */
  uint8_t save;     /* mpeghb200_stereo_save_dev: 0 keep, 1 save the last window, 2 save zeros (:1955-1963) */
  uint8_t used[128];     /* ms_used / cplx_pred_used per (window, band): long [band], short [16 * window + band] */
  int16_t alpha_re[128]; /* alpha_q_re, same indexing */
  int16_t alpha_im[128];
} mpeghb200_stereo_pair;
/* Scale-factor-band tables of one sampling rate, shared by the spectral tools (stereo, IGF).
 * sfb_top_*: upper band edges of one window (ptr_sfb_1024 / ptr_sfb_128 of the stream's sampling rate,
 * decoder/ia_core_coder_rom.c:1393); they must end at 1024 / 128. */
typedef struct mpeghb200_bands mpeghb200_bands;
mpeghb200_status mpeghb200_bands_create(mpeghb200_ctx *ctx, const int16_t *sfb_top_long, int n_long,
                                        const int16_t *sfb_top_short, int n_short, mpeghb200_bands **out);
void mpeghb200_bands_destroy(mpeghb200_bands *b);
size_t mpeghb200_stereo_state_floats(void);
mpeghb200_status mpeghb200_stereo_dev(mpeghb200_ctx *ctx, const mpeghb200_bands *bands, float *d_coef,
                                      const mpeghb200_stereo_pair *d_pairs, const float *d_state, int n_pairs,
                                      void *cuda_stream);
mpeghb200_status mpeghb200_stereo_save_dev(mpeghb200_ctx *ctx, const mpeghb200_bands *bands, const float *d_coef,
                                           const mpeghb200_stereo_pair *d_pairs, float *d_state, int n_pairs,
                                           void *cuda_stream);
/* MDST filter table of the complex-prediction tool: curr [4][2][2][7], prev [2][2][7] (host, no GPU needed). */
void mpeghb200_stereo_mdst_rom(float *curr112, float *prev28);

/* ---- intelligent gap filling (IGF), mono path, on the spectra, in place -----------------------------------
 * Replaces ia_core_coder_igf_mono (decoder/ia_core_coder_igf_dec.c:2479: whitening :913, energies :1083, gains
 * and fill :1321) and ia_core_coder_igf_tnf (:2297) for FD long and eight-short blocks.  The host keeps the
 * bitstream side: ia_core_coder_get_igf_levels / _get_igf_data (:1901, :2048; arithmetic decoding, the level
 * dequantisation with pow) and the source tiles igf_input_spec it builds while parsing
 * (decoder/ia_core_coder_arith_dec.c:686-790).  Position in the frame: where ia_core_coder_data_process calls
 * igf_mono (before or after TNS, igf_after_tns_synth).
 * d_tiles [n][4][1024]: the unit's four source tiles (tile t of the reference lives at igf_input_spec +
 * 2048 * t; the first 1024 floats of each are used by FD blocks), consumed (whitened in place is NOT written
 * back).  d_tnf_state [n_units][2048], zero-initialised, indexed by side.unit: igf_tnf_spec_float / igf_tnf_mask,
 * which persist in the reference handle.  d_seed_out [n]: seed_value[ch] after the frame (the noise-filling
 * parser of the next frame continues from it). */
typedef struct mpeghb200_igf_side {
  int32_t unit;   /* channel-frame index into d_coef [n_units][1024] and d_tnf_state */
  uint32_t seed;  /* seed_value[ch] before the frame */
  uint8_t win_type; /* igf_win_type: 0 long block, 1 eight short */
  uint8_t sfb_start, sfb_stop, igf_min; /* igf_grid[win_type] of the element (ia_core_coder_igf_init :2171) */
  uint8_t use_high_res, use_whitening, use_tnf;
  uint8_t all_zero, is_tnf; /* igf_all_zero, igf_is_tnf of the frame */
  uint8_t num_groups;
  uint8_t pad[2];
  uint8_t group_len[8];
  uint8_t tile_num[4];        /* igf_tile_num */
  uint8_t whitening_level[4]; /* igf_whitening_level */
  float levels[128];          /* igf_levels_curr_float: long [band], short [16 * group + band] */
} mpeghb200_igf_side;
mpeghb200_status mpeghb200_igf_dev(mpeghb200_ctx *ctx, const mpeghb200_bands *bands, float *d_coef,
                                   const float *d_tiles, const mpeghb200_igf_side *d_side, float *d_tnf_state,
                                   uint32_t *d_seed_out, int n_units, void *cuda_stream);
/* Joint-stereo IGF of channel pairs (igf_independent_tiling == 0 with a common window): replaces
 * ia_core_coder_igf_apply_stereo (decoder/ia_core_coder_igf_dec.c:1661; calc :1169, proc :1431).  Pair p uses the
 * side records d_side[2p] (left) and [2p+1] (right) -- same grid, window type and grouping, each channel's own
 * unit, seed, levels, tile numbers and whitening levels, already synchronised the way ia_core_coder_sync_data
 * (:1615) does -- the tiles d_tiles[2p], [2p+1] and the stereo band mask d_mask[p][128] per (group, band)
 * (cplx_pred_used or ms_used; long [band], short [16 * group + band]; zeros without a mask).  It fills both
 * spectra and always leaves the TNF buffers behind.  The frame then continues as
 * ia_core_coder_data_process_igf_joint_stereo_* (decoder/ia_core_coder_ext_ch_ele.c:1500,1590) does:
 * mpeghb200_stereo_dev over the IGF band range on the spectra AND on the TNF spectra (unit index 2 * unit into
 * d_tnf_state viewed as [.][1024]), then mpeghb200_igf_tnf_dev (= ia_core_coder_igf_tnf, :2297) per channel. */
mpeghb200_status mpeghb200_igf_stereo_dev(mpeghb200_ctx *ctx, const mpeghb200_bands *bands, float *d_coef,
                                          const float *d_tiles, const mpeghb200_igf_side *d_side,
                                          const uint8_t *d_mask, float *d_tnf_state, uint32_t *d_seed_out,
                                          int n_pairs, void *cuda_stream);
mpeghb200_status mpeghb200_igf_tnf_dev(mpeghb200_ctx *ctx, const mpeghb200_bands *bands, float *d_coef,
                                       const mpeghb200_igf_side *d_side, float *d_tnf_state, int n_units,
                                       void *cuda_stream);
/* whitening gain tables [2][64] (host, no GPU needed): pinned against the reference ROM by the tests */
void mpeghb200_igf_whitening_rom(float *out128);

/* ---- multichannel coding tool: inverse rotation / prediction boxes on pairs of channel spectra -------------------
 * Replaces impeghd_mc_process (decoder/impeghd_multichannel.c:899; rotation :781, prediction :817) and the window
 * loop around it (decoder/ia_core_coder_process.c:1043-1075) -- SURVEY.md section 8f rank 3, the step ahead of the
 * seam.  A box works in place on two rows of a [unit][1024] spectrum array: downmix / residual in, the two channels
 * out.  The boxes of one stream-frame form a group (d_group_first[g] .. d_group_first[g + 1]) and are applied in
 * order; groups are independent.  Parsing (impeghd_mc_parse) and stereo filling stay on the host;
 * mpeghb200_mct_resolve is the host half: band -> bin resolution per window exactly like the reference
 * (coef_idx / mask: [n_windows][bands_per_window], sfb_offset: the band-end table of the block type). */
#define MPEGHB200_MCT_MAX_SEG 64
typedef struct mpeghb200_mct_box {
  int32_t unit_l, unit_r; /* rows of the spectrum array */
  int32_t op;             /* 0 / 1: prediction with pred_dir 0 / 1 (coef = alpha_q), 2: rotation (coef = angle index 0..64) */
  int32_t n_seg;
  int16_t start[MPEGHB200_MCT_MAX_SEG], len[MPEGHB200_MCT_MAX_SEG], coef[MPEGHB200_MCT_MAX_SEG];
} mpeghb200_mct_box;
mpeghb200_status mpeghb200_mct_resolve(int signal_type, int pred_dir, int fullband, int n_windows, int bins_per_window,
                                       int bands_per_window, int total_sfb, const int16_t *sfb_offset,
                                       const int32_t *coef_idx, const int32_t *mask, int unit_l, int unit_r,
                                       mpeghb200_mct_box *out);
mpeghb200_status mpeghb200_mct_dev(mpeghb200_ctx *ctx, float *d_spec, const mpeghb200_mct_box *d_boxes,
                                   const int32_t *d_group_first, int n_groups, void *cuda_stream);
/* the angle tables impeghd_mc_index_to_sin_alpha / _cos_alpha [65] as this library builds them */
void mpeghb200_mct_rom(float *sin_alpha, float *cos_alpha);

/* ---- element-wise steps between the renderers and the limiter -----------------
 * Bus add: d_out[i] = d_a[i] + d_b[i] over n_floats samples (a multiple of 4; d_out may alias d_a or d_b): the
 * bed + objects / bed + HOA / HOA + objects sums of decoder/ia_core_coder_decode_main.c:2185-2195, 2203-2213,
 * 2276-2302.  Channel mask: channel c of d_audio [n_channels][1024] becomes zeros where d_on[c] == 0 -- the
 * per-channel outcome of impeghd_mdp_dec_process (decode_main.c:906-1070); the metadata group / preset logic that
 * produces d_on stays on the host. */
mpeghb200_status mpeghb200_bus_add_dev(mpeghb200_ctx *ctx, const float *d_a, const float *d_b, float *d_out,
                                       size_t n_floats, void *cuda_stream);
mpeghb200_status mpeghb200_channel_mask_dev(mpeghb200_ctx *ctx, float *d_audio, const uint8_t *d_on, int n_channels,
                                            void *cuda_stream);
/* Planar <-> interleaved, per stream: to_planar != 0 turns d_in [stream][n_samples][n_ch] into d_out
 * [stream][n_ch][n_samples], 0 the reverse (the interleaved copy impegh_dec_peak_limiter builds around the limiter,
 * decoder/ia_core_coder_decode_main.c:1429-1460).  n_ch <= 64 (33 KB of shared memory at most); d_in != d_out. */
mpeghb200_status mpeghb200_interleave_dev(mpeghb200_ctx *ctx, const float *d_in, float *d_out, int n_ch, int n_samples,
                                          int to_planar, int n_streams, void *cuda_stream);
/* Per-sample gain curves: d_audio[c][s] *= d_gains[d_seq_of_ch[c]][s] for every
 * channel with d_seq_of_ch[c] >= 0 -- the time-domain application of single-band DRC gain sequences,
 * impd_drc_apply_gains_and_add (decoder/impd_drc_process.c:68-175).  Gain decoding, DRC set selection and the
 * spline interpolation into the per-sample curve lpcm_gains (libm) stay on the host: d_gains [n_seq][1024]. */
mpeghb200_status mpeghb200_gain_apply_dev(mpeghb200_ctx *ctx, float *d_audio, const float *d_gains,
                                          const int32_t *d_seq_of_ch, int n_channels, void *cuda_stream);
/* The same in the STFT domain the decoder's domain switcher produces (sub_band_domain_mode STFT256: 4 time slots x
 * 256 bands per channel-frame, re / im planes [channel][4][256], im[slot][0] = the Nyquist bin):
 * impd_drc_apply_gains_subband (decoder/impd_drc_process.c:243-385), called from impd_drc_fd_process (:610).
 * Single-band groups multiply by the slot's gain, multi-band groups (2..4 bands) by the ordered sum of overlap weight
 * x sequence gain.  d_gains [n_seq][4]: per gain sequence the value the reference picks from lpcm_gains at the slot
 * centre, lpcm_gains[MAX_SIGNAL_DELAY - gain_delay_samples + offset + 256 slot + 127] (:328, :345); d_weights
 * [rows][256]: overlap_weight of the bands of each multi-band group (impd_drc_generate_overlap_weights), rows of a
 * group consecutive.  One call per selected DRC set, in the reference's order (:640-651). */
typedef struct mpeghb200_drc_sb_channel {
  int32_t first_seq;  /* first gain sequence of the channel's group (gain_idx_grp, :296-304); < 0: no DRC on this channel */
  int32_t band_count; /* band_count_of_ch_group: 1 = single band, 2..4 = multi-band */
  int32_t weight_row; /* first row of the group's overlap weights in d_weights (multi-band only) */
  int32_t reserved;
} mpeghb200_drc_sb_channel;
mpeghb200_status mpeghb200_gain_subband_dev(mpeghb200_ctx *ctx, float *d_re, float *d_im,
                                            const mpeghb200_drc_sb_channel *d_chan, const float *d_gains,
                                            const float *d_weights, int n_channels, void *cuda_stream);
/* The whole STFT-domain DRC step of a frame in one launch, in place on the time signal d_audio [n_channels][1024]:
 * domain switcher time -> frequency, the sub-band gains of n_sets (0..3) selected DRC sets in application order
 * (d_chan [n_sets][n_channels]), frequency -> time -- decoder/ia_core_coder_decode_main.c:2128-2141, i.e.
 * impeghd_domain_switcher_process (decoder/ia_core_coder_process.c:505) with ds_t2f = 1, impd_drc_fd_process
 * (decoder/impd_drc_process.c:610), impeghd_domain_switcher_process with ds_t2f = 0.  Every channel takes the
 * transform round trip, as in the reference.  d_state [n_channels][512] (zero at stream start): the previous input
 * slot and the overlap of the output (stft_in_buf / stft_out_buf of ia_ds_state_struct). */
mpeghb200_status mpeghb200_drc_stft_dev(mpeghb200_ctx *ctx, float *d_audio, float *d_state,
                                        const mpeghb200_drc_sb_channel *d_chan, int n_sets, const float *d_gains,
                                        const float *d_weights, int n_channels, void *cuda_stream);

/* The two halves of the domain switcher on their own (impeghd_domain_switcher_process with ds_t2f = 1 / 0,
 * decoder/ia_core_coder_process.c:556-585, :589-619), for callers that apply the gains as a separate step
 * (mpeghb200_gain_subband_dev on the two planes): to_frequency != 0 reads d_audio [n_channels][1024] and writes d_spec
 * [n_channels][2048] = re [4][256] | im [4][256] in the reference's packed form (time_sample_stft: im[slot][0] carries the
 * Nyquist bin); 0 reads d_spec and writes d_audio.  d_state [n_channels][512] as above: the first half belongs to the
 * forward direction (previous input slot), the second to the inverse one (output overlap); each call touches its own. */
mpeghb200_status mpeghb200_stft_dev(mpeghb200_ctx *ctx, float *d_audio, float *d_spec, float *d_state, int to_frequency,
                                    int n_channels, void *cuda_stream);

/* Time-domain multi-band DRC of one selected set, in place on d_audio [n_channels][1024]: phase-alignment all-pass
 * cascade, Linkwitz-Riley band split (2, 3 or 4 bands), per-band gain curves, band sum --
 * impd_drc_filter_banks_process (decoder/impd_drc_process.c:399; banks decoder/impd_drc_filter_bank.c:367-613) and the
 * multiply-and-sum of impd_drc_apply_gains_and_add (:68-225), chained as in impd_drc_td_process (:529).  The host
 * designs the coefficients (impd_drc_init_all_filter_banks) and interpolates the gains: d_gains [n_seq][1024], the
 * bands of a channel group consecutive.  d_banks: one record per channel group plus, if needed, one for the channels
 * outside every group (only its cascade applies).  d_state [n_channels][mpeghb200_drc_td_state_floats()], zero at
 * stream start (the x_p / y_p memories of ia_iir_filter_struct). */
typedef struct mpeghb200_drc_biquad {
  float a1, a2, b0, b1, b2;
} mpeghb200_drc_biquad;
typedef struct mpeghb200_drc_td_bank {
  int32_t n_bands;   /* 1..4 */
  int32_t n_cascade; /* all-pass sections the reference runs: 0, or str_all_pass_cascade.num_filter + 1 (:604) */
  /* 2 bands: lp, hp; 3: lp_1, hp_1, ap_2, lp_2, hp_2; 4: lp_1, hp_1, ap_2_low, ap_2_high, lp_3_low, hp_3_low, lp_3_high, hp_3_high */
  mpeghb200_drc_biquad f[8];
  mpeghb200_drc_biquad ap[9]; /* str_ap_cascade_filt[0 .. n_cascade - 1] */
} mpeghb200_drc_td_bank;
typedef struct mpeghb200_drc_td_channel {
  int32_t bank;      /* index into d_banks */
  int32_t first_seq; /* first gain sequence of the channel's group; < 0: channel outside every group */
} mpeghb200_drc_td_channel;
size_t mpeghb200_drc_td_state_floats(void);
mpeghb200_status mpeghb200_drc_td_dev(mpeghb200_ctx *ctx, float *d_audio, float *d_state,
                                      const mpeghb200_drc_td_channel *d_chan, const mpeghb200_drc_td_bank *d_banks,
                                      const float *d_gains, int n_channels, void *cuda_stream);

/* ---- LTPF: long-term post-filter of FD channels, in place on the time signal after FD synthesis -------------
 * Replaces ia_core_coder_usac_ltpf_process (decoder/ia_core_coder_ltpf.c:395), called per FD channel right after
 * ia_core_coder_fd_frm_dec (decoder/ia_core_coder_ext_ch_ele.c:1977); SURVEY.md section 8f rank 3.  The side record
 * carries the frame's ltpf_data fields as parsed (decoder/ia_core_coder_acelp_bitparse.c:554-571); a frame without
 * LTPF data (data_present == 0) still has to be passed while the channel's previous frame had it (fade-out) and
 * costs a copy otherwise.  d_state [channel][mpeghb200_ltpf_state_floats(sample_rate)], zero-initialised, indexed
 * by side.unit: previous parameters, the last 6 inputs and the last pitch_max + 4 outputs. */
typedef struct mpeghb200_ltpf_side {
  int32_t unit;           /* channel-frame index into d_audio [n_units][1024] and d_state */
  uint16_t pitch_lag_idx; /* 9 bits */
  uint8_t data_present;
  uint8_t gain_idx;       /* 2 bits */
} mpeghb200_ltpf_side;
size_t mpeghb200_ltpf_state_floats(int sample_rate);
mpeghb200_status mpeghb200_ltpf_dev(mpeghb200_ctx *ctx, int sample_rate, float *d_audio,
                                    const mpeghb200_ltpf_side *d_side, float *d_state, int n_units,
                                    void *cuda_stream);

/* ---- output resampler ------------------------------------------------------------------------------------
 * Replaces impeghd_resample (decoder/impeghd_resampler.c:93): fac_up/fac_down = 2/1, 3/1 or 3/2
 * (impeghd_resampler_get_sampling_ratio, decoder/impeghd_api.c:132).  filter_up = impeghd_resampler_filter
 * (x2) or impeghd_resampler_filter_3 (x3), filter_down = impeghd_resampler_filter (only for fac_down = 2),
 * 60 taps each (decoder/impeghd_resampler_rom.h).  One unit = one channel-frame: d_in [n_units][1024] ->
 * d_out [n_units][1024 * fac_up / fac_down]; d_state [n_units][mpeghb200_resampler_state_floats], zero-initialised
 * by the caller (ia_resampler_struct's delay lines and polyphase memory). */
typedef struct mpeghb200_resampler mpeghb200_resampler;
mpeghb200_status mpeghb200_resampler_create(mpeghb200_ctx *ctx, int fac_up, int fac_down, const float *filter_up,
                                            const float *filter_down, mpeghb200_resampler **out);
void mpeghb200_resampler_destroy(mpeghb200_resampler *r);
size_t mpeghb200_resampler_state_floats(const mpeghb200_resampler *r);
int mpeghb200_resampler_out_len(const mpeghb200_resampler *r);
mpeghb200_status mpeghb200_resample_dev(mpeghb200_ctx *ctx, const mpeghb200_resampler *r, const float *d_in,
                                        float *d_out, float *d_state, int n_units, void *cuda_stream);

/* ---- chains: one BASELINE.json configuration end to end behind one host-buffer call --------------------------
 * A chain owns everything one configuration of the post-parse path needs for n_streams streams on one GPU -- the
 * per-stream state of every stage, device-resident intermediates, two frame slots of input / output staging and
 * three CUDA streams (upload, compute, download) -- and runs the stages in the reference's order
 * (decoder/ia_core_coder_decode_main.c:2736-2893, ia_core_coder_dec_ext_ele_proc :1949-2316):
 *
 *   core channels (dequantised spectra + FD side words -> FD synthesis, or time signals as they are)
 *     = bed channels | object signals | HOA transport channels        (ia_signals_3d: num_ch, num_audio_obj,
 *                                                                       num_hoa_transport_ch)
 *   objects  -> VBAP gains + ramped mix into the layout (:2166-2216)     mpeghb200_obj_render_dev
 *   HOA      -> spatial decode -> NFC + render matrix (:2217-2302)       mpeghb200_hoa_spatial_dev, _hoa_render_dev
 *   bus adds bed + objects + HOA (ia_add_flt loops :2185-2302)           mpeghb200_bus_add_dev
 *   format converter on the mix (:2752-2766)                             mpeghb200_format_conv_dev
 *   binaural renderer (:2826-2860), a block every fft_size/2 samples     mpeghb200_binaural_block_dev
 *   loudness gain + peak limiter (:2861-2867, impegh_dec_ln_pl_process)  mpeghb200_limiter_dev
 *   PCM packer (:2895, ia_core_coder_samples_sat)                        mpeghb200_pcm_pack_dev
 *
 * With the binaural renderer on, the reference's limiter runs over time_sample_vector[0..1] while the packer reads
 * ptr_binaural_output (:2861 vs :2895), i.e. it never touches the binaural signal; the chain therefore packs the
 * renderer's output directly, like the reference does in effect.
 *
 * Only what has to cross the host link does: the core channels and the per-frame side records go up, packed PCM
 * comes down (2, 3 or 4 bytes per output sample instead of 4 per intermediate sample of every stage).
 *
 * Host buffer layout of one frame (n_frames of them back to back in one submit):
 *   core     [n_bed_ch + n_obj + n_hoa_transport regions][stream][channel of the region][1024] float
 *            (region-major: all streams' bed channels, then all streams' object signals, then HOA transport)
 *   fd_side  the same arrangement, one MPEGHB200_FD_SIDE word per channel (core_is_spectra only)
 *   obj_side [stream][n_obj] mpeghb200_obj_side, hoa_side [stream] mpeghb200_hoa_spatial_side
 *   pcm      [stream][pcm_bytes_per_frame]; pcm_bytes [stream] valid bytes (start-up delay trimming), or NULL.
 *            With the binaural renderer: one PCM record per completed block, [stream][block_len * 2 * bytes]. */
typedef struct mpeghb200_chain mpeghb200_chain;
typedef struct mpeghb200_chain_desc {
  int32_t n_streams;
  int32_t core_is_spectra; /* 1: FD synthesis in front (spectra + side words); 0: time-domain core channels */
  int32_t n_bed_ch, n_obj, n_hoa_transport;
  int32_t pcm_bits;        /* 16, 24, 32 */
  int32_t limiter;         /* 0: no limiter stage, 1: limiter on, 2: stage present with limiter_on = 0 (pure delay) */
  int32_t start_delay;     /* delay_in_samples of every stream at start (the reference trims the limiter's 240) */
  float loudness_gain;     /* (float)pow(10, dB / 20) from the host, 1.0f = off (decode_main.c:1533-1544) */
  int32_t max_frames_per_call; /* frames one submit may carry (sizes the two staging slots); 0 = one PCM record */
  const mpeghb200_obj_layout *obj_layout;     /* n_obj > 0 */
  const mpeghb200_hoa_spatial *hoa_spatial;   /* n_hoa_transport > 0 */
  const mpeghb200_hoa_renderer *hoa_renderer; /* n_hoa_transport > 0 */
  const mpeghb200_format_conv *format_conv;   /* or NULL */
  const mpeghb200_binaural *binaural;         /* or NULL */
  const int32_t *out_ch_map;                  /* as mpeghb200_pcm_pack_dev, or NULL */
  const int32_t *binaural_set_of_stream;      /* HOST [n_streams]: BRIR filter set per stream, or NULL (set 0) */
} mpeghb200_chain_desc;
typedef struct mpeghb200_chain_frame {
  const float *core;
  const uint32_t *fd_side;
  const mpeghb200_obj_side *obj_side;
  const mpeghb200_hoa_spatial_side *hoa_side;
  uint8_t *pcm;
  int32_t *pcm_bytes;
} mpeghb200_chain_frame;
mpeghb200_status mpeghb200_chain_open(mpeghb200_ctx *ctx, const mpeghb200_chain_desc *desc, mpeghb200_chain **out);
void mpeghb200_chain_close(mpeghb200_chain *c);
int mpeghb200_chain_out_channels(const mpeghb200_chain *c);
/* floats of one frame's core buffer; bytes of one PCM record per stream (a frame, or a binaural block) */
size_t mpeghb200_chain_core_floats(const mpeghb200_chain *c);
size_t mpeghb200_chain_pcm_record_bytes(const mpeghb200_chain *c);
int mpeghb200_chain_frames_per_record(const mpeghb200_chain *c); /* 1, or block_len / 1024 with the binaural renderer */
/* HOST buffers, n_frames (1..max_frames_per_call <= MPEGHB200_CHAIN_MAX_FRAMES) consecutive frames per call; with the binaural renderer
 * n_frames must be a multiple of frames_per_record.  submit enqueues and returns; wait blocks until the OLDEST
 * submitted call is complete (its pcm / pcm_bytes filled).  At most two calls in flight: the download of one
 * overlaps the upload of the next.  Page-locked host buffers (mpeghb200_host_alloc) are DMA'd in place.
 * execute = drain, submit, wait. */
#define MPEGHB200_CHAIN_MAX_FRAMES 8
mpeghb200_status mpeghb200_chain_submit(mpeghb200_chain *c, const mpeghb200_chain_frame *host, int n_frames);
mpeghb200_status mpeghb200_chain_wait(mpeghb200_chain *c);
mpeghb200_status mpeghb200_chain_execute(mpeghb200_chain *c, const mpeghb200_chain_frame *host, int n_frames);
/* The same stages on DEVICE buffers of the caller (same layouts, n_frames frames), on the caller's stream, without
 * copies: what bench.py times with the inputs already resident in HBM. */
mpeghb200_status mpeghb200_chain_run_dev(mpeghb200_chain *c, const mpeghb200_chain_frame *dev, int n_frames,
                                         void *cuda_stream);

#ifdef __cplusplus
}
#endif
#endif /* MPEGHB200_H */
