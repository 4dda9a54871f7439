/* seam_interpose.c -- libmpeghgpu.so (and the LD_PRELOAD form libmpeghb200_seam.so): the core-coder stages.
 *
 * The reference decoder library keeps its public API (ia_mpegh_dec_create / _init / _execute / _delete,
 * decoder/impeghd_api.h:47-51) and its whole parse half.  This shared object stands in front of it in the symbol
 * lookup order and takes over the post-parse signal functions by ELF symbol interposition -- the reference library
 * is used binary-unmodified, no source patch is needed:
 *
 *   ia_core_coder_fd_frm_dec             decoder/ia_core_coder_imdct.c:832      -> mpeghb200_fd_synth_dev
 *   ia_core_coder_tns_apply              decoder/ia_core_coder_tns.c:204        -> mpeghb200_tns_resolve + _tns_dev
 *   ia_core_coder_usac_ltpf_process      decoder/ia_core_coder_ltpf.c:395       -> mpeghb200_ltpf_dev
 *   impeghd_mc_process                   decoder/impeghd_multichannel.c:899     -> mpeghb200_mct_dev
 *   impeghd_peak_limiter_init/_process   decoder/impeghd_peak_limiter.c:146,190 -> mpeghb200_limiter_dev
 *   impeghd_format_converter_process     decoder/ia_core_coder_process.c:641    -> mpeghb200_format_conv_dev
 *                                        (seam_interpose_fc.c)
 *   impeghd_obj_renderer_dec             decoder/impeghd_obj_ren_dec.c:1319     -> mpeghb200_obj_render_dev
 *   ia_core_coder_samples_sat            decoder/ia_core_coder_decode_main.c:209 -> mpeghb200_pcm_pack_block_dev
 *                                        (seam_interpose_obj.c; the packer is `static` in the reference, so it is
 *                                        reachable only when the reference is built with external linkage for it)
 *
 * An interposer here does not compute.  It describes its call in a seam_req and seam_submit()s it; the back-end of
 * its kind later serves ALL the requests the handles of a batch have parked with one gather / upload / launch /
 * download / scatter (seam_core.c).  A lone ia_mpegh_dec_execute is a batch of one.
 *
 * State ownership: the FD overlap buffers and the LTPF memories stay in the reference handle (they make the round
 * trip with the call) because the reference's LPD / FAC transition code, which stays on the host, reads and writes
 * them too; channel-frames in an LPD transition (td_frame_prev or FAC data present) are forwarded to the reference
 * function.  The limiter's delay line and sliding-maximum history live on the device, one slot per
 * ia_peak_limiter_struct, reset whenever the reference re-initialises that struct.
 *
 * No CPU fallback: if the CUDA context cannot be created the process is stopped with a message; a stage is
 * never silently computed by the reference instead.  MPEGHB200_SEAM_STATS=1 prints the call counters at exit.
 *
 * Built only where the reference headers are available (csrc/Makefile); struct layouts come from the reference's
 * own headers, read in place.
 */
#define _GNU_SOURCE
#include <dlfcn.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <impeghd_type_def.h>
#include <impeghd_fft_ifft.h>
#include "impeghd_intrinsics_flt.h"
#include "ia_core_coder_cnst.h"
#include <ia_core_coder_constants.h>
#include "ia_core_coder_acelp_info.h"
#include "ia_core_coder_acelp_com.h"
#include <ia_core_coder_basic_ops32.h>
#include <ia_core_coder_basic_ops40.h>
#include "ia_core_coder_bitbuffer.h"
#include "ia_core_coder_defines.h"
#include "ia_core_coder_function_selector.h"
#include "ia_core_coder_interface.h"
#include "ia_core_coder_info.h"
#include "ia_core_coder_tns_usac.h"
#include "impeghd_tbe_dec.h"
#include "ia_core_coder_igf_data_struct.h"
#include "ia_core_coder_stereo_lpd.h"
#include "ia_core_coder_main.h"
#include "impeghd_error_codes.h"
#include "impeghd_peak_limiter_struct_def.h"
#include "impeghd_hoa_common_values.h"
#include "impeghd_hoa_matrix_struct.h"
#include "impeghd_hoa_matrix.h"
#include "impd_drc_common.h"
#include "impd_drc_extr_delta_coded_info.h"
#include "impd_drc_struct.h"
#include "ia_core_coder_config.h"

#include "seam_internal.h"
#include <pthread.h>

#define FRAME 1024

/* ---- FD synthesis ------------------------------------------------------------------------------------ */
static seam_buf b_coef, b_ovl, b_out, b_side;

static void fd_gather(int i, void *arg)
{
  seam_req **rq = (seam_req **)arg;
  ia_usac_data_struct *u = (ia_usac_data_struct *)rq[i]->p[0];
  const int ch = (int)rq[i]->v[0];
  memcpy((float *)b_coef.h + (size_t)i * FRAME, u->coef[ch], FRAME * sizeof(float));
  memcpy((float *)b_ovl.h + (size_t)i * FRAME, u->overlap_data_ptr[ch], FRAME * sizeof(float));
  ((uint32_t *)b_side.h)[i] = MPEGHB200_FD_SIDE(u->window_sequence[ch], u->window_shape[ch], u->window_shape_prev[ch],
                                                u->prev_aliasing_symmetry[ch], u->curr_aliasing_symmetry[ch]);
}
static void fd_scatter(int i, void *arg)
{
  seam_req **rq = (seam_req **)arg;
  ia_usac_data_struct *u = (ia_usac_data_struct *)rq[i]->p[0];
  const int ch = (int)rq[i]->v[0];
  memcpy(u->time_sample_vector[ch], (float *)b_out.h + (size_t)i * FRAME, FRAME * sizeof(float));
  memcpy(u->overlap_data_ptr[ch], (float *)b_ovl.h + (size_t)i * FRAME, FRAME * sizeof(float));
}

/* p[0] = usac_data, v[0] = channel */
static void be_fd(seam_req **rq, int n)
{
  mpeghb200_ctx *ctx = seam_ctx();
  const size_t fb = (size_t)n * FRAME * sizeof(float);
  seam_need(&b_coef, fb), seam_need(&b_ovl, fb), seam_need(&b_out, fb), seam_need(&b_side, (size_t)n * sizeof(uint32_t));
  seam_parallel_for(n, fd_gather, rq);
  CK(mpeghb200_h2d(ctx, b_coef.d, b_coef.h, fb));
  CK(mpeghb200_h2d(ctx, b_ovl.d, b_ovl.h, fb));
  CK(mpeghb200_h2d(ctx, b_side.d, b_side.h, (size_t)n * sizeof(uint32_t)));
  CK(mpeghb200_fd_synth_dev(ctx, (const float *)b_coef.d, (const uint32_t *)b_side.d, (float *)b_ovl.d, (float *)b_out.d,
                            n, NULL));
  CK(mpeghb200_d2h(ctx, b_out.h, b_out.d, fb));
  CK(mpeghb200_d2h(ctx, b_ovl.h, b_ovl.d, fb));
  seam_parallel_for(n, fd_scatter, rq);
  seam_n.fd_gpu += (unsigned long)n;
}

IA_ERRORCODE ia_core_coder_fd_frm_dec(ia_usac_data_struct *usac_data, WORD32 i_ch, WORD32 ele_id)
{
  typedef IA_ERRORCODE (*fn_t)(ia_usac_data_struct *, WORD32, WORD32);
  seam_req r;
  if (usac_data->td_frame_prev[i_ch] || usac_data->fac_data_present[i_ch] || usac_data->ccfl != FRAME)
  {
    /* LPD -> FD transition frame: FAC / BPF / TBE hand-over stays with the host LPD decoder */
    static fn_t real;
    if (!real)
      real = (fn_t)dlsym(RTLD_NEXT, "ia_core_coder_fd_frm_dec");
    __sync_fetch_and_add(&seam_n.fd_host, 1);
    return real(usac_data, i_ch, ele_id);
  }
  r.kind = RQ_FD;
  r.p[0] = usac_data;
  r.v[0] = i_ch;
  seam_submit(&r);
  return IA_MPEGH_DEC_NO_ERROR;
}

/* ---- TNS: the whole of ia_core_coder_tns_apply for one channel ------------------------------------------ */
static seam_buf b_tns_spec, b_tns_filt, b_tns_first;

/* p[0] = spec (1024 floats, in place), p[1] = filter records, p[2] = chain lengths, v[0] = n_filters, v[1] = n_chains */
static void be_tns(seam_req **rq, int n)
{
  mpeghb200_ctx *ctx = seam_ctx();
  const size_t fb = (size_t)n * FRAME * sizeof(float);
  int i, k, n_filt = 0, n_chain = 0;
  mpeghb200_tns_filter *hf;
  int32_t *first;
  for (i = 0; i < n; i++)
    n_filt += (int)rq[i]->v[0], n_chain += (int)rq[i]->v[1];
  seam_need(&b_tns_spec, fb);
  seam_need(&b_tns_filt, (size_t)n_filt * sizeof(mpeghb200_tns_filter));
  seam_need(&b_tns_first, (size_t)(n_chain + 1) * sizeof(int32_t));
  hf = (mpeghb200_tns_filter *)b_tns_filt.h;
  first = (int32_t *)b_tns_first.h;
  n_filt = n_chain = 0;
  for (i = 0; i < n; i++)
  {
    const mpeghb200_tns_filter *f = (const mpeghb200_tns_filter *)rq[i]->p[1];
    const int32_t *len = (const int32_t *)rq[i]->p[2];
    int done = 0;
    memcpy((float *)b_tns_spec.h + (size_t)i * FRAME, rq[i]->p[0], FRAME * sizeof(float));
    for (k = 0; k < (int)rq[i]->v[1]; k++)
    {
      first[n_chain++] = n_filt + done;
      done += len[k];
    }
    for (k = 0; k < (int)rq[i]->v[0]; k++)
    {
      hf[n_filt] = f[k];
      hf[n_filt++].unit = i;
    }
  }
  first[n_chain] = n_filt;
  CK(mpeghb200_h2d(ctx, b_tns_spec.d, b_tns_spec.h, fb));
  CK(mpeghb200_h2d(ctx, b_tns_filt.d, hf, (size_t)n_filt * sizeof(mpeghb200_tns_filter)));
  CK(mpeghb200_h2d(ctx, b_tns_first.d, first, (size_t)(n_chain + 1) * sizeof(int32_t)));
  CK(mpeghb200_tns_dev(ctx, (float *)b_tns_spec.d, (const mpeghb200_tns_filter *)b_tns_filt.d,
                       (const int32_t *)b_tns_first.d, n_chain, NULL));
  CK(mpeghb200_d2h(ctx, b_tns_spec.h, b_tns_spec.d, fb));
  for (i = 0; i < n; i++)
    memcpy(rq[i]->p[0], (float *)b_tns_spec.h + (size_t)i * FRAME, FRAME * sizeof(float));
  seam_n.tns += (unsigned long)n_filt;
}

IA_ERRORCODE ia_core_coder_tns_apply(ia_usac_data_struct *usac_data, FLOAT32 *spec, WORD32 nbands,
                                     ia_sfb_info_struct *pstr_sfb_info, ia_tns_frame_info_struct *pstr_tns,
                                     ia_usac_igf_config_struct *igf_config)
{
  mpeghb200_tns_channel ch;
  mpeghb200_tns_filter filt[24];
  int32_t chain_len[8];
  int n_filters = 0, n_chains = 0, w, f, k;
  const int idx = pstr_sfb_info->islong ? 0 : 1;
  const int tns_max = (*usac_data->tns_max_bands_tbl_usac)[usac_data->sampling_rate_idx][idx];
  mpeghb200_status st;
  seam_req r;
  memset(&ch, 0, sizeof(ch));
  ch.unit = 0;
  ch.is_long = pstr_sfb_info->islong ? 1 : 0;
  ch.max_sfb = nbands;
  ch.igf_active = igf_config->igf_active;
  ch.igf_after_tns_synth = igf_config->igf_after_tns_synth;
  ch.igf_sfb_start = igf_config->igf_grid[idx].igf_sfb_start;
  ch.igf_sfb_stop = igf_config->igf_grid[idx].igf_sfb_stop;
  for (w = 0; w < pstr_tns->n_subblocks && w < 8; w++)
  {
    const ia_tns_info_struct *ti = &pstr_tns->str_tns_info[w];
    ch.n_filt[w] = ti->n_filt;
    ch.coef_res[w] = ti->coef_res;
    for (f = 0; f < ti->n_filt && f < 3; f++)
    {
      ch.filt[w][f].order = ti->str_filter[f].order;
      ch.filt[w][f].start_band = ti->str_filter[f].start_band;
      ch.filt[w][f].stop_band = ti->str_filter[f].stop_band;
      ch.filt[w][f].direction = ti->str_filter[f].direction;
      for (k = 0; k < 16 && k < TNS_MAX_ORDER; k++)
        ch.filt[w][f].coef[k] = ti->str_filter[f].coef[k];
    }
  }
  st = mpeghb200_tns_resolve(&ch, pstr_sfb_info->ptr_sfb_tbl, pstr_sfb_info->sfb_per_sbk, tns_max, filt, 24, chain_len, 8,
                             &n_filters, &n_chains);
  if (st != MPEGHB200_OK)
    return IA_MPEGH_DEC_EXE_FATAL_INVALID_TNS_SFB;
  if (n_filters == 0)
    return IA_MPEGH_DEC_NO_ERROR;
  r.kind = RQ_TNS;
  r.p[0] = spec, r.p[1] = filt, r.p[2] = chain_len;
  r.v[0] = n_filters, r.v[1] = n_chains;
  seam_submit(&r);
  return IA_MPEGH_DEC_NO_ERROR;
}

/* ---- MCT: one window of one box per call (the window loop is the caller's, process.c:1043-1075); computed inside
 * the interposer (not request-ised), one at a time ---- */
static struct
{
  float *d_spec;
  mpeghb200_mct_box *d_box;
  int32_t *d_first;
} g_mct;

VOID impeghd_mc_process(ia_multichannel_data *ptr_mc_data, FLOAT32 *ptr_dmx, FLOAT32 *ptr_res,
                        WORD32 *ptr_coeff_sfb_idx, WORD32 *ptr_mask, WORD32 bands_per_window, WORD32 total_sfb,
                        WORD32 pair, const WORD16 *ptr_sfb_offset, WORD32 num_samples)
{
  mpeghb200_mct_box box;
  const int32_t first[2] = {0, 1};
  const int fullband = !ptr_mc_data->bandwise_coeff_flag[pair] && !ptr_mc_data->mask_flag[pair];
  mpeghb200_ctx *ctx;
  if (!ptr_mc_data->use_tool)
    return;
  ctx = seam_ctx();
  CK(mpeghb200_mct_resolve(ptr_mc_data->signal_type, ptr_mc_data->pred_dir[pair], fullband, 1, num_samples,
                           bands_per_window, total_sfb, ptr_sfb_offset, ptr_coeff_sfb_idx, ptr_mask, 0, 1, &box));
  if (box.n_seg == 0)
    return;
  seam_inline_lock();
  if (!g_mct.d_spec)
  {
    CK(mpeghb200_dev_alloc(ctx, 2 * FRAME * sizeof(float), (void **)&g_mct.d_spec));
    CK(mpeghb200_dev_alloc(ctx, sizeof(mpeghb200_mct_box), (void **)&g_mct.d_box));
    CK(mpeghb200_dev_alloc(ctx, 2 * sizeof(int32_t), (void **)&g_mct.d_first));
  }
  CK(mpeghb200_h2d(ctx, g_mct.d_spec, ptr_dmx, (size_t)num_samples * sizeof(float)));
  CK(mpeghb200_h2d(ctx, g_mct.d_spec + FRAME, ptr_res, (size_t)num_samples * sizeof(float)));
  CK(mpeghb200_h2d(ctx, g_mct.d_box, &box, sizeof(box)));
  CK(mpeghb200_h2d(ctx, g_mct.d_first, first, sizeof(first)));
  CK(mpeghb200_mct_dev(ctx, g_mct.d_spec, g_mct.d_box, g_mct.d_first, 1, NULL));
  CK(mpeghb200_d2h(ctx, ptr_dmx, g_mct.d_spec, (size_t)num_samples * sizeof(float)));
  CK(mpeghb200_d2h(ctx, ptr_res, g_mct.d_spec + FRAME, (size_t)num_samples * sizeof(float)));
  seam_n.mct++;
  seam_inline_unlock();
}

/* ---- LTPF: the filter state stays in the reference handle (the LPD path shares it), round trip per call ---- */
static seam_buf b_ltpf_state, b_ltpf_side, b_ltpf_audio;

/* p[0] = ia_ltpf_data_str, p[1] = time_sample_vector */
static void ltpf_gather(int i, void *arg)
{
  seam_req **rq = (seam_req **)arg;
  const size_t ns = mpeghb200_ltpf_state_floats(48000);
  const ia_ltpf_data_str *l = (const ia_ltpf_data_str *)rq[i]->p[0];
  const int hist = l->pitch_max + LTPF_NUM_FILT_COEF1 / 2;
  float *st = (float *)b_ltpf_state.h + (size_t)i * ns;
  mpeghb200_ltpf_side *sd = (mpeghb200_ltpf_side *)b_ltpf_side.h + i;
  int32_t iv;
  memset(st, 0, ns * sizeof(float));
  iv = l->prev_pitch, memcpy(&st[0], &iv, 4);
  iv = l->prev_pitch_fr, memcpy(&st[1], &iv, 4);
  iv = l->prev_gain_idx, memcpy(&st[2], &iv, 4);
  st[3] = l->prev_gain;
  memcpy(&st[4], l->prev_in_buf, 6 * sizeof(float));
  memcpy(&st[16], l->prev_out_buf, (size_t)hist * sizeof(float));
  sd->unit = i;
  sd->data_present = (uint8_t)l->data_present;
  sd->pitch_lag_idx = (uint16_t)l->pitch_lag_idx;
  sd->gain_idx = (uint8_t)l->gain_idx;
  memcpy((float *)b_ltpf_audio.h + (size_t)i * FRAME, rq[i]->p[1], FRAME * sizeof(float));
}
static void ltpf_scatter(int i, void *arg)
{
  seam_req **rq = (seam_req **)arg;
  const size_t ns = mpeghb200_ltpf_state_floats(48000);
  ia_ltpf_data_str *l = (ia_ltpf_data_str *)rq[i]->p[0];
  const int hist = l->pitch_max + LTPF_NUM_FILT_COEF1 / 2;
  const float *st = (const float *)b_ltpf_state.h + (size_t)i * ns;
  int32_t iv;
  memcpy(&iv, &st[0], 4), l->prev_pitch = iv;
  memcpy(&iv, &st[1], 4), l->prev_pitch_fr = iv;
  memcpy(&iv, &st[2], 4), l->prev_gain_idx = iv;
  l->prev_gain = st[3];
  memcpy(l->prev_in_buf, &st[4], 6 * sizeof(float));
  memcpy(l->prev_out_buf, &st[16], (size_t)hist * sizeof(float));
  memcpy(rq[i]->p[1], (float *)b_ltpf_audio.h + (size_t)i * FRAME, FRAME * sizeof(float));
}

static void be_ltpf(seam_req **rq, int n)
{
  mpeghb200_ctx *ctx = seam_ctx();
  const size_t ns = mpeghb200_ltpf_state_floats(48000);
  seam_need(&b_ltpf_state, (size_t)n * ns * sizeof(float));
  seam_need(&b_ltpf_side, (size_t)n * sizeof(mpeghb200_ltpf_side));
  seam_need(&b_ltpf_audio, (size_t)n * FRAME * sizeof(float));
  seam_parallel_for(n, ltpf_gather, rq);
  CK(mpeghb200_h2d(ctx, b_ltpf_state.d, b_ltpf_state.h, (size_t)n * ns * sizeof(float)));
  CK(mpeghb200_h2d(ctx, b_ltpf_side.d, b_ltpf_side.h, (size_t)n * sizeof(mpeghb200_ltpf_side)));
  CK(mpeghb200_h2d(ctx, b_ltpf_audio.d, b_ltpf_audio.h, (size_t)n * FRAME * sizeof(float)));
  CK(mpeghb200_ltpf_dev(ctx, 48000, (float *)b_ltpf_audio.d, (const mpeghb200_ltpf_side *)b_ltpf_side.d,
                        (float *)b_ltpf_state.d, n, NULL));
  CK(mpeghb200_d2h(ctx, b_ltpf_audio.h, b_ltpf_audio.d, (size_t)n * FRAME * sizeof(float)));
  CK(mpeghb200_d2h(ctx, b_ltpf_state.h, b_ltpf_state.d, (size_t)n * ns * sizeof(float)));
  seam_parallel_for(n, ltpf_scatter, rq);
  seam_n.ltpf += (unsigned long)n;
}

VOID ia_core_coder_usac_ltpf_process(ia_ltpf_data_str *ptr_ltpf_data, FLOAT32 *time_sample_vector)
{
  typedef VOID (*fn_t)(ia_ltpf_data_str *, FLOAT32 *);
  seam_req r;
  if (ptr_ltpf_data->frame_len != FRAME || ptr_ltpf_data->pitch_min != 128)
  { /* not a 48 kHz / 1024-sample FD channel: the LPD path's own use of the filter stays on the host */
    static fn_t real;
    if (!real)
      real = (fn_t)dlsym(RTLD_NEXT, "ia_core_coder_usac_ltpf_process");
    real(ptr_ltpf_data, time_sample_vector);
    return;
  }
  r.kind = RQ_LTPF;
  r.p[0] = ptr_ltpf_data, r.p[1] = time_sample_vector;
  seam_submit(&r);
}

/* ---- peak limiter: device-resident state per ia_peak_limiter_struct ------------------------------------ */
static seam_pool lim_pool[MAX_NUM_CHANNELS + 1]; /* by channel count (the slot size depends on it) */
static seam_buf b_lim_in, b_lim_out;
#define MAX_DIRTY 4096
static const void *lim_dirty[MAX_DIRTY]; /* structs the reference (re-)initialised since their slot was last used */
static int n_lim_dirty;
static pthread_mutex_t lim_mu = PTHREAD_MUTEX_INITIALIZER;

static int lim_take_dirty(const void *key)
{
  int i, hit = 0;
  pthread_mutex_lock(&lim_mu);
  for (i = 0; i < n_lim_dirty; i++)
    if (lim_dirty[i] == key)
    {
      lim_dirty[i] = lim_dirty[--n_lim_dirty];
      hit = 1;
      break;
    }
  pthread_mutex_unlock(&lim_mu);
  return hit;
}

IA_ERRORCODE impeghd_peak_limiter_init(ia_peak_limiter_struct *pstr_peak_limiter, UWORD32 channel_num,
                                       UWORD32 sample_rate, FLOAT32 *ptr_buff)
{
  typedef IA_ERRORCODE (*fn_t)(ia_peak_limiter_struct *, UWORD32, UWORD32, FLOAT32 *);
  static fn_t real;
  int i, known = 0;
  if (!real)
    real = (fn_t)dlsym(RTLD_NEXT, "impeghd_peak_limiter_init");
  /* the reference (re-)initialised this instance: its device state restarts too */
  pthread_mutex_lock(&lim_mu);
  for (i = 0; i < n_lim_dirty; i++)
    known |= lim_dirty[i] == pstr_peak_limiter;
  if (!known && n_lim_dirty < MAX_DIRTY)
    lim_dirty[n_lim_dirty++] = pstr_peak_limiter;
  pthread_mutex_unlock(&lim_mu);
  return real(pstr_peak_limiter, channel_num, sample_rate, ptr_buff);
}

typedef struct
{
  seam_req **rq;
  size_t bytes;
} lim_arg;
static void lim_gather(int i, void *arg)
{
  const lim_arg *a = (const lim_arg *)arg;
  memcpy((char *)b_lim_in.h + (size_t)i * a->bytes, a->rq[i]->p[1], a->bytes);
}
static void lim_scatter(int i, void *arg)
{
  const lim_arg *a = (const lim_arg *)arg;
  memcpy(a->rq[i]->p[1], (const char *)b_lim_out.h + (size_t)i * a->bytes, a->bytes);
}

/* one group: n requests with the same channel count / rate / mode.  p[0] = ia_peak_limiter_struct, p[1] = samples
 * (interleaved [1024][n_ch], in place) */
static void lim_group(seam_req **rq, int n)
{
  mpeghb200_ctx *ctx = seam_ctx();
  const ia_peak_limiter_struct *l0 = (const ia_peak_limiter_struct *)rq[0]->p[0];
  const int n_ch = (int)l0->num_channels;
  const size_t fb = (size_t)n * n_ch * FRAME * sizeof(float);
  mpeghb200_limiter_params prm;
  seam_pool *pool = &lim_pool[n_ch];
  const void **keys = (const void **)malloc(sizeof(void *) * (size_t)n);
  float *base;
  size_t stride;
  lim_arg la;
  int i, fresh = 0;
  CK(mpeghb200_limiter_params_init(&prm, n_ch, (int)l0->sample_rate, (int)l0->limiter_on));
  stride = mpeghb200_limiter_state_floats(&prm);
  pool->stride = stride * sizeof(float);
  for (i = 0; i < n; i++)
    keys[i] = rq[i]->p[0];
  seam_need(&b_lim_in, fb), seam_need(&b_lim_out, fb);
  /* the caller hands over interleaved [sample][channel] (impegh_dec_peak_limiter); the kernel is planar: the frames
   * travel as they are and are turned on the device */
  la.rq = rq, la.bytes = (size_t)n_ch * FRAME * sizeof(float);
  seam_parallel_for(n, lim_gather, &la);
  CK(mpeghb200_h2d(ctx, b_lim_in.d, b_lim_in.h, fb));
  CK(mpeghb200_interleave_dev(ctx, (const float *)b_lim_in.d, (float *)b_lim_out.d, n_ch, FRAME, 1, n, NULL));
  base = (float *)seam_pool_range(pool, keys, n, &fresh);
  if (base)
  {
    if (fresh)
      CK(mpeghb200_limiter_state_init_dev(ctx, &prm, base, n, NULL));
    for (i = 0; i < n; i++)
      if (lim_take_dirty(keys[i]) && !fresh)
        CK(mpeghb200_limiter_state_init_dev(ctx, &prm, base + (size_t)i * stride, 1, NULL));
    CK(mpeghb200_limiter_dev(ctx, &prm, (const float *)b_lim_out.d, (float *)b_lim_in.d, base, NULL, n, NULL));
  }
  else
  { /* the batch does not line up with the slot order: one launch per handle, each on its own slot */
    for (i = 0; i < n; i++)
    {
      float *st = (float *)seam_pool_one(pool, keys[i], &fresh);
      if (lim_take_dirty(keys[i]) || fresh)
        CK(mpeghb200_limiter_state_init_dev(ctx, &prm, st, 1, NULL));
      CK(mpeghb200_limiter_dev(ctx, &prm, (const float *)b_lim_out.d + (size_t)i * n_ch * FRAME,
                               (float *)b_lim_in.d + (size_t)i * n_ch * FRAME, st, NULL, 1, NULL));
    }
  }
  CK(mpeghb200_interleave_dev(ctx, (const float *)b_lim_in.d, (float *)b_lim_out.d, n_ch, FRAME, 0, n, NULL));
  CK(mpeghb200_d2h(ctx, b_lim_out.h, b_lim_out.d, fb));
  seam_parallel_for(n, lim_scatter, &la);
  free(keys);
  seam_n.lim += (unsigned long)n;
}

static void be_lim(seam_req **rq, int n)
{
  /* split into runs of equal (channel count, rate, mode); a batch of like streams is a single run */
  int a = 0, b;
  while (a < n)
  {
    const ia_peak_limiter_struct *la = (const ia_peak_limiter_struct *)rq[a]->p[0];
    for (b = a + 1; b < n; b++)
    {
      const ia_peak_limiter_struct *lb = (const ia_peak_limiter_struct *)rq[b]->p[0];
      if (lb->num_channels != la->num_channels || lb->sample_rate != la->sample_rate || lb->limiter_on != la->limiter_on)
        break;
    }
    lim_group(rq + a, b - a);
    a = b;
  }
}

VOID impeghd_peak_limiter_process(ia_peak_limiter_struct *pstr_peak_limiter, FLOAT32 *ptr_samples,
                                  UWORD32 frame_size)
{
  const int n_ch = (int)pstr_peak_limiter->num_channels;
  seam_req r;
  if (frame_size != FRAME || n_ch < 1 || n_ch > MAX_NUM_CHANNELS)
  {
    fprintf(stderr, "mpeghb200 seam: limiter frame of %u samples x %d channels is outside the GPU path\n",
            (unsigned)frame_size, n_ch);
    abort();
  }
  r.kind = RQ_LIM;
  r.p[0] = pstr_peak_limiter, r.p[1] = ptr_samples;
  seam_submit(&r);
}

__attribute__((constructor)) static void seam_register_core(void)
{
  seam_register(RQ_FD, be_fd, "fd_frm_dec");
  seam_register(RQ_TNS, be_tns, "tns_apply");
  seam_register(RQ_LTPF, be_ltpf, "ltpf");
  seam_register(RQ_LIM, be_lim, "limiter");
}
