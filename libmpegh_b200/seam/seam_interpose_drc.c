/* seam_interpose_drc.c -- libmpeghgpu.so: application of the decoded DRC gains.
 *
 *   impeghd_domain_switcher_process decoder/ia_core_coder_process.c:505 -> mpeghb200_stft_dev        (time <-> STFT256)
 *   impd_drc_apply_gains_subband    decoder/impd_drc_process.c:243   -> mpeghb200_gain_subband_dev   (STFT256 domain)
 *   impd_drc_td_process             decoder/impd_drc_process.c:529   -> mpeghb200_drc_td_dev         (time domain, the whole step)
 *   impd_drc_filter_banks_process   decoder/impd_drc_process.c:399  \
 *   impd_drc_apply_gains_and_add    decoder/impd_drc_process.c:68   /  -> the same, for callers of the two halves
 *
 * Gain decoding, DRC set selection and the interpolation into per-sample gain curves (impd_drc_get_gain and below: libm
 * splines) stay in the reference; these interposers read the finished curves out of ia_drc_gain_buffer_struct the way
 * the reference's own application code does (lpcm_gains + MAX_SIGNAL_DELAY - gain_delay_samples [+ one frame in
 * DELAY_MODE_LOW]), the channel groups out of ia_drc_instructions_struct, the overlap weights out of
 * ia_drc_overlap_params_struct and the filter-bank coefficients out of ia_filter_banks_struct.
 *
 * Time domain: the reference splits a frame into band signals (filter_banks_process) and then multiplies and sums them
 * (apply_gains_and_add); the kernel does both at once.  impd_drc_td_process, the function the decoder calls, is taken
 * over as a whole: the reference's impd_drc_get_gain for every selected set, then one request per set -- this works
 * against the stock reference library.  The two halves are interposed as well for callers that use them directly: the
 * first call only records its arguments, the second one files the request; impd_drc_apply_gains_and_add is `static` in
 * the reference, so that pair is reachable only when the reference library in use exports it (the build oracle/Makefile
 * makes with -Dstatic= on the one file) and is forwarded otherwise.  The filter memories live on the device, one slot
 * per (ia_drc_gain_dec_struct, selected set).
 */
#define _GNU_SOURCE
#include <stdlib.h>
#include <string.h>

#include <impeghd_type_def.h>
#include "impd_drc_common.h"
#include "impd_drc_extr_delta_coded_info.h"
#include "impd_drc_struct.h"
#include "impeghd_intrinsics_flt.h"
#include "ia_core_coder_cnst.h"
#include "ia_core_coder_defines.h"
#include "ia_core_coder_bitbuffer.h"
#include "ia_core_coder_interface.h"
#include "ia_core_coder_info.h"
#include <ia_core_coder_rom.h>
#include "ia_core_coder_dec.h"
#include "ia_core_coder_acelp_info.h"
#include "ia_core_coder_channelinfo.h"
#include "ia_core_coder_channel.h"
#include "ia_core_coder_env_extr.h"
#include "ia_core_coder_igf_data_struct.h"
#include "impeghd_memory_standards.h"
#include "ia_core_coder_stereo_lpd.h"
#include "ia_core_coder_tns_usac.h"
#include "ia_core_coder_windows.h"
#include "impeghd_hoa_common_values.h"
#include "impeghd_hoa_matrix_struct.h"
#include "impeghd_hoa_matrix.h"
#include "ia_core_coder_config.h"
#include "ia_core_coder_struct_def.h"
#include "impeghd_hoa_dec_struct.h"
#include "impeghd_error_codes.h"
#include <impeghd_fft_ifft.h>
#include "impeghd_mhas_parse.h"
#include "impeghd_multichannel.h"
#include "impeghd_tbe_dec.h"
#include "impeghd_binaural.h"
#include "impeghd_ele_interaction_intrfc.h"
#include "impeghd_hoa_data_types.h"
#include "impeghd_hoa_frame_params.h"
#include "impeghd_hoa_nfc_filtering.h"
#include "impeghd_hoa_space_positions.h"
#include "impeghd_hoa_simple_mtrx.h"
#include "impeghd_hoa_render_mtrx.h"
#include "impeghd_hoa_renderer.h"
#include "impeghd_hoa_spatial_decoder_struct.h"
#include "impeghd_hoa_spatial_decoder.h"
#include "impeghd_hoa_decoder.h"
#include "impeghd_3d_vec_struct_def.h"
#include "impeghd_cicp_defines.h"
#include "impeghd_cicp_struct_def.h"
#include "impeghd_oam_dec_defines.h"
#include "impeghd_oam_dec_struct_def.h"
#include "impeghd_obj_ren_dec_defines.h"
#include "impeghd_obj_ren_dec_struct_def.h"
#include "impeghd_uni_drc_struct.h"
#include "ia_core_coder_main.h"
#include "ia_core_coder_arith_dec.h"
#include "ia_core_coder_bit_extract.h"
#include "ia_core_coder_func_def.h"
#include "impeghd_binaural_renderer.h"
#include "impeghd_format_conv_defines.h"
#include "impeghd_format_conv_rom.h"
#include "impeghd_format_conv_data.h"
#include "impeghd_peak_limiter_struct_def.h"
#include "impeghd_resampler.h"
#include "ia_core_coder_struct.h"
#include "ia_core_coder_create.h"
#include "impeghd_cicp_2_geometry_rom.h"
#include "impeghd_cicp_2_geometry.h"
#include "impeghd_obj_ren_dec.h"


#include <dlfcn.h>
#include <stdio.h>
#include "impd_drc_filter_bank.h"
#include "impd_drc_gain_dec.h"
#include "impd_drc_multi_band.h"
#include "impd_drc_process_audio.h"
#include "seam_internal.h"

#define FRAME 1024
#define SB 256
#define SLOTS 4
#define MAX_SEQ (CHANNEL_GROUP_COUNT_MAX * BAND_COUNT_MAX)

static int gain_offset(const ia_drc_params_struct *p)
{
  return MAX_SIGNAL_DELAY - p->gain_delay_samples + (p->delay_mode == DELAY_MODE_LOW ? p->drc_frame_size : 0);
}

/* ---- sub-band gains ---------------------------------------------------------------------------------------------- */
typedef struct
{
  int n_ch, n_seq, n_rows;
  FLOAT32 **re, **im; /* rows sig_idx = 0 .. n_ch - 1 */
  mpeghb200_drc_sb_channel chan[MAX_CHANNEL_COUNT];
  float gains[MAX_SEQ][SLOTS];
  const float *wrow[MAX_SEQ]; /* overlap_weight[256] of every (multi-band group, band) */
} sb_job;
static seam_buf b_sb_re, b_sb_im, b_sb_chan, b_sb_gain, b_sb_w;

static void be_drc_sb(seam_req **rq, int n)
{
  mpeghb200_ctx *ctx = seam_ctx();
  int N = 0, Q = 0, R = 0, i, c, k, c0 = 0, q0 = 0, r0 = 0;
  for (i = 0; i < n; i++)
  {
    const sb_job *j = (const sb_job *)rq[i]->p[0];
    N += j->n_ch, Q += j->n_seq, R += j->n_rows;
  }
  seam_need(&b_sb_re, (size_t)N * FRAME * sizeof(float)), seam_need(&b_sb_im, (size_t)N * FRAME * sizeof(float));
  seam_need(&b_sb_chan, (size_t)N * sizeof(mpeghb200_drc_sb_channel));
  seam_need(&b_sb_gain, (size_t)(Q ? Q : 1) * SLOTS * sizeof(float)), seam_need(&b_sb_w, (size_t)(R ? R : 1) * SB * sizeof(float));
  for (i = 0; i < n; i++)
  {
    const sb_job *j = (const sb_job *)rq[i]->p[0];
    for (c = 0; c < j->n_ch; c++)
    {
      mpeghb200_drc_sb_channel *d = (mpeghb200_drc_sb_channel *)b_sb_chan.h + c0 + c;
      *d = j->chan[c];
      if (d->first_seq >= 0)
        d->first_seq += q0, d->weight_row += r0;
      memcpy((float *)b_sb_re.h + (size_t)(c0 + c) * FRAME, j->re[c], FRAME * sizeof(float));
      memcpy((float *)b_sb_im.h + (size_t)(c0 + c) * FRAME, j->im[c], FRAME * sizeof(float));
    }
    memcpy((float *)b_sb_gain.h + (size_t)q0 * SLOTS, j->gains, (size_t)j->n_seq * SLOTS * sizeof(float));
    for (k = 0; k < j->n_rows; k++)
      memcpy((float *)b_sb_w.h + (size_t)(r0 + k) * SB, j->wrow[k], SB * sizeof(float));
    c0 += j->n_ch, q0 += j->n_seq, r0 += j->n_rows;
  }
  CK(mpeghb200_h2d(ctx, b_sb_re.d, b_sb_re.h, (size_t)N * FRAME * sizeof(float)));
  CK(mpeghb200_h2d(ctx, b_sb_im.d, b_sb_im.h, (size_t)N * FRAME * sizeof(float)));
  CK(mpeghb200_h2d(ctx, b_sb_chan.d, b_sb_chan.h, (size_t)N * sizeof(mpeghb200_drc_sb_channel)));
  CK(mpeghb200_h2d(ctx, b_sb_gain.d, b_sb_gain.h, (size_t)(Q ? Q : 1) * SLOTS * sizeof(float)));
  CK(mpeghb200_h2d(ctx, b_sb_w.d, b_sb_w.h, (size_t)(R ? R : 1) * SB * sizeof(float)));
  CK(mpeghb200_gain_subband_dev(ctx, (float *)b_sb_re.d, (float *)b_sb_im.d, (const mpeghb200_drc_sb_channel *)b_sb_chan.d,
                                (const float *)b_sb_gain.d, (const float *)b_sb_w.d, N, NULL));
  CK(mpeghb200_d2h(ctx, b_sb_re.h, b_sb_re.d, (size_t)N * FRAME * sizeof(float)));
  CK(mpeghb200_d2h(ctx, b_sb_im.h, b_sb_im.d, (size_t)N * FRAME * sizeof(float)));
  for (i = 0, c0 = 0; i < n; i++)
  {
    const sb_job *j = (const sb_job *)rq[i]->p[0];
    for (c = 0; c < j->n_ch; c++)
    {
      memcpy(j->re[c], (float *)b_sb_re.h + (size_t)(c0 + c) * FRAME, FRAME * sizeof(float));
      memcpy(j->im[c], (float *)b_sb_im.h + (size_t)(c0 + c) * FRAME, FRAME * sizeof(float));
    }
    c0 += j->n_ch;
  }
  seam_n.drc_sb += (unsigned long)N;
}

IA_ERRORCODE impd_drc_apply_gains_subband(FLOAT32 *deinterleaved_audio_re[], FLOAT32 *deinterleaved_audio_im[],
                                          ia_drc_instructions_struct *pstr_drc_instruction_arr,
                                          ia_drc_gain_dec_struct *pstr_drc_gain_dec, WORD32 sel_drc_idx)
{
  ia_drc_params_struct *p = &pstr_drc_gain_dec->ia_drc_params_struct;
  const ia_drc_gain_buffer_struct *gb = &pstr_drc_gain_dec->drc_gain_buffers.pstr_gain_buf[sel_drc_idx];
  const WORD32 ii = p->sel_drc_array[sel_drc_idx].drc_instrns_idx;
  const ia_drc_instructions_struct *ins;
  int first[CHANNEL_GROUP_COUNT_MAX], row[CHANNEL_GROUP_COUNT_MAX], g, b, c, s, off;
  sb_job j;
  seam_req r;
  if (p->sub_band_domain_mode != SUBBAND_DOMAIN_MODE_STFT256 || p->drc_frame_size != FRAME ||
      (ii >= 0 && pstr_drc_instruction_arr[ii].num_drc_ch_groups > CHANNEL_GROUP_COUNT_MAX))
  { /* the QMF domains are not part of the decoder's signal path (the domain switcher is STFT256): forwarded */
    typedef IA_ERRORCODE (*fn_t)(FLOAT32 **, FLOAT32 **, ia_drc_instructions_struct *, ia_drc_gain_dec_struct *, WORD32);
    static fn_t real;
    if (!real)
      real = (fn_t)seam_real("impd_drc_apply_gains_subband");
    seam_n.drc_host++;
    return real(deinterleaved_audio_re, deinterleaved_audio_im, pstr_drc_instruction_arr, pstr_drc_gain_dec, sel_drc_idx);
  }
  if (ii < 0)
    return IA_MPEGH_DEC_NO_ERROR;
  ins = &pstr_drc_instruction_arr[ii];
  if (ins->drc_set_id <= 0)
    return IA_MPEGH_DEC_NO_ERROR;
  for (g = 0; g < ins->num_drc_ch_groups; g++)
    if (ins->band_count_of_ch_group[g] < 1 || ins->band_count_of_ch_group[g] > BAND_COUNT_MAX)
    { /* not a band count the parser produces: leave it to the reference */
      typedef IA_ERRORCODE (*fn_t)(FLOAT32 **, FLOAT32 **, ia_drc_instructions_struct *, ia_drc_gain_dec_struct *, WORD32);
      seam_n.drc_host++;
      return ((fn_t)seam_real("impd_drc_apply_gains_subband"))(deinterleaved_audio_re, deinterleaved_audio_im,
                                                                       pstr_drc_instruction_arr, pstr_drc_gain_dec,
                                                                       sel_drc_idx);
    }
  if (p->num_ch_process == -1) /* :318-325 */
    p->num_ch_process = ins->audio_num_chan;
  if (ins->audio_num_chan < p->channel_offset + p->num_ch_process)
    return IA_MPEGD_DRC_INIT_FATAL_UNSUPPORTED_CH_COUNT;
  off = gain_offset(p);
  memset(&j, 0, sizeof(j));
  for (g = 0; g < ins->num_drc_ch_groups; g++)
  {
    const int nb = ins->band_count_of_ch_group[g] < 1 ? 1 : ins->band_count_of_ch_group[g];
    first[g] = j.n_seq, row[g] = j.n_rows;
    for (b = 0; b < nb; b++)
    {
      const float *l = gb->pstr_buf_interp[first[g] + b].lpcm_gains + off;
      for (s = 0; s < SLOTS; s++)
        j.gains[j.n_seq][s] = l[s * SB + (SB - 1) / 2]; /* the slot centre (:328, :345) */
      j.n_seq++;
      if (nb > 1)
        j.wrow[j.n_rows++] =
            pstr_drc_gain_dec->str_overlap_params.str_grp_overlap_params[g].str_band_overlap_params[b].overlap_weight;
    }
  }
  j.n_ch = p->num_ch_process;
  j.re = deinterleaved_audio_re, j.im = deinterleaved_audio_im;
  for (c = 0; c < j.n_ch; c++)
  {
    g = ins->channel_group_of_ch[p->channel_offset + c];
    j.chan[c].first_seq = g >= 0 ? first[g] : -1;
    j.chan[c].band_count = g >= 0 ? ins->band_count_of_ch_group[g] : 0;
    j.chan[c].weight_row = g >= 0 ? row[g] : 0;
  }
  r.kind = RQ_DRCSB;
  r.p[0] = &j;
  seam_submit(&r);
  return IA_MPEGH_DEC_NO_ERROR;
}

/* ---- time domain: filter banks + gains + band sum ---------------------------------------------------------------- */
typedef struct
{
  int n_ch, n_seq, n_banks;
  const void *key;
  FLOAT32 **audio; /* rows of the frame, channel_offset already taken off */
  mpeghb200_drc_td_channel chan[MAX_CHANNEL_COUNT];
  mpeghb200_drc_td_bank bank[10];
  const float *gain[MAX_SEQ]; /* 1024 samples each */
} td_job;
static seam_buf b_td_x, b_td_chan, b_td_bank, b_td_gain;
static seam_pool td_pool[MAX_CHANNEL_COUNT + 1]; /* by channel count of the request */

static void td_run(td_job *j, float *d_state)
{
  mpeghb200_ctx *ctx = seam_ctx();
  int c, k;
  seam_need(&b_td_x, (size_t)j->n_ch * FRAME * sizeof(float)), seam_need(&b_td_chan, sizeof(j->chan));
  seam_need(&b_td_bank, sizeof(j->bank)), seam_need(&b_td_gain, (size_t)(j->n_seq ? j->n_seq : 1) * FRAME * sizeof(float));
  for (c = 0; c < j->n_ch; c++)
    memcpy((float *)b_td_x.h + (size_t)c * FRAME, j->audio[c], FRAME * sizeof(float));
  for (k = 0; k < j->n_seq; k++)
    memcpy((float *)b_td_gain.h + (size_t)k * FRAME, j->gain[k], FRAME * sizeof(float));
  memcpy(b_td_chan.h, j->chan, sizeof(j->chan)), memcpy(b_td_bank.h, j->bank, sizeof(j->bank));
  CK(mpeghb200_h2d(ctx, b_td_x.d, b_td_x.h, (size_t)j->n_ch * FRAME * sizeof(float)));
  CK(mpeghb200_h2d(ctx, b_td_chan.d, b_td_chan.h, sizeof(j->chan)));
  CK(mpeghb200_h2d(ctx, b_td_bank.d, b_td_bank.h, sizeof(j->bank)));
  CK(mpeghb200_h2d(ctx, b_td_gain.d, b_td_gain.h, (size_t)(j->n_seq ? j->n_seq : 1) * FRAME * sizeof(float)));
  CK(mpeghb200_drc_td_dev(ctx, (float *)b_td_x.d, d_state, (const mpeghb200_drc_td_channel *)b_td_chan.d,
                          (const mpeghb200_drc_td_bank *)b_td_bank.d, (const float *)b_td_gain.d, j->n_ch, NULL));
  CK(mpeghb200_d2h(ctx, b_td_x.h, b_td_x.d, (size_t)j->n_ch * FRAME * sizeof(float)));
  for (c = 0; c < j->n_ch; c++)
    memcpy(j->audio[c], (float *)b_td_x.h + (size_t)c * FRAME, FRAME * sizeof(float));
}

/* the banks differ from handle to handle (they are designed from the stream's DRC configuration), so a request is
 * its own launch; the filter memories of its channels sit in a device slot of their own */
static void be_drc_td(seam_req **rq, int n)
{
  int i, fresh;
  for (i = 0; i < n; i++)
  {
    td_job *j = (td_job *)rq[i]->p[0];
    seam_pool *pool = &td_pool[j->n_ch];
    float *st;
    pool->stride = (size_t)j->n_ch * mpeghb200_drc_td_state_floats() * sizeof(float);
    fresh = 0;
    st = (float *)seam_pool_one(pool, j->key, &fresh);
    if (fresh)
      CK(mpeghb200_dev_memset(seam_ctx(), st, 0, pool->stride));
    td_run(j, st);
    seam_n.drc_td += (unsigned long)j->n_ch;
  }
}

static void set_f(mpeghb200_drc_biquad *d, const ia_iir_filter_struct *s)
{
  d->a1 = s->a1, d->a2 = s->a2, d->b0 = s->b0, d->b1 = s->b1, d->b2 = s->b2;
}
static void bank_of(mpeghb200_drc_td_bank *d, const ia_drc_filter_bank_struct *b, int n_bands)
{
  int i;
  memset(d, 0, sizeof(*d));
  d->n_bands = n_bands;
  d->n_cascade = b->str_all_pass_cascade.num_filter > 0 ? b->str_all_pass_cascade.num_filter + 1 : 0; /* filter_bank.c:604 */
  if (d->n_cascade > CASCADE_ALLPASS_COUNT_MAX)
    d->n_cascade = CASCADE_ALLPASS_COUNT_MAX;
  for (i = 0; i < d->n_cascade; i++)
    set_f(&d->ap[i], &b->str_all_pass_cascade.str_ap_cascade_filt[i].str_ap_stage);
  if (n_bands == 2)
    set_f(&d->f[0], &b->str_two_band_filt_bank.str_lp), set_f(&d->f[1], &b->str_two_band_filt_bank.str_hp);
  else if (n_bands == 3)
  {
    const ia_three_band_filt_struct *t = &b->str_three_band_filt_bank;
    set_f(&d->f[0], &t->str_lp_stage_1), set_f(&d->f[1], &t->str_hp_stage_1), set_f(&d->f[2], &t->str_ap_stage_2);
    set_f(&d->f[3], &t->str_lp_stage_2), set_f(&d->f[4], &t->str_hp_stage_2);
  }
  else if (n_bands == 4)
  {
    const ia_four_band_filt_struct *t = &b->str_four_band_filt_bank;
    set_f(&d->f[0], &t->str_lp_stage_1), set_f(&d->f[1], &t->str_hp_stage_1);
    set_f(&d->f[2], &t->str_ap_stage_2_low), set_f(&d->f[3], &t->str_ap_stage_2_high);
    set_f(&d->f[4], &t->str_lp_stage_3_low), set_f(&d->f[5], &t->str_hp_stage_3_low);
    set_f(&d->f[6], &t->str_lp_stage_3_high), set_f(&d->f[7], &t->str_hp_stage_3_high);
  }
}

typedef VOID (*add_fn)(ia_drc_gain_dec_struct *, FLOAT32 **, ia_drc_instructions_struct *, WORD32, WORD32);
typedef IA_ERRORCODE (*banks_fn)(FLOAT32 **, ia_drc_instructions_struct *, ia_drc_gain_dec_struct *, WORD32);
static add_fn real_add(void)
{
  static add_fn f;
  static int looked;
  if (!looked)
    f = (add_fn)dlsym(RTLD_NEXT, "impd_drc_apply_gains_and_add"), looked = 1;
  return f; /* NULL: `static` in the reference library in use */
}
/* MPEGHB200_SEAM_DRC_TD_HOST=1 leaves the pair in the reference even where it could be taken over (tests) */
static int td_enabled(void)
{
  static int on = -1;
  if (on < 0)
    on = real_add() != NULL && !getenv("MPEGHB200_SEAM_DRC_TD_HOST");
  return on;
}

/* what the pending filter_banks_process call of this thread was given (no yield between the two calls) */
static __thread FLOAT32 **tl_audio;
static __thread const ia_drc_gain_dec_struct *tl_dec;
static __thread int tl_sel = -1;

/* a selected set the kernel covers: a positive set id, at most 8 channel groups of 1 .. 4 bands, and only the multi-band
 * set may split (the others are gain-only, :417-424) */
static int td_shape_ok(const ia_drc_instructions_struct *arr, const ia_drc_gain_dec_struct *dec, int sel, int n_ch)
{
  const ia_drc_params_struct *p = &dec->ia_drc_params_struct;
  const int ii = p->sel_drc_array[sel].drc_instrns_idx;
  const ia_drc_instructions_struct *ins;
  int g;
  if (ii < 0 || p->drc_frame_size != FRAME || n_ch < 1 || n_ch > MAX_CHANNEL_COUNT)
    return 0;
  ins = &arr[ii];
  if (ins->drc_set_id <= 0 || ins->num_drc_ch_groups < 0 || ins->num_drc_ch_groups > 8 ||
      ins->audio_num_chan < p->channel_offset + n_ch)
    return 0;
  for (g = 0; g < ins->num_drc_ch_groups; g++)
  {
    const int nb = ins->band_count_of_ch_group[g];
    if (nb < 1 || nb > 4 || (p->multiband_sel_drc_idx != sel && nb != 1))
      return 0;
  }
  return 1;
}
static int td_on_gpu(const ia_drc_instructions_struct *arr, const ia_drc_gain_dec_struct *dec, int sel)
{
  return td_enabled() && td_shape_ok(arr, dec, sel, dec->ia_drc_params_struct.num_ch_process);
}

/* the request of one selected set: channel records, banks, gain curves */
static void td_build(td_job *j, ia_drc_gain_dec_struct *dec, ia_drc_instructions_struct *arr, int sel, FLOAT32 **audio)
{
  const ia_drc_params_struct *p = &dec->ia_drc_params_struct;
  const ia_drc_gain_buffer_struct *gb = &dec->drc_gain_buffers.pstr_gain_buf[sel];
  const ia_drc_instructions_struct *ins = &arr[p->sel_drc_array[sel].drc_instrns_idx];
  const int pass = p->multiband_sel_drc_idx != sel, off = gain_offset(p);
  int first[8], g, b, c;
  memset(j, 0, sizeof(*j));
  j->n_ch = p->num_ch_process, j->audio = audio;
  j->key = (const char *)dec + sel;
  for (g = 0; g < ins->num_drc_ch_groups; g++)
  {
    const int nb = ins->band_count_of_ch_group[g];
    first[g] = j->n_seq;
    for (b = 0; b < nb; b++)
      j->gain[j->n_seq++] = gb->pstr_buf_interp[first[g] + b].lpcm_gains + off;
    if (pass) /* a set that is not the multi-band one: no split, no phase alignment (:417-424) */
      j->bank[g].n_bands = 1;
    else
      bank_of(&j->bank[g], &dec->ia_filter_banks_struct.str_drc_filter_bank[g], nb);
  }
  g = ins->num_drc_ch_groups; /* the bank of the channels outside every group: its cascade only (:452-457) */
  if (pass)
    j->bank[g].n_bands = 1;
  else
    bank_of(&j->bank[g], &dec->ia_filter_banks_struct.str_drc_filter_bank[g], 1);
  j->n_banks = g + 1;
  for (c = 0; c < j->n_ch; c++)
  {
    g = ins->channel_group_of_ch[p->channel_offset + c];
    j->chan[c].bank = g >= 0 ? g : ins->num_drc_ch_groups;
    j->chan[c].first_seq = g >= 0 ? first[g] : -1;
  }
}

IA_ERRORCODE impd_drc_filter_banks_process(FLOAT32 *audio_io_buf[], ia_drc_instructions_struct *pstr_drc_instruction_arr,
                                           ia_drc_gain_dec_struct *pstr_drc_gain_dec, WORD32 sel_drc_idx)
{
  if (!td_on_gpu(pstr_drc_instruction_arr, pstr_drc_gain_dec, sel_drc_idx))
  {
    static banks_fn real;
    if (!real)
      real = (banks_fn)seam_real("impd_drc_filter_banks_process");
    tl_sel = -1;
    seam_n.drc_host++;
    return real(audio_io_buf, pstr_drc_instruction_arr, pstr_drc_gain_dec, sel_drc_idx);
  }
  tl_audio = audio_io_buf, tl_dec = pstr_drc_gain_dec, tl_sel = sel_drc_idx;
  return IA_MPEGH_DEC_NO_ERROR;
}

/* `static` in the stock reference (see the file header) */
VOID impd_drc_apply_gains_and_add(ia_drc_gain_dec_struct *pstr_drc_gain_dec, FLOAT32 *channel_audio[],
                                  ia_drc_instructions_struct *pstr_drc_instruction_arr, WORD32 impd_apply_gains,
                                  WORD32 sel_drc_idx)
{
  td_job j;
  seam_req r;
  if (tl_sel != sel_drc_idx || tl_dec != pstr_drc_gain_dec || impd_apply_gains != 1 || tl_audio != channel_audio)
  { /* not the pair recorded above (the bank split ran in the reference): the reference finishes it */
    tl_sel = -1;
    real_add()(pstr_drc_gain_dec, channel_audio, pstr_drc_instruction_arr, impd_apply_gains, sel_drc_idx);
    return;
  }
  tl_sel = -1;
  td_build(&j, pstr_drc_gain_dec, pstr_drc_instruction_arr, sel_drc_idx, channel_audio);
  r.kind = RQ_DRCTD;
  r.p[0] = &j;
  seam_submit(&r);
}

/* The whole time-domain step (decoder/impd_drc_process.c:529): the reference's own gain decoding for every selected set
 * (impd_drc_get_gain: exported, stays reference code), then banks + gains + band sum of every set on the GPU.  This is
 * the form that works against the stock reference library, where impd_drc_apply_gains_and_add cannot be reached. */
IA_ERRORCODE impd_drc_td_process(FLOAT32 *audio_in_out_buf[], ia_drc_gain_dec_struct *pstr_drc_gain_dec,
                                 ia_drc_config *pstr_drc_config, ia_drc_gain_struct *pstr_drc_gain, FLOAT32 ln_gain_db,
                                 FLOAT32 boost_fac, FLOAT32 compress_fac, WORD32 drc_characteristic_target)
{
  typedef IA_ERRORCODE (*td_fn)(FLOAT32 **, ia_drc_gain_dec_struct *, ia_drc_config *, ia_drc_gain_struct *, FLOAT32, FLOAT32,
                                FLOAT32, WORD32);
  typedef IA_ERRORCODE (*gain_fn)(ia_drc_gain_buffers_struct *, ia_drc_gain_dec_struct *, ia_drc_config *,
                                  ia_drc_gain_struct *, FLOAT32, FLOAT32, WORD32, FLOAT32, WORD32);
  static gain_fn get_gain;
  ia_drc_instructions_struct *arr = pstr_drc_config->str_drc_instruction_str;
  ia_drc_params_struct *p = &pstr_drc_gain_dec->ia_drc_params_struct;
  int sel, ok = !getenv("MPEGHB200_SEAM_DRC_TD_HOST"), npc = p->num_ch_process;
  IA_ERRORCODE err;
  if (!pstr_drc_config->apply_drc)
    return IA_MPEGH_DEC_NO_ERROR;
  for (sel = p->drc_set_counter - 1; ok && sel >= 0; sel--)
  { /* every selected set must be one the kernel covers, with the channel count the reference would arrive at (:562-573) */
    const int ii = p->sel_drc_array[sel].drc_instrns_idx;
    if (ii < 0)
      ok = 0;
    else
    {
      if (npc == -1)
        npc = arr[ii].audio_num_chan;
      ok = arr[ii].audio_num_chan >= p->channel_offset + npc && td_shape_ok(arr, pstr_drc_gain_dec, sel, npc);
    }
  }
  if (!ok)
  {
    static td_fn real;
    if (!real)
      real = (td_fn)seam_real("impd_drc_td_process");
    seam_n.drc_host++;
    return real(audio_in_out_buf, pstr_drc_gain_dec, pstr_drc_config, pstr_drc_gain, ln_gain_db, boost_fac, compress_fac,
                drc_characteristic_target);
  }
  if (!get_gain)
    get_gain = (gain_fn)seam_real("impd_drc_get_gain");
  for (sel = p->drc_set_counter - 1; sel >= 0; sel--)
    if ((err = get_gain(&pstr_drc_gain_dec->drc_gain_buffers, pstr_drc_gain_dec, pstr_drc_config, pstr_drc_gain, compress_fac,
                        boost_fac, drc_characteristic_target, ln_gain_db, sel)) != IA_MPEGH_DEC_NO_ERROR)
      return err;
  for (sel = p->drc_set_counter - 1; sel >= 0; sel--)
  {
    td_job j;
    seam_req r;
    if (p->num_ch_process == -1)
      p->num_ch_process = arr[p->sel_drc_array[sel].drc_instrns_idx].audio_num_chan;
    td_build(&j, pstr_drc_gain_dec, arr, sel, audio_in_out_buf);
    r.kind = RQ_DRCTD;
    r.p[0] = &j;
    seam_submit(&r);
  }
  return IA_MPEGH_DEC_NO_ERROR;
}

/* ---- the domain switcher around sub-band DRC (impeghd_domain_switcher_process, ia_core_coder_process.c:505) ------- */
/* p[0] = ia_usac_data_struct, p[1] = ia_ds_state_struct; v[0] = n_ch, v[1] = ds_t2f */
static seam_buf b_ds_x, b_ds_spec, b_ds_state;

static void ds_group(seam_req **rq, int n, int t2f)
{
  mpeghb200_ctx *ctx = seam_ctx();
  int N = 0, i, c, c0;
  for (i = 0; i < n; i++)
    N += (int)rq[i]->v[0];
  seam_need(&b_ds_x, (size_t)N * FRAME * sizeof(float)), seam_need(&b_ds_spec, (size_t)N * 2 * FRAME * sizeof(float));
  seam_need(&b_ds_state, (size_t)N * 512 * sizeof(float));
  for (i = 0, c0 = 0; i < n; c0 += (int)rq[i]->v[0], i++)
  {
    ia_usac_data_struct *u = (ia_usac_data_struct *)rq[i]->p[0];
    const ia_ds_state_struct *ds = (const ia_ds_state_struct *)rq[i]->p[1];
    for (c = 0; c < (int)rq[i]->v[0]; c++)
    {
      float *st = (float *)b_ds_state.h + (size_t)(c0 + c) * 512;
      memcpy(st, ds->stft_in_buf[c], 256 * sizeof(float)), memcpy(st + 256, ds->stft_out_buf[c], 256 * sizeof(float));
      if (t2f)
        memcpy((float *)b_ds_x.h + (size_t)(c0 + c) * FRAME, u->time_sample_vector[c], FRAME * sizeof(float));
      else
        memcpy((float *)b_ds_spec.h + (size_t)(c0 + c) * 2 * FRAME, u->time_sample_stft[c], 2 * FRAME * sizeof(float));
    }
  }
  if (t2f)
    CK(mpeghb200_h2d(ctx, b_ds_x.d, b_ds_x.h, (size_t)N * FRAME * sizeof(float)));
  else
    CK(mpeghb200_h2d(ctx, b_ds_spec.d, b_ds_spec.h, (size_t)N * 2 * FRAME * sizeof(float)));
  CK(mpeghb200_h2d(ctx, b_ds_state.d, b_ds_state.h, (size_t)N * 512 * sizeof(float)));
  CK(mpeghb200_stft_dev(ctx, (float *)b_ds_x.d, (float *)b_ds_spec.d, (float *)b_ds_state.d, t2f, N, NULL));
  if (t2f)
    CK(mpeghb200_d2h(ctx, b_ds_spec.h, b_ds_spec.d, (size_t)N * 2 * FRAME * sizeof(float)));
  else
    CK(mpeghb200_d2h(ctx, b_ds_x.h, b_ds_x.d, (size_t)N * FRAME * sizeof(float)));
  CK(mpeghb200_d2h(ctx, b_ds_state.h, b_ds_state.d, (size_t)N * 512 * sizeof(float)));
  for (i = 0, c0 = 0; i < n; c0 += (int)rq[i]->v[0], i++)
  {
    ia_usac_data_struct *u = (ia_usac_data_struct *)rq[i]->p[0];
    ia_ds_state_struct *ds = (ia_ds_state_struct *)rq[i]->p[1];
    for (c = 0; c < (int)rq[i]->v[0]; c++)
    {
      const float *st = (const float *)b_ds_state.h + (size_t)(c0 + c) * 512;
      if (t2f)
      {
        memcpy(ds->stft_in_buf[c], st, 256 * sizeof(float));
        memcpy(u->time_sample_stft[c], (float *)b_ds_spec.h + (size_t)(c0 + c) * 2 * FRAME, 2 * FRAME * sizeof(float));
      }
      else
      {
        memcpy(ds->stft_out_buf[c], st + 256, 256 * sizeof(float));
        memcpy(u->time_sample_vector[c], (float *)b_ds_x.h + (size_t)(c0 + c) * FRAME, FRAME * sizeof(float));
      }
    }
  }
  seam_n.drc_stft += (unsigned long)N;
}

static void be_ds(seam_req **rq, int n)
{
  int a = 0, b;
  while (a < n)
  {
    for (b = a + 1; b < n && rq[b]->v[1] == rq[a]->v[1]; b++)
      ;
    ds_group(rq + a, b - a, (int)rq[a]->v[1]);
    a = b;
  }
}

IA_ERRORCODE impeghd_domain_switcher_process(ia_mpegh_dec_api_struct *p_obj_mpegh_dec, WORD32 *num_chn_out, UWORD32 ds_t2f,
                                             FLOAT32 *scratch_buf)
{
  ia_mpegh_dec_state_struct *state = p_obj_mpegh_dec->p_state_mpeghd;
  ia_dec_data_struct *dec = (ia_dec_data_struct *)state->pstr_dec_data;
  ia_ds_state_struct *ds = &state->state_domain_switcher;
  seam_req r;
  int n_ch;
  (void)scratch_buf;
  if (ds->ds_params == NULL)
    return IA_MPEGH_DEC_EXE_FATAL_DOMAIN_SWITCHER_PROCESS_FAILED;
  *num_chn_out = ds->ds_params->num_out_ch;
  n_ch = ds_t2f ? ds->ds_params->num_in_ch : ds->ds_params->num_out_ch;
  if (n_ch > 0)
  {
    r.kind = RQ_DS;
    r.p[0] = &dec->str_usac_data, r.p[1] = ds;
    r.v[0] = n_ch, r.v[1] = ds_t2f != 0;
    seam_submit(&r);
  }
  state->ui_out_bytes = 1024 * *num_chn_out * sizeof(FLOAT32); /* :622 */
  return IA_MPEGH_DEC_NO_ERROR;
}

/* ---- mid-stream re-initialisation (a changed configuration packet, decoder/ia_core_coder_decode_main.c:2519-2533) ---
 * ia_core_coder_decoder_flush_api clears the handle's persistent structs in place and initialises the decoder again:
 * the device state keyed by their addresses is the past. */
IA_ERRORCODE ia_core_coder_decoder_flush_api(ia_mpegh_dec_api_struct *p_obj_mpegh_dec)
{
  typedef IA_ERRORCODE (*fn_t)(ia_mpegh_dec_api_struct *);
  static fn_t real;
  int i;
  if (!real)
    real = (fn_t)seam_real("ia_core_coder_decoder_flush_api");
  for (i = 0; i < 4; i++) /* NUM_MEMTABS: persistent, scratch, input, output */
    if (p_obj_mpegh_dec->pp_mem_mpeghd && p_obj_mpegh_dec->pp_mem_mpeghd[i] && p_obj_mpegh_dec->p_mem_info_mpeghd)
      seam_release_range(p_obj_mpegh_dec->pp_mem_mpeghd[i],
                         (const char *)p_obj_mpegh_dec->pp_mem_mpeghd[i] + p_obj_mpegh_dec->p_mem_info_mpeghd[i].ui_size);
  return real(p_obj_mpegh_dec);
}

__attribute__((constructor)) static void seam_register_drc(void)
{
  seam_register(RQ_DS, be_ds, "domain_switcher");
  seam_register(RQ_DRCSB, be_drc_sb, "drc_apply_gains_subband");
  seam_register(RQ_DRCTD, be_drc_td, "drc filter banks + gains");
}
