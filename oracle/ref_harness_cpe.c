/* oracle/ref_harness_cpe.c -- TEST INFRASTRUCTURE ONLY.
 *
 * Phase 2 of one channel-pair element through the UNMODIFIED reference: the parse-phase downmix update
 * ia_core_coder_cplx_prev_mdct_dmx (decoder/ia_core_coder_ext_ch_ele.c:703) followed by
 * ia_core_coder_data_process (:1695): stereo tools (M/S, complex prediction), TNS before or after them
 * (tns_on_lr), spectrum save, FD synthesis and the (inactive) LTPF pass, on a hand-filled ia_usac_data_struct
 * that persists across frames in the handle.  Contains no reference code, only calls into it.
 */
#include <stdlib.h>
#include <string.h>

#include <impeghd_type_def.h>
#include "impd_drc_common.h"
#include "impd_drc_extr_delta_coded_info.h"
#include "impd_drc_struct.h"
#include "impeghd_intrinsics_flt.h"
#include "ia_core_coder_bitbuffer.h"
#include "ia_core_coder_cnst.h"
#include "ia_core_coder_constants.h"
#include "ia_core_coder_defines.h"
#include "ia_core_coder_acelp_info.h"
#include "ia_core_coder_interface.h"
#include "ia_core_coder_info.h"
#include "ia_core_coder_tns_usac.h"
#include "ia_core_coder_rom.h"
#include "impeghd_tbe_dec.h"
#include "ia_core_coder_igf_data_struct.h"
#include "ia_core_coder_stereo_lpd.h"
#include "impeghd_error_codes.h"
#include "impeghd_hoa_common_values.h"
#include "impeghd_hoa_matrix_struct.h"
#include "impeghd_hoa_matrix.h"
#include "ia_core_coder_config.h"
#include "ia_core_coder_main.h"
#include "ia_core_coder_arith_dec.h"
#include "ia_core_coder_bit_extract.h"
#include "ia_core_coder_func_def.h"
#include "ia_core_coder_ltpf.h"

#include "oracle.h"

extern IA_ERRORCODE ia_core_coder_data_process(ia_usac_data_struct *usac_data, WORD32 elem_idx, WORD32 chan_offset,
                                               WORD32 nr_core_coder_channels,
                                               ia_usac_tmp_core_coder_struct *pstr_core_coder);

typedef struct
{
  ia_usac_data_struct usac;
  FLOAT32 scratch_float[4 * 4096];
  FLOAT32 fft_scratch[8 * 4096];
  FLOAT32 x_ac[2048];
  FLOAT32 ltpf_scratch[2][8192];
  WORD32 tns_max[16][2];
} ref_cpe_handle;

/* what ia_core_coder_info_init (static, ia_core_coder_create.c:128) derives from the sampling-rate ROM entry */
static void fill_sfb_info(ia_sfb_info_struct *info, int islong, const WORD16 *tbl, int n_sfb, WORD16 *width)
{
  int w, sfb, k = 0, n = 0;
  memset(info, 0, sizeof(*info));
  info->islong = islong;
  info->samp_per_bk = 1024;
  info->max_win_len = islong ? 1 : 8;
  info->bins_per_sbk = 1024 / info->max_win_len;
  info->ptr_sfb_tbl = tbl;
  info->sfb_per_sbk = n_sfb;
  info->sfb_width = width;
  for (sfb = 0; sfb < n_sfb; sfb++)
    width[sfb] = (WORD16)(tbl[sfb] - (sfb ? tbl[sfb - 1] : 0));
  info->num_groups = 1;
  info->group_len[0] = 1;
  for (w = 0; w < info->max_win_len; w++)
  {
    info->sfb_per_bk += n_sfb;
    for (sfb = 0; sfb < n_sfb; sfb++)
      info->sfb_idx_tbl[k + sfb] = (WORD16)(tbl[sfb] + n);
    k += n_sfb;
    n += info->bins_per_sbk;
  }
}

void *ref_cpe_create(int sampling_rate_idx, int tns_max_long, int tns_max_short)
{
  ref_cpe_handle *h = (ref_cpe_handle *)calloc(1, sizeof(ref_cpe_handle));
  ia_usac_data_struct *u;
  const ia_usac_samp_rate_info *sr = &ia_core_coder_samp_rate_info[sampling_rate_idx];
  int ch;
  if (!h)
    return 0;
  u = &h->usac;
  u->ccfl = 1024;
  u->sampling_rate_idx = sampling_rate_idx;
  u->scratch_buffer_float = h->scratch_float;
  u->ptr_fft_scratch = h->fft_scratch;
  u->x_ac_dec_float = h->x_ac;
  h->tns_max[sampling_rate_idx][0] = tns_max_long;
  h->tns_max[sampling_rate_idx][1] = tns_max_short;
  u->tns_max_bands_tbl_usac = &h->tns_max;
  fill_sfb_info(&u->str_only_long_info, 1, sr->ptr_sfb_1024, sr->num_sfb_1024, u->sfb_width_long);
  fill_sfb_info(&u->str_eight_short_info, 0, sr->ptr_sfb_128, sr->num_sfb_128, u->sfb_width_short);
  u->pstr_usac_winmap[0] = u->pstr_usac_winmap[1] = u->pstr_usac_winmap[3] = u->pstr_usac_winmap[4] =
      &u->str_only_long_info;
  u->pstr_usac_winmap[2] = &u->str_eight_short_info;
  for (ch = 0; ch < 2; ch++)
  {
    u->coef[ch] = u->arr_coef[ch];
    u->coef_save_float[ch] = u->arr_coef_save_float[ch];
    u->group_dis[ch] = u->arr_group_dis[ch];
    u->ms_used[ch] = u->arr_ms_used[ch];
    u->pstr_tns[ch] = &u->arr_str_tns[ch];
    u->str_tddec[ch] = &u->arr_str_tddec[ch];
    ia_core_coder_ltpf_init(&u->str_tddec[ch]->ltpf_data, 48000, 1024, h->ltpf_scratch[ch]);
  }
  return h;
}

void ref_cpe_destroy(void *p) { free(p); }

/* sfb tables of the sampling-rate index, for the tests (band edges are ROM data of the reference) */
int ref_cpe_sfb_table(int sampling_rate_idx, int islong, short *out, int cap)
{
  const ia_usac_samp_rate_info *sr = &ia_core_coder_samp_rate_info[sampling_rate_idx];
  const WORD16 *t = islong ? sr->ptr_sfb_1024 : sr->ptr_sfb_128;
  int n = islong ? sr->num_sfb_1024 : sr->num_sfb_128, i;
  for (i = 0; i < n && i < cap; i++)
    out[i] = t[i];
  return n;
}

static void fill_tns(ia_tns_frame_info_struct *tns, int n_win, const int *n_filt, const int *filt)
{
  int w, f, k;
  memset(tns, 0, sizeof(*tns));
  tns->n_subblocks = n_win;
  for (w = 0; w < n_win; w++)
  {
    tns->str_tns_info[w].n_filt = n_filt[w];
    for (f = 0; f < 3; f++)
    {
      const int *p = filt + (w * 3 + f) * 20;
      ia_tns_filter_struct *q = &tns->str_tns_info[w].str_filter[f];
      tns->str_tns_info[w].coef_res = p[4] ? p[4] : tns->str_tns_info[w].coef_res;
      q->order = p[0], q->start_band = p[1], q->stop_band = p[2], q->direction = p[3];
      for (k = 0; k < 15; k++)
        q->coef[k] = (WORD16)p[5 + k];
    }
  }
}

/* One frame of the pair.  sd: stereo side info (window_shape_prev is taken from the handle's history, which the
 * reference maintains itself; sd->window_shape_prev must agree).  win_group[8]: group index of every short
 * window (ignored for long blocks).  tns_n_filt [2][8], tns_filt [2][8][3][20] in ref_tns_apply's layout,
 * max_sfb[2] = nbands of the TNS band clamp.  coef_l/r are consumed; out_l/r = time_sample_vector;
 * spec_l/r (may be NULL) = the handle's coef_save_float after the frame. */
int ref_cpe_process(void *p, const oracle_stereo_side *sd, const int *win_group, int tns_on_lr, const int *tns_n_filt,
                    const int *tns_filt, const int *max_sfb, const float *coef_l, const float *coef_r, float *out_l,
                    float *out_r, float *spec_l, float *spec_r)
{
  ref_cpe_handle *h = (ref_cpe_handle *)p;
  ia_usac_data_struct *u = &h->usac;
  ia_usac_tmp_core_coder_struct cc;
  ia_sfb_info_struct *info;
  const int n_win = sd->is_short ? 8 : 1;
  int ch, w, g, sfb, err, n_groups = 1;

  memset(&cc, 0, sizeof(cc));
  for (ch = 0; ch < 2; ch++)
  {
    u->window_sequence[ch] = sd->window_sequence;
    u->window_shape[ch] = sd->window_shape;
    if (u->window_shape_prev[ch] != sd->window_shape_prev)
      return -100;
    u->pstr_sfb_info[ch] = u->pstr_usac_winmap[sd->window_sequence];
    fill_tns(u->pstr_tns[ch], n_win, tns_n_filt + ch * 8, tns_filt + ch * 8 * 3 * 20);
    cc.max_sfb[ch] = (UWORD8)max_sfb[ch];
  }
  info = u->pstr_sfb_info[0];
  /* grouping: group_dis[g] = index one past the last window of group g (ia_core_coder_ics_info :779-788) */
  if (sd->is_short)
  {
    n_groups = 0;
    for (w = 0; w < 8; w++)
      if (w == 7 || win_group[w + 1] != win_group[w])
        u->group_dis[0][n_groups] = u->group_dis[1][n_groups] = (UWORD8)(w + 1), n_groups++;
  }
  else
    u->group_dis[0][0] = u->group_dis[1][0] = 1;
  memset(u->arr_ms_used[0], 0, sizeof(u->arr_ms_used[0]));
  memset(u->cplx_pred_used[0], 0, sizeof(u->cplx_pred_used[0]));
  memset(u->alpha_q_re[0], 0, sizeof(u->alpha_q_re[0]));
  memset(u->alpha_q_im[0], 0, sizeof(u->alpha_q_im[0]));
  for (g = 0, w = 0; g < n_groups; g++)
  { /* first window of the group carries the group's side info */
    const int base = sd->is_short ? 16 * w : 0;
    for (sfb = 0; sfb < info->sfb_per_sbk; sfb++)
    {
      u->arr_ms_used[0][g * info->sfb_per_sbk + sfb] = sd->used[base + sfb];
      u->cplx_pred_used[0][g][sfb] = sd->used[base + sfb];
      u->alpha_q_re[0][g][sfb] = sd->alpha_re[base + sfb];
      u->alpha_q_im[0][g][sfb] = sd->alpha_im[base + sfb];
    }
    if (sd->is_short)
      while (w < 8 && win_group[w] == g)
        w++;
  }
  cc.common_window = 1;
  cc.tns_on_lr = tns_on_lr;
  cc.ms_mask_present[0] = sd->tool;
  cc.max_sfb_ste = sd->last_band;
  cc.max_sfb_ste_clear = sd->last_band;
  cc.pred_dir = sd->pred_dir;
  cc.complex_coef = sd->complex_coef;
  cc.use_prev_frame = sd->use_prev_frame;
  memcpy(u->coef[0], coef_l, 1024 * sizeof(float));
  memcpy(u->coef[1], coef_r, 1024 * sizeof(float));

  /* parse phase (:1396-1406): previous-frame downmix from last frame's saved spectra */
  ia_core_coder_cplx_prev_mdct_dmx(info, u->coef_save_float[0], u->coef_save_float[1], u->dmx_re_prev[0], sd->pred_dir,
                                   !sd->prev_dmx);
  err = (int)ia_core_coder_data_process(u, 0, 0, 2, &cc);
  if (err)
    return err;
  memcpy(out_l, u->time_sample_vector[0], 1024 * sizeof(float));
  memcpy(out_r, u->time_sample_vector[1], 1024 * sizeof(float));
  if (spec_l)
    memcpy(spec_l, u->coef_save_float[0], 1024 * sizeof(float));
  if (spec_r)
    memcpy(spec_r, u->coef_save_float[1], 1024 * sizeof(float));
  return 0;
}

/* MDST filter ROM (decoder/ia_core_coder_rom.c:953-1022) in the oracle's / product's class order */
void ref_mdst_rom(float *curr /*[4][2][2][7]*/, float *prev /*[2][2][7]*/)
{
  const FLOAT32 *const(*cur[4])[2] = {ia_core_coder_mdst_fcoeff_longshort_curr, ia_core_coder_mdst_fcoeff_start_curr,
                                      ia_core_coder_mdst_fcoeff_stop_cur, ia_core_coder_mdst_fcoeff_stopstart_cur};
  const FLOAT32 *const *prv[2] = {ia_core_coder_mdst_fcoeff_l_s_start_left_prev,
                                  ia_core_coder_mdst_fcoeff_stop_stopstart_left_prev};
  int c, a, b, k;
  for (c = 0; c < 4; c++)
    for (a = 0; a < 2; a++)
      for (b = 0; b < 2; b++)
        for (k = 0; k < 7; k++)
          curr[((c * 2 + a) * 2 + b) * 7 + k] = cur[c][a][b][k];
  for (c = 0; c < 2; c++)
    for (a = 0; a < 2; a++)
      for (k = 0; k < 7; k++)
        prev[(c * 2 + a) * 7 + k] = prv[c][a][k];
}

/* ------------------------------------------------------------------------------------------------
 * IGF, mono path: ia_core_coder_igf_init (decoder/ia_core_coder_igf_dec.c:2171) derives the grid from the
 * configuration indices, then per frame ia_core_coder_igf_mono (:2479) + ia_core_coder_igf_tnf (:2297), called
 * the way ia_core_coder_data_process does (:1880-1890).
 * ---------------------------------------------------------------------------------------------- */
#include "ia_core_coder_igf_dec.h"

typedef struct
{
  ref_cpe_handle cpe;
} ref_igf_handle;

void *ref_igf_create(int sampling_rate_idx, int sample_rate, int igf_start_index, int igf_stop_index, int use_high_res,
                     int use_whitening, int use_enf)
{
  ref_cpe_handle *h = (ref_cpe_handle *)ref_cpe_create(sampling_rate_idx, 0, 0);
  ia_usac_dec_element_config_struct ele;
  if (!h)
    return 0;
  memset(&ele, 0, sizeof(ele));
  ele.enhanced_noise_filling = 1;
  ele.str_usac_ele_igf_init_config.igf_use_enf = (WORD8)use_enf;
  ele.str_usac_ele_igf_init_config.igf_use_whitening = (WORD8)use_whitening;
  ele.str_usac_ele_igf_init_config.igf_independent_tilling = 1;
  ele.str_usac_ele_igf_init_config.igf_start_index = (WORD8)igf_start_index;
  ele.str_usac_ele_igf_init_config.igf_stop_index = (WORD8)igf_stop_index;
  ele.str_usac_ele_igf_init_config.igf_use_high_resolution = (WORD8)use_high_res;
  ia_core_coder_igf_init(&h->usac, &ele, ID_USAC_SCE, 0, (UINT32)sample_rate, 0);
  return h;
}

void ref_igf_destroy(void *p) { free(p); }

/* out[2][4] = {sfb_start, sfb_stop, igf_min, num_tiles} of the long and the short grid */
void ref_igf_grid(void *p, int *out)
{
  ref_cpe_handle *h = (ref_cpe_handle *)p;
  int g;
  for (g = 0; g < 2; g++)
  {
    const ia_usac_igf_grid_config_struct *q = &h->usac.igf_config[0].igf_grid[g];
    out[4 * g + 0] = q->igf_sfb_start, out[4 * g + 1] = q->igf_sfb_stop;
    out[4 * g + 2] = q->igf_min, out[4 * g + 3] = q->igf_num_tiles;
  }
}

/* one channel-frame: coef [1024] in place, tiles [4][1024] (source tiles, window-major inside a tile);
 * returns seed_value after the frame */
unsigned int ref_igf_process(void *p, const oracle_igf_side *sd, float *coef, const float *tiles)
{
  ref_cpe_handle *h = (ref_cpe_handle *)p;
  ia_usac_data_struct *u = &h->usac;
  ia_usac_igf_config_struct *cfg = &u->igf_config[0];
  ia_usac_igf_dec_data_struct *d = &u->igf_dec_data[0];
  ia_usac_igf_bitstream_struct *bs = &d->igf_bitstream[0];
  const int win = sd->win_type == 1 ? EIGHT_SHORT_SEQUENCE : ONLY_LONG_SEQUENCE;
  WORD16 group_len[8];
  int t, g, sfb;
  for (t = 0; t < 4; t++)
  {
    memset(&d->igf_input_spec[t * MAX_IGF_LEN], 0, MAX_IGF_LEN * sizeof(FLOAT32));
    memcpy(&d->igf_input_spec[t * MAX_IGF_LEN], tiles + t * 1024, 1024 * sizeof(FLOAT32));
    bs->igf_tile_num[t] = (WORD8)sd->tile_num[t];
    bs->igf_whitening_level[t] = sd->whitening_level[t];
  }
  bs->igf_all_zero = sd->all_zero;
  bs->igf_is_tnf = (WORD8)sd->is_tnf;
  for (g = 0; g < 8; g++)
  {
    group_len[g] = sd->group_len[g];
    for (sfb = 0; sfb < (sd->win_type == 1 ? 16 : 64) && sfb < NSFB_LONG; sfb++)
      bs->igf_levels_curr_float[g][sfb] = sd->levels[(sd->win_type == 1 ? 16 * g : 0) + sfb];
  }
  u->seed_value[0] = sd->seed;
  ia_core_coder_igf_mono(u, cfg, 0, sd->win_type, sd->win_type, sd->num_groups, group_len,
                         sd->win_type == 1 ? 128 : 1024, win, coef, h->fft_scratch);
  ia_core_coder_igf_tnf(u, cfg, 0, sd->win_type, sd->win_type, win, coef, h->fft_scratch);
  return u->seed_value[0];
}

/* ------------------------------------------------------------------------------------------------
 * IGF, joint-stereo path of a channel pair, in the order of ia_core_coder_data_process_igf_joint_stereo_*
 * (decoder/ia_core_coder_ext_ch_ele.c:1500,1590): ia_core_coder_igf_apply_stereo (igf_dec.c:1661) with an M/S-type
 * band mask, optionally impeghd_ms_stereo (:1427) over the IGF band range on the filled spectra and on the TNF
 * buffers, then ia_core_coder_igf_tnf (:2297) per channel.  Handle from ref_igf_create.
 * mask: per (group, band) [n_groups][sfb_per_sbk]; tiles_* [4][1024]; seeds_out[2].
 * ---------------------------------------------------------------------------------------------- */
extern VOID impeghd_ms_stereo(ia_sfb_info_struct *info, WORD8 *group, WORD8 *mask, WORD32 first_band, WORD32 last_band,
                              FLOAT32 *right_spec, FLOAT32 *left_spec);

static void fill_igf_channel(ia_usac_data_struct *u, int ch, const oracle_igf_side *sd, const float *tiles)
{
  ia_usac_igf_dec_data_struct *d = &u->igf_dec_data[ch];
  ia_usac_igf_bitstream_struct *bs = &d->igf_bitstream[0];
  int t, g, sfb;
  for (t = 0; t < 4; t++)
  {
    memset(&d->igf_input_spec[t * MAX_IGF_LEN], 0, MAX_IGF_LEN * sizeof(FLOAT32));
    memcpy(&d->igf_input_spec[t * MAX_IGF_LEN], tiles + t * 1024, 1024 * sizeof(FLOAT32));
    bs->igf_tile_num[t] = (WORD8)sd->tile_num[t];
    bs->igf_whitening_level[t] = sd->whitening_level[t];
  }
  bs->igf_all_zero = sd->all_zero;
  bs->igf_is_tnf = (WORD8)sd->is_tnf;
  for (g = 0; g < 8; g++)
    for (sfb = 0; sfb < (sd->win_type == 1 ? 16 : 64) && sfb < NSFB_LONG; sfb++)
      bs->igf_levels_curr_float[g][sfb] = sd->levels[(sd->win_type == 1 ? 16 * g : 0) + sfb];
  u->seed_value[ch] = sd->seed;
}

void ref_igf_stereo_process(void *p, const oracle_igf_side *sl, const oracle_igf_side *sr, const unsigned char *mask,
                            int ms_after, float *coef_l, float *coef_r, const float *tiles_l, const float *tiles_r,
                            unsigned int *seeds_out)
{
  ref_cpe_handle *h = (ref_cpe_handle *)p;
  ia_usac_data_struct *u = &h->usac;
  ia_usac_igf_config_struct *cfg = &u->igf_config[0];
  const int win = sl->win_type == 1 ? EIGHT_SHORT_SEQUENCE : ONLY_LONG_SEQUENCE;
  ia_sfb_info_struct *info = u->pstr_usac_winmap[win];
  WORD16 group_len[8];
  UWORD8 group_dis[8];
  int g, ch, acc = 0;
  fill_igf_channel(u, 0, sl, tiles_l);
  fill_igf_channel(u, 1, sr, tiles_r);
  memcpy(u->coef[0], coef_l, 1024 * sizeof(float));
  memcpy(u->coef[1], coef_r, 1024 * sizeof(float));
  u->window_sequence[0] = u->window_sequence[1] = win;
  for (g = 0; g < 8; g++)
  {
    group_len[g] = sl->group_len[g];
    acc += sl->group_len[g];
    group_dis[g] = (UWORD8)acc;
  }
  ia_core_coder_igf_apply_stereo(u, 0, cfg, sl->win_type, (const WORD8 *)mask, 1, sl->num_groups, group_len,
                                 sl->win_type == 1 ? 128 : 1024, info->sfb_per_sbk);
  if (ms_after)
  {
    impeghd_ms_stereo(info, (WORD8 *)group_dis, (WORD8 *)mask, sl->sfb_start, sl->sfb_stop, u->coef[1], u->coef[0]);
    if (cfg->use_tnf && win != EIGHT_SHORT_SEQUENCE)
      impeghd_ms_stereo(info, (WORD8 *)group_dis, (WORD8 *)mask, sl->sfb_start, sl->sfb_stop,
                        u->igf_dec_data[1].igf_tnf_spec_float, u->igf_dec_data[0].igf_tnf_spec_float);
  }
  for (ch = 0; ch < 2; ch++)
    ia_core_coder_igf_tnf(u, cfg, ch, sl->win_type, sl->win_type, win, u->coef[ch], h->fft_scratch);
  memcpy(coef_l, u->coef[0], 1024 * sizeof(float));
  memcpy(coef_r, u->coef[1], 1024 * sizeof(float));
  seeds_out[0] = u->seed_value[0], seeds_out[1] = u->seed_value[1];
}

/* ------------------------------------------------------------------------------------------------
 * LTPF: ia_core_coder_ltpf_init (decoder/ia_core_coder_ltpf.c:356) + ia_core_coder_usac_ltpf_process (:395)
 * ---------------------------------------------------------------------------------------------- */
typedef struct
{
  ia_ltpf_data_str d;
  FLOAT32 scratch[16384];
} ref_ltpf_handle;

void *ref_ltpf_create(int sample_rate)
{
  ref_ltpf_handle *h = (ref_ltpf_handle *)calloc(1, sizeof(ref_ltpf_handle));
  if (h)
    ia_core_coder_ltpf_init(&h->d, sample_rate, 1024, h->scratch);
  return h;
}
void ref_ltpf_destroy(void *p) { free(p); }
/* out5: pitch_min, pitch_fr1, pitch_fr2, pitch_res, pitch_max */
void ref_ltpf_cfg(void *p, int *out5)
{
  ref_ltpf_handle *h = (ref_ltpf_handle *)p;
  out5[0] = h->d.pitch_min, out5[1] = h->d.pitch_fr1, out5[2] = h->d.pitch_fr2, out5[3] = h->d.pitch_res;
  out5[4] = h->d.pitch_max;
}
void ref_ltpf_process(void *p, int data_present, int pitch_lag_idx, int gain_idx, float *x)
{
  ref_ltpf_handle *h = (ref_ltpf_handle *)p;
  h->d.data_present = data_present;
  h->d.pitch_lag_idx = pitch_lag_idx;
  h->d.gain_idx = gain_idx;
  ia_core_coder_usac_ltpf_process(&h->d, x);
}
