/* oracle/ref_harness_hoasp.c -- TEST INFRASTRUCTURE ONLY.
 *
 * Function-level entry points into the UNMODIFIED reference HOA spatial decoder (linked into
 * oracle/_ref/libmpeghref.so): impeghd_hoa_spatial_init (decoder/impeghd_hoa_spatial_decoder.c:154) and the
 * four signal stages that impeghd_hoa_spatial_process (:663) runs after the side-info decode:
 *   impeghd_hoa_inv_dyn_correction_process   decoder/impeghd_hoa_inverse_dyn_correction.c:75
 *   impeghd_hoa_ambience_synthesis_process   decoder/impeghd_hoa_amb_syn.c:76
 *   impeghd_hoa_ch_reassignment_process      decoder/impeghd_hoa_ch_reassignment.c:71
 *   impeghd_hoa_pre_dom_sound_syn_process    decoder/impeghd_hoa_pre_dom_sound_syn.c:72
 * The frame parameters are written straight into frame_params_handle (the bitstream-side decode stays host
 * code in the product too).  No reference code, only calls.
 */
#include <stdlib.h>
#include <string.h>
#include <math.h>

#include <impeghd_type_def.h>
#include "impd_drc_common.h"
#include "impd_drc_extr_delta_coded_info.h"
#include "impd_drc_struct.h"
#include "impeghd_hoa_common_values.h"
#include "impeghd_hoa_common_functions.h"
#include "impeghd_hoa_data_types.h"
#include "impeghd_error_codes.h"
#include "impeghd_hoa_frame_params.h"
#include "impeghd_hoa_rom.h"
#include "ia_core_coder_bitbuffer.h"
#include "impeghd_hoa_matrix_struct.h"
#include "impeghd_hoa_matrix.h"
#include "ia_core_coder_cnst.h"
#include "ia_core_coder_config.h"
#include "impeghd_hoa_spatial_decoder_struct.h"
#include "impeghd_hoa_amb_syn.h"
#include "impeghd_hoa_ch_reassignment.h"
#include "impeghd_hoa_config.h"
#include "impeghd_hoa_dec_struct.h"
#include "impeghd_hoa_frame.h"
#include "impeghd_hoa_init.h"
#include "impeghd_hoa_inverse_dyn_correction.h"
#include "impeghd_hoa_pre_dom_sound_syn.h"

#include "hoasp_side.h"

typedef struct
{
  ia_spatial_dec_str dec;
  ia_hoa_config_struct cfg;
  WORD32 is_exception[HOA_MAXIMUM_NUM_PERC_CODERS];
  char scratch[4 << 20];
} ref_hoasp_handle;

/* cfg[12] = {order, n_transport, n_addnl_coders, min_amb_order, coded_v_vec_length, spat_interpolation_time,
 *            spat_interpolation_method, pred_dir_sigs, num_bits_per_sf, use_phase_shift_decorr, 0, 0} */
void *ref_hoasp_create(const int *c, int *err_out)
{
  ref_hoasp_handle *h = (ref_hoasp_handle *)calloc(1, sizeof(ref_hoasp_handle));
  IA_ERRORCODE err;
  if (!h)
    return 0;
  h->cfg.order = c[0];
  h->cfg.num_coeffs = (c[0] + 1) * (c[0] + 1);
  h->cfg.num_transport_ch = c[1];
  h->cfg.num_addnl_coders = c[2];
  h->cfg.min_amb_order = c[3];
  h->cfg.min_coeffs_for_amb = (c[3] + 1) * (c[3] + 1);
  h->cfg.coded_v_vec_length = c[4];
  h->cfg.spat_interpolation_time = c[5];
  h->cfg.spat_interpolation_method = c[6];
  h->cfg.pred_dir_sigs = c[7];
  h->cfg.num_bits_per_sf = c[8];
  h->cfg.use_phase_shift_decorr = c[9];
  h->cfg.frame_length = 1024;
  h->cfg.core_coder_frame_length = 1024;
  h->cfg.max_order_to_be_transmitted = c[0];
  err = impeghd_hoa_spatial_init(&h->dec, &h->cfg, h->scratch);
  h->dec.use_phase_shift_decorr = c[9];
  if (err_out)
    *err_out = (int)err;
  return h;
}
void ref_hoasp_destroy(void *p) { free(p); }

/* info[4] = {low_order_mat_sz, vec_start, n_coef, 0}; tables sized for the maxima:
 * order_matrix [49*49], fine [900*49] (row stride 49), coarse [49*49] (row stride n_coef),
 * wins [4][1024] = fade_in_vec, fade_out_vec, fade_in_dir, fade_out_dir, dyn_pow [7][1024], dyn_win [1024],
 * pow2 [128] = 2^0..2^63 then 2^-0..2^-63 as the reference tabulates them */
void ref_hoasp_tables(void *p, int *info, float *order_matrix, float *fine, float *coarse, float *wins,
                      float *dyn_pow, float *dyn_win, float *pow2)
{
  ref_hoasp_handle *h = (ref_hoasp_handle *)p;
  ia_spatial_dec_pre_dom_sound_syn_str *s = &h->dec.pre_dom_sound_syn_handle;
  int i;
  info[0] = h->dec.amb_syn_handle.low_order_mat_sz;
  info[1] = h->dec.vec_start;
  info[2] = h->cfg.num_coeffs;
  info[3] = 0;
  memcpy(order_matrix, h->dec.amb_syn_handle.order_matrix, sizeof(float) * 49 * 49);
  memcpy(fine, h->dec.transpose_mod_mtx_grid_pts, sizeof(float) * 900 * 49);
  memcpy(coarse, s->dir_based_pre_dom_sound_syn_handle.mode_mt_all_coarse_grid_points, sizeof(float) * 49 * 49);
  memcpy(wins, s->fade_in_win_for_vec_sigs, sizeof(float) * 1024);
  memcpy(wins + 1024, s->fade_out_win_for_vec_sigs, sizeof(float) * 1024);
  memcpy(wins + 2048, s->fade_in_win_for_dir_sigs, sizeof(float) * 1024);
  memcpy(wins + 3072, s->fade_out_win_for_dir_sigs, sizeof(float) * 1024);
  memcpy(dyn_pow, ia_hoa_inv_dyn_win_pow_table, sizeof(float) * 7 * 1024);
  memcpy(dyn_win, h->dec.inv_dyn_corr_win_func, sizeof(float) * 1024);
  for (i = 0; i < 64; i++)
  {
    pow2[i] = impeghd_hoa_pow_2(i);
    pow2[64 + i] = impeghd_hoa_pow_2(-i);
  }
}

/* One frame: in [n_transport][1024] -> out [n_coef][1024]; the decoder state inside h advances. */
int ref_hoasp_process(void *p, const hoasp_side *sd, const float *in, float *out)
{
  ref_hoasp_handle *h = (ref_hoasp_handle *)p;
  ia_spatial_dec_str *d = &h->dec;
  ia_hoa_dec_frame_param_str *fp = &d->frame_params_handle;
  WORD32 *reassign = (WORD32 *)h->scratch;
  int nc = (int)h->cfg.num_coeffs, T = (int)h->cfg.num_transport_ch, i, r;
  IA_ERRORCODE err;
  static __thread float inbuf[HOASP_MAX_CH * 1024];

  fp->num_vec_predom_sounds = sd->n_vec;
  fp->num_dir_predom_sounds = sd->n_dir;
  fp->num_enable_coeff = sd->n_enable;
  fp->num_disable_coeff = sd->n_disable;
  fp->ptr_is_exception = h->is_exception;
  for (i = 0; i < T; i++)
  {
    fp->exponents[i] = sd->exponents[i];
    h->is_exception[i] = sd->is_exception[i];
    fp->amb_hoa_assign[i] = sd->amb_hoa_assign[i];
    if (sd->prev_gain_reset[i] > 0.0f)
      d->inv_dyn_corr_prev_gain[i] = sd->prev_gain_reset[i];
  }
  for (i = 0; i < HOASP_MAX_CH; i++)
  {
    fp->vectors[i].index = sd->vec_index[i];
    memcpy(fp->vectors[i].sig_id, sd->vec[i], sizeof(float) * HOASP_MAX_COEF);
  }
  for (i = 0; i <= HOASP_MAX_CH; i++)
  {
    fp->active_and_grid_dir_indices[i].ch_idx = sd->dir_ch[i];
    fp->active_and_grid_dir_indices[i].dir_id = sd->dir_id[i];
  }
  for (i = 0; i < HOASP_MAX_COEF; i++)
  {
    fp->amb_coeff_indices_to_enable[i] = (UWORD32)sd->enable[i];
    fp->amb_coeff_indices_to_disable[i] = (UWORD32)sd->disable[i];
    fp->non_en_dis_able_act_hoa_coeff_indices[i] = (UWORD32)sd->non_en_dis[i];
  }
  for (i = 0; i < nc; i++)
    fp->prediction_type_vec[i] = sd->pred_type[i];
  for (r = 0; r < HOASP_MAX_PRED; r++)
    for (i = 0; i < nc; i++)
    {
      fp->prediction_indices_mat[r * nc + i] = sd->pred_idx[r][i];
      fp->quant_pred_fact_mat[r * nc + i] = sd->pred_quant[r][i];
    }

  /* the lines of impeghd_hoa_spatial_process after the side-info decode (:690-718) */
  d->scratch = (WORD8 *)h->scratch + HOA_MAXIMUM_NUM_PERC_CODERS * sizeof(*reassign);
  memset(reassign, 0xFF, sizeof(*reassign) * HOA_MAXIMUM_NUM_PERC_CODERS);
  for (i = 0; i < T; i++)
    if (fp->amb_hoa_assign[i] != 0)
      reassign[fp->amb_hoa_assign[i] - 1] = i;
  memcpy(inbuf, in, sizeof(float) * T * 1024);
  err = impeghd_hoa_inv_dyn_correction_process(inbuf, d);
  if (!err)
    err = impeghd_hoa_ambience_synthesis_process(d, reassign);
  if (!err)
    err = impeghd_hoa_ch_reassignment_process(d);
  if (!err)
    err = impeghd_hoa_pre_dom_sound_syn_process(d, out);
  return (int)err;
}
