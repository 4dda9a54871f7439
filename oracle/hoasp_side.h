/* oracle/hoasp_side.h -- TEST INFRASTRUCTURE: per-frame HOA spatial-decoder side info as the harness and the
 * oracle take it; same layout as mpeghb200_hoa_spatial_side (include/mpeghb200.h).  Contents = what
 * impeghd_hoa_spatial_decode_frame_side_info (decoder/impeghd_hoa_spatial_decoder.c:291) leaves in
 * ia_hoa_dec_frame_param_str (decoder/impeghd_hoa_frame_params.h:38-62), stale entries included. */
#ifndef HOASP_SIDE_H
#define HOASP_SIDE_H
#include <stdint.h>
#define HOASP_MAX_CH 16
#define HOASP_MAX_COEF 49
#define HOASP_MAX_PRED 4
typedef struct
{
  int32_t n_vec, n_dir, n_enable, n_disable;
  int32_t exponents[HOASP_MAX_CH];
  int32_t is_exception[HOASP_MAX_CH];
  float prev_gain_reset[HOASP_MAX_CH]; /* > 0: overwrite the carried gain first (hoa_independency_flag) */
  int32_t amb_hoa_assign[HOASP_MAX_CH];
  int32_t vec_index[HOASP_MAX_CH];     /* vectors[].index, all 16 slots */
  int32_t dir_ch[HOASP_MAX_CH + 1];    /* active_and_grid_dir_indices[].ch_idx, slots 0..16 */
  int32_t dir_id[HOASP_MAX_CH + 1];
  int32_t pad0[2];
  int32_t enable[HOASP_MAX_COEF];      /* amb_coeff_indices_to_enable, -1 padded */
  int32_t disable[HOASP_MAX_COEF];
  int32_t non_en_dis[HOASP_MAX_COEF];  /* non_en_dis_able_act_hoa_coeff_indices */
  int32_t pred_type[HOASP_MAX_COEF];
  int32_t pred_idx[HOASP_MAX_PRED][HOASP_MAX_COEF];   /* prediction_indices_mat, row stride n_coef in the reference */
  int32_t pred_quant[HOASP_MAX_PRED][HOASP_MAX_COEF]; /* quant_pred_fact_mat */
  float vec[HOASP_MAX_CH][HOASP_MAX_COEF];            /* vectors[].sig_id */
} hoasp_side;
#endif
