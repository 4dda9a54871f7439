/* oracle/ref_harness_hoa.c -- TEST INFRASTRUCTURE ONLY.
 *
 * Function-level entry points into the UNMODIFIED reference HOA renderer (linked into
 * oracle/_ref/libmpeghref.so): impeghd_hoa_ren_renderer_init (impeghd_hoa_renderer.c:399, render-matrix design),
 * impeghd_hoa_ren_nfc_init (:92) and impeghd_hoa_ren_block_process (:181).  No reference code, only calls.
 */
#include <stdlib.h>
#include <string.h>
#include <math.h>

#include <impeghd_type_def.h>
#include "impd_drc_common.h"
#include "impd_drc_extr_delta_coded_info.h"
#include "impd_drc_struct.h"
#include "ia_core_coder_bitbuffer.h"
#include "ia_core_coder_cnst.h"
#include "impeghd_hoa_common_values.h"
#include "impeghd_hoa_common_functions.h"
#include "impeghd_hoa_data_types.h"
#include "impeghd_error_codes.h"
#include "impeghd_hoa_frame_params.h"
#include "impeghd_hoa_matrix_struct.h"
#include "impeghd_hoa_matrix.h"
#include "impeghd_hoa_rom.h"
#include "ia_core_coder_config.h"
#include "impeghd_hoa_spatial_decoder_struct.h"
#include "impeghd_hoa_amb_syn.h"
#include "impeghd_intrinsics_flt.h"
#include "impeghd_tbe_dec.h"
#include "impeghd_hoa_nfc_filtering.h"
#include "impeghd_hoa_space_positions.h"
#include "impeghd_hoa_simple_mtrx.h"
#include "impeghd_hoa_render_mtrx.h"
#include "impeghd_hoa_renderer.h"

typedef struct
{
  ia_render_hoa_str rh;
  ia_speaker_config_3d spk;
  FLOAT32 trans[64 * 4096];
  char scratch[8 << 20];
} ref_hoa_ren_handle;

/* Designs the render matrix for CICP layout `cicp` and HOA order `order` with the reference's own init.
 * nfc_radius > 0 enables near-field compensation (impeghd_hoa_ren_nfc_init) with that radius in metres. */
void *ref_hoa_ren_create(int cicp, int order, int lfe_enabled, float nfc_radius, int sample_rate, int *err_out)
{
  ref_hoa_ren_handle *h = (ref_hoa_ren_handle *)calloc(1, sizeof(ref_hoa_ren_handle));
  WORD32 orders[1];
  IA_ERRORCODE err;
  if (!h)
    return 0;
  orders[0] = order;
  err = impeghd_hoa_ren_renderer_init(&h->rh, cicp, &h->spk, orders, 1, lfe_enabled, h->scratch);
  if (!err)
  {
    h->rh.mtrx_selected = 0;
    h->rh.use_nfc = 0;
    if (nfc_radius > 0.0f)
    {
      err = impeghd_hoa_ren_nfc_init(&h->rh, (UWORD32)order, nfc_radius, (UWORD32)sample_rate);
      h->rh.use_nfc = 1;
    }
  }
  if (err_out)
    *err_out = (int)err;
  return h;
}
void ref_hoa_ren_destroy(void *p) { free(p); }

/* info[4] = {num_out_channels, rows, cols, ldspk_dist as float bits}; mtrx [rows*cols] row-major */
void ref_hoa_ren_matrix(void *p, int *info, float *mtrx)
{
  ref_hoa_ren_handle *h = (ref_hoa_ren_handle *)p;
  ia_render_hoa_render_mtrx_str *m = &h->rh.v_render_mtrx[h->rh.mtrx_selected < 0 ? 0 : h->rh.mtrx_selected];
  info[0] = h->rh.num_out_channels;
  info[1] = m->sm_mtrx.rows;
  info[2] = m->sm_mtrx.cols;
  memcpy(&info[3], &h->rh.ldspk_dist, 4);
  if (mtrx)
    memcpy(mtrx, m->sm_mtrx.mtrx, sizeof(float) * m->sm_mtrx.rows * m->sm_mtrx.cols);
}

/* per coefficient channel mn: out[mn][16] = {global_gain, n2, n1, (a1,a2,b1,b2) x 3, (a1,b1)} */
void ref_hoa_ren_nfc_coeffs(void *p, int n_coef, float *out)
{
  ref_hoa_ren_handle *h = (ref_hoa_ren_handle *)p;
  int mn, n;
  for (mn = 0; mn < n_coef; mn++)
  {
    ia_render_hoa_nfc_filtering_str *f = &h->rh.nfc_filter[mn];
    float *o = out + mn * 17;
    memset(o, 0, 17 * sizeof(float));
    o[0] = f->global_gain;
    o[1] = (float)f->num_iir_2_filts;
    o[2] = (float)f->num_iir_1_filts;
    for (n = 0; n < f->num_iir_2_filts && n < 3; n++)
    {
      o[3 + 4 * n + 0] = f->str_iir_2[n].a1;
      o[3 + 4 * n + 1] = f->str_iir_2[n].a2;
      o[3 + 4 * n + 2] = f->str_iir_2[n].b1;
      o[3 + 4 * n + 3] = f->str_iir_2[n].b2;
    }
    o[15] = f->str_iir_1.a1;
    o[16] = f->str_iir_1.b1;
  }
}

/* one frame: in [rows][n] (copied: NFC filters in place), out [num_out_channels][n] */
int ref_hoa_ren_process(void *p, const float *in, int rows, int n, int clip, float *out)
{
  ref_hoa_ren_handle *h = (ref_hoa_ren_handle *)p;
  static __thread float tmp[64 * 4096];
  memcpy(tmp, in, sizeof(float) * rows * n);
  return (int)impeghd_hoa_ren_block_process(&h->rh, tmp, rows, (UWORD32)n, clip, h->trans, out);
}
