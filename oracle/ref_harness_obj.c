/* oracle/ref_harness_obj.c -- TEST INFRASTRUCTURE ONLY.
 *
 * Function-level entry points into the UNMODIFIED reference object renderer (linked into
 * oracle/_ref/libmpeghref.so): impeghd_obj_renderer_dec_init (impeghd_obj_ren_dec.c:1057) and
 * impeghd_obj_renderer_dec (:1319).  Contains no reference code, only calls into it.
 */
#include <stdlib.h>
#include <string.h>

#include <impeghd_type_def.h>
#include "impd_drc_common.h"
#include "impd_drc_extr_delta_coded_info.h"
#include "impd_drc_struct.h"
#include "impeghd_3d_vec_struct_def.h"
#include "impeghd_cicp_defines.h"
#include "impeghd_cicp_struct_def.h"
#include "impeghd_cicp_2_geometry.h"
#include "impeghd_cicp_2_geometry_rom.h"
#include "impeghd_error_codes.h"
#include "ia_core_coder_bitbuffer.h"
#include "ia_core_coder_cnst.h"
#include "impeghd_hoa_common_values.h"
#include "impeghd_hoa_matrix_struct.h"
#include "impeghd_hoa_matrix.h"
#include "impeghd_intrinsics_flt.h"
#include "ia_core_coder_config.h"
#include "impeghd_mhas_parse.h"
#include "impeghd_binaural.h"
#include "impeghd_ele_interaction_intrfc.h"
#include "impeghd_oam_dec_defines.h"
#include "impeghd_oam_dec_struct_def.h"
#include "impeghd_obj_ren_dec_defines.h"
#include "impeghd_obj_ren_dec_struct_def.h"
#include "impeghd_obj_ren_dec.h"
#include "impeghd_3d_vec_basic_ops.h"

typedef struct
{
  ia_obj_ren_dec_state_struct st;
  ia_local_setup_struct local_setup;
  ia_oam_dec_config_struct md_cfg;
  ia_speaker_config_3d spk;
  char scratch[1 << 20];
} ref_obj_handle;

void *ref_obj_create(int cicp_idx, int *err_out)
{
  ref_obj_handle *h = (ref_obj_handle *)calloc(1, sizeof(ref_obj_handle));
  IA_ERRORCODE err;
  if (!h)
    return 0;
  h->st.cicp_out_idx = cicp_idx;
  h->st.pstr_local_setup = &h->local_setup;
  h->st.str_obj_md_dec_state.p_obj_md_cfg = &h->md_cfg;
  h->md_cfg.frame_length = 1024;
  err = impeghd_obj_renderer_dec_init(&h->st, &h->spk, h->scratch);
  if (err_out)
    *err_out = (int)err;
  return h;
}
void ref_obj_destroy(void *p) { free(p); }

/* Layout descriptor read back from the initialised state.
 * counts[8] = {num_cicp_speakers, num_non_lfe_ls, num_imag_ls, num_triangles, height_spks_present, 0,0,0}
 * lfe[44], tri[44][3], inv[44][9] (row-major rows 0..2 = x,y,z), cart[44][3], dmx[44][44]. */
void ref_obj_layout(void *p, int *counts, int *lfe, int *tri, float *inv, float *cart, float *dmx)
{
  ref_obj_handle *h = (ref_obj_handle *)p;
  ia_obj_ren_dec_state_struct *s = &h->st;
  int i, j;
  memset(counts, 0, 8 * sizeof(int));
  counts[0] = s->num_cicp_speakers;
  counts[1] = s->num_non_lfe_ls;
  counts[2] = s->num_imag_ls;
  counts[3] = s->num_triangles;
  counts[4] = s->height_spks_present;
  for (i = 0; i < CICP_MAX_NUM_LS; i++)
  {
    lfe[i] = (i < s->num_cicp_speakers && s->ptr_cicp_ls_geo[i]) ? s->ptr_cicp_ls_geo[i]->lfe_flag : 0;
    for (j = 0; j < 3; j++)
    {
      tri[i * 3 + j] = s->ls_triangle[i].spk_index[j];
      inv[i * 9 + j * 3 + 0] = s->ls_inv_mtx[i].row[j].x;
      inv[i * 9 + j * 3 + 1] = s->ls_inv_mtx[i].row[j].y;
      inv[i * 9 + j * 3 + 2] = s->ls_inv_mtx[i].row[j].z;
    }
    cart[i * 3 + 0] = s->non_lfe_ls_str[i].ls_cart_coord.x;
    cart[i * 3 + 1] = s->non_lfe_ls_str[i].ls_cart_coord.y;
    cart[i * 3 + 2] = s->non_lfe_ls_str[i].ls_cart_coord.z;
    for (j = 0; j < CICP_MAX_NUM_LS; j++)
      dmx[i * CICP_MAX_NUM_LS + j] = s->downmix_mtx[i][j];
  }
}

/* first_frame_flag as the reference initialises it (1 for the first MAX_NUM_OAM_OBJS entries only) */
void ref_obj_first_flags(void *p, int *flags, int n)
{
  ref_obj_handle *h = (ref_obj_handle *)p;
  int i;
  for (i = 0; i < n; i++)
    flags[i] = h->st.first_frame_flag[i];
}

/* One frame.  params[n_obj][7] = {azimuth, elevation, radius, gain, spread_width, spread_height,
 * spread_depth} (the *_descaled floats); in [n_obj][1024]; out [num_cicp_speakers][1024] (zeroed here like
 * impeghd_obj_md_dec_ren_process does, decode_main.c:1374).  gains_out (optional) [n_obj][44] receives the
 * final gains of each object (state->initial_gains after the call). */
int ref_obj_render(void *p, const float *params, int n_obj, const float *in, float *out, float *gains_out)
{
  ref_obj_handle *h = (ref_obj_handle *)p;
  ia_oam_dec_state_struct *md = &h->st.str_obj_md_dec_state;
  IA_ERRORCODE err;
  int o;
  md->num_objects = n_obj;
  md->sub_frame_number = 1;
  for (o = 0; o < n_obj; o++)
  {
    const float *q = params + o * 7;
    md->azimuth_descaled[o] = q[0];
    md->elevation_descaled[o] = q[1];
    md->radius_descaled[o] = q[2];
    md->gain_descaled[o] = q[3];
    md->spread_width_descaled[o] = q[4];
    md->spread_height_descaled[o] = q[5];
    md->spread_depth_descaled[o] = q[6];
  }
  memset(out, 0, sizeof(float) * 1024 * h->st.num_cicp_speakers);
  err = impeghd_obj_renderer_dec(&h->st, (FLOAT32 *)in, out, 1024);
  if (gains_out)
    for (o = 0; o < n_obj; o++)
      memcpy(gains_out + o * CICP_MAX_NUM_LS, h->st.initial_gains[o], sizeof(float) * CICP_MAX_NUM_LS);
  return (int)err;
}

/* One core frame with OAM sub-frames: the loop of impeghd_obj_md_dec_ren_process (ia_core_coder_decode_main.c:1383-1410)
 * around the reference's impeghd_obj_renderer_dec.  params [n_sub][n_obj][7] = the metadata sets of the frame's
 * sub-frames as the OAM parser would have left them (index sub * n_obj + obj), present [n_sub] =
 * sub_frame_obj_md_present; frame_len = 256 or 512 (n_sub = 1024 / frame_len).  used_set [n_sub] (optional) receives
 * the index of the set each call rendered with (sub_frame_number - 1). */
int ref_obj_render_subframes(void *p, const float *params, const int *present, int n_obj, int frame_len, const float *in,
                             float *out, int *used_set)
{
  ref_obj_handle *h = (ref_obj_handle *)p;
  ia_oam_dec_state_struct *md = &h->st.str_obj_md_dec_state;
  const int num_iter = 1024 / frame_len;
  IA_ERRORCODE err = 0;
  int i, o;
  md->num_objects = n_obj;
  h->md_cfg.frame_length = frame_len;
  for (i = 0; i < num_iter; i++)
    for (o = 0; o < n_obj; o++)
    {
      const float *q = params + ((size_t)i * n_obj + o) * 7;
      const int k = i * n_obj + o;
      md->azimuth_descaled[k] = q[0];
      md->elevation_descaled[k] = q[1];
      md->radius_descaled[k] = q[2];
      md->gain_descaled[k] = q[3];
      md->spread_width_descaled[k] = q[4];
      md->spread_height_descaled[k] = q[5];
      md->spread_depth_descaled[k] = q[6];
    }
  memset(out, 0, sizeof(float) * 1024 * h->st.num_cicp_speakers);
  md->sub_frame_number = 0;
  for (i = 0; i < num_iter; i++)
  {
    if (present[i])
    {
      md->sub_frame_number++;
      if (md->sub_frame_number > num_iter)
        md->sub_frame_number = md->sub_frame_number % num_iter;
    }
    else if (i == 0)
    {
      md->sub_frame_number = num_iter;
    }
    if (used_set)
      used_set[i] = md->sub_frame_number - 1;
    err |= impeghd_obj_renderer_dec(&h->st, (FLOAT32 *)in + i * frame_len, out + i * frame_len, 1024);
  }
  h->md_cfg.frame_length = 1024;
  return (int)err;
}
