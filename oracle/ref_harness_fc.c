/* oracle/ref_harness_fc.c -- TEST INFRASTRUCTURE ONLY.
 *
 * Function-level entry points into the UNMODIFIED reference format converter (linked into
 * oracle/_ref/libmpeghref.so):
 *   impeghd_format_conv_init_data      decoder/impeghd_format_conv_init.c:477  (rule-based downmix matrix + EQs)
 *   impeghd_format_conv_dmx_init       decoder/impeghd_format_conv_init.c:81
 *   impeghd_format_converter_process   decoder/ia_core_coder_process.c:641     (STFT, active downmix, OLA)
 * The pointer wiring of impeghd_format_conv_init (decoder/ia_core_coder_decode_main.c:519, which needs a parsed
 * audio specific config) is redone here on bare structs.  No reference code, only calls.
 */
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

IA_ERRORCODE impeghd_format_converter_process(ia_mpegh_dec_api_struct *p_obj_mpegh_dec, WORD32 *num_chn_out,
                                              FLOAT32 *scratch_buf);

typedef struct
{
  ia_mpegh_dec_api_struct api;
  ia_mpegh_dec_state_struct state;
  ia_format_conv_struct fc;
  ia_speaker_config_3d spk;
  ia_cicp_ls_geo_str azi_elev[CICP_MAX_CH];
  WORD32 in_names[FC_MAX_CHANNELS];
  FLOAT32 init_scratch[4096];
  ia_dec_data_struct *dec;
  FLOAT32 *scratch;
  ia_obj_ren_dec_state_struct obj;
  ia_local_setup_struct local_setup;
  ia_oam_dec_config_struct md_cfg;
} ref_fc_handle;

/* cicp_in/cicp_out: CICP loudspeaker layout indices; returns the handle, err_out = reference error code */
void *ref_fc_create(int cicp_in, int cicp_out, int samp_rate, int *err_out)
{
  ref_fc_handle *h = (ref_fc_handle *)calloc(1, sizeof(ref_fc_handle));
  ia_format_conv_struct *fc;
  ia_format_conv_param *p;
  int i, j, err;
  if (!h)
    return 0;
  h->dec = (ia_dec_data_struct *)calloc(1, sizeof(ia_dec_data_struct));
  h->scratch = (FLOAT32 *)calloc(1 << 20, sizeof(FLOAT32));
  fc = &h->fc;
  p = &fc->fc_params;
  h->api.p_state_mpeghd = &h->state;
  h->state.pstr_dec_data = h->dec;
  p->active_dmx_stft = &fc->active_dmx_stft;
  p->downmix_mat = (FLOAT32 **)&fc->downmix_mat_ptr[0];
  h->state.state_format_conv.stft_out_buf = (FLOAT32 **)&fc->stft_output_buf_ptr[0];
  h->state.state_format_conv.stft_in_buf = (FLOAT32 **)&fc->stft_input_buf_ptr[0];
  for (i = 0; i < FC_MAX_CHANNELS; i++)
  {
    fc->stft_input_buf_ptr[i] = &fc->stft_input_buf[i * FC_STFT_FRAME];
    fc->downmix_mat_ptr[i] = &fc->downmix_mat[i * FC_MAX_CHANNELS];
    fc->stft_output_buf_ptr[i] = &fc->stft_output_buf[i * FC_STFT_FRAME];
  }
  p->active_dmx_stft->downmix_mat = (FLOAT32 ***)&fc->active_downmix_mat_ptr[0];
  for (i = 0; i < FC_MAX_CHANNELS; i++)
  {
    p->active_dmx_stft->downmix_mat[i] = (FLOAT32 **)&fc->active_downmix_mat_ch_ptr[i * FC_MAX_CHANNELS];
    for (j = 0; j < FC_MAX_CHANNELS; j++)
      p->active_dmx_stft->downmix_mat[i][j] =
          (FLOAT32 *)&fc->active_downmix_mat_freq_ptr[(i * FC_MAX_CHANNELS * FC_BANDS) + j * FC_BANDS];
  }
  p->data_state = &fc->format_conv_params_int;
  h->state.state_format_conv.fc_params = p;
  p->samp_rate = samp_rate;
  p->num_in_ch = impgehd_cicp_get_num_ls[cicp_in];
  p->num_out_ch = impgehd_cicp_get_num_ls[cicp_out];
  for (i = 0; i < p->num_in_ch; i++)
    h->in_names[i] = ia_cicp_idx_ls_set_map_tbl[cicp_in][i];
  p->in_ch = h->in_names;
  p->out_ch = ia_cicp_idx_ls_set_map_tbl[cicp_out];
  h->spk.cicp_spk_layout_idx = cicp_in;
  h->spk.spk_layout_type = 0;
  err = (int)impeghd_format_conv_init_data(h->init_scratch, h->azi_elev, p, &h->spk);
  if ((unsigned)err == IA_MPEGH_FORMAT_CONV_INIT_FATAL_INVALID_CHANNEL_NUM)
  {
    /* the rule table does not cover this input layout: the reference pans every input loudspeaker with VBAP
     * instead (ia_core_coder_decode_main.c:659-735); same calls in the same order */
    ia_cicp_ls_geo_str *in_geo[CICP_MAX_NUM_LS], *out_geo[CICP_MAX_NUM_LS];
    const WORD32 *in_names, *out_names;
    WORD32 n_spk_in, n_lfe_in, n_spk_out, n_lfe_out, chn, o;
    h->obj.str_obj_md_dec_state.p_obj_md_cfg = &h->md_cfg;
    h->obj.str_obj_md_dec_state.num_objects = p->num_out_ch;
    h->obj.pstr_local_setup = &h->local_setup;
    h->obj.cicp_out_idx = cicp_out;
    err = (int)impeghd_obj_renderer_dec_init(&h->obj, &h->spk, (pVOID)h->scratch);
    impeghd_cicpidx_2_ls_geometry(cicp_in, (const ia_cicp_ls_geo_str **)&in_geo[0], &n_spk_in, &n_lfe_in, &in_names);
    impeghd_cicpidx_2_ls_geometry(cicp_out, (const ia_cicp_ls_geo_str **)&out_geo[0], &n_spk_out, &n_lfe_out,
                                  &out_names);
    for (chn = 0; chn < n_spk_in && !err; chn++)
    {
      h->obj.str_obj_md_dec_state.gain_descaled[chn] = 1;
      h->obj.str_obj_md_dec_state.elevation_descaled[chn] = in_geo[chn]->ls_elevation;
      h->obj.str_obj_md_dec_state.azimuth_descaled[chn] = in_geo[chn]->ls_azimuth;
      h->obj.str_obj_md_dec_state.radius_descaled[chn] = 1;
      h->obj.str_obj_md_dec_state.spread_depth_descaled[chn] = 0;
      h->obj.str_obj_md_dec_state.spread_height_descaled[chn] = 0;
      h->obj.str_obj_md_dec_state.spread_width_descaled[chn] = 0;
      err = (int)impegh_obj_ren_vbap_process(&h->obj, chn);
      for (o = 0; o < p->num_out_ch; o++)
        p->downmix_mat[chn][o] = h->obj.final_gains[o];
    }
    if (!err)
      err = (int)impeghd_format_conv_map_lfes(p, (const ia_cicp_ls_geo_str **)&in_geo[0],
                                              (const ia_cicp_ls_geo_str **)&out_geo[0]);
    if (!err)
      err = (int)impeghd_format_conv_post_proc_dmx_mtx(p);
  }
  if (!err)
    err = (int)impeghd_format_conv_dmx_init(p->active_dmx_stft, p);
  if (err_out)
    *err_out = err;
  return h;
}
void ref_fc_destroy(void *q)
{
  ref_fc_handle *h = (ref_fc_handle *)q;
  free(h->dec);
  free(h->scratch);
  free(h);
}
void ref_fc_info(void *q, int *info)
{
  ref_fc_handle *h = (ref_fc_handle *)q;
  info[0] = h->fc.fc_params.num_in_ch;
  info[1] = h->fc.fc_params.num_out_ch;
}
/* D [n_out][n_in][257] */
void ref_fc_get_matrix(void *q, float *D)
{
  ref_fc_handle *h = (ref_fc_handle *)q;
  ia_format_conv_dmx_state *a = &h->fc.active_dmx_stft;
  int o, i;
  for (o = 0; o < a->num_out_ch; o++)
    for (i = 0; i < a->num_in_ch; i++)
      memcpy(D + ((size_t)o * a->num_in_ch + i) * FC_BANDS, a->downmix_mat[o][i], sizeof(float) * FC_BANDS);
}
/* replace shape and matrix (arbitrary test matrices), state reset */
void ref_fc_set_matrix(void *q, int n_in, int n_out, const float *D)
{
  ref_fc_handle *h = (ref_fc_handle *)q;
  ia_format_conv_dmx_state *a = &h->fc.active_dmx_stft;
  int o, i;
  h->fc.fc_params.num_in_ch = a->num_in_ch = n_in;
  h->fc.fc_params.num_out_ch = a->num_out_ch = n_out;
  for (o = 0; o < n_out; o++)
    for (i = 0; i < n_in; i++)
      memcpy(a->downmix_mat[o][i], D + ((size_t)o * n_in + i) * FC_BANDS, sizeof(float) * FC_BANDS);
  memset(a->target_energy_prev, 0, sizeof(a->target_energy_prev));
  memset(a->realized_energy_prev, 0, sizeof(a->realized_energy_prev));
  memset(h->fc.stft_input_buf, 0, sizeof(h->fc.stft_input_buf));
  memset(h->fc.stft_output_buf, 0, sizeof(h->fc.stft_output_buf));
}
/* one frame: in [n_in][1024] -> out [n_out][1024] (the reference works in place on time_sample_vector) */
int ref_fc_process(void *q, const float *in, float *out)
{
  ref_fc_handle *h = (ref_fc_handle *)q;
  ia_usac_data_struct *u = &h->dec->str_usac_data;
  int n_in = h->fc.fc_params.num_in_ch, n_out = 0, ch, err;
  for (ch = 0; ch < n_in; ch++)
    memcpy(u->time_sample_vector[ch], in + (size_t)ch * 1024, sizeof(float) * 1024);
  err = (int)impeghd_format_converter_process(&h->api, &n_out, h->scratch);
  for (ch = 0; ch < n_out; ch++)
    memcpy(out + (size_t)ch * 1024, u->time_sample_vector[ch], sizeof(float) * 1024);
  return err;
}

/* ---- domain switcher (impeghd_domain_switcher_process, decoder/ia_core_coder_process.c:505) -------------------- */
IA_ERRORCODE impeghd_domain_switcher_process(ia_mpegh_dec_api_struct *p_obj_mpegh_dec, WORD32 *num_chn_out,
                                             UWORD32 ds_t2f, FLOAT32 *scratch_buf);
typedef struct
{
  ia_mpegh_dec_api_struct api;
  ia_mpegh_dec_state_struct state;
  ia_ds_param params;
  FLOAT32 *in_ptr[MAX_NUM_CHANNELS], *out_ptr[MAX_NUM_CHANNELS];
  FLOAT32 in_buf[MAX_NUM_CHANNELS][FC_STFT_FRAME], out_buf[MAX_NUM_CHANNELS][FC_STFT_FRAME];
  ia_dec_data_struct *dec;
  FLOAT32 *scratch, *stft;
  int n_ch;
} ref_ds_handle;

void *ref_ds_create(int n_ch)
{
  ref_ds_handle *h = (ref_ds_handle *)calloc(1, sizeof(ref_ds_handle));
  h->dec = (ia_dec_data_struct *)calloc(1, sizeof(ia_dec_data_struct));
  h->scratch = (FLOAT32 *)calloc(1 << 20, sizeof(FLOAT32));
  h->stft = (FLOAT32 *)calloc((size_t)MAX_NUM_CHANNELS * FC_STFT_FRAMEx2 * 4, sizeof(FLOAT32));
  h->n_ch = n_ch;
  h->api.p_state_mpeghd = &h->state;
  h->state.pstr_dec_data = h->dec;
  h->params.num_in_ch = h->params.num_out_ch = n_ch;
  h->state.state_domain_switcher.ds_params = &h->params;
  h->state.state_domain_switcher.stft_in_buf = h->in_ptr;
  h->state.state_domain_switcher.stft_out_buf = h->out_ptr;
  for (int c = 0; c < MAX_NUM_CHANNELS; c++)
  {
    h->in_ptr[c] = h->in_buf[c], h->out_ptr[c] = h->out_buf[c];
    h->dec->str_usac_data.time_sample_stft[c] = h->stft + (size_t)c * FC_STFT_FRAMEx2 * 4; /* decode_main.c:2121 */
  }
  return h;
}
void ref_ds_destroy(void *q)
{
  ref_ds_handle *h = (ref_ds_handle *)q;
  free(h->dec), free(h->scratch), free(h->stft), free(h);
}
/* x [n_ch][1024] -> re, im [n_ch][1024] (4 slots x 256) */
int ref_ds_t2f(void *q, const float *x, float *re, float *im)
{
  ref_ds_handle *h = (ref_ds_handle *)q;
  ia_usac_data_struct *u = &h->dec->str_usac_data;
  int n_out = 0, err;
  for (int c = 0; c < h->n_ch; c++) memcpy(u->time_sample_vector[c], x + (size_t)c * 1024, 4096);
  err = (int)impeghd_domain_switcher_process(&h->api, &n_out, 1, h->scratch);
  for (int c = 0; c < h->n_ch; c++)
  {
    memcpy(re + (size_t)c * 1024, u->time_sample_stft[c], 4096);
    memcpy(im + (size_t)c * 1024, u->time_sample_stft[c] + 1024, 4096);
  }
  return err;
}
int ref_ds_f2t(void *q, const float *re, const float *im, float *y)
{
  ref_ds_handle *h = (ref_ds_handle *)q;
  ia_usac_data_struct *u = &h->dec->str_usac_data;
  int n_out = 0, err;
  for (int c = 0; c < h->n_ch; c++)
  {
    memcpy(u->time_sample_stft[c], re + (size_t)c * 1024, 4096);
    memcpy(u->time_sample_stft[c] + 1024, im + (size_t)c * 1024, 4096);
  }
  err = (int)impeghd_domain_switcher_process(&h->api, &n_out, 0, h->scratch);
  for (int c = 0; c < h->n_ch; c++) memcpy(y + (size_t)c * 1024, u->time_sample_vector[c], 4096);
  return err;
}
