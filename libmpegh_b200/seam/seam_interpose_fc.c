/* seam_interpose_fc.c -- libmpeghb200_seam.so, format converter.
 *
 * impeghd_format_converter_process (decoder/ia_core_coder_process.c:641) is what the reference calls when the
 * output loudspeaker layout differs from the stream's (testbench option -cicp:).  This definition takes its place:
 * same signature, same error returns for an uninitialised converter, the STFT-domain active downmix done by
 * mpeghb200_format_conv_dev.  The downmix tensor is read once from the handle the reference's own init designed
 * (active_dmx_stft->downmix_mat); the overlap buffers and smoothed energies stay in the reference handle and make
 * the round trip with every call, like the FD overlap state in seam_interpose.c.
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


#include "seam_internal.h"

#define FRAME 1024
#define ERB 58

/* converters by content: handles that decode like streams to the same layout share one (the downmix tensor is a pure
 * function of the layout pair) */
#define MAX_FC 64
static struct
{
  int n_in, n_out;
  unsigned long long hash;
  mpeghb200_format_conv *fc;
} conv[MAX_FC];
static int n_conv;
static struct
{
  float *d_in, *d_out, *d_state;
  float *h_state, *h_D;
  size_t state_floats;
} c;

IA_ERRORCODE
impeghd_format_converter_process(ia_mpegh_dec_api_struct *p_obj_mpegh_dec, WORD32 *num_chn_out, FLOAT32 *scratch_buf)
{
  ia_mpegh_dec_state_struct *st = p_obj_mpegh_dec->p_state_mpeghd;
  ia_format_conv_state_struct *fcs = &st->state_format_conv;
  ia_dec_data_struct *dec = (ia_dec_data_struct *)st->pstr_dec_data;
  ia_usac_data_struct *u = &dec->str_usac_data;
  ia_format_conv_dmx_state *a;
  mpeghb200_ctx *ctx;
  int n_in, n_out, ch, k;
  float *p;
  mpeghb200_format_conv *fc;
  (void)scratch_buf;
  if (fcs->fc_params == NULL)
    return IA_MPEGH_FORMAT_CONV_INIT_FATAL_INVALID_PARAM;
  n_in = fcs->fc_params->num_in_ch;
  n_out = fcs->fc_params->num_out_ch;
  *num_chn_out = n_out;
  a = fcs->fc_params->active_dmx_stft;
  if (n_out <= 0 || n_in <= 0)
    return IA_MPEGH_FORMAT_CONV_INIT_FATAL_INVALID_CHANNEL_NUM;
  ctx = seam_ctx();
  seam_inline_lock();
  if (!c.d_in)
  {
    CK(mpeghb200_dev_alloc(ctx, (size_t)FC_MAX_CHANNELS * FRAME * sizeof(float), (void **)&c.d_in));
    CK(mpeghb200_dev_alloc(ctx, (size_t)FC_MAX_CHANNELS * FRAME * sizeof(float), (void **)&c.d_out));
    CK(mpeghb200_dev_alloc(ctx, (size_t)FC_MAX_CHANNELS * (512 + 2 * ERB) * sizeof(float), (void **)&c.d_state));
    c.h_state = (float *)malloc((size_t)FC_MAX_CHANNELS * (512 + 2 * ERB) * sizeof(float));
    c.h_D = (float *)malloc((size_t)FC_MAX_CHANNELS * FC_MAX_CHANNELS * FC_BANDS * sizeof(float));
  }
  { /* take over the downmix tensor the reference designed (first frame with this layout pair: create the converter) */
    unsigned long long h = 1469598103934665603ull;
    const unsigned char *b;
    size_t nb = (size_t)n_out * n_in * FC_BANDS * sizeof(float), q;
    int o, i;
    for (o = 0; o < n_out; o++)
      for (i = 0; i < n_in; i++)
        memcpy(c.h_D + ((size_t)o * n_in + i) * FC_BANDS, a->downmix_mat[o][i], sizeof(float) * FC_BANDS);
    b = (const unsigned char *)c.h_D;
    for (q = 0; q < nb; q++)
      h = (h ^ b[q]) * 1099511628211ull;
    for (k = 0; k < n_conv; k++)
      if (conv[k].n_in == n_in && conv[k].n_out == n_out && conv[k].hash == h)
        break;
    if (k == n_conv)
    {
      if (n_conv == MAX_FC)
      {
        mpeghb200_format_conv_destroy(conv[0].fc);
        conv[0] = conv[--n_conv];
        k = n_conv;
      }
      conv[k].n_in = n_in, conv[k].n_out = n_out, conv[k].hash = h;
      CK(mpeghb200_format_conv_create(ctx, n_in, n_out, c.h_D, &conv[k].fc));
      n_conv++;
    }
    fc = conv[k].fc;
    c.state_floats = mpeghb200_format_conv_state_floats(fc);
  }
  /* state layout of mpeghb200_format_conv_dev: in_prev [n_in][256], out_prev [n_out][256], t_prev, r_prev [n_out][58] */
  p = c.h_state;
  for (ch = 0; ch < n_in; ch++, p += FC_STFT_FRAME)
    memcpy(p, fcs->stft_in_buf[ch], FC_STFT_FRAME * sizeof(float));
  for (ch = 0; ch < n_out; ch++, p += FC_STFT_FRAME)
    memcpy(p, fcs->stft_out_buf[ch], FC_STFT_FRAME * sizeof(float));
  for (ch = 0; ch < n_out; ch++, p += ERB)
    memcpy(p, a->target_energy_prev[ch], ERB * sizeof(float));
  for (ch = 0; ch < n_out; ch++, p += ERB)
    memcpy(p, a->realized_energy_prev[ch], ERB * sizeof(float));
  CK(mpeghb200_h2d(ctx, c.d_state, c.h_state, c.state_floats * sizeof(float)));
  for (ch = 0; ch < n_in; ch++)
    CK(mpeghb200_h2d(ctx, c.d_in + (size_t)ch * FRAME, u->time_sample_vector[ch], FRAME * sizeof(float)));
  CK(mpeghb200_format_conv_dev(ctx, fc, c.d_in, c.d_out, c.d_state, 1, NULL));
  for (ch = 0; ch < n_out; ch++)
    CK(mpeghb200_d2h(ctx, u->time_sample_vector[ch], c.d_out + (size_t)ch * FRAME, FRAME * sizeof(float)));
  CK(mpeghb200_d2h(ctx, c.h_state, c.d_state, c.state_floats * sizeof(float)));
  p = c.h_state;
  for (ch = 0; ch < n_in; ch++, p += FC_STFT_FRAME)
    memcpy(fcs->stft_in_buf[ch], p, FC_STFT_FRAME * sizeof(float));
  for (ch = 0; ch < n_out; ch++, p += FC_STFT_FRAME)
    memcpy(fcs->stft_out_buf[ch], p, FC_STFT_FRAME * sizeof(float));
  for (ch = 0; ch < n_out; ch++, p += ERB)
    memcpy(a->target_energy_prev[ch], p, ERB * sizeof(float));
  for (ch = 0; ch < n_out; ch++, p += ERB)
    memcpy(a->realized_energy_prev[ch], p, ERB * sizeof(float));
  st->ui_out_bytes = 1024 * n_out * sizeof(FLOAT32);
  seam_n.fc++;
  seam_inline_unlock();
  return IA_MPEGH_DEC_NO_ERROR;
}
