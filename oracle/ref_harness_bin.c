/* oracle/ref_harness_bin.c -- TEST INFRASTRUCTURE ONLY.
 *
 * Function-level entry points into the UNMODIFIED reference binaural renderer and its long transforms
 * (linked into oracle/_ref/libmpeghref.so):
 *   impeghd_cplx_ifft_8k                      decoder/impeghd_ifft.c:1264
 *   impeghd_cplx_fft_16k                      decoder/impeghd_fft.c:1415
 *   impeghd_binaural_renderer_init            decoder/impeghd_binaural_renderer.c:346   (<= 21 inputs)
 *   impeghd_binaural_struct_process_ld_ola    decoder/impeghd_binaural_renderer.c:543
 * No reference code, only calls.  For more than 21 inputs (MAX_NUM_BRIR_PAIRS limits the init path, the
 * persistent struct holds 32) the persistent struct is filled here with the reference's own transform;
 * tests check that this fill equals what impeghd_binaural_renderer_init produces for 21 inputs.
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <impeghd_type_def.h>
#include "impd_drc_common.h"
#include "impd_drc_extr_delta_coded_info.h"
#include "impd_drc_struct.h"
#include "impeghd_error_codes.h"
#include "impeghd_intrinsics_flt.h"
#include "ia_core_coder_defines.h"
#include "ia_core_coder_bitbuffer.h"
#include "ia_core_coder_interface.h"
#include "ia_core_coder_info.h"
#include "ia_core_coder_rom.h"
#include "ia_core_coder_dec.h"
#include "ia_core_coder_apicmd_standards.h"
#include "ia_core_coder_api_defs.h"
#include "ia_core_coder_channel.h"
#include "ia_core_coder_channelinfo.h"
#include "ia_core_coder_cnst.h"
#include <ia_core_coder_constants.h>
#include "ia_core_coder_definitions.h"
#include "ia_core_coder_env_extr.h"
#include "ia_core_coder_acelp_info.h"
#include "impeghd_memory_standards.h"
#include "ia_core_coder_tns_usac.h"
#include "ia_core_coder_igf_data_struct.h"
#include "ia_core_coder_stereo_lpd.h"
#include "impeghd_tbe_dec.h"
#include "ia_core_coder_main.h"
#include "ia_core_coder_arith_dec.h"
#include "ia_core_coder_func_def.h"
#include "impeghd_hoa_common_values.h"
#include "impeghd_hoa_matrix_struct.h"
#include "impeghd_hoa_matrix.h"
#include "ia_core_coder_config.h"
#include "impeghd_binaural.h"
#include "impeghd_binaural_filter_design.h"
#include "impeghd_binaural_renderer.h"

void ref_cplx_ifft_8k(float *re, float *im, int n)
{
  float *scratch = (float *)calloc(8 * 16384, sizeof(float));
  impeghd_cplx_ifft_8k(re, im, scratch, n);
  free(scratch);
}
/* the same with every float of the scratch preset to `fill` (what an earlier user of the shared scratch left there) */
void ref_cplx_ifft_8k_fill(float *re, float *im, int n, float fill)
{
  float *scratch = (float *)malloc(8 * 16384 * sizeof(float));
  int i;
  for (i = 0; i < 8 * 16384; i++)
    scratch[i] = fill;
  impeghd_cplx_ifft_8k(re, im, scratch, n);
  free(scratch);
}
void ref_cplx_fft_16k(float *re, float *im, int n)
{
  float *scratch = (float *)calloc(8 * 16384, sizeof(float));
  impeghd_cplx_fft_16k(re, im, scratch, n);
  free(scratch);
}

typedef struct
{
  ia_binaural_render_struct ren; /* holds str_binaural (persistent memory) */
  ia_binaural_scratch scratch;
  float *src[MAX_NUM_BRIR_CHANNELS];
  float *dst[BINAURAL_NB_OUTPUT];
  float out[BINAURAL_NB_OUTPUT][BINAURAL_FFT_MAXLEN];
  float work[BINAURAL_FFT_MAXLEN];
} ref_bin_handle;

/* taps_direct [2][n_in][len], taps_diffuse [n_blocks][2][len], direct_fc [2][n_in], diffuse_fc [2][n_blocks],
 * inv_diffuse_weight [n_in].  via_init != 0: go through impeghd_binaural_renderer_init (n_in <= 21). */
void *ref_bin_create(int n_in, int len, int n_blocks, const float *taps_direct, const float *taps_diffuse,
                     const float *direct_fc, const float *diffuse_fc, const float *inv_diffuse_weight,
                     int via_init, int *err_out)
{
  ref_bin_handle *h = (ref_bin_handle *)calloc(1, sizeof(ref_bin_handle));
  int out, inp, b, i, err = 0;
  if (!h)
    return 0;
  if (via_init)
  {
    ia_binaural_renderer *br = (ia_binaural_renderer *)calloc(1, sizeof(ia_binaural_renderer));
    ia_binaural_in_stream_cfg_str cfg;
    ia_td_binaural_ren_param_str *p = &br->binaural_rep.td_binaural_ren_param;
    br->ptr_scratch = (FLOAT32 *)&h->scratch;
    br->num_binaural_representation = 1;
    br->ptr_binaural_rep[0] = &br->binaural_rep;
    br->binaural_rep.brir_sampling_frequency = 48000;
    br->binaural_rep.is_hoa_data = 1; /* identity channel map, no geometry lookup */
    br->binaural_rep.binaural_data_format_id = 2;
    br->binaural_rep.brir_pairs = (WORD16)n_in;
    p->num_channel = n_in;
    p->len_direct = len;
    p->num_diffuse_block = n_blocks;
    for (out = 0; out < 2; out++)
    {
      for (inp = 0; inp < n_in; inp++)
      {
        p->ptr_direct_fc[out][inp] = direct_fc[out * n_in + inp];
        memcpy(p->ptr_taps_direct[out][inp], taps_direct + ((size_t)out * n_in + inp) * len, sizeof(float) * len);
      }
      for (b = 0; b < n_blocks; b++)
      {
        p->ptr_diffuse_fc[out][b] = diffuse_fc[out * n_blocks + b];
        memcpy(p->ptr_taps_diffuse[b][out], taps_diffuse + ((size_t)b * 2 + out) * len, sizeof(float) * len);
      }
    }
    for (inp = 0; inp < n_in; inp++)
      p->ptr_inv_diffuse_weight[inp] = inv_diffuse_weight[inp];
    cfg.fs_input = 48000;
    cfg.num_speaker_expected = n_in;
    cfg.cicp_spk_idx = 0;
    cfg.num_channel_is_input = n_in;
    err = (int)impeghd_binaural_renderer_init(br, &cfg, &h->ren);
    free(br);
  }
  else
  {
    /* the same steps as impeghd_binaural_struct_init/_set (:177,:244), using the reference transform */
    ia_binaural_ren_pers_mem_str *pb = &h->ren.str_binaural;
    int fft = 2;
    while (fft < 2 * len)
      fft <<= 1;
    pb->fft_size = pb->fft_len = pb->ifft_len = fft;
    pb->num_input = n_in;
    pb->num_output = 2;
    pb->diffuse_blocks = n_blocks;
    pb->direct_size = pb->diffuse_size = len;
    pb->filter_size = fft / 2;
    pb->spec_mul_size = fft;
    pb->processed_samples = pb->output_samples_size = fft - fft / 2;
    for (out = 0; out < 2; out++)
      for (inp = 0; inp < n_in; inp++)
      {
        float *t = pb->taps_direct[out][inp];
        float *xi = h->scratch.xi, *tmp = h->scratch.complex_scratch;
        float fc = direct_fc[out * n_in + inp];
        memset(t, 0, sizeof(float) * fft);
        memcpy(t, taps_direct + ((size_t)out * n_in + inp) * len, sizeof(float) * (len < fft / 2 ? len : fft / 2));
        memset(xi, 0, sizeof(float) * fft);
        impeghd_cplx_ifft_8k(t, xi, h->scratch.fft_ifft_scratch, fft);
        memcpy(tmp, t, sizeof(float) * fft);
        t[0] = tmp[0];
        t[1] = tmp[fft / 2];
        for (i = 1; i < fft / 2; i++)
        {
          t[2 * i] = tmp[i];
          t[2 * i + 1] = -1 * xi[i];
        }
        if (fc > 1.0f)
          fc = 1.0f;
        pb->direct_process_size[out][inp] = (WORD32)(fc * pb->filter_size);
      }
    for (b = 0; b < n_blocks; b++)
      for (out = 0; out < 2; out++)
      {
        float *t = pb->taps_diffuse[b][out];
        float *xi = h->scratch.xi, *tmp = h->scratch.complex_scratch;
        float fc = diffuse_fc[out * n_blocks + b];
        memset(t, 0, sizeof(float) * fft);
        memcpy(t, taps_diffuse + ((size_t)b * 2 + out) * len, sizeof(float) * (len < fft / 2 ? len : fft / 2));
        memset(xi, 0, sizeof(float) * fft);
        impeghd_cplx_ifft_8k(t, xi, h->scratch.fft_ifft_scratch, fft);
        memcpy(tmp, t, sizeof(float) * fft);
        t[0] = tmp[0];
        t[1] = tmp[fft / 2];
        for (i = 1; i < fft / 2; i++)
        {
          t[2 * i] = tmp[i];
          t[2 * i + 1] = -1 * xi[i];
        }
        if (fc > 1.0f)
          fc = 1.0f;
        pb->diffuse_process_size[b][out] = (WORD32)(fc * pb->filter_size);
      }
    for (inp = 0; inp < n_in; inp++)
      pb->diffuse_weight[inp] = inv_diffuse_weight[inp];
  }
  for (out = 0; out < 2; out++)
    h->dst[out] = h->out[out];
  if (err_out)
    *err_out = err;
  return h;
}
void ref_bin_destroy(void *p) { free(p); }

/* info[8] = {fft_size, filter_size, processed_samples, num_input, diffuse_blocks, spec_mul_size, write_idx,
 * memory_index} */
void ref_bin_info(void *p, int *info)
{
  ia_binaural_ren_pers_mem_str *pb = &((ref_bin_handle *)p)->ren.str_binaural;
  info[0] = pb->fft_size, info[1] = pb->filter_size, info[2] = pb->processed_samples, info[3] = pb->num_input;
  info[4] = pb->diffuse_blocks, info[5] = pb->spec_mul_size, info[6] = pb->write_idx, info[7] = pb->memory_index;
}
/* spectra as the reference keeps them: direct [2][n_in][fft], diffuse [n_blocks][2][fft],
 * sizes: direct_len [2][n_in], diffuse_len [n_blocks][2] */
void ref_bin_spectra(void *p, float *direct, float *diffuse, int *direct_len, int *diffuse_len)
{
  ia_binaural_ren_pers_mem_str *pb = &((ref_bin_handle *)p)->ren.str_binaural;
  int out, inp, b, fft = pb->fft_size, n_in = pb->num_input;
  for (out = 0; out < 2; out++)
    for (inp = 0; inp < n_in; inp++)
    {
      memcpy(direct + ((size_t)out * n_in + inp) * fft, pb->taps_direct[out][inp], sizeof(float) * fft);
      direct_len[out * n_in + inp] = pb->direct_process_size[out][inp];
    }
  for (b = 0; b < pb->diffuse_blocks; b++)
    for (out = 0; out < 2; out++)
    {
      memcpy(diffuse + ((size_t)b * 2 + out) * fft, pb->taps_diffuse[b][out], sizeof(float) * fft);
      diffuse_len[b * 2 + out] = pb->diffuse_process_size[b][out];
    }
}

/* One call of impeghd_binaural_struct_process_ld_ola with n samples per input: in [n_in][n], out [2][>= n_out].
 * Returns the number of output samples (0 until a block of processed_samples inputs has accumulated). */
int ref_bin_process(void *p, const float *in, int n, float *out, int out_stride)
{
  ref_bin_handle *h = (ref_bin_handle *)p;
  ia_binaural_ren_pers_mem_str *pb = &h->ren.str_binaural;
  int inp, o, got;
  for (inp = 0; inp < pb->num_input; inp++)
    h->src[inp] = (float *)in + (size_t)inp * n;
  got = impeghd_binaural_struct_process_ld_ola(h->src, h->dst, n, h->work, pb, (FLOAT32 *)&h->scratch);
  for (o = 0; o < 2; o++)
    memcpy(out + (size_t)o * out_stride, h->out[o], sizeof(float) * got);
  return got;
}
