/* seam_interpose_ren.c -- libmpeghgpu.so: HOA loudspeaker renderer and output resampler.
 *
 *   impeghd_hoa_ren_block_process   decoder/impeghd_hoa_renderer.c:181   -> mpeghb200_hoa_render_dev
 *   impeghd_resample                decoder/impeghd_resampler.c:93       -> mpeghb200_resample_dev
 *
 * HOA renderer: the render matrix and the near-field-compensation filters are what the reference's own init
 * (impeghd_hoa_ren_renderer_init, impeghd_hoa_ren_nfc_init) left in ia_render_hoa_str; renderers are cached by
 * content, so every handle that renders the same order to the same layout shares one.  The NFC IIR memories stay in
 * the reference struct (x1 / x2 / y1 / y2 of every cell) and make the round trip with the call, like the FD overlap
 * buffers: a flush or re-init by the reference needs no bookkeeping here.
 *
 * Resampler: ratio and filter tables are the reference's (impeghd_resampler_filter, impeghd_resampler_filter_3, read
 * from the reference library's ROM); the delay lines and polyphase memory of ia_resampler_struct make the round trip.
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
#include "impeghd_resampler_rom.h"
#include "seam_internal.h"

/* the reference library's filter ROM; weak so that this object loads in front of a reference library that only
 * becomes visible later (LD_PRELOAD ahead of a dlopen): looked up again at first use */
extern const FLOAT32 impeghd_resampler_filter[RESAMPLE_FILTER_LEN] __attribute__((weak));
extern const FLOAT32 impeghd_resampler_filter_3[RESAMPLE_FILTER_LEN] __attribute__((weak));

static const float *rs_rom(int three)
{
  const float *p = three ? impeghd_resampler_filter_3 : impeghd_resampler_filter;
  if (!p)
    p = (const float *)dlsym(RTLD_DEFAULT, three ? "impeghd_resampler_filter_3" : "impeghd_resampler_filter");
  if (!p)
  {
    fprintf(stderr, "mpeghb200 seam: the reference library's resampler filter tables are not visible\n");
    abort();
  }
  return p;
}

#define FRAME 1024

/* ---- HOA renderer ------------------------------------------------------------------------------------------------ */
#define MAX_REN 64
static struct
{
  unsigned long long hash;
  int n_out, n_coef, use_nfc, clip;
  mpeghb200_hoa_renderer *h;
} ren[MAX_REN];
static int n_ren;
static seam_buf b_hr_in, b_hr_out, b_hr_state;

static void nfc_coeffs(const ia_render_hoa_str *rh, int n_coef, float *out /* [n_coef][17] */)
{
  int mn, n;
  for (mn = 0; mn < n_coef; mn++)
  {
    const ia_render_hoa_nfc_filtering_str *f = &rh->nfc_filter[mn];
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

/* renderer of this handle's configuration, by content (matrix, NFC coefficients, clip flag) */
static int renderer_of(const ia_render_hoa_str *rh, int n_coef, int clip)
{
  static float M[64 * 64], nfc[64 * 17];
  const ia_render_hoa_render_mtrx_str *m = &rh->v_render_mtrx[rh->mtrx_selected];
  const int n_out = rh->num_out_channels, cols_a = m->sm_mtrx.cols;
  unsigned long long h = 1469598103934665603ull;
  const unsigned char *b;
  size_t q;
  int o, k;
  for (o = 0; o < n_out; o++)
    memcpy(M + o * n_coef, m->sm_mtrx.mtrx + o * cols_a, sizeof(float) * (size_t)n_coef);
  if (rh->use_nfc)
    nfc_coeffs(rh, n_coef, nfc);
  b = (const unsigned char *)M;
  for (q = 0; q < sizeof(float) * (size_t)n_out * n_coef; q++)
    h = (h ^ b[q]) * 1099511628211ull;
  b = (const unsigned char *)nfc;
  for (q = 0; rh->use_nfc && q < sizeof(float) * 17 * (size_t)n_coef; q++)
    h = (h ^ b[q]) * 1099511628211ull;
  for (k = 0; k < n_ren; k++)
    if (ren[k].hash == h && ren[k].n_out == n_out && ren[k].n_coef == n_coef && ren[k].use_nfc == (rh->use_nfc != 0) &&
        ren[k].clip == clip)
      return k;
  if (n_ren == MAX_REN)
  {
    fprintf(stderr, "mpeghb200 seam: more than %d distinct HOA renderers in one process\n", MAX_REN);
    abort();
  }
  ren[n_ren].hash = h, ren[n_ren].n_out = n_out, ren[n_ren].n_coef = n_coef, ren[n_ren].use_nfc = rh->use_nfc != 0;
  ren[n_ren].clip = clip;
  CK(mpeghb200_hoa_renderer_create(seam_ctx(), M, n_out, n_coef, rh->use_nfc ? nfc : NULL, clip, &ren[n_ren].h));
  return n_ren++;
}

/* one group: same renderer.  p[0] = ia_render_hoa_str, p[1] = in [n_coef][1024], p[2] = out [n_out][1024], v[0] = clip */
static void hoa_ren_group(seam_req **rq, int n, int R)
{
  mpeghb200_ctx *ctx = seam_ctx();
  const int n_coef = ren[R].n_coef, n_out = ren[R].n_out, use_nfc = ren[R].use_nfc;
  const size_t in_b = (size_t)n * n_coef * FRAME * sizeof(float), out_b = (size_t)n * n_out * FRAME * sizeof(float);
  const size_t st_f = mpeghb200_hoa_nfc_state_floats(ren[R].h);
  int i, mn, c;
  seam_need(&b_hr_in, in_b), seam_need(&b_hr_out, out_b), seam_need(&b_hr_state, (size_t)n * st_f * sizeof(float) + 16);
  for (i = 0; i < n; i++)
  {
    ia_render_hoa_str *rh = (ia_render_hoa_str *)rq[i]->p[0];
    memcpy((float *)b_hr_in.h + (size_t)i * n_coef * FRAME, rq[i]->p[1], (size_t)n_coef * FRAME * sizeof(float));
    for (mn = 0; use_nfc && mn < n_coef; mn++)
    { /* state record: {(x1, x2, y1, y2) x 3, (x1, y1)} */
      const ia_render_hoa_nfc_filtering_str *f = &rh->nfc_filter[mn];
      float *s = (float *)b_hr_state.h + (size_t)i * st_f + (size_t)mn * 14;
      for (c = 0; c < 3; c++)
        s[4 * c] = f->str_iir_2[c].x1, s[4 * c + 1] = f->str_iir_2[c].x2, s[4 * c + 2] = f->str_iir_2[c].y1,
              s[4 * c + 3] = f->str_iir_2[c].y2;
      s[12] = f->str_iir_1.x1, s[13] = f->str_iir_1.y1;
    }
  }
  CK(mpeghb200_h2d(ctx, b_hr_in.d, b_hr_in.h, in_b));
  if (use_nfc)
    CK(mpeghb200_h2d(ctx, b_hr_state.d, b_hr_state.h, (size_t)n * st_f * sizeof(float)));
  CK(mpeghb200_hoa_render_dev(ctx, ren[R].h, (const float *)b_hr_in.d, (float *)b_hr_out.d, (float *)b_hr_state.d, n, NULL));
  CK(mpeghb200_d2h(ctx, b_hr_out.h, b_hr_out.d, out_b));
  if (use_nfc)
    CK(mpeghb200_d2h(ctx, b_hr_state.h, b_hr_state.d, (size_t)n * st_f * sizeof(float)));
  for (i = 0; i < n; i++)
  {
    ia_render_hoa_str *rh = (ia_render_hoa_str *)rq[i]->p[0];
    memcpy(rq[i]->p[2], (float *)b_hr_out.h + (size_t)i * n_out * FRAME, (size_t)n_out * FRAME * sizeof(float));
    for (mn = 0; use_nfc && mn < n_coef; mn++)
    {
      ia_render_hoa_nfc_filtering_str *f = &rh->nfc_filter[mn];
      const float *s = (const float *)b_hr_state.h + (size_t)i * st_f + (size_t)mn * 14;
      for (c = 0; c < 3; c++)
        f->str_iir_2[c].x1 = s[4 * c], f->str_iir_2[c].x2 = s[4 * c + 1], f->str_iir_2[c].y1 = s[4 * c + 2],
        f->str_iir_2[c].y2 = s[4 * c + 3];
      f->str_iir_1.x1 = s[12], f->str_iir_1.y1 = s[13];
    }
  }
  seam_n.hoa_ren += (unsigned long)n;
}

static void be_hoa_ren(seam_req **rq, int n)
{
  int a = 0, b;
  while (a < n)
  {
    const int R = (int)rq[a]->v[1];
    for (b = a + 1; b < n && (int)rq[b]->v[1] == R; b++)
      ;
    hoa_ren_group(rq + a, b - a, R);
    a = b;
  }
}

IA_ERRORCODE impeghd_hoa_ren_block_process(pVOID handle, pFLOAT32 in_buf, WORD32 rows, UWORD32 cols,
                                           WORD32 is_clipping_protect_enabled, pVOID scratch, pFLOAT32 out)
{
  ia_render_hoa_str *rh = (ia_render_hoa_str *)handle;
  seam_req r;
  WORD32 ch;
  (void)scratch;
  /* the reference's own argument checks (:196-205) */
  if (rh->mtrx_selected == -1)
    return IA_MPEGH_HOA_INIT_FATAL_INVALID_MATRIX_SELECTED;
  ch = (rh->v_render_mtrx[rh->mtrx_selected].order + 1) * (rh->v_render_mtrx[rh->mtrx_selected].order + 1);
  if (ch != rows)
    return IA_MPEGH_HOA_INIT_FATAL_INVALID_MATRIX_SELECTED;
  if (cols == 0)
    return IA_MPEGH_DEC_NO_ERROR;
  if (cols != FRAME || rh->num_out_channels > 24 || rh->num_out_channels < 1)
  {
    typedef IA_ERRORCODE (*fn_t)(pVOID, pFLOAT32, WORD32, UWORD32, WORD32, pVOID, pFLOAT32);
    static fn_t real;
    if (!real)
      real = (fn_t)dlsym(RTLD_NEXT, "impeghd_hoa_ren_block_process");
    __sync_fetch_and_add(&seam_n.hoa_ren_host, 1);
    return real(handle, in_buf, rows, cols, is_clipping_protect_enabled, scratch, out);
  }
  r.kind = RQ_HOAREN;
  r.p[0] = rh, r.p[1] = in_buf, r.p[2] = out;
  r.v[0] = is_clipping_protect_enabled != 0;
  seam_inline_lock(); /* the renderer cache is shared by the parse threads */
  r.v[1] = renderer_of(rh, rows, is_clipping_protect_enabled != 0);
  seam_inline_unlock();
  seam_submit(&r);
  return IA_MPEGH_DEC_NO_ERROR;
}

/* ---- output resampler ------------------------------------------------------------------------------------------- */
static mpeghb200_resampler *rs_handle[4][3]; /* [fac_up][fac_down] */
static seam_buf b_rs_in, b_rs_out, b_rs_state;

/* one group: same ratio.  p[0] = ia_resampler_struct, p[1] = in [1024], p[2] = out */
static void rs_group(seam_req **rq, int n, int up, int down)
{
  mpeghb200_ctx *ctx = seam_ctx();
  mpeghb200_resampler *h;
  size_t st_f;
  int n_out, i, hist = 59 / up;
  if (!rs_handle[up][down])
    CK(mpeghb200_resampler_create(ctx, up, down, rs_rom(up == 3), rs_rom(0), &rs_handle[up][down]));
  h = rs_handle[up][down];
  st_f = mpeghb200_resampler_state_floats(h);
  n_out = mpeghb200_resampler_out_len(h);
  seam_need(&b_rs_in, (size_t)n * FRAME * sizeof(float)), seam_need(&b_rs_out, (size_t)n * n_out * sizeof(float));
  seam_need(&b_rs_state, (size_t)n * st_f * sizeof(float));
  memset(b_rs_state.h, 0, (size_t)n * st_f * sizeof(float));
  for (i = 0; i < n; i++)
  { /* state record: in_hist [32] = delay_samples_upsamp, up_hist [60] = delay_samples, poly_mem [4] */
    const ia_resampler_struct *s = (const ia_resampler_struct *)rq[i]->p[0];
    float *st = (float *)b_rs_state.h + (size_t)i * st_f;
    memcpy(st, s->delay_samples_upsamp, sizeof(float) * (size_t)hist);
    memcpy(st + 32, s->delay_samples, sizeof(float) * RESAMPLE_DELAY);
    st[92] = s->poly_filt_mem[0], st[93] = s->poly_filt_mem[1];
    memcpy((float *)b_rs_in.h + (size_t)i * FRAME, rq[i]->p[1], FRAME * sizeof(float));
  }
  CK(mpeghb200_h2d(ctx, b_rs_in.d, b_rs_in.h, (size_t)n * FRAME * sizeof(float)));
  CK(mpeghb200_h2d(ctx, b_rs_state.d, b_rs_state.h, (size_t)n * st_f * sizeof(float)));
  CK(mpeghb200_resample_dev(ctx, h, (const float *)b_rs_in.d, (float *)b_rs_out.d, (float *)b_rs_state.d, n, NULL));
  CK(mpeghb200_d2h(ctx, b_rs_out.h, b_rs_out.d, (size_t)n * n_out * sizeof(float)));
  CK(mpeghb200_d2h(ctx, b_rs_state.h, b_rs_state.d, (size_t)n * st_f * sizeof(float)));
  for (i = 0; i < n; i++)
  {
    ia_resampler_struct *s = (ia_resampler_struct *)rq[i]->p[0];
    const float *st = (const float *)b_rs_state.h + (size_t)i * st_f;
    memcpy(s->delay_samples_upsamp, st, sizeof(float) * (size_t)hist);
    memcpy(s->delay_samples, st + 32, sizeof(float) * RESAMPLE_DELAY);
    s->poly_filt_mem[0] = st[92], s->poly_filt_mem[1] = st[93];
    memcpy(rq[i]->p[2], (float *)b_rs_out.h + (size_t)i * n_out, (size_t)n_out * sizeof(float));
  }
  seam_n.resample += (unsigned long)n;
}

static void be_rs(seam_req **rq, int n)
{
  int a = 0, b;
  while (a < n)
  {
    for (b = a + 1; b < n && rq[b]->v[0] == rq[a]->v[0] && rq[b]->v[1] == rq[a]->v[1]; b++)
      ;
    rs_group(rq + a, b - a, (int)rq[a]->v[0], (int)rq[a]->v[1]);
    a = b;
  }
}

VOID impeghd_resample(ia_resampler_struct *pstr_resampler, FLOAT32 *ptr_input, FLOAT32 *ptr_output, FLOAT32 *ptr_scratch)
{
  const int up = pstr_resampler->fac_up, down = pstr_resampler->fac_down;
  seam_req r;
  if (pstr_resampler->input_length != FRAME || !((up == 2 && down == 1) || (up == 3 && down == 1) || (up == 3 && down == 2)))
  {
    typedef VOID (*fn_t)(ia_resampler_struct *, FLOAT32 *, FLOAT32 *, FLOAT32 *);
    static fn_t real;
    if (!real)
      real = (fn_t)dlsym(RTLD_NEXT, "impeghd_resample");
    real(pstr_resampler, ptr_input, ptr_output, ptr_scratch);
    return;
  }
  r.kind = RQ_RESAMPLE;
  r.p[0] = pstr_resampler, r.p[1] = ptr_input, r.p[2] = ptr_output;
  r.v[0] = up, r.v[1] = down;
  seam_submit(&r);
}

__attribute__((constructor)) static void seam_register_ren(void)
{
  seam_register(RQ_HOAREN, be_hoa_ren, "hoa_ren_block_process");
  seam_register(RQ_RESAMPLE, be_rs, "resample");
}
