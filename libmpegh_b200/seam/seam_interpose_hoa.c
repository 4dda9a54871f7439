/* seam_interpose_hoa.c -- libmpeghgpu.so: HOA spatial decoding.
 *
 * impeghd_hoa_spatial_process (decoder/impeghd_hoa_spatial_decoder.c:663) first decodes the frame's side info with a
 * `static` host function (:262, stays the reference's) and then runs four signal stages in a row:
 *   impeghd_hoa_inv_dyn_correction_process   decoder/impeghd_hoa_inverse_dyn_correction.c:75
 *   impeghd_hoa_ambience_synthesis_process   decoder/impeghd_hoa_amb_syn.c:76
 *   impeghd_hoa_ch_reassignment_process      decoder/impeghd_hoa_ch_reassignment.c:71
 *   impeghd_hoa_pre_dom_sound_syn_process    decoder/impeghd_hoa_pre_dom_sound_syn.c:72
 * These four are interposed: the first three only note their arguments, the last one hands the whole spatial decode of
 * the frame (transport channels in, HOA coefficient signals out) to mpeghb200_hoa_spatial_dev.  The side record
 * (mpeghb200_hoa_spatial_side) is filled from what the reference's side-info decoder left in
 * ia_hoa_dec_frame_param_str; the decoder tables from what impeghd_hoa_spatial_init left in ia_spatial_dec_str.
 *
 * State: everything the decoder carries between frames lives on the device, one slot per ia_spatial_dec_str.  The one
 * thing the reference writes into that state from outside the four stages is the gain reset of an independency frame
 * (inv_dyn_corr_prev_gain[ch] = 2^gain_corr_prev_amp_exp, spatial_decoder.c:683-689): after every frame the interposer
 * leaves a sentinel in that host array, so a value that is not the sentinel at the next frame IS such a reset (or the
 * initial gain) and travels in the side record.
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
#include "impeghd_hoa_common_functions.h"
#include "impeghd_hoa_rom.h"
#include "seam_internal.h"

#define FRAME 1024
#define SENTINEL_BITS 0x7FC0B200u

/* ---- decoders by content ---- */
#define MAX_HOASP 64
static struct
{
  unsigned long long hash;
  int cfg[10];
  mpeghb200_hoa_spatial *h;
  seam_pool pool;
} dec[MAX_HOASP];
static int n_dec;
static seam_buf b_hs_in, b_hs_out, b_hs_side;

/* the reference library's ROM, found at run time: this object may be loaded in front of a reference library that only
 * arrives later through dlopen(), so no load-time data references */
static const float *rom(const char *name)
{
  const float *p = (const float *)dlsym(RTLD_DEFAULT, name);
  if (!p)
  {
    fprintf(stderr, "mpeghb200 seam: the reference library's HOA table %s is not visible\n", name);
    abort();
  }
  return p;
}

static int decoder_of(ia_spatial_dec_str *d)
{
  const ia_hoa_config_struct *c = d->ia_hoa_config;
  const ia_spatial_dec_pre_dom_sound_syn_str *s = &d->pre_dom_sound_syn_handle;
  static float wins[4 * FRAME], pow2[128];
  const float *dyn_pow, *p2, *p2i;
  mpeghb200_hoa_spatial_desc desc;
  unsigned long long h = 1469598103934665603ull;
  const int nc = (int)c->num_coeffs;
  int cfg[10], k, i;
  size_t q;
  const unsigned char *b;
  cfg[0] = c->order, cfg[1] = c->num_transport_ch, cfg[2] = c->num_addnl_coders;
  cfg[3] = d->amb_syn_handle.low_order_mat_sz, cfg[4] = d->vec_start, cfg[5] = c->coded_v_vec_length;
  cfg[6] = c->spat_interpolation_time, cfg[7] = c->pred_dir_sigs, cfg[8] = c->num_bits_per_sf;
  cfg[9] = d->use_phase_shift_decorr;
  /* the tables are functions of the configuration and of two init-time designs: hash those */
  b = (const unsigned char *)d->amb_syn_handle.order_matrix;
  for (q = 0; q < sizeof(float) * (size_t)cfg[3] * cfg[3]; q++)
    h = (h ^ b[q]) * 1099511628211ull;
  b = (const unsigned char *)s->dir_based_pre_dom_sound_syn_handle.mode_mt_all_coarse_grid_points;
  for (q = 0; q < sizeof(float) * (size_t)nc * nc; q++)
    h = (h ^ b[q]) * 1099511628211ull;
  b = (const unsigned char *)s->fade_in_win_for_dir_sigs;
  for (q = 0; q < sizeof(float) * FRAME; q++)
    h = (h ^ b[q]) * 1099511628211ull;
  for (k = 0; k < n_dec; k++)
    if (dec[k].hash == h && !memcmp(dec[k].cfg, cfg, sizeof(cfg)))
      return k;
  if (n_dec == MAX_HOASP)
  {
    fprintf(stderr, "mpeghb200 seam: more than %d distinct HOA configurations in one process\n", MAX_HOASP);
    abort();
  }
  dyn_pow = rom("ia_hoa_inv_dyn_win_pow_table");
  p2 = rom("ia_hoa_pow_2_table"), p2i = rom("ia_hoa_pow_2_inv_table"); /* what impeghd_hoa_pow_2() indexes */
  memcpy(wins, s->fade_in_win_for_vec_sigs, sizeof(float) * FRAME);
  memcpy(wins + FRAME, s->fade_out_win_for_vec_sigs, sizeof(float) * FRAME);
  memcpy(wins + 2 * FRAME, s->fade_in_win_for_dir_sigs, sizeof(float) * FRAME);
  memcpy(wins + 3 * FRAME, s->fade_out_win_for_dir_sigs, sizeof(float) * FRAME);
  for (i = 0; i < 64; i++)
    pow2[i] = p2[i], pow2[64 + i] = p2i[i];
  memset(&desc, 0, sizeof(desc));
  desc.order = cfg[0], desc.n_transport = cfg[1], desc.n_addnl_coders = cfg[2], desc.low_order_mat_sz = cfg[3];
  desc.vec_start = cfg[4], desc.coded_v_vec_length = cfg[5], desc.spat_interp_time = cfg[6], desc.pred_dir_sigs = cfg[7];
  desc.num_bits_per_sf = cfg[8], desc.use_phase_shift_decorr = cfg[9];
  desc.order_matrix = d->amb_syn_handle.order_matrix;
  desc.fine_grid_mode_matrix = d->transpose_mod_mtx_grid_pts;
  desc.coarse_grid_mode_matrix = s->dir_based_pre_dom_sound_syn_handle.mode_mt_all_coarse_grid_points;
  desc.fade_windows = wins;
  desc.dyn_win_pow_table = dyn_pow;
  desc.dyn_win_func = d->inv_dyn_corr_win_func;
  desc.pow2_table = pow2;
  dec[n_dec].hash = h;
  memcpy(dec[n_dec].cfg, cfg, sizeof(cfg));
  CK(mpeghb200_hoa_spatial_create(seam_ctx(), &desc, &dec[n_dec].h));
  dec[n_dec].pool.stride = mpeghb200_hoa_spatial_state_bytes(dec[n_dec].h);
  return n_dec++;
}

static void fill_side(ia_spatial_dec_str *d, mpeghb200_hoa_spatial_side *sd)
{
  const ia_hoa_dec_frame_param_str *fp = &d->frame_params_handle;
  const int nc = (int)d->ia_hoa_config->num_coeffs, T = (int)d->ia_hoa_config->num_transport_ch;
  int i, r;
  memset(sd, 0, sizeof(*sd));
  sd->n_vec = (int32_t)fp->num_vec_predom_sounds, sd->n_dir = (int32_t)fp->num_dir_predom_sounds;
  sd->n_enable = (int32_t)fp->num_enable_coeff, sd->n_disable = (int32_t)fp->num_disable_coeff;
  for (i = 0; i < T && i < MPEGHB200_HOASP_MAX_CH; i++)
  {
    uint32_t bits;
    sd->exponents[i] = fp->exponents[i];
    sd->is_exception[i] = fp->ptr_is_exception ? fp->ptr_is_exception[i] : 0;
    sd->amb_hoa_assign[i] = (int32_t)fp->amb_hoa_assign[i];
    memcpy(&bits, &d->inv_dyn_corr_prev_gain[i], 4);
    sd->prev_gain_reset[i] = bits == SENTINEL_BITS ? 0.0f : d->inv_dyn_corr_prev_gain[i];
    bits = SENTINEL_BITS;
    memcpy(&d->inv_dyn_corr_prev_gain[i], &bits, 4);
  }
  for (i = 0; i < MPEGHB200_HOASP_MAX_CH; i++)
  {
    sd->vec_index[i] = (int32_t)fp->vectors[i].index;
    memcpy(sd->vec[i], fp->vectors[i].sig_id, sizeof(float) * MPEGHB200_HOASP_MAX_COEF);
  }
  for (i = 0; i <= MPEGHB200_HOASP_MAX_CH; i++)
  {
    sd->dir_ch[i] = (int32_t)fp->active_and_grid_dir_indices[i].ch_idx;
    sd->dir_id[i] = (int32_t)fp->active_and_grid_dir_indices[i].dir_id;
  }
  for (i = 0; i < MPEGHB200_HOASP_MAX_COEF; i++)
  {
    sd->enable[i] = (int32_t)fp->amb_coeff_indices_to_enable[i];
    sd->disable[i] = (int32_t)fp->amb_coeff_indices_to_disable[i];
    sd->non_en_dis[i] = (int32_t)fp->non_en_dis_able_act_hoa_coeff_indices[i];
  }
  for (i = 0; i < nc; i++)
    sd->pred_type[i] = (int32_t)fp->prediction_type_vec[i];
  for (r = 0; r < MPEGHB200_HOASP_MAX_PRED; r++)
    for (i = 0; i < nc; i++)
    {
      sd->pred_idx[r][i] = (int32_t)fp->prediction_indices_mat[r * nc + i];
      sd->pred_quant[r][i] = fp->quant_pred_fact_mat[r * nc + i];
    }
}

/* one group: same decoder.  p[0] = ia_spatial_dec_str, p[1] = in [T][1024], p[2] = out [n_coef][1024], p[3] = side */
static void hoasp_group(seam_req **rq, int n, int D)
{
  mpeghb200_ctx *ctx = seam_ctx();
  const int T = dec[D].cfg[1], nc = mpeghb200_hoa_spatial_num_coeffs(dec[D].h);
  const size_t in_b = (size_t)n * T * FRAME * sizeof(float), out_b = (size_t)n * nc * FRAME * sizeof(float);
  const size_t stride = dec[D].pool.stride;
  const void **keys = (const void **)malloc(sizeof(void *) * (size_t)n);
  char *base;
  int i, fresh = 0;
  seam_need(&b_hs_in, in_b), seam_need(&b_hs_out, out_b), seam_need(&b_hs_side, (size_t)n * sizeof(mpeghb200_hoa_spatial_side));
  for (i = 0; i < n; i++)
  {
    keys[i] = rq[i]->p[0];
    memcpy((float *)b_hs_in.h + (size_t)i * T * FRAME, rq[i]->p[1], (size_t)T * FRAME * sizeof(float));
    memcpy((mpeghb200_hoa_spatial_side *)b_hs_side.h + i, rq[i]->p[3], sizeof(mpeghb200_hoa_spatial_side));
  }
  CK(mpeghb200_h2d(ctx, b_hs_in.d, b_hs_in.h, in_b));
  CK(mpeghb200_h2d(ctx, b_hs_side.d, b_hs_side.h, (size_t)n * sizeof(mpeghb200_hoa_spatial_side)));
  base = (char *)seam_pool_range(&dec[D].pool, keys, n, &fresh);
  if (base)
  {
    if (fresh)
      CK(mpeghb200_hoa_spatial_state_init_dev(ctx, dec[D].h, base, n, NULL));
    CK(mpeghb200_hoa_spatial_dev(ctx, dec[D].h, (const mpeghb200_hoa_spatial_side *)b_hs_side.d, (const float *)b_hs_in.d,
                                 (float *)b_hs_out.d, base, n, NULL));
  }
  else
  {
    for (i = 0; i < n; i++)
    {
      char *st = (char *)seam_pool_one(&dec[D].pool, keys[i], &fresh);
      if (fresh)
        CK(mpeghb200_hoa_spatial_state_init_dev(ctx, dec[D].h, st, 1, NULL));
      CK(mpeghb200_hoa_spatial_dev(ctx, dec[D].h, (const mpeghb200_hoa_spatial_side *)b_hs_side.d + i,
                                   (const float *)b_hs_in.d + (size_t)i * T * FRAME,
                                   (float *)b_hs_out.d + (size_t)i * nc * FRAME, st, 1, NULL));
    }
  }
  (void)stride;
  CK(mpeghb200_d2h(ctx, b_hs_out.h, b_hs_out.d, out_b));
  for (i = 0; i < n; i++)
    memcpy(rq[i]->p[2], (float *)b_hs_out.h + (size_t)i * nc * FRAME, (size_t)nc * FRAME * sizeof(float));
  free(keys);
  seam_n.hoa_spatial += (unsigned long)n;
}

static void be_hoasp(seam_req **rq, int n)
{
  int a = 0, b;
  while (a < n)
  {
    const int D = (int)rq[a]->v[0];
    for (b = a + 1; b < n && (int)rq[b]->v[0] == D; b++)
      ;
    hoasp_group(rq + a, b - a, D);
    a = b;
  }
}

/* can this decoder's frames run on the GPU?  (decided per handle, identically for all four stages) */
static int on_gpu(const ia_spatial_dec_str *d)
{
  const ia_hoa_config_struct *c = d->ia_hoa_config;
  return c && c->frame_length == FRAME && c->order >= 1 && c->order <= 6 && c->num_transport_ch >= 1 &&
         c->num_transport_ch <= MPEGHB200_HOASP_MAX_CH && !d->use_phase_shift_decorr;
}

static __thread FLOAT32 *tl_in;                 /* transport signals of the frame under way on this thread */
static __thread const ia_spatial_dec_str *tl_d; /* ... and whose they are (no yield between the four stages) */

IA_ERRORCODE impeghd_hoa_inv_dyn_correction_process(FLOAT32 *ptr_inp, ia_spatial_dec_str *dec_handle)
{
  if (!on_gpu(dec_handle))
  {
    typedef IA_ERRORCODE (*fn_t)(FLOAT32 *, ia_spatial_dec_str *);
    static fn_t real;
    if (!real)
      real = (fn_t)dlsym(RTLD_NEXT, "impeghd_hoa_inv_dyn_correction_process");
    return real(ptr_inp, dec_handle);
  }
  tl_in = ptr_inp, tl_d = dec_handle;
  return IA_MPEGH_DEC_NO_ERROR;
}

IA_ERRORCODE impeghd_hoa_ambience_synthesis_process(ia_spatial_dec_str *dec_handle, pWORD32 reassign_vec)
{
  if (!on_gpu(dec_handle))
  {
    typedef IA_ERRORCODE (*fn_t)(ia_spatial_dec_str *, pWORD32);
    static fn_t real;
    if (!real)
      real = (fn_t)dlsym(RTLD_NEXT, "impeghd_hoa_ambience_synthesis_process");
    return real(dec_handle, reassign_vec);
  }
  if (dec_handle->scratch == NULL) /* the reference's own check (amb_syn.c:89) */
    return IA_MPEGH_HOA_EXE_FATAL_SCRATCH_NULL;
  return IA_MPEGH_DEC_NO_ERROR; /* the kernel derives the reassignment from amb_hoa_assign in the side record */
}

IA_ERRORCODE impeghd_hoa_ch_reassignment_process(ia_spatial_dec_str *dec_handle)
{
  if (!on_gpu(dec_handle))
  {
    typedef IA_ERRORCODE (*fn_t)(ia_spatial_dec_str *);
    static fn_t real;
    if (!real)
      real = (fn_t)dlsym(RTLD_NEXT, "impeghd_hoa_ch_reassignment_process");
    return real(dec_handle);
  }
  return IA_MPEGH_DEC_NO_ERROR;
}

IA_ERRORCODE impeghd_hoa_pre_dom_sound_syn_process(ia_spatial_dec_str *dec_handle, pFLOAT32 ptr_out)
{
  mpeghb200_hoa_spatial_side side;
  seam_req r;
  if (!on_gpu(dec_handle))
  {
    typedef IA_ERRORCODE (*fn_t)(ia_spatial_dec_str *, pFLOAT32);
    static fn_t real;
    if (!real)
      real = (fn_t)dlsym(RTLD_NEXT, "impeghd_hoa_pre_dom_sound_syn_process");
    __sync_fetch_and_add(&seam_n.hoa_spatial_host, 1);
    return real(dec_handle, ptr_out);
  }
  if (tl_d != dec_handle || !tl_in)
  {
    fprintf(stderr, "mpeghb200 seam: HOA synthesis stage called without the inverse dynamic correction before it\n");
    abort();
  }
  fill_side(dec_handle, &side);
  r.kind = RQ_HOASP;
  r.p[0] = dec_handle, r.p[1] = tl_in, r.p[2] = ptr_out, r.p[3] = &side;
  seam_inline_lock(); /* the decoder cache is shared by the parse threads */
  r.v[0] = decoder_of(dec_handle);
  seam_inline_unlock();
  tl_in = NULL, tl_d = NULL;
  seam_submit(&r);
  return IA_MPEGH_DEC_NO_ERROR;
}

__attribute__((constructor)) static void seam_register_hoa(void) { seam_register(RQ_HOASP, be_hoasp, "hoa_spatial"); }
