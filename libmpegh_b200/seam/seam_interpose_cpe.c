/* seam_interpose_cpe.c -- libmpeghgpu.so: the spectral tools ia_core_coder_data_process runs ahead of TNS / FD synthesis.
 *
 *   impeghd_ms_stereo                  decoder/ia_core_coder_ext_ch_ele.c:1427  -> mpeghb200_stereo_dev (tool 1)
 *   ia_core_coder_cplx_pred_upmixing   decoder/ia_core_coder_ext_ch_ele.c:561   -> mpeghb200_stereo_dev (tool 3)
 *   ia_core_coder_igf_mono             decoder/ia_core_coder_igf_dec.c:2479     -> mpeghb200_igf_dev (with its TNF step)
 *   ia_core_coder_igf_tnf              decoder/ia_core_coder_igf_dec.c:2297     -> mpeghb200_igf_tnf_dev
 *   ia_core_coder_igf_apply_stereo     decoder/ia_core_coder_igf_dec.c:1661     -> mpeghb200_igf_stereo_dev
 *
 * The side records the GPU entry points take (mpeghb200_stereo_pair, mpeghb200_igf_side) are filled here from what the
 * reference's parser left in ia_usac_data_struct: masks, alpha_q, window grouping, the IGF grid of the element, levels,
 * tile numbers, whitening levels, noise seeds.  The carried state stays in the reference handle and makes the round
 * trip with the call: the saved spectra of the complex-prediction tool (coef_save_float), the TNF spectrum and mask
 * (igf_tnf_spec_float / igf_tnf_mask), the seeds.  ia_core_coder_data_process itself -- the order of the tools -- is the
 * reference's code, so every variant (TNS before or after stereo, IGF before or after TNS, independent or joint
 * tiling, the stereo tools on the IGF source tiles and on the TNF spectra) reaches the GPU through these five.
 *
 * ia_core_coder_cplx_pred_upmixing is `static` in the reference: like ia_core_coder_samples_sat it is reachable only in
 * a reference build that gives it external linkage (oracle/Makefile, -Dstatic= on the one file); against the stock
 * library complex prediction stays in the reference, everything else above is taken over.
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
#include "ia_core_coder_igf_dec.h"
#include "seam_internal.h"

#define FRAME 1024
#define ST_STATE 2048 /* mpeghb200_stereo_state_floats(): the two saved spectra */

/* ---- band tables by (long table, short table); a tool that only knows one block type gets a one-band dummy for the other */
#define MAX_BANDS 64
static struct
{
  const WORD16 *tl, *ts;
  int nl, ns;
  mpeghb200_bands *h;
} bands[MAX_BANDS];
static int n_bands;
static const int16_t one_long[1] = {1024}, one_short[1] = {128};

/* under seam_inline_lock */
static int bands_of(const WORD16 *tl, int nl, const WORD16 *ts, int ns)
{
  int k;
  for (k = 0; k < n_bands; k++)
    if (bands[k].tl == tl && bands[k].ts == ts && bands[k].nl == nl && bands[k].ns == ns)
      return k;
  if (n_bands == MAX_BANDS)
  {
    fprintf(stderr, "mpeghb200 seam: more than %d band-table pairs in one process\n", MAX_BANDS);
    abort();
  }
  CK(mpeghb200_bands_create(seam_ctx(), tl ? (const int16_t *)tl : one_long, tl ? nl : 1,
                            ts ? (const int16_t *)ts : one_short, ts ? ns : 1, &bands[k].h));
  bands[k].tl = tl, bands[k].ts = ts, bands[k].nl = nl, bands[k].ns = ns;
  return n_bands++;
}
static int bands_of_info(const ia_sfb_info_struct *info)
{
  int k;
  seam_inline_lock();
  k = info->islong ? bands_of(info->ptr_sfb_tbl, info->sfb_per_sbk, NULL, 0)
                   : bands_of(NULL, 0, info->ptr_sfb_tbl, info->sfb_per_sbk);
  seam_inline_unlock();
  return k;
}
static int bands_of_usac(ia_usac_data_struct *u)
{
  const ia_sfb_info_struct *l = u->pstr_usac_winmap[ONLY_LONG_SEQUENCE], *s = u->pstr_usac_winmap[EIGHT_SHORT_SEQUENCE];
  int k;
  seam_inline_lock();
  k = bands_of(l->ptr_sfb_tbl, l->sfb_per_sbk, s->ptr_sfb_tbl, s->sfb_per_sbk);
  seam_inline_unlock();
  return k;
}

/* ---- stereo tools: p[0] = left, p[1] = right spectrum (in place), p[2] = pair record, p[3], p[4] = saved spectra of
 * the pair or NULL; v[0] = band tables ---- */
static seam_buf b_st_coef, b_st_pair, b_st_state;

static void stereo_group(seam_req **rq, int n, int K)
{
  mpeghb200_ctx *ctx = seam_ctx();
  const size_t cb = (size_t)n * 2 * FRAME * sizeof(float), sb = (size_t)n * ST_STATE * sizeof(float);
  int i;
  seam_need(&b_st_coef, cb), seam_need(&b_st_pair, (size_t)n * sizeof(mpeghb200_stereo_pair)), seam_need(&b_st_state, sb);
  for (i = 0; i < n; i++)
  {
    mpeghb200_stereo_pair *pr = (mpeghb200_stereo_pair *)b_st_pair.h + i;
    float *st = (float *)b_st_state.h + (size_t)i * ST_STATE;
    memcpy(pr, rq[i]->p[2], sizeof(*pr));
    pr->left = 2 * i, pr->right = 2 * i + 1, pr->slot = i;
    memcpy((float *)b_st_coef.h + (size_t)(2 * i) * FRAME, rq[i]->p[0], FRAME * sizeof(float));
    memcpy((float *)b_st_coef.h + (size_t)(2 * i + 1) * FRAME, rq[i]->p[1], FRAME * sizeof(float));
    if (rq[i]->p[3])
      memcpy(st, rq[i]->p[3], FRAME * sizeof(float)), memcpy(st + FRAME, rq[i]->p[4], FRAME * sizeof(float));
    else
      memset(st, 0, ST_STATE * sizeof(float));
  }
  CK(mpeghb200_h2d(ctx, b_st_coef.d, b_st_coef.h, cb));
  CK(mpeghb200_h2d(ctx, b_st_pair.d, b_st_pair.h, (size_t)n * sizeof(mpeghb200_stereo_pair)));
  CK(mpeghb200_h2d(ctx, b_st_state.d, b_st_state.h, sb));
  CK(mpeghb200_stereo_dev(ctx, bands[K].h, (float *)b_st_coef.d, (const mpeghb200_stereo_pair *)b_st_pair.d,
                          (const float *)b_st_state.d, n, NULL));
  CK(mpeghb200_d2h(ctx, b_st_coef.h, b_st_coef.d, cb));
  for (i = 0; i < n; i++)
  {
    memcpy(rq[i]->p[0], (float *)b_st_coef.h + (size_t)(2 * i) * FRAME, FRAME * sizeof(float));
    memcpy(rq[i]->p[1], (float *)b_st_coef.h + (size_t)(2 * i + 1) * FRAME, FRAME * sizeof(float));
  }
  seam_n.stereo += (unsigned long)n;
}

static void be_stereo(seam_req **rq, int n)
{
  int a = 0, b;
  while (a < n)
  {
    for (b = a + 1; b < n && rq[b]->v[0] == rq[a]->v[0]; b++)
      ;
    stereo_group(rq + a, b - a, (int)rq[a]->v[0]);
    a = b;
  }
}

/* window w of a block belongs to the first group whose end index (group_dis) exceeds it */
static void expand_groups(const ia_sfb_info_struct *info, const UWORD8 *group_dis, int *group_of_win)
{
  int w, g = 0;
  for (w = 0; w < info->max_win_len; w++)
  {
    while (w >= group_dis[g])
      g++;
    group_of_win[w] = g;
  }
}

VOID impeghd_ms_stereo(ia_sfb_info_struct *info, WORD8 *group, WORD8 *mask, WORD32 first_band, WORD32 last_band,
                       FLOAT32 *right_spec, FLOAT32 *left_spec)
{
  mpeghb200_stereo_pair pr;
  seam_req r;
  int gw[8], w, sfb;
  const int per = info->sfb_per_sbk;
  if (last_band > per || first_band < 0 || per > (info->islong ? 64 : 16) || info->max_win_len > 8)
  {
    typedef VOID (*fn_t)(ia_sfb_info_struct *, WORD8 *, WORD8 *, WORD32, WORD32, FLOAT32 *, FLOAT32 *);
    static fn_t real;
    if (!real)
      real = (fn_t)dlsym(RTLD_NEXT, "impeghd_ms_stereo");
    seam_n.stereo_host++;
    real(info, group, mask, first_band, last_band, right_spec, left_spec);
    return;
  }
  memset(&pr, 0, sizeof(pr));
  pr.tool = 1, pr.is_short = !info->islong;
  pr.first_band = (uint8_t)first_band, pr.last_band = (uint8_t)last_band;
  expand_groups(info, (const UWORD8 *)group, gw);
  for (w = 0; w < info->max_win_len; w++)
    for (sfb = 0; sfb < per; sfb++)
      pr.used[(info->islong ? 0 : 16 * w) + sfb] = mask[gw[w] * per + sfb] != 0;
  r.kind = RQ_STEREO;
  r.p[0] = left_spec, r.p[1] = right_spec, r.p[2] = &pr, r.p[3] = NULL, r.p[4] = NULL;
  r.v[0] = bands_of_info(info);
  seam_submit(&r);
}

/* `static` in the stock reference (see the file header) */
VOID ia_core_coder_cplx_pred_upmixing(ia_usac_data_struct *usac_data, FLOAT32 *left_spec, FLOAT32 *right_spec,
                                      ia_usac_tmp_core_coder_struct *pstr_core_coder, WORD32 chn, WORD32 pred_dir,
                                      WORD32 complex_coef, WORD32 sfb_start, WORD32 sfb_stop)
{
  const ia_sfb_info_struct *info = usac_data->pstr_sfb_info[chn];
  mpeghb200_stereo_pair pr;
  seam_req r;
  int gw[8], w, sfb, i, any = 0;
  const int per = info->sfb_per_sbk;
  memset(&pr, 0, sizeof(pr));
  pr.tool = 3, pr.is_short = !info->islong;
  pr.window_sequence = (uint8_t)usac_data->window_sequence[chn];
  pr.window_shape = (uint8_t)usac_data->window_shape[chn];
  pr.window_shape_prev = (uint8_t)usac_data->window_shape_prev[chn];
  pr.pred_dir = (uint8_t)pred_dir, pr.complex_coef = (uint8_t)complex_coef;
  pr.use_prev_frame = (uint8_t)pstr_core_coder->use_prev_frame;
  pr.first_band = (uint8_t)sfb_start, pr.last_band = (uint8_t)sfb_stop;
  /* the parse phase has already decided between last frame's downmix and zeros (ia_core_coder_cplx_prev_mdct_dmx,
   * :703): an all-zero dmx_re_prev is the second case (or a downmix that is zero anyway) */
  for (i = 0; i < FRAME && !any; i++)
    any = usac_data->dmx_re_prev[chn][i] != 0.0f;
  pr.prev_dmx = (uint8_t)any;
  expand_groups(info, usac_data->group_dis[chn], gw);
  for (w = 0; w < info->max_win_len; w++)
    for (sfb = 0; sfb < per; sfb++)
    {
      const int k = (info->islong ? 0 : 16 * w) + sfb;
      pr.used[k] = usac_data->cplx_pred_used[chn][gw[w]][sfb];
      pr.alpha_re[k] = (int16_t)usac_data->alpha_q_re[chn][gw[w]][sfb];
      pr.alpha_im[k] = (int16_t)usac_data->alpha_q_im[chn][gw[w]][sfb];
    }
  r.kind = RQ_STEREO;
  r.p[0] = left_spec, r.p[1] = right_spec, r.p[2] = &pr;
  r.p[3] = usac_data->coef_save_float[chn], r.p[4] = usac_data->coef_save_float[chn + 1];
  r.v[0] = bands_of_usac(usac_data);
  seam_submit(&r);
}

/* ---- IGF ---- */
static void fill_igf(mpeghb200_igf_side *sd, ia_usac_data_struct *u, const ia_usac_igf_config_struct *cfg, int chn,
                     int win_type, int num_groups, const WORD16 *group_len)
{
  const ia_usac_igf_bitstream_struct *bs = &u->igf_dec_data[chn].igf_bitstream[0];
  const ia_usac_igf_grid_config_struct *g = &cfg->igf_grid[win_type];
  int k, sfb, t;
  memset(sd, 0, sizeof(*sd));
  sd->seed = u->seed_value[chn];
  sd->win_type = (uint8_t)win_type;
  sd->sfb_start = g->igf_sfb_start, sd->sfb_stop = g->igf_sfb_stop, sd->igf_min = g->igf_min;
  sd->use_high_res = cfg->use_high_res != 0, sd->use_whitening = cfg->use_whitening != 0, sd->use_tnf = cfg->use_tnf != 0;
  sd->all_zero = bs->igf_all_zero != 0, sd->is_tnf = (uint8_t)bs->igf_is_tnf;
  sd->num_groups = (uint8_t)num_groups;
  for (k = 0; k < num_groups && k < 8; k++)
    sd->group_len[k] = (uint8_t)group_len[k];
  for (t = 0; t < 4; t++)
    sd->tile_num[t] = (uint8_t)bs->igf_tile_num[t], sd->whitening_level[t] = (uint8_t)bs->igf_whitening_level[t];
  for (k = 0; k < (win_type == 1 ? 8 : 1); k++)
    for (sfb = 0; sfb < (win_type == 1 ? 16 : 64) && sfb < NSFB_LONG; sfb++)
      sd->levels[(win_type == 1 ? 16 * k : 0) + sfb] = bs->igf_levels_curr_float[k][sfb];
}

static seam_buf b_ig_coef, b_ig_tiles, b_ig_side, b_ig_tnf, b_ig_seed, b_ig_mask;

static void tnf_in(float *st, const ia_usac_igf_dec_data_struct *d)
{
  int i;
  memcpy(st, d->igf_tnf_spec_float, FRAME * sizeof(float));
  for (i = 0; i < FRAME; i++)
    st[FRAME + i] = d->igf_tnf_mask[i] ? 1.0f : 0.0f;
}
static void tnf_out(ia_usac_igf_dec_data_struct *d, const float *st)
{
  int i;
  memcpy(d->igf_tnf_spec_float, st, FRAME * sizeof(float));
  for (i = 0; i < FRAME; i++)
    d->igf_tnf_mask[i] = st[FRAME + i] != 0.0f;
}

/* p[0] = coef (in place), p[1] = ia_usac_igf_dec_data_struct, p[2] = side record, p[3] = &seed_value[chn];
 * v[0] = band tables, v[1] = 1: TNF step only */
static void igf_group(seam_req **rq, int n, int K, int tnf_only)
{
  mpeghb200_ctx *ctx = seam_ctx();
  const size_t cb = (size_t)n * FRAME * sizeof(float), tb = (size_t)n * 4 * FRAME * sizeof(float),
               fb = (size_t)n * 2 * FRAME * sizeof(float);
  int i, t;
  seam_need(&b_ig_coef, cb), seam_need(&b_ig_tiles, tb), seam_need(&b_ig_side, (size_t)n * sizeof(mpeghb200_igf_side));
  seam_need(&b_ig_tnf, fb), seam_need(&b_ig_seed, (size_t)n * sizeof(uint32_t));
  for (i = 0; i < n; i++)
  {
    ia_usac_igf_dec_data_struct *d = (ia_usac_igf_dec_data_struct *)rq[i]->p[1];
    mpeghb200_igf_side *sd = (mpeghb200_igf_side *)b_ig_side.h + i;
    memcpy(sd, rq[i]->p[2], sizeof(*sd));
    sd->unit = i;
    memcpy((float *)b_ig_coef.h + (size_t)i * FRAME, rq[i]->p[0], FRAME * sizeof(float));
    if (!tnf_only)
      for (t = 0; t < 4; t++)
        memcpy((float *)b_ig_tiles.h + ((size_t)i * 4 + t) * FRAME, &d->igf_input_spec[t * MAX_IGF_LEN], FRAME * sizeof(float));
    tnf_in((float *)b_ig_tnf.h + (size_t)i * 2 * FRAME, d);
  }
  CK(mpeghb200_h2d(ctx, b_ig_coef.d, b_ig_coef.h, cb));
  CK(mpeghb200_h2d(ctx, b_ig_side.d, b_ig_side.h, (size_t)n * sizeof(mpeghb200_igf_side)));
  CK(mpeghb200_h2d(ctx, b_ig_tnf.d, b_ig_tnf.h, fb));
  if (tnf_only)
    CK(mpeghb200_igf_tnf_dev(ctx, bands[K].h, (float *)b_ig_coef.d, (const mpeghb200_igf_side *)b_ig_side.d,
                             (float *)b_ig_tnf.d, n, NULL));
  else
  {
    CK(mpeghb200_h2d(ctx, b_ig_tiles.d, b_ig_tiles.h, tb));
    CK(mpeghb200_igf_dev(ctx, bands[K].h, (float *)b_ig_coef.d, (const float *)b_ig_tiles.d,
                         (const mpeghb200_igf_side *)b_ig_side.d, (float *)b_ig_tnf.d, (uint32_t *)b_ig_seed.d, n, NULL));
    CK(mpeghb200_d2h(ctx, b_ig_seed.h, b_ig_seed.d, (size_t)n * sizeof(uint32_t)));
  }
  CK(mpeghb200_d2h(ctx, b_ig_coef.h, b_ig_coef.d, cb));
  CK(mpeghb200_d2h(ctx, b_ig_tnf.h, b_ig_tnf.d, fb));
  for (i = 0; i < n; i++)
  {
    memcpy(rq[i]->p[0], (float *)b_ig_coef.h + (size_t)i * FRAME, FRAME * sizeof(float));
    tnf_out((ia_usac_igf_dec_data_struct *)rq[i]->p[1], (const float *)b_ig_tnf.h + (size_t)i * 2 * FRAME);
    if (!tnf_only)
      *(UWORD32 *)rq[i]->p[3] = ((const uint32_t *)b_ig_seed.h)[i];
  }
  if (tnf_only)
    seam_n.igf_tnf += (unsigned long)n;
  else
    seam_n.igf += (unsigned long)n;
}

static void be_igf(seam_req **rq, int n)
{
  int a = 0, b;
  while (a < n)
  {
    for (b = a + 1; b < n && rq[b]->v[0] == rq[a]->v[0] && rq[b]->v[1] == rq[a]->v[1]; b++)
      ;
    igf_group(rq + a, b - a, (int)rq[a]->v[0], (int)rq[a]->v[1]);
    a = b;
  }
}

/* the TNF step of (usac_data, chn) has been done together with the mono fill on this thread */
static __thread const void *tl_tnf_u;
static __thread int tl_tnf_ch = -1;

static int igf_on_gpu(const ia_usac_igf_config_struct *cfg, int win_type, int frame_id, int bins_per_sbk)
{ /* FD blocks only (frame id 0 / 1 = window type; 2, 3 are the TCX grids of the LPD path) */
  return frame_id == win_type && (win_type == 0 || win_type == 1) && bins_per_sbk == (win_type == 1 ? 128 : 1024) &&
         cfg->igf_grid[win_type].igf_sfb_stop <= (win_type == 1 ? 16 : 64);
}

VOID ia_core_coder_igf_mono(ia_usac_data_struct *usac_data, ia_usac_igf_config_struct *igf_config, WORD32 chn,
                            WORD32 igf_win_type, WORD32 igf_frame_id, WORD32 num_groups, WORD16 *group_len,
                            WORD32 bins_per_sbk, WORD32 win, FLOAT32 *coef, FLOAT32 *scratch)
{
  mpeghb200_igf_side sd;
  seam_req r;
  if (!igf_on_gpu(igf_config, igf_win_type, igf_frame_id, bins_per_sbk))
  {
    typedef VOID (*fn_t)(ia_usac_data_struct *, ia_usac_igf_config_struct *, WORD32, WORD32, WORD32, WORD32, WORD16 *,
                         WORD32, WORD32, FLOAT32 *, FLOAT32 *);
    static fn_t real;
    if (!real)
      real = (fn_t)dlsym(RTLD_NEXT, "ia_core_coder_igf_mono");
    seam_n.igf_host++;
    real(usac_data, igf_config, chn, igf_win_type, igf_frame_id, num_groups, group_len, bins_per_sbk, win, coef, scratch);
    return;
  }
  fill_igf(&sd, usac_data, igf_config, chn, igf_win_type, num_groups, group_len);
  r.kind = RQ_IGF;
  r.p[0] = coef, r.p[1] = &usac_data->igf_dec_data[chn], r.p[2] = &sd, r.p[3] = &usac_data->seed_value[chn];
  r.v[0] = bands_of_usac(usac_data), r.v[1] = 0;
  seam_submit(&r);
  tl_tnf_u = usac_data, tl_tnf_ch = chn; /* ia_core_coder_igf_tnf follows at once (:1867, :1912): already applied */
}

VOID ia_core_coder_igf_tnf(ia_usac_data_struct *usac_data, ia_usac_igf_config_struct *igf_config, WORD32 chn,
                           WORD32 igf_frame_id, WORD32 igf_win_type, WORD32 win, FLOAT32 *coef, FLOAT32 *scratch)
{
  mpeghb200_igf_side sd;
  seam_req r;
  if (tl_tnf_u == usac_data && tl_tnf_ch == chn)
  {
    tl_tnf_u = NULL, tl_tnf_ch = -1;
    return;
  }
  if (!igf_on_gpu(igf_config, igf_win_type, igf_frame_id, igf_win_type == 1 ? 128 : 1024))
  {
    typedef VOID (*fn_t)(ia_usac_data_struct *, ia_usac_igf_config_struct *, WORD32, WORD32, WORD32, WORD32, FLOAT32 *,
                         FLOAT32 *);
    static fn_t real;
    if (!real)
      real = (fn_t)dlsym(RTLD_NEXT, "ia_core_coder_igf_tnf");
    seam_n.igf_host++;
    real(usac_data, igf_config, chn, igf_frame_id, igf_win_type, win, coef, scratch);
    return;
  }
  if (!(igf_config->use_tnf && igf_win_type != 1 && usac_data->igf_dec_data[chn].igf_bitstream[0].igf_is_tnf == 1))
    return; /* nothing to do (:2320-2323) */
  fill_igf(&sd, usac_data, igf_config, chn, igf_win_type, 1, (const WORD16[]){1});
  r.kind = RQ_IGF;
  r.p[0] = coef, r.p[1] = &usac_data->igf_dec_data[chn], r.p[2] = &sd, r.p[3] = NULL;
  r.v[0] = bands_of_usac(usac_data), r.v[1] = 1;
  seam_submit(&r);
}

/* joint-stereo fill: p[0] = usac_data, p[2] = side records [2], p[3] = mask [128]; v[0] = band tables, v[1] = chn */
static void igf_stereo_group(seam_req **rq, int n, int K)
{
  mpeghb200_ctx *ctx = seam_ctx();
  const size_t cb = (size_t)n * 2 * FRAME * sizeof(float), tb = (size_t)n * 8 * FRAME * sizeof(float),
               fb = (size_t)n * 4 * FRAME * sizeof(float);
  int i, c, t;
  seam_need(&b_ig_coef, cb), seam_need(&b_ig_tiles, tb), seam_need(&b_ig_side, (size_t)n * 2 * sizeof(mpeghb200_igf_side));
  seam_need(&b_ig_tnf, fb), seam_need(&b_ig_seed, (size_t)n * 2 * sizeof(uint32_t)), seam_need(&b_ig_mask, (size_t)n * 128);
  for (i = 0; i < n; i++)
  {
    ia_usac_data_struct *u = (ia_usac_data_struct *)rq[i]->p[0];
    const int chn = (int)rq[i]->v[1];
    memcpy((uint8_t *)b_ig_mask.h + (size_t)i * 128, rq[i]->p[3], 128);
    for (c = 0; c < 2; c++)
    {
      const ia_usac_igf_dec_data_struct *d = &u->igf_dec_data[chn + c];
      mpeghb200_igf_side *sd = (mpeghb200_igf_side *)b_ig_side.h + 2 * i + c;
      memcpy(sd, (const mpeghb200_igf_side *)rq[i]->p[2] + c, sizeof(*sd));
      sd->unit = 2 * i + c;
      memcpy((float *)b_ig_coef.h + (size_t)(2 * i + c) * FRAME, u->coef[chn + c], FRAME * sizeof(float));
      for (t = 0; t < 4; t++)
        memcpy((float *)b_ig_tiles.h + ((size_t)(2 * i + c) * 4 + t) * FRAME, &d->igf_input_spec[t * MAX_IGF_LEN],
               FRAME * sizeof(float));
      tnf_in((float *)b_ig_tnf.h + (size_t)(2 * i + c) * 2 * FRAME, d);
    }
  }
  CK(mpeghb200_h2d(ctx, b_ig_coef.d, b_ig_coef.h, cb));
  CK(mpeghb200_h2d(ctx, b_ig_tiles.d, b_ig_tiles.h, tb));
  CK(mpeghb200_h2d(ctx, b_ig_side.d, b_ig_side.h, (size_t)n * 2 * sizeof(mpeghb200_igf_side)));
  CK(mpeghb200_h2d(ctx, b_ig_tnf.d, b_ig_tnf.h, fb));
  CK(mpeghb200_h2d(ctx, b_ig_mask.d, b_ig_mask.h, (size_t)n * 128));
  CK(mpeghb200_igf_stereo_dev(ctx, bands[K].h, (float *)b_ig_coef.d, (const float *)b_ig_tiles.d,
                              (const mpeghb200_igf_side *)b_ig_side.d, (const uint8_t *)b_ig_mask.d, (float *)b_ig_tnf.d,
                              (uint32_t *)b_ig_seed.d, n, NULL));
  CK(mpeghb200_d2h(ctx, b_ig_coef.h, b_ig_coef.d, cb));
  CK(mpeghb200_d2h(ctx, b_ig_tnf.h, b_ig_tnf.d, fb));
  CK(mpeghb200_d2h(ctx, b_ig_seed.h, b_ig_seed.d, (size_t)n * 2 * sizeof(uint32_t)));
  for (i = 0; i < n; i++)
  {
    ia_usac_data_struct *u = (ia_usac_data_struct *)rq[i]->p[0];
    const int chn = (int)rq[i]->v[1];
    for (c = 0; c < 2; c++)
    {
      memcpy(u->coef[chn + c], (float *)b_ig_coef.h + (size_t)(2 * i + c) * FRAME, FRAME * sizeof(float));
      tnf_out(&u->igf_dec_data[chn + c], (const float *)b_ig_tnf.h + (size_t)(2 * i + c) * 2 * FRAME);
      u->seed_value[chn + c] = ((const uint32_t *)b_ig_seed.h)[2 * i + c];
    }
  }
  seam_n.igf_stereo += (unsigned long)n;
}

static void be_igf_stereo(seam_req **rq, int n)
{
  int a = 0, b;
  while (a < n)
  {
    for (b = a + 1; b < n && rq[b]->v[0] == rq[a]->v[0]; b++)
      ;
    igf_stereo_group(rq + a, b - a, (int)rq[a]->v[0]);
    a = b;
  }
}

VOID ia_core_coder_igf_apply_stereo(ia_usac_data_struct *usac_data, WORD32 chn, ia_usac_igf_config_struct *igf_config,
                                    const WORD32 igf_win_type, const WORD8 *mask, const WORD32 hasmask,
                                    const WORD32 num_groups, const WORD16 *group_len, const WORD32 bins_per_sbk,
                                    const WORD32 sfb_per_sbk)
{
  ia_usac_igf_bitstream_struct *bl = &usac_data->igf_dec_data[chn].igf_bitstream[0];
  ia_usac_igf_bitstream_struct *br = &usac_data->igf_dec_data[chn + 1].igf_bitstream[0];
  mpeghb200_igf_side sd[2];
  uint8_t m[128];
  seam_req r;
  int g, sfb;
  if (!igf_on_gpu(igf_config, igf_win_type, igf_win_type, bins_per_sbk) || sfb_per_sbk > (igf_win_type == 1 ? 16 : 64) ||
      num_groups > 8)
  {
    typedef VOID (*fn_t)(ia_usac_data_struct *, WORD32, ia_usac_igf_config_struct *, const WORD32, const WORD8 *,
                         const WORD32, const WORD32, const WORD16 *, const WORD32, const WORD32);
    static fn_t real;
    if (!real)
      real = (fn_t)dlsym(RTLD_NEXT, "ia_core_coder_igf_apply_stereo");
    seam_n.igf_host++;
    real(usac_data, chn, igf_config, igf_win_type, mask, hasmask, num_groups, group_len, bins_per_sbk, sfb_per_sbk);
    return;
  }
  if (bl->igf_all_zero && br->igf_all_zero)
    return; /* :1690 */
  /* ia_core_coder_sync_data (:1615): a silent channel takes the other one's tiling, whitening and TNF flag */
  if (bl->igf_all_zero == 1 && br->igf_all_zero == 0)
  {
    memcpy(bl->igf_tile_num, br->igf_tile_num, sizeof(bl->igf_tile_num[0]) * 4);
    memcpy(bl->igf_whitening_level, br->igf_whitening_level, sizeof(bl->igf_whitening_level[0]) * 4);
    bl->igf_is_tnf = br->igf_is_tnf;
  }
  else if (bl->igf_all_zero == 0 && br->igf_all_zero == 1)
  {
    memcpy(br->igf_tile_num, bl->igf_tile_num, sizeof(br->igf_tile_num[0]) * 4);
    memcpy(br->igf_whitening_level, bl->igf_whitening_level, sizeof(br->igf_whitening_level[0]) * 4);
    br->igf_is_tnf = bl->igf_is_tnf;
  }
  fill_igf(&sd[0], usac_data, igf_config, chn, igf_win_type, num_groups, group_len);
  fill_igf(&sd[1], usac_data, igf_config, chn + 1, igf_win_type, num_groups, group_len);
  memset(m, 0, sizeof(m));
  for (g = 0; g < num_groups; g++)
    for (sfb = 0; sfb < sfb_per_sbk; sfb++)
    {
      int v = 0;
      if (hasmask == 3)
        v = usac_data->cplx_pred_used[chn][g][sfb];
      else if (hasmask > 0)
        v = mask[g * sfb_per_sbk + sfb];
      m[(igf_win_type == 1 ? 16 * g : 0) + sfb] = (uint8_t)(v != 0);
    }
  r.kind = RQ_IGFST;
  r.p[0] = usac_data, r.p[2] = sd, r.p[3] = m;
  r.v[0] = bands_of_usac(usac_data), r.v[1] = chn;
  seam_submit(&r);
}

__attribute__((constructor)) static void seam_register_cpe(void)
{
  seam_register(RQ_STEREO, be_stereo, "ms_stereo / cplx_pred_upmixing");
  seam_register(RQ_IGF, be_igf, "igf_mono / igf_tnf");
  seam_register(RQ_IGFST, be_igf_stereo, "igf_apply_stereo");
}
