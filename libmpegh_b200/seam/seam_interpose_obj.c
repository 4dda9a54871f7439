/* seam_interpose_obj.c -- libmpeghgpu.so: object renderer and PCM packer.
 *
 *   impeghd_obj_renderer_dec   decoder/impeghd_obj_ren_dec.c:1319       -> mpeghb200_obj_render_dev
 *   ia_core_coder_samples_sat  decoder/ia_core_coder_decode_main.c:209  -> mpeghb200_pcm_pack_block_dev
 *
 * Object renderer: what the reference's own init left in ia_obj_ren_dec_state_struct (triangles, inverse matrices,
 * loudspeaker vectors, imaginary-speaker downmix) becomes an mpeghb200_obj_layout -- the side record of SURVEY.md row
 * a1 filled from a REAL handle --, the frame's OAM values (the *_descaled arrays the reference's OAM parser wrote)
 * become mpeghb200_obj_side records through the library's host trigonometry, previous gains and first-frame flags
 * live on the device, one slot per renderer state.  OAM sub-frames (frame length 512 / 256: two / four calls per
 * core frame, each with its own metadata set) go through mpeghb200_obj_render_sub_dev.
 *
 * PCM packer: stateless apart from the start-up delay counter, which stays in the handle.  The function is `static`
 * in the reference; it is reachable for interposition only in a reference build that gives it external linkage
 * (oracle/Makefile builds such a variant for the tests: -Dstatic= on that one file, no source change).
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
#include "seam_internal.h"

#define FRAME 1024

/* ---- layouts by content ---- */
#define MAX_LAYOUTS 64
static struct
{
  mpeghb200_obj_layout_desc desc;
  mpeghb200_obj_layout *h;
  seam_pool pool[MPEGHB200_OBJ_MAX_OBJECTS + 1]; /* device state by object count */
} lay[MAX_LAYOUTS];
static int n_lay;
static seam_buf b_obj_in, b_obj_out, b_obj_side, b_obj_scratch, b_obj_first, b_obj_set;

static void fill_desc(const ia_obj_ren_dec_state_struct *s, mpeghb200_obj_layout_desc *d)
{
  int i, j;
  memset(d, 0, sizeof(*d));
  d->n_spk = s->num_cicp_speakers;
  d->n_non_lfe = s->num_non_lfe_ls;
  d->n_imag = s->num_imag_ls;
  d->n_tri = s->num_triangles;
  d->height_present = s->height_spks_present;
  for (i = 0; i < CICP_MAX_NUM_LS && i < MPEGHB200_OBJ_MAX_LS; i++)
  {
    d->lfe[i] = (i < s->num_cicp_speakers && s->ptr_cicp_ls_geo[i]) ? s->ptr_cicp_ls_geo[i]->lfe_flag : 0;
    for (j = 0; j < 3; j++)
    {
      d->tri[i][j] = s->ls_triangle[i].spk_index[j];
      d->inv[i][j * 3 + 0] = s->ls_inv_mtx[i].row[j].x;
      d->inv[i][j * 3 + 1] = s->ls_inv_mtx[i].row[j].y;
      d->inv[i][j * 3 + 2] = s->ls_inv_mtx[i].row[j].z;
    }
    d->cart[i][0] = s->non_lfe_ls_str[i].ls_cart_coord.x;
    d->cart[i][1] = s->non_lfe_ls_str[i].ls_cart_coord.y;
    d->cart[i][2] = s->non_lfe_ls_str[i].ls_cart_coord.z;
    for (j = 0; j < CICP_MAX_NUM_LS && j < MPEGHB200_OBJ_MAX_LS; j++)
      d->dmx[i][j] = s->downmix_mtx[i][j];
  }
}

static int layout_of(const ia_obj_ren_dec_state_struct *s)
{
  static mpeghb200_obj_layout_desc d; /* back-ends run one at a time */
  int k;
  fill_desc(s, &d);
  for (k = 0; k < n_lay; k++)
    if (!memcmp(&lay[k].desc, &d, sizeof(d)))
      return k;
  if (n_lay == MAX_LAYOUTS)
  {
    fprintf(stderr, "mpeghb200 seam: more than %d distinct loudspeaker layouts in one process\n", MAX_LAYOUTS);
    abort();
  }
  lay[n_lay].desc = d;
  CK(mpeghb200_obj_layout_create(seam_ctx(), &d, &lay[n_lay].h));
  return n_lay++;
}

typedef struct
{
  seam_req **rq;
  const void **keys;
  int n_obj, n_spk, flen;
} obj_arg;
static void obj_gather(int i, void *arg)
{
  const obj_arg *a = (const obj_arg *)arg;
  const ia_obj_ren_dec_state_struct *s = (const ia_obj_ren_dec_state_struct *)a->rq[i]->p[0];
  const ia_oam_dec_state_struct *md = &s->str_obj_md_dec_state;
  const int n_obj = a->n_obj, k0 = (int)a->rq[i]->v[1];
  float params[MPEGHB200_OBJ_MAX_OBJECTS][7];
  int o;
  a->keys[i] = s;
  ((int32_t *)b_obj_set.h)[i] = k0 / n_obj; /* the metadata set: sub_frame_number - 1 */
  for (o = 0; o < n_obj; o++)
  {
    const int k = k0 + o;
    params[o][0] = md->azimuth_descaled[k], params[o][1] = md->elevation_descaled[k];
    params[o][2] = md->radius_descaled[k], params[o][3] = md->gain_descaled[k];
    params[o][4] = md->spread_width_descaled[k], params[o][5] = md->spread_height_descaled[k];
    params[o][6] = md->spread_depth_descaled[k];
  }
  mpeghb200_obj_side_from_polar(&params[0][0], n_obj, (mpeghb200_obj_side *)b_obj_side.h + (size_t)i * n_obj);
  for (o = 0; o < n_obj; o++) /* the sub-frame's samples to the head of the staging rows */
    memcpy((float *)b_obj_in.h + ((size_t)i * n_obj + o) * FRAME, (const float *)a->rq[i]->p[1] + (size_t)o * FRAME,
           (size_t)a->flen * sizeof(float));
}
static void obj_scatter(int i, void *arg)
{ /* the reference accumulates into rows its caller zeroed: every sample is written once */
  const obj_arg *a = (const obj_arg *)arg;
  int o;
  for (o = 0; o < a->n_spk; o++)
    memcpy((float *)a->rq[i]->p[2] + (size_t)o * FRAME, (float *)b_obj_out.h + ((size_t)i * a->n_spk + o) * FRAME,
           (size_t)a->flen * sizeof(float));
}

/* one group: same layout, same object count, same OAM frame length.  p[0] = ia_obj_ren_dec_state_struct, p[1] = in
 * (rows of 1024, already offset to the sub-frame), p[2] = out (likewise), v[0] = OAM frame length (1024, or a
 * sub-frame of 512 / 256: decode_main.c:1383-1410), v[1] = index of the first object's metadata
 * ((sub_frame_number - 1) * num_objects, impeghd_obj_ren_dec.c:1331) */
static void obj_group(seam_req **rq, int n, int L, int n_obj)
{
  const int flen = (int)rq[0]->v[0];
  mpeghb200_ctx *ctx = seam_ctx();
  const int n_spk = lay[L].desc.n_spk;
  const size_t in_b = (size_t)n * n_obj * FRAME * sizeof(float), out_b = (size_t)n * n_spk * FRAME * sizeof(float);
  const size_t side_b = (size_t)n * n_obj * sizeof(mpeghb200_obj_side);
  const size_t st_f = mpeghb200_obj_state_floats(lay[L].h, n_obj), sc_f = mpeghb200_obj_scratch_floats(lay[L].h, n_obj);
  seam_pool *pool = &lay[L].pool[n_obj];
  const void **keys = (const void **)malloc(sizeof(void *) * (size_t)n);
  obj_arg oa;
  float *base;
  int i, o, fresh = 0;
  pool->stride = st_f * sizeof(float);
  seam_need(&b_obj_in, in_b), seam_need(&b_obj_out, out_b), seam_need(&b_obj_side, side_b);
  seam_need(&b_obj_scratch, (size_t)n * sc_f * sizeof(float));
  seam_need(&b_obj_first, (size_t)n_obj * sizeof(int32_t));
  seam_need(&b_obj_set, (size_t)n * sizeof(int32_t));
  oa.rq = rq, oa.keys = keys, oa.n_obj = n_obj, oa.n_spk = n_spk, oa.flen = flen;
  seam_parallel_for(n, obj_gather, &oa);
  CK(mpeghb200_h2d(ctx, b_obj_in.d, b_obj_in.h, in_b));
  CK(mpeghb200_h2d(ctx, b_obj_side.d, b_obj_side.h, side_b));
  CK(mpeghb200_h2d(ctx, b_obj_set.d, b_obj_set.h, (size_t)n * sizeof(int32_t)));
  base = (float *)seam_pool_range(pool, keys, n, &fresh);
  for (i = 0; i < n; i++)
  {
    float *st = base ? base + (size_t)i * st_f : (float *)seam_pool_one(pool, keys[i], &fresh);
    if (fresh)
    { /* a renderer state seen for the first time: its first_frame_flag[] as the reference's init left them */
      const ia_obj_ren_dec_state_struct *s = (const ia_obj_ren_dec_state_struct *)keys[i];
      int32_t *ff = (int32_t *)b_obj_first.h;
      for (o = 0; o < n_obj; o++)
        ff[o] = s->first_frame_flag[o];
      CK(mpeghb200_h2d(ctx, b_obj_first.d, ff, (size_t)n_obj * sizeof(int32_t)));
      CK(mpeghb200_obj_state_init_dev(ctx, lay[L].h, n_obj, (const int32_t *)b_obj_first.d, st, 1, NULL));
    }
    if (!base)
      CK(mpeghb200_obj_render_sub_dev(ctx, lay[L].h, n_obj, (const mpeghb200_obj_side *)b_obj_side.d + (size_t)i * n_obj,
                                      (const float *)b_obj_in.d + (size_t)i * n_obj * FRAME,
                                      (float *)b_obj_out.d + (size_t)i * n_spk * FRAME, st, (float *)b_obj_scratch.d, 1,
                                      flen, 0, (const int32_t *)b_obj_set.d + i, NULL));
  }
  if (base)
    CK(mpeghb200_obj_render_sub_dev(ctx, lay[L].h, n_obj, (const mpeghb200_obj_side *)b_obj_side.d,
                                    (const float *)b_obj_in.d, (float *)b_obj_out.d, base, (float *)b_obj_scratch.d, n,
                                    flen, 0, (const int32_t *)b_obj_set.d, NULL));
  CK(mpeghb200_d2h(ctx, b_obj_out.h, b_obj_out.d, out_b));
  seam_parallel_for(n, obj_scatter, &oa);
  free(keys);
  seam_n.obj += (unsigned long)n;
}

static void be_obj(seam_req **rq, int n)
{
  int a = 0, b;
  while (a < n)
  {
    const ia_obj_ren_dec_state_struct *sa = (const ia_obj_ren_dec_state_struct *)rq[a]->p[0];
    const int L = layout_of(sa), n_obj = sa->str_obj_md_dec_state.num_objects;
    for (b = a + 1; b < n; b++)
    {
      const ia_obj_ren_dec_state_struct *sb = (const ia_obj_ren_dec_state_struct *)rq[b]->p[0];
      if (rq[b]->v[0] != rq[a]->v[0] || sb->str_obj_md_dec_state.num_objects != n_obj ||
          sb->cicp_out_idx != sa->cicp_out_idx ||
          sb->num_cicp_speakers != sa->num_cicp_speakers || layout_of(sb) != L)
        break;
    }
    obj_group(rq + a, b - a, L, n_obj);
    a = b;
  }
}

IA_ERRORCODE
impeghd_obj_renderer_dec(ia_obj_ren_dec_state_struct *ptr_obj_ren_dec_state, FLOAT32 *p_in_buf, FLOAT32 *p_out_buf,
                         WORD32 cc_frame_len)
{
  typedef IA_ERRORCODE (*fn_t)(ia_obj_ren_dec_state_struct *, FLOAT32 *, FLOAT32 *, WORD32);
  const ia_oam_dec_state_struct *md = &ptr_obj_ren_dec_state->str_obj_md_dec_state;
  seam_req r;
  if (md->num_objects <= 0)
    return IA_MPEGH_DEC_NO_ERROR;
  const int flen = md->p_obj_md_cfg->frame_length;
  if (cc_frame_len != FRAME || (flen != 256 && flen != 512 && flen != FRAME) || md->sub_frame_number < 1 ||
      md->sub_frame_number * md->num_objects > MAX_OAM_FRAMES * MAX_NUM_OAM_OBJS ||
      md->num_objects > MPEGHB200_OBJ_MAX_OBJECTS || ptr_obj_ren_dec_state->num_cicp_speakers > MPEGHB200_OBJ_MAX_LS)
  { /* a shape outside the GPU path (core frame length other than 1024, more than 32 objects): forwarded, counted */
    static fn_t real;
    if (!real)
      real = (fn_t)dlsym(RTLD_NEXT, "impeghd_obj_renderer_dec");
    __sync_fetch_and_add(&seam_n.obj_host, 1);
    return real(ptr_obj_ren_dec_state, p_in_buf, p_out_buf, cc_frame_len);
  }
  r.kind = RQ_OBJ;
  r.p[0] = ptr_obj_ren_dec_state, r.p[1] = p_in_buf, r.p[2] = p_out_buf;
  r.v[0] = flen, r.v[1] = (md->sub_frame_number - 1) * md->num_objects;
  seam_submit(&r);
  return IA_MPEGH_DEC_NO_ERROR;
}

/* ---- PCM packer ------------------------------------------------------------------------------------------------- */
static seam_buf b_pcm_in, b_pcm_out, b_pcm_delay;

typedef struct
{
  seam_req **rq;
  int n_ch, ns, bps;
} pcm_arg;
static void pcm_gather(int i, void *arg)
{
  const pcm_arg *a = (const pcm_arg *)arg;
  FLOAT32 **rows = (FLOAT32 **)a->rq[i]->p[1];
  int c;
  for (c = 0; c < a->n_ch; c++)
    memcpy((float *)b_pcm_in.h + ((size_t)i * a->n_ch + c) * a->ns, rows[c], (size_t)a->ns * sizeof(float));
  ((int32_t *)b_pcm_delay.h)[i] = *(const WORD32 *)a->rq[i]->p[3];
}
static void pcm_scatter(int i, void *arg)
{
  const pcm_arg *a = (const pcm_arg *)arg;
  seam_req *r = a->rq[i];
  WORD32 *delay = (WORD32 *)r->p[3], *out_bytes = (WORD32 *)r->p[4], *m = (WORD32 *)r->p[2];
  const int n_ch = a->n_ch, ns = a->ns, bps = a->bps, d = *delay;
  const int nb = d <= ns ? n_ch * (ns - d) * bps : 0;
  int c;
  memcpy(r->p[0], (const unsigned char *)b_pcm_out.h + (size_t)i * n_ch * ns * bps, (size_t)nb);
  *delay = d <= ns ? 0 : d - ns; /* decode_main.c:223-232 */
  for (c = 0; c < n_ch; c++)    /* the reference rewrites -1 entries of the caller's map in place (:236) */
    if (m[c] == -1)
      m[c] = c;
  *out_bytes = (d <= ns ? n_ch * (ns - d) : 0) * ((int)r->v[3] >> 3); /* :313 */
}

/* p[0] = outbuffer, p[1] = out_samples (FLOAT32 **), p[2] = out_ch_map, p[3] = ptr_delay_samples, p[4] = out_bytes,
 * v[0] = num_samples_out, v[1] = pcmsize, v[2] = num_channel_out */
static void pcm_group(seam_req **rq, int n)
{
  mpeghb200_ctx *ctx = seam_ctx();
  const int ns = (int)rq[0]->v[0], bits = (int)rq[0]->v[1], n_ch = (int)rq[0]->v[2], bps = bits / 8;
  const size_t in_b = (size_t)n * n_ch * ns * sizeof(float), out_b = (size_t)n * n_ch * ns * bps;
  int32_t map[MAX_NUM_CHANNELS];
  pcm_arg pa;
  int c;
  seam_need(&b_pcm_in, in_b), seam_need(&b_pcm_out, out_b), seam_need(&b_pcm_delay, (size_t)n * sizeof(int32_t));
  for (c = 0; c < n_ch; c++)
    map[c] = ((const WORD32 *)rq[0]->p[2])[c];
  pa.rq = rq, pa.n_ch = n_ch, pa.ns = ns, pa.bps = bps;
  seam_parallel_for(n, pcm_gather, &pa);
  CK(mpeghb200_h2d(ctx, b_pcm_in.d, b_pcm_in.h, in_b));
  CK(mpeghb200_h2d(ctx, b_pcm_delay.d, b_pcm_delay.h, (size_t)n * sizeof(int32_t)));
  CK(mpeghb200_pcm_pack_block_dev(ctx, (const float *)b_pcm_in.d, n_ch, ns, ns, bits, map, (int32_t *)b_pcm_delay.d,
                                  (uint8_t *)b_pcm_out.d, NULL, n, NULL));
  CK(mpeghb200_d2h(ctx, b_pcm_out.h, b_pcm_out.d, out_b));
  seam_parallel_for(n, pcm_scatter, &pa);
  seam_n.pcm += (unsigned long)n;
}

static void be_pcm(seam_req **rq, int n)
{
  int a = 0, b, c;
  while (a < n)
  {
    for (b = a + 1; b < n; b++)
    {
      int same = rq[b]->v[0] == rq[a]->v[0] && rq[b]->v[1] == rq[a]->v[1] && rq[b]->v[2] == rq[a]->v[2];
      for (c = 0; same && c < (int)rq[a]->v[2]; c++)
      {
        const WORD32 ma = ((const WORD32 *)rq[a]->p[2])[c], mb = ((const WORD32 *)rq[b]->p[2])[c];
        same = (ma == -1 ? c : ma) == (mb == -1 ? c : mb);
      }
      if (!same)
        break;
    }
    pcm_group(rq + a, b - a);
    a = b;
  }
}

VOID ia_core_coder_samples_sat(WORD8 *outbuffer, WORD32 num_samples_out, WORD32 pcmsize, FLOAT32 **out_samples,
                               WORD32 *out_ch_map, WORD32 *ptr_delay_samples, WORD32 *out_bytes,
                               WORD32 num_channel_out)
{
  seam_req r;
  const int bps = (pcmsize == 16 || pcmsize == 24) ? pcmsize / 8 : 4;
  if (num_samples_out <= 0 || num_samples_out % 256 || num_channel_out <= 0 || num_channel_out > MAX_NUM_CHANNELS)
  {
    if (num_samples_out == 0 && num_channel_out > 0 && num_channel_out <= MAX_NUM_CHANNELS)
    { /* nothing to pack (binaural renderer between two blocks): the bookkeeping of :219-237, :313 only */
      int c;
      for (c = 0; c < num_channel_out; c++)
        if (out_ch_map[c] == -1)
          out_ch_map[c] = c;
      if (*ptr_delay_samples <= 0)
        *ptr_delay_samples = 0;
      *out_bytes = 0;
      return;
    }
    fprintf(stderr, "mpeghb200 seam: PCM block of %d samples x %d channels is outside the GPU path\n", num_samples_out,
            num_channel_out);
    abort();
  }
  r.kind = RQ_PCM;
  r.p[0] = outbuffer, r.p[1] = out_samples, r.p[2] = out_ch_map, r.p[3] = ptr_delay_samples, r.p[4] = out_bytes;
  r.v[0] = num_samples_out, r.v[1] = bps * 8, r.v[2] = num_channel_out, r.v[3] = pcmsize;
  seam_submit(&r);
}

__attribute__((constructor)) static void seam_register_obj(void)
{
  seam_register(RQ_OBJ, be_obj, "obj_renderer");
  seam_register(RQ_PCM, be_pcm, "samples_sat");
}
