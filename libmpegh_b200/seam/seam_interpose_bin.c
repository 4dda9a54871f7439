/* seam_interpose_bin.c -- libmpeghgpu.so: time-domain binaural renderer.
 *
 *   impeghd_binaural_struct_process_ld_ola   decoder/impeghd_binaural_renderer.c:543   -> mpeghb200_binaural_block_dev
 *
 * The filters are the reference's own: the packed spectra, multiply lengths and diffuse weights that
 * impeghd_binaural_renderer_init left in ia_binaural_ren_pers_mem_str go to the device as they are
 * (mpeghb200_binaural_create_from_state), once per distinct content -- handles that loaded the same BRIR file share one
 * renderer and one launch.  The sample collector (input_mem_buf / memory_index) stays in the reference struct and is
 * run here exactly as the reference runs it; the overlap tails and the diffuse downmix ring live on the device, one slot
 * per reference struct, starting from the all-zero state the reference's init leaves.
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

/* ---- renderers by content ---- */
#define MAX_BIN 64
static struct
{
  unsigned long long hash;
  int fft_size, n_in, n_blocks;
  mpeghb200_binaural *h;
  seam_pool pool;
} ren[MAX_BIN];
static int n_ren;
static seam_buf b_bin_in, b_bin_out;

/* which renderer a reference struct maps to, remembered per struct (hashing 2 x n_in x fft_size floats once) */
#define MAX_KNOWN 4096
static struct
{
  const ia_binaural_ren_pers_mem_str *p;
  int ren;
  float probe[4];
} known[MAX_KNOWN];
static int n_known;

static void probe_of(const ia_binaural_ren_pers_mem_str *b, float *probe)
{
  probe[0] = b->taps_direct[0][0][0], probe[1] = b->taps_direct[1][b->num_input - 1][b->fft_size / 2];
  probe[2] = b->taps_direct[0][0][3], probe[3] = (float)(b->fft_size * 64 + b->num_input * 4 + b->diffuse_blocks);
}

static unsigned long long fnv(unsigned long long h, const void *p, size_t n)
{
  const unsigned char *b = (const unsigned char *)p;
  size_t q;
  for (q = 0; q < n; q++)
    h = (h ^ b[q]) * 1099511628211ull;
  return h;
}

/* under seam_inline_lock */
static int renderer_of(const ia_binaural_ren_pers_mem_str *b)
{
  float probe[4];
  unsigned long long h = 1469598103934665603ull;
  int k, o, i, j;
  probe_of(b, probe);
  for (k = 0; k < n_known; k++)
    if (known[k].p == b && !memcmp(known[k].probe, probe, sizeof(probe)))
      return known[k].ren;
  for (o = 0; o < 2; o++)
    for (i = 0; i < b->num_input; i++)
      h = fnv(h, b->taps_direct[o][i], sizeof(float) * (size_t)b->fft_size);
  for (j = 0; j < b->diffuse_blocks; j++)
    for (o = 0; o < 2; o++)
      h = fnv(h, b->taps_diffuse[j][o], sizeof(float) * (size_t)b->fft_size);
  h = fnv(h, b->direct_process_size, sizeof(b->direct_process_size));
  h = fnv(h, b->diffuse_process_size, sizeof(b->diffuse_process_size));
  h = fnv(h, b->diffuse_weight, sizeof(float) * (size_t)b->num_input);
  for (k = 0; k < n_ren; k++)
    if (ren[k].hash == h && ren[k].fft_size == b->fft_size && ren[k].n_in == b->num_input &&
        ren[k].n_blocks == b->diffuse_blocks)
      break;
  if (k == n_ren)
  {
    if (n_ren == MAX_BIN)
    {
      fprintf(stderr, "mpeghb200 seam: more than %d distinct binaural filter sets in one process\n", MAX_BIN);
      abort();
    }
    CK(mpeghb200_binaural_create_from_state(
        seam_ctx(), b->fft_size, b->num_input, b->diffuse_blocks, &b->taps_direct[0][0][0], BINAURAL_FFT_MAXLEN,
        (size_t)MAX_NUM_BRIR_CHANNELS * BINAURAL_FFT_MAXLEN, &b->taps_diffuse[0][0][0], BINAURAL_FFT_MAXLEN,
        &b->direct_process_size[0][0], MAX_NUM_BRIR_CHANNELS, &b->diffuse_process_size[0][0], b->diffuse_weight,
        &ren[k].h));
    ren[k].hash = h, ren[k].fft_size = b->fft_size, ren[k].n_in = b->num_input, ren[k].n_blocks = b->diffuse_blocks;
    ren[k].pool.stride = mpeghb200_binaural_state_floats(ren[k].h) * sizeof(float);
    n_ren++;
  }
  if (n_known < MAX_KNOWN)
  {
    known[n_known].p = b, known[n_known].ren = k;
    memcpy(known[n_known].probe, probe, sizeof(probe));
    n_known++;
  }
  return k;
}

/* one group: same renderer.  p[0] = ia_binaural_ren_pers_mem_str (input_mem_buf holds the block), p[1], p[2] = where
 * the two ears' B samples go */
static void bin_group(seam_req **rq, int n, int R)
{
  mpeghb200_ctx *ctx = seam_ctx();
  const int n_in = ren[R].n_in, B = ren[R].fft_size / 2;
  const size_t in_b = (size_t)n * n_in * B * sizeof(float), out_b = (size_t)n * 2 * B * sizeof(float);
  const void **keys = (const void **)malloc(sizeof(void *) * (size_t)n);
  char *base;
  int i, c, fresh = 0;
  seam_need(&b_bin_in, in_b), seam_need(&b_bin_out, out_b);
  for (i = 0; i < n; i++)
  {
    const ia_binaural_ren_pers_mem_str *b = (const ia_binaural_ren_pers_mem_str *)rq[i]->p[0];
    keys[i] = b;
    for (c = 0; c < n_in; c++)
      memcpy((float *)b_bin_in.h + ((size_t)i * n_in + c) * B, b->input_mem_buf[c], (size_t)B * sizeof(float));
  }
  CK(mpeghb200_h2d(ctx, b_bin_in.d, b_bin_in.h, in_b));
  base = (char *)seam_pool_range(&ren[R].pool, keys, n, &fresh);
  if (base)
  {
    if (fresh) /* what impeghd_binaural_struct_set leaves: empty tail, empty diffuse ring, write index 0 */
      CK(mpeghb200_dev_memset(ctx, base, 0, (size_t)n * ren[R].pool.stride));
    CK(mpeghb200_binaural_block_dev(ctx, ren[R].h, (const float *)b_bin_in.d, 0, 1.0f, 1.0f, (float *)b_bin_out.d,
                                    (float *)base, NULL, n, NULL));
  }
  else
  {
    for (i = 0; i < n; i++)
    {
      char *st = (char *)seam_pool_one(&ren[R].pool, keys[i], &fresh);
      if (fresh)
        CK(mpeghb200_dev_memset(ctx, st, 0, ren[R].pool.stride));
      CK(mpeghb200_binaural_block_dev(ctx, ren[R].h, (const float *)b_bin_in.d + (size_t)i * n_in * B, 0, 1.0f, 1.0f,
                                      (float *)b_bin_out.d + (size_t)i * 2 * B, (float *)st, NULL, 1, NULL));
    }
  }
  CK(mpeghb200_d2h(ctx, b_bin_out.h, b_bin_out.d, out_b));
  for (i = 0; i < n; i++)
  {
    memcpy(rq[i]->p[1], (float *)b_bin_out.h + (size_t)i * 2 * B, (size_t)B * sizeof(float));
    memcpy(rq[i]->p[2], (float *)b_bin_out.h + (size_t)i * 2 * B + B, (size_t)B * sizeof(float));
  }
  free(keys);
  seam_n.binaural += (unsigned long)n;
}

static void be_bin(seam_req **rq, int n)
{
  int a = 0, b;
  while (a < n)
  {
    const int R = (int)rq[a]->v[0];
    for (b = a + 1; b < n && (int)rq[b]->v[0] == R; b++)
      ;
    bin_group(rq + a, b - a, R);
    a = b;
  }
}

/* The reference's own bookkeeping around the block transform (:558-582, :664-692): samples are collected in
 * input_mem_buf until a block of processed_samples = fft_size / 2 is complete; every complete block is one request. */
WORD32 impeghd_binaural_struct_process_ld_ola(FLOAT32 **ptr_src, FLOAT32 **ptr_dst, WORD32 samples_len, FLOAT32 *ptr_work,
                                              ia_binaural_ren_pers_mem_str *ptr_binaural, FLOAT32 *ptr_scratch_buf)
{
  ia_binaural_ren_pers_mem_str *b = ptr_binaural;
  const WORD32 B = b->processed_samples;
  WORD32 n_in_samps, n_out = 0, idx = 0, len, inp, R;
  const int fft = b->fft_size;
  const int size_ok = fft == 2048 || fft == 8192 || fft == 16384 || (fft >= 16 && fft <= 1024 && !(fft & (fft - 1)));
  if (!size_ok || B != fft / 2 || b->num_output != 2 || b->num_input < 1 || b->num_input > 32 || b->diffuse_blocks < 0 ||
      b->diffuse_blocks > 1)
  { /* the 4096-point renderer, whose reference output is not a function of the input, and the
       two-diffuse-block ring the reference indexes out of bounds, stay where they are */
    typedef WORD32 (*fn_t)(FLOAT32 **, FLOAT32 **, WORD32, FLOAT32 *, ia_binaural_ren_pers_mem_str *, FLOAT32 *);
    static fn_t real;
    if (!real)
      real = (fn_t)dlsym(RTLD_NEXT, "impeghd_binaural_struct_process_ld_ola");
    seam_n.binaural_host++;
    return real(ptr_src, ptr_dst, samples_len, ptr_work, ptr_binaural, ptr_scratch_buf);
  }
  if (samples_len + b->memory_index < B)
  {
    for (inp = 0; inp < b->num_input; inp++)
      memmove(&b->input_mem_buf[inp][b->memory_index], ptr_src[inp], sizeof(FLOAT32) * (size_t)samples_len);
    b->memory_index += samples_len;
    return 0;
  }
  seam_inline_lock(); /* the renderer cache is shared by the parse threads */
  R = renderer_of(b);
  seam_inline_unlock();
  n_in_samps = samples_len + b->memory_index;
  while (n_in_samps >= B)
  {
    seam_req r;
    len = B - b->memory_index;
    for (inp = 0; inp < b->num_input; inp++)
      memmove(&b->input_mem_buf[inp][b->memory_index], &ptr_src[inp][idx], sizeof(FLOAT32) * (size_t)len);
    r.kind = RQ_BINAURAL;
    r.p[0] = b, r.p[1] = &ptr_dst[0][n_out], r.p[2] = &ptr_dst[1][n_out];
    r.v[0] = R;
    seam_submit(&r);
    b->write_idx = (b->write_idx + 1) % (b->diffuse_blocks + 1); /* mirrored; the ring itself is on the device */
    n_out += B, n_in_samps -= B, samples_len -= len, idx += len;
    b->memory_index = 0;
  }
  for (inp = 0; inp < b->num_input; inp++)
    memmove(b->input_mem_buf[inp], &ptr_src[inp][idx], sizeof(FLOAT32) * (size_t)samples_len);
  b->memory_index = samples_len;
  return n_out;
}

__attribute__((constructor)) static void seam_register_bin(void) { seam_register(RQ_BINAURAL, be_bin, "binaural_ld_ola"); }
