/* oracle/ref_harness.c -- TEST INFRASTRUCTURE ONLY.
 *
 * Function-level entry points into the UNMODIFIED reference decoder (compiled from /root/reference in
 * place by oracle/Makefile into oracle/_ref/libmpeghref.so).  Every export here takes raw float / int
 * arrays ("inputs + state in -> outputs + state out"), fills the reference's own structs and calls the
 * reference's own functions, so that tests can pin both the plain-C restatement (oracle_*.c) and the CUDA
 * path against the real thing.  Nothing in the product library may link or load this.
 *
 * This file contains no reference code, only calls into it.
 */
#include <stdlib.h>
#include <string.h>
#include <math.h>

#include <impeghd_type_def.h>
#include <impeghd_fft_ifft.h>
#include "impeghd_intrinsics_flt.h"
#include "ia_core_coder_cnst.h"
#include <ia_core_coder_constants.h>
#include "ia_core_coder_acelp_info.h"
#include "ia_core_coder_acelp_com.h"
#include <ia_core_coder_basic_ops32.h>
#include <ia_core_coder_basic_ops40.h>
#include "ia_core_coder_bitbuffer.h"
#include "ia_core_coder_defines.h"
#include "ia_core_coder_function_selector.h"
#include "ia_core_coder_interface.h"
#include "ia_core_coder_info.h"
#include "ia_core_coder_tns_usac.h"
#include "impeghd_tbe_dec.h"
#include "ia_core_coder_igf_data_struct.h"
#include "ia_core_coder_stereo_lpd.h"
#include "ia_core_coder_main.h"
#include "ia_core_coder_func_def.h"
#include "ia_core_coder_rom.h"
#include "ia_core_coder_arith_dec.h"
#include "ia_core_coder_windows.h"
#include "ia_core_coder_vec_baisc_ops.h"
#include "impeghd_error_codes.h"

/* ------------------------------------------------------------------------------------------------
 * FD synthesis: ia_core_coder_fd_frm_dec (ia_core_coder_imdct.c:832)
 * ---------------------------------------------------------------------------------------------- */
typedef struct
{
  ia_usac_data_struct usac;
  FLOAT32 scratch_float[4 * 4096];
  FLOAT32 fft_scratch[8 * 4096];
} ref_fd_handle;

void *ref_fd_create(void)
{
  ref_fd_handle *h = (ref_fd_handle *)calloc(1, sizeof(ref_fd_handle));
  int ch;
  if (!h)
    return 0;
  h->usac.ccfl = 1024;
  h->usac.scratch_buffer_float = h->scratch_float;
  h->usac.ptr_fft_scratch = h->fft_scratch;
  for (ch = 0; ch < MAX_NUM_CHANNELS; ch++)
  {
    h->usac.coef[ch] = h->usac.arr_coef[ch];
    h->usac.str_tddec[ch] = &h->usac.arr_str_tddec[ch];
  }
  return h;
}

void ref_fd_destroy(void *p) { free(p); }

/* One channel-frame.  coef[1024] (copied: the reference transforms in place), overlap[1024] in/out,
 * out[1024].  FD-only channel history (td_frame_prev = 0, no FAC). Returns the reference error code. */
int ref_fd_frm_dec(void *p, const float *coef, float *overlap, float *out, int window_sequence,
                   int window_shape, int window_shape_prev, int prev_aliasing_symmetry,
                   int curr_aliasing_symmetry)
{
  ref_fd_handle *h = (ref_fd_handle *)p;
  ia_usac_data_struct *u = &h->usac;
  const int ch = 0;
  int err;
  memcpy(u->coef[ch], coef, 1024 * sizeof(float));
  memcpy(u->overlap_data_ptr[ch], overlap, 1024 * sizeof(float));
  u->window_sequence[ch] = window_sequence;
  u->window_shape[ch] = window_shape;
  u->window_shape_prev[ch] = window_shape_prev;
  u->prev_aliasing_symmetry[ch] = (WORD16)prev_aliasing_symmetry;
  u->curr_aliasing_symmetry[ch] = (WORD16)curr_aliasing_symmetry;
  u->td_frame_prev[ch] = 0;
  u->fac_data_present[ch] = 0;
  err = ia_core_coder_fd_frm_dec(u, ch, 0);
  memcpy(out, u->time_sample_vector[ch], 1024 * sizeof(float));
  memcpy(overlap, u->overlap_data_ptr[ch], 1024 * sizeof(float));
  return err;
}

/* n_frames consecutive frames of n_ch channels, state carried; layout [frame][ch][1024].
 * side[frame][ch][5] = {window_sequence, window_shape, window_shape_prev, prev_sym, curr_sym}. */
int ref_fd_frm_dec_batch(void *p, const float *coef, float *overlap /*[n_ch][1024]*/, float *out,
                         const int *side, int n_frames, int n_ch)
{
  int f, c, err = 0;
  for (f = 0; f < n_frames; f++)
    for (c = 0; c < n_ch; c++)
    {
      const int *s = side + ((size_t)f * n_ch + c) * 5;
      size_t o = ((size_t)f * n_ch + c) * 1024;
      err |= ref_fd_frm_dec(p, coef + o, overlap + (size_t)c * 1024, out + o, s[0], s[1], s[2], s[3],
                            s[4]);
    }
  return err;
}

/* Plain complex transforms (impeghd_ifft.c:297, impeghd_fft.c:297). */
void ref_rad2_cplx_ifft(float *re, float *im, int n)
{
  static __thread FLOAT32 scratch[8 * 4096];
  impeghd_rad2_cplx_ifft(re, im, n, scratch);
}
void ref_rad2_cplx_fft(float *re, float *im, int n)
{
  static __thread FLOAT32 scratch[8 * 4096];
  impeghd_rad2_cplx_fft(re, im, n, scratch);
}

/* ------------------------------------------------------------------------------------------------
 * Peak limiter: impeghd_peak_limiter_init / _process (impeghd_peak_limiter.c:146,190) with the planar <->
 * interleaved transposes of impegh_dec_peak_limiter (ia_core_coder_decode_main.c:1429).
 * ---------------------------------------------------------------------------------------------- */
#include "impeghd_peak_limiter_struct_def.h"
IA_ERRORCODE impeghd_peak_limiter_init(ia_peak_limiter_struct *peak_limiter, UWORD32 num_channels,
                                       UWORD32 sample_rate, FLOAT32 *buffer);
VOID impeghd_peak_limiter_process(ia_peak_limiter_struct *peak_limiter, FLOAT32 *samples,
                                  UWORD32 frame_len);

typedef struct
{
  ia_peak_limiter_struct lim;
  FLOAT32 inter[MAX_NUM_CHANNELS * 4096];
} ref_limiter_handle;

void *ref_limiter_create(int n_ch, int sample_rate, int limiter_on)
{
  ref_limiter_handle *h = (ref_limiter_handle *)calloc(1, sizeof(ref_limiter_handle));
  if (!h)
    return 0;
  impeghd_peak_limiter_init(&h->lim, (UWORD32)n_ch, (UWORD32)sample_rate, &h->lim.buffer[0]);
  h->lim.limiter_on = (UWORD32)limiter_on;
  return h;
}
void ref_limiter_destroy(void *p) { free(p); }

/* planar [n_ch][n] in place, one frame */
void ref_limiter_process(void *p, float *planar, int n)
{
  ref_limiter_handle *h = (ref_limiter_handle *)p;
  int nch = (int)h->lim.num_channels, i, j;
  for (i = 0; i < nch; i++)
    for (j = 0; j < n; j++)
      h->inter[j * nch + i] = planar[i * n + j];
  impeghd_peak_limiter_process(&h->lim, h->inter, (UWORD32)n);
  for (i = 0; i < nch; i++)
    for (j = 0; j < n; j++)
      planar[i * n + j] = h->inter[j * nch + i];
}
void ref_limiter_constants(void *p, float *out3)
{
  ref_limiter_handle *h = (ref_limiter_handle *)p;
  out3[0] = h->lim.attack_constant;
  out3[1] = h->lim.release_constant;
  out3[2] = h->lim.limit_threshold;
}

/* ------------------------------------------------------------------------------------------------
 * TNS: ia_core_coder_tns_apply (ia_core_coder_tns.c:204) on one channel's spectrum.
 * filt [n_win][3][20] ints = {order, start_band, stop_band, direction, coef_res, coef[15]} (order 0 = unused),
 * sfb_tbl = upper band edges of one window, spec [1024] in place.
 * ---------------------------------------------------------------------------------------------- */
int ref_tns_apply(float *spec, int islong, const int *n_filt, const int *filt, const short *sfb_tbl, int sfb_per_sbk,
                  int tns_max_band, int nbands)
{
  static __thread ia_usac_data_struct usac;
  static __thread FLOAT32 scratch[4096];
  static __thread WORD32 max_tbl[16][2];
  ia_sfb_info_struct info;
  ia_tns_frame_info_struct tns;
  ia_usac_igf_config_struct igf;
  int w, f, k, n_win = islong ? 1 : 8;
  memset(&info, 0, sizeof(info));
  memset(&tns, 0, sizeof(tns));
  memset(&igf, 0, sizeof(igf));
  info.islong = islong;
  info.ptr_sfb_tbl = sfb_tbl;
  info.sfb_per_sbk = sfb_per_sbk;
  info.bins_per_sbk = islong ? 1024 : 128;
  max_tbl[0][0] = max_tbl[0][1] = tns_max_band;
  usac.tns_max_bands_tbl_usac = &max_tbl;
  usac.sampling_rate_idx = 0;
  usac.scratch_buffer_float = scratch;
  tns.n_subblocks = n_win;
  for (w = 0; w < n_win; w++)
  {
    tns.str_tns_info[w].n_filt = n_filt[w];
    for (f = 0; f < 3; f++)
    {
      const int *p = filt + (w * 3 + f) * 20;
      ia_tns_filter_struct *q = &tns.str_tns_info[w].str_filter[f];
      tns.str_tns_info[w].coef_res = p[4] ? p[4] : tns.str_tns_info[w].coef_res;
      q->order = p[0], q->start_band = p[1], q->stop_band = p[2], q->direction = p[3];
      for (k = 0; k < 15; k++)
        q->coef[k] = (WORD16)p[5 + k];
    }
  }
  return (int)ia_core_coder_tns_apply(&usac, spec, nbands, &info, &tns, &igf);
}
