/* seam_interpose.c -- libmpeghb200_seam.so: the drop-in at library level.
 *
 * The reference decoder library keeps its public API (ia_mpegh_dec_create / _init / _execute / _delete,
 * decoder/impeghd_api.h:47-51) and its whole parse half.  This shared object is placed in front of it in the
 * link / LD_PRELOAD order and takes over the post-parse signal functions by ELF symbol interposition -- the
 * reference library is used binary-unmodified, no source patch is needed:
 *
 *   ia_core_coder_fd_frm_dec             decoder/ia_core_coder_imdct.c:832      -> mpeghb200_fd_synth_dev
 *   impeghd_tns_decode_coeff_ar_filter   decoder/ia_core_coder_tns.c:83         -> mpeghb200_tns_dev
 *   ia_core_coder_usac_ltpf_process      decoder/ia_core_coder_ltpf.c:395       -> mpeghb200_ltpf_dev
 *   impeghd_mc_process                   decoder/impeghd_multichannel.c:899     -> mpeghb200_mct_dev
 *   impeghd_format_converter_process     decoder/ia_core_coder_process.c:641    -> mpeghb200_format_conv_dev
 *                                        (seam_interpose_fc.c)
 *   impeghd_peak_limiter_init/_process   decoder/impeghd_peak_limiter.c:146,190 -> mpeghb200_limiter_dev
 *
 * Every call here handles ONE handle's frame, so it moves a few KB per launch: this object exists to prove that
 * a stock application (the reference testbench decoding the smoke_test_suite streams) produces identical PCM
 * with these stages on the GPU.  Throughput comes from the batched sessions of include/mpeghb200.h, which run
 * the same kernels over thousands of handles per launch.
 *
 * State ownership: the FD overlap buffers stay in the reference handle (uploaded and downloaded around the
 * kernel) because the reference's LPD / FAC transition code, which stays on the host, reads and writes them
 * too; channel-frames in an LPD transition (td_frame_prev or FAC data present) are forwarded to the reference
 * function.  The limiter's delay line and sliding-maximum history live on the device, one slot per
 * ia_peak_limiter_struct, reset whenever the reference re-initialises that struct.
 *
 * No CPU fallback: if the CUDA context cannot be created the process is stopped with a message; a stage is
 * never silently computed by the reference instead.  MPEGHB200_SEAM_STATS=1 prints the call counters at exit.
 *
 * Built only where the reference headers are available (csrc/Makefile, target seam); struct layouts come from
 * the reference's own headers, read in place.
 */
#define _GNU_SOURCE
#include <dlfcn.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

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
#include "impeghd_error_codes.h"
#include "impeghd_peak_limiter_struct_def.h"
#include "impeghd_hoa_common_values.h"
#include "impeghd_hoa_matrix_struct.h"
#include "impeghd_hoa_matrix.h"
#include "impd_drc_common.h"
#include "impd_drc_extr_delta_coded_info.h"
#include "impd_drc_struct.h"
#include "ia_core_coder_config.h"

#include "seam_internal.h"

#define FRAME 1024
#define MAX_LIM 8

static struct
{
  mpeghb200_ctx *ctx;
  float *d_coef, *d_ovl, *d_out;
  uint32_t *d_side;
  /* TNS scratch */
  mpeghb200_tns_filter *d_filt;
  int32_t *d_chain;
  /* limiter */
  float *d_lim_in, *d_lim_out;
  struct
  {
    const ia_peak_limiter_struct *key;
    float *d_state;
    mpeghb200_limiter_params p;
    int fresh;
  } lim[MAX_LIM];
  float *h_planar;
  float *d_ltpf_state;
  mpeghb200_ltpf_side *d_ltpf_side;
  /* MCT scratch: two spectrum rows, one box, one group */
  float *d_mct_spec;
  mpeghb200_mct_box *d_mct_box;
  int32_t *d_mct_first;
  unsigned long n_fd_gpu, n_fd_host, n_tns, n_lim, n_ltpf, n_fc, n_mct;
} g;

void seam_die(const char *what, mpeghb200_status st)
{
  fprintf(stderr, "mpeghb200 seam: %s failed (0x%08x): %s\n", what, (unsigned)st, mpeghb200_last_error(g.ctx));
  abort();
}
static void seam_stats(void)
{
  if (getenv("MPEGHB200_SEAM_STATS"))
    fprintf(stderr,
            "mpeghb200 seam: fd_frm_dec gpu=%lu host(LPD transition)=%lu tns_filters gpu=%lu limiter_frames gpu=%lu "
            "ltpf_frames gpu=%lu format_conv_frames gpu=%lu mct_boxes gpu=%lu kernel_launches=%llu\n",
            g.n_fd_gpu, g.n_fd_host, g.n_tns, g.n_lim, g.n_ltpf, g.n_fc, g.n_mct,
            g.ctx ? (unsigned long long)mpeghb200_launch_count(g.ctx) : 0ull);
}

static void seam_init(void)
{
  const char *dev = getenv("MPEGHB200_DEVICE");
  if (g.ctx)
    return;
  CK(mpeghb200_create(&g.ctx, dev ? atoi(dev) : 0));
  CK(mpeghb200_dev_alloc(g.ctx, FRAME * sizeof(float), (void **)&g.d_coef));
  CK(mpeghb200_dev_alloc(g.ctx, FRAME * sizeof(float), (void **)&g.d_ovl));
  CK(mpeghb200_dev_alloc(g.ctx, FRAME * sizeof(float), (void **)&g.d_out));
  CK(mpeghb200_dev_alloc(g.ctx, sizeof(uint32_t), (void **)&g.d_side));
  CK(mpeghb200_dev_alloc(g.ctx, sizeof(mpeghb200_tns_filter), (void **)&g.d_filt));
  CK(mpeghb200_dev_alloc(g.ctx, 2 * sizeof(int32_t), (void **)&g.d_chain));
  CK(mpeghb200_dev_alloc(g.ctx, MAX_NUM_CHANNELS * FRAME * sizeof(float), (void **)&g.d_lim_in));
  CK(mpeghb200_dev_alloc(g.ctx, MAX_NUM_CHANNELS * FRAME * sizeof(float), (void **)&g.d_lim_out));
  CK(mpeghb200_dev_alloc(g.ctx, 4096 * sizeof(float), (void **)&g.d_ltpf_state));
  CK(mpeghb200_dev_alloc(g.ctx, sizeof(mpeghb200_ltpf_side), (void **)&g.d_ltpf_side));
  CK(mpeghb200_dev_alloc(g.ctx, 2 * FRAME * sizeof(float), (void **)&g.d_mct_spec));
  CK(mpeghb200_dev_alloc(g.ctx, sizeof(mpeghb200_mct_box), (void **)&g.d_mct_box));
  CK(mpeghb200_dev_alloc(g.ctx, 2 * sizeof(int32_t), (void **)&g.d_mct_first));
  g.h_planar = (float *)malloc(MAX_NUM_CHANNELS * FRAME * sizeof(float));
  atexit(seam_stats);
}

mpeghb200_ctx *seam_ctx(void)
{
  seam_init();
  return g.ctx;
}
void seam_count_fc(void) { g.n_fc++; }

/* ---- FD synthesis ------------------------------------------------------------------------------------ */
IA_ERRORCODE ia_core_coder_fd_frm_dec(ia_usac_data_struct *usac_data, WORD32 i_ch, WORD32 ele_id)
{
  typedef IA_ERRORCODE (*fn_t)(ia_usac_data_struct *, WORD32, WORD32);
  uint32_t side;
  if (usac_data->td_frame_prev[i_ch] || usac_data->fac_data_present[i_ch] || usac_data->ccfl != FRAME)
  {
    /* LPD -> FD transition frame: FAC / BPF / TBE hand-over stays with the host LPD decoder */
    static fn_t real;
    if (!real)
      real = (fn_t)dlsym(RTLD_NEXT, "ia_core_coder_fd_frm_dec");
    g.n_fd_host++;
    return real(usac_data, i_ch, ele_id);
  }
  seam_init();
  side = MPEGHB200_FD_SIDE(usac_data->window_sequence[i_ch], usac_data->window_shape[i_ch],
                           usac_data->window_shape_prev[i_ch], usac_data->prev_aliasing_symmetry[i_ch],
                           usac_data->curr_aliasing_symmetry[i_ch]);
  CK(mpeghb200_h2d(g.ctx, g.d_coef, usac_data->coef[i_ch], FRAME * sizeof(float)));
  CK(mpeghb200_h2d(g.ctx, g.d_ovl, usac_data->overlap_data_ptr[i_ch], FRAME * sizeof(float)));
  CK(mpeghb200_h2d(g.ctx, g.d_side, &side, sizeof(side)));
  CK(mpeghb200_fd_synth_dev(g.ctx, g.d_coef, g.d_side, g.d_ovl, g.d_out, 1, NULL));
  CK(mpeghb200_d2h(g.ctx, usac_data->time_sample_vector[i_ch], g.d_out, FRAME * sizeof(float)));
  CK(mpeghb200_d2h(g.ctx, usac_data->overlap_data_ptr[i_ch], g.d_ovl, FRAME * sizeof(float)));
  g.n_fd_gpu++;
  return IA_MPEGH_DEC_NO_ERROR;
}

/* ---- TNS: one filter over one spectral range (ia_core_coder_tns_apply resolves the bands on the host) ---- */
VOID impeghd_tns_decode_coeff_ar_filter(WORD32 order, WORD32 coef_res, WORD16 *coef, FLOAT32 *ptr_spec,
                                        WORD32 size, WORD32 inc)
{
  mpeghb200_tns_filter f;
  int32_t chain[2] = {0, 1};
  if (order <= 0 || size <= 0)
    return;
  seam_init();
  memset(&f, 0, sizeof(f));
  f.unit = 0;
  f.start = 0;
  f.size = size;
  f.inc = inc < 0 ? -1 : 1;
  f.order = order;
  CK(mpeghb200_tns_lpc_from_coef(order, coef_res, coef, f.lpc));
  CK(mpeghb200_h2d(g.ctx, g.d_coef, ptr_spec, (size_t)size * sizeof(float)));
  CK(mpeghb200_h2d(g.ctx, g.d_filt, &f, sizeof(f)));
  CK(mpeghb200_h2d(g.ctx, g.d_chain, chain, sizeof(chain)));
  CK(mpeghb200_tns_dev(g.ctx, g.d_coef, g.d_filt, g.d_chain, 1, NULL));
  CK(mpeghb200_d2h(g.ctx, ptr_spec, g.d_coef, (size_t)size * sizeof(float)));
  g.n_tns++;
}

/* ---- MCT: one window of one box per call (the window loop is the caller's, process.c:1043-1075) ---- */
VOID impeghd_mc_process(ia_multichannel_data *ptr_mc_data, FLOAT32 *ptr_dmx, FLOAT32 *ptr_res,
                        WORD32 *ptr_coeff_sfb_idx, WORD32 *ptr_mask, WORD32 bands_per_window, WORD32 total_sfb,
                        WORD32 pair, const WORD16 *ptr_sfb_offset, WORD32 num_samples)
{
  mpeghb200_mct_box box;
  const int32_t first[2] = {0, 1};
  const int fullband = !ptr_mc_data->bandwise_coeff_flag[pair] && !ptr_mc_data->mask_flag[pair];
  if (!ptr_mc_data->use_tool)
    return;
  seam_init();
  CK(mpeghb200_mct_resolve(ptr_mc_data->signal_type, ptr_mc_data->pred_dir[pair], fullband, 1, num_samples,
                           bands_per_window, total_sfb, ptr_sfb_offset, ptr_coeff_sfb_idx, ptr_mask, 0, 1, &box));
  if (box.n_seg == 0)
    return;
  CK(mpeghb200_h2d(g.ctx, g.d_mct_spec, ptr_dmx, (size_t)num_samples * sizeof(float)));
  CK(mpeghb200_h2d(g.ctx, g.d_mct_spec + FRAME, ptr_res, (size_t)num_samples * sizeof(float)));
  CK(mpeghb200_h2d(g.ctx, g.d_mct_box, &box, sizeof(box)));
  CK(mpeghb200_h2d(g.ctx, g.d_mct_first, first, sizeof(first)));
  CK(mpeghb200_mct_dev(g.ctx, g.d_mct_spec, g.d_mct_box, g.d_mct_first, 1, NULL));
  CK(mpeghb200_d2h(g.ctx, ptr_dmx, g.d_mct_spec, (size_t)num_samples * sizeof(float)));
  CK(mpeghb200_d2h(g.ctx, ptr_res, g.d_mct_spec + FRAME, (size_t)num_samples * sizeof(float)));
  g.n_mct++;
}

/* ---- LTPF: the filter state stays in the reference handle (the LPD path shares it), round trip per call ---- */
VOID ia_core_coder_usac_ltpf_process(ia_ltpf_data_str *ptr_ltpf_data, FLOAT32 *time_sample_vector)
{
  typedef VOID (*fn_t)(ia_ltpf_data_str *, FLOAT32 *);
  const int hist = ptr_ltpf_data->pitch_max + LTPF_NUM_FILT_COEF1 / 2;
  float st[4096];
  mpeghb200_ltpf_side sd;
  int32_t iv;
  size_t n_state;
  if (ptr_ltpf_data->frame_len != FRAME || ptr_ltpf_data->pitch_min != 128)
  { /* not a 48 kHz / 1024-sample FD channel: the LPD path's own use of the filter stays on the host */
    static fn_t real;
    if (!real)
      real = (fn_t)dlsym(RTLD_NEXT, "ia_core_coder_usac_ltpf_process");
    real(ptr_ltpf_data, time_sample_vector);
    return;
  }
  seam_init();
  n_state = mpeghb200_ltpf_state_floats(48000);
  memset(st, 0, sizeof(st));
  iv = ptr_ltpf_data->prev_pitch, memcpy(&st[0], &iv, 4);
  iv = ptr_ltpf_data->prev_pitch_fr, memcpy(&st[1], &iv, 4);
  iv = ptr_ltpf_data->prev_gain_idx, memcpy(&st[2], &iv, 4);
  st[3] = ptr_ltpf_data->prev_gain;
  memcpy(&st[4], ptr_ltpf_data->prev_in_buf, 6 * sizeof(float));
  memcpy(&st[16], ptr_ltpf_data->prev_out_buf, hist * sizeof(float));
  sd.unit = 0;
  sd.data_present = (uint8_t)ptr_ltpf_data->data_present;
  sd.pitch_lag_idx = (uint16_t)ptr_ltpf_data->pitch_lag_idx;
  sd.gain_idx = (uint8_t)ptr_ltpf_data->gain_idx;
  CK(mpeghb200_h2d(g.ctx, g.d_ltpf_state, st, n_state * sizeof(float)));
  CK(mpeghb200_h2d(g.ctx, g.d_ltpf_side, &sd, sizeof(sd)));
  CK(mpeghb200_h2d(g.ctx, g.d_out, time_sample_vector, FRAME * sizeof(float)));
  CK(mpeghb200_ltpf_dev(g.ctx, 48000, g.d_out, g.d_ltpf_side, g.d_ltpf_state, 1, NULL));
  CK(mpeghb200_d2h(g.ctx, time_sample_vector, g.d_out, FRAME * sizeof(float)));
  CK(mpeghb200_d2h(g.ctx, st, g.d_ltpf_state, n_state * sizeof(float)));
  memcpy(&iv, &st[0], 4), ptr_ltpf_data->prev_pitch = iv;
  memcpy(&iv, &st[1], 4), ptr_ltpf_data->prev_pitch_fr = iv;
  memcpy(&iv, &st[2], 4), ptr_ltpf_data->prev_gain_idx = iv;
  ptr_ltpf_data->prev_gain = st[3];
  memcpy(ptr_ltpf_data->prev_in_buf, &st[4], 6 * sizeof(float));
  memcpy(ptr_ltpf_data->prev_out_buf, &st[16], hist * sizeof(float));
  g.n_ltpf++;
}

/* ---- peak limiter ---------------------------------------------------------------------------------- */
static int lim_slot(const ia_peak_limiter_struct *key, int create)
{
  int i;
  for (i = 0; i < MAX_LIM; i++)
    if (g.lim[i].key == key)
      return i;
  if (!create)
    return -1;
  for (i = 0; i < MAX_LIM; i++)
    if (!g.lim[i].key)
    {
      g.lim[i].key = key;
      g.lim[i].fresh = 1;
      return i;
    }
  fprintf(stderr, "mpeghb200 seam: more than %d limiter instances\n", MAX_LIM);
  abort();
}

IA_ERRORCODE impeghd_peak_limiter_init(ia_peak_limiter_struct *pstr_peak_limiter, UWORD32 channel_num,
                                       UWORD32 sample_rate, FLOAT32 *ptr_buff)
{
  typedef IA_ERRORCODE (*fn_t)(ia_peak_limiter_struct *, UWORD32, UWORD32, FLOAT32 *);
  static fn_t real;
  int s;
  if (!real)
    real = (fn_t)dlsym(RTLD_NEXT, "impeghd_peak_limiter_init");
  s = lim_slot(pstr_peak_limiter, 0);
  if (s >= 0)
    g.lim[s].fresh = 1; /* the reference re-initialised this instance: restart the device state too */
  return real(pstr_peak_limiter, channel_num, sample_rate, ptr_buff);
}

VOID impeghd_peak_limiter_process(ia_peak_limiter_struct *pstr_peak_limiter, FLOAT32 *ptr_samples,
                                  UWORD32 frame_size)
{
  const int n_ch = (int)pstr_peak_limiter->num_channels;
  int s, c;
  UWORD32 j;
  if (frame_size != FRAME || n_ch < 1 || n_ch > MAX_NUM_CHANNELS)
  {
    fprintf(stderr, "mpeghb200 seam: limiter frame of %u samples x %d channels is outside the GPU path\n",
            (unsigned)frame_size, n_ch);
    abort();
  }
  seam_init();
  s = lim_slot(pstr_peak_limiter, 1);
  if (g.lim[s].fresh || g.lim[s].p.n_ch != n_ch)
  {
    if (g.lim[s].d_state)
      CK(mpeghb200_dev_free(g.ctx, g.lim[s].d_state));
    CK(mpeghb200_limiter_params_init(&g.lim[s].p, n_ch, (int)pstr_peak_limiter->sample_rate,
                                     (int)pstr_peak_limiter->limiter_on));
    CK(mpeghb200_dev_alloc(g.ctx, mpeghb200_limiter_state_floats(&g.lim[s].p) * sizeof(float),
                           (void **)&g.lim[s].d_state));
    CK(mpeghb200_limiter_state_init_dev(g.ctx, &g.lim[s].p, g.lim[s].d_state, 1, NULL));
    g.lim[s].fresh = 0;
  }
  /* the caller hands over interleaved [sample][channel] (impegh_dec_peak_limiter); the kernel is planar */
  for (c = 0; c < n_ch; c++)
    for (j = 0; j < FRAME; j++)
      g.h_planar[c * FRAME + j] = ptr_samples[j * n_ch + c];
  CK(mpeghb200_h2d(g.ctx, g.d_lim_in, g.h_planar, (size_t)n_ch * FRAME * sizeof(float)));
  CK(mpeghb200_limiter_dev(g.ctx, &g.lim[s].p, g.d_lim_in, g.d_lim_out, g.lim[s].d_state, NULL, 1, NULL));
  CK(mpeghb200_d2h(g.ctx, g.h_planar, g.d_lim_out, (size_t)n_ch * FRAME * sizeof(float)));
  for (c = 0; c < n_ch; c++)
    for (j = 0; j < FRAME; j++)
      ptr_samples[j * n_ch + c] = g.h_planar[c * FRAME + j];
  g.n_lim++;
}
