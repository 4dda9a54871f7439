/* ref_harness_drc.c -- TEST INFRASTRUCTURE: function-level entry point into the UNMODIFIED reference build for the
 * sub-band DRC gain application (decoder/impd_drc_process.c:243).  Compiled into oracle/_ref/libmpeghref.so only. */
#include <stdlib.h>
#include <string.h>

#include <impeghd_type_def.h>
#include "impd_drc_common.h"
#include "impeghd_error_codes.h"
#include "impd_drc_extr_delta_coded_info.h"
#include "impd_drc_struct.h"
#include "impd_drc_filter_bank.h"
#include "impd_drc_gain_dec.h"
#include "impd_drc_multi_band.h"
#include "impd_drc_process_audio.h"

IA_ERRORCODE impd_drc_apply_gains_subband(FLOAT32 *deinterleaved_audio_re[], FLOAT32 *deinterleaved_audio_im[],
                                          ia_drc_instructions_struct *pstr_drc_instruction_arr,
                                          ia_drc_gain_dec_struct *pstr_drc_gain_dec, WORD32 sel_drc_idx);

/* same flat arguments as oracle_drc_gains_subband, plus where the reference shall find the gains inside lpcm_gains:
 * delay_mode (DELAY_MODE_LOW adds one frame) and gain_delay_samples */
int ref_drc_gains_subband(float *re, float *im, int n_ch, const int *group_of_ch, const int *band_count, int n_groups,
                          const float *gains, const float *weights, int n_slots, int n_sb, int mode, int delay_mode,
                          int gain_delay)
{
  ia_drc_gain_dec_struct *dec = (ia_drc_gain_dec_struct *)calloc(1, sizeof(*dec));
  ia_drc_instructions_struct *ins = (ia_drc_instructions_struct *)calloc(1, sizeof(*ins));
  ia_drc_params_struct *p = &dec->ia_drc_params_struct;
  FLOAT32 *pre[MAX_CHANNEL_COUNT], *pim[MAX_CHANNEL_COUNT];
  const int ds = 1024 / n_slots;
  int n_seq = 0, err;
  for (int g = 0; g < n_groups; ++g) n_seq += band_count[g];
  p->sub_band_domain_mode = mode;
  p->drc_frame_size = 1024;
  p->delay_mode = delay_mode;
  p->gain_delay_samples = gain_delay;
  p->channel_offset = 0;
  p->num_ch_process = -1;
  p->sel_drc_array[0].drc_instrns_idx = 0;
  ins->drc_set_id = 1;
  ins->audio_num_chan = n_ch;
  ins->num_drc_ch_groups = n_groups;
  for (int g = 0; g < n_groups; ++g)
  {
    ins->band_count_of_ch_group[g] = band_count[g];
    for (int b = 0; b < band_count[g] && b < BAND_COUNT_MAX; ++b)
      memcpy(dec->str_overlap_params.str_grp_overlap_params[g].str_band_overlap_params[b].overlap_weight,
             weights + (g * 4 + b) * 256, 256 * sizeof(float));
  }
  for (int c = 0; c < n_ch; ++c)
  {
    ins->channel_group_of_ch[c] = group_of_ch[c];
    pre[c] = re + (long)c * n_slots * n_sb;
    pim[c] = im + (long)c * n_slots * n_sb;
  }
  dec->drc_gain_buffers.pstr_gain_buf[0].pstr_buf_interp =
      (ia_drc_interp_buf_struct *)calloc(n_seq ? n_seq : 1, sizeof(ia_drc_interp_buf_struct));
  dec->drc_gain_buffers.pstr_gain_buf[0].buf_interp_cnt = n_seq;
  for (int k = 0; k < n_seq; ++k)
  {
    float *l = dec->drc_gain_buffers.pstr_gain_buf[0].pstr_buf_interp[k].lpcm_gains + MAX_SIGNAL_DELAY - gain_delay +
               (delay_mode == DELAY_MODE_LOW ? 1024 : 0);
    for (int s = 0; s < n_slots; ++s) l[s * ds + (ds - 1) / 2] = gains[k * n_slots + s];
  }
  err = impd_drc_apply_gains_subband(pre, pim, ins, dec, 0);
  free(dec->drc_gain_buffers.pstr_gain_buf[0].pstr_buf_interp);
  free(ins);
  free(dec);
  return err;
}

/* ---- time-domain DRC through impd_drc_td_process (decoder/impd_drc_process.c:529) ---------------------------------
 * The whole step the decoder runs per frame: impd_drc_get_gain (the reference's own gain-buffer bookkeeping: it advances
 * lpcm_gains by one frame; the harness has no spline nodes to decode, so it places the frame's gain curves one frame
 * ahead, where the advance picks them up), then filter banks, gains and band sum.  Same arguments as ref_drctd_* of
 * ref_harness_ns.c, which calls the two inner functions directly. */
typedef struct
{
  float a1, a2, b0, b1, b2;
} hw_biquad;
typedef struct
{
  int n_bands, n_cascade;
  hw_biquad f[8], ap[9];
} hw_bank; /* == oracle_drc_bank */
typedef struct
{
  ia_drc_gain_dec_struct dec;
  ia_drc_config *cfg;
  ia_drc_gain_struct *gain;
  FLOAT32 *band_ptr[MAX_CHANNEL_COUNT * 4];
  FLOAT32 *bands;
  int n_ch, n_seq;
} ref_drctdw_handle;

static void hw_set(ia_iir_filter_struct *d, const hw_biquad *s)
{
  d->a0 = 1.0f, d->a1 = s->a1, d->a2 = s->a2, d->b0 = s->b0, d->b1 = s->b1, d->b2 = s->b2;
}

void *ref_drctd_whole_create(int n_ch, const int *group_of_ch, int n_groups, const void *banks_v)
{
  const hw_bank *banks = (const hw_bank *)banks_v;
  ref_drctdw_handle *h = (ref_drctdw_handle *)calloc(1, sizeof(*h));
  ia_drc_params_struct *p = &h->dec.ia_drc_params_struct;
  ia_drc_instructions_struct *ins;
  h->cfg = (ia_drc_config *)calloc(1, sizeof(ia_drc_config));
  h->gain = (ia_drc_gain_struct *)calloc(1, sizeof(ia_drc_gain_struct));
  ins = &h->cfg->str_drc_instruction_str[0];
  h->n_ch = n_ch;
  h->bands = (FLOAT32 *)calloc((size_t)MAX_CHANNEL_COUNT * 4 * 1024, sizeof(FLOAT32));
  for (int i = 0; i < MAX_CHANNEL_COUNT * 4; ++i) h->band_ptr[i] = h->bands + (size_t)i * 1024;
  h->dec.audio_band_buffer.non_interleaved_audio = h->band_ptr;
  h->cfg->apply_drc = 1;
  p->drc_frame_size = 1024;
  p->drc_set_counter = 1;
  p->channel_offset = 0;
  p->num_ch_process = n_ch;
  p->multiband_sel_drc_idx = 0;
  p->sel_drc_array[0].drc_instrns_idx = 0;
  ins->drc_set_id = 1;
  ins->audio_num_chan = n_ch;
  ins->num_drc_ch_groups = n_groups;
  for (int c = 0; c < n_ch; ++c) ins->channel_group_of_ch[c] = group_of_ch[c];
  for (int g = 0; g <= n_groups; ++g)
  {
    ia_drc_filter_bank_struct *b = &h->dec.ia_filter_banks_struct.str_drc_filter_bank[g];
    const hw_bank *s = &banks[g];
    if (g < n_groups)
    {
      ins->band_count_of_ch_group[g] = s->n_bands;
      h->n_seq += s->n_bands;
    }
    b->num_bands = s->n_bands;
    b->str_all_pass_cascade.num_filter = s->n_cascade > 0 ? s->n_cascade - 1 : 0;
    for (int i = 0; i < s->n_cascade && i < CASCADE_ALLPASS_COUNT_MAX; ++i)
      hw_set(&b->str_all_pass_cascade.str_ap_cascade_filt[i].str_ap_stage, &s->ap[i]);
    hw_set(&b->str_two_band_filt_bank.str_lp, &s->f[0]);
    hw_set(&b->str_two_band_filt_bank.str_hp, &s->f[1]);
    hw_set(&b->str_three_band_filt_bank.str_lp_stage_1, &s->f[0]);
    hw_set(&b->str_three_band_filt_bank.str_hp_stage_1, &s->f[1]);
    hw_set(&b->str_three_band_filt_bank.str_ap_stage_2, &s->f[2]);
    hw_set(&b->str_three_band_filt_bank.str_lp_stage_2, &s->f[3]);
    hw_set(&b->str_three_band_filt_bank.str_hp_stage_2, &s->f[4]);
    hw_set(&b->str_four_band_filt_bank.str_lp_stage_1, &s->f[0]);
    hw_set(&b->str_four_band_filt_bank.str_hp_stage_1, &s->f[1]);
    hw_set(&b->str_four_band_filt_bank.str_ap_stage_2_low, &s->f[2]);
    hw_set(&b->str_four_band_filt_bank.str_ap_stage_2_high, &s->f[3]);
    hw_set(&b->str_four_band_filt_bank.str_lp_stage_3_low, &s->f[4]);
    hw_set(&b->str_four_band_filt_bank.str_hp_stage_3_low, &s->f[5]);
    hw_set(&b->str_four_band_filt_bank.str_lp_stage_3_high, &s->f[6]);
    hw_set(&b->str_four_band_filt_bank.str_hp_stage_3_high, &s->f[7]);
  }
  h->dec.drc_gain_buffers.pstr_gain_buf[0].pstr_buf_interp =
      (ia_drc_interp_buf_struct *)calloc(h->n_seq ? h->n_seq : 1, sizeof(ia_drc_interp_buf_struct));
  h->dec.drc_gain_buffers.pstr_gain_buf[0].buf_interp_cnt = h->n_seq;
  return h;
}
void ref_drctd_whole_destroy(void *q)
{
  ref_drctdw_handle *h = (ref_drctdw_handle *)q;
  free(h->dec.drc_gain_buffers.pstr_gain_buf[0].pstr_buf_interp);
  free(h->bands), free(h->cfg), free(h->gain), free(h);
}
/* x [n_ch][1024] in place; gains [n_seq][1024] */
int ref_drctd_whole_process(void *q, float *x, const float *gains)
{
  ref_drctdw_handle *h = (ref_drctdw_handle *)q;
  FLOAT32 *rows[MAX_CHANNEL_COUNT];
  for (int c = 0; c < h->n_ch; ++c) rows[c] = x + (size_t)c * 1024;
  for (int k = 0; k < h->n_seq; ++k)
    memcpy(h->dec.drc_gain_buffers.pstr_gain_buf[0].pstr_buf_interp[k].lpcm_gains + MAX_SIGNAL_DELAY + 1024,
           gains + (size_t)k * 1024, 4096);
  return (int)impd_drc_td_process(rows, &h->dec, h->cfg, h->gain, 0.0f, 1.0f, 1.0f, 0);
}
