/* oracle/ref_harness_ns.c -- TEST INFRASTRUCTURE ONLY.
 *
 * Entry points into `static` functions of the unmodified reference.  oracle/Makefile compiles the reference
 * files that hold them a second time with -Dstatic= (linkage only; the code is unchanged) and links this
 * harness with them into oracle/_ref/libmpeghref_ns.so.  Contains no reference code, only calls into it.
 */
#include <stdlib.h>
#include <string.h>
#include <impeghd_type_def.h>

/* ia_core_coder_samples_sat (ia_core_coder_decode_main.c:209) */
VOID ia_core_coder_samples_sat(WORD8 *outbuffer, WORD32 num_samples_out, WORD32 pcmsize,
                               FLOAT32 **out_samples, WORD32 *out_ch_map, WORD32 *ptr_delay_samples,
                               WORD32 *out_bytes, WORD32 num_channel_out);

/* planar [n_ch][n] floats -> interleaved PCM bytes.  out_ch_map[n_ch] may hold -1 (identity), is not
 * modified; *delay_samples is updated like the reference does.  Returns the byte count reported. */
int ref_samples_sat(const float *planar, int n, int n_ch, int pcmsize, const int *out_ch_map,
                    int *delay_samples, unsigned char *out)
{
  FLOAT32 *rows[64];
  WORD32 map[64];
  WORD32 out_bytes = 0;
  int c;
  for (c = 0; c < n_ch; c++)
  {
    rows[c] = (FLOAT32 *)(planar + (size_t)c * n);
    map[c] = out_ch_map[c];
  }
  ia_core_coder_samples_sat((WORD8 *)out, n, pcmsize, rows, map, delay_samples, &out_bytes, n_ch);
  return out_bytes;
}

/* ---- time-domain multi-band DRC: impd_drc_filter_banks_process (impd_drc_process.c:399) followed by the static
 * impd_drc_apply_gains_and_add (:68), as impd_drc_td_process (:529) chains them for one selected DRC set --------- */
#include "impd_drc_common.h"
#include "impeghd_error_codes.h"
#include "impd_drc_extr_delta_coded_info.h"
#include "impd_drc_struct.h"
#include "impd_drc_filter_bank.h"
#include "impd_drc_gain_dec.h"
#include "impd_drc_multi_band.h"
#include "impd_drc_process_audio.h"

IA_ERRORCODE impd_drc_filter_banks_process(FLOAT32 *audio_io_buf[], ia_drc_instructions_struct *pstr_drc_instruction_arr,
                                           ia_drc_gain_dec_struct *pstr_drc_gain_dec, WORD32 sel_drc_idx);
VOID impd_drc_apply_gains_and_add(ia_drc_gain_dec_struct *pstr_drc_gain_dec, FLOAT32 *channel_audio[],
                                  ia_drc_instructions_struct *pstr_drc_instruction_arr, WORD32 impd_apply_gains,
                                  WORD32 sel_drc_idx);

typedef struct
{
  float a1, a2, b0, b1, b2;
} h_biquad;
typedef struct
{
  int n_bands, n_cascade;
  h_biquad f[8], ap[9];
} h_bank; /* == oracle_drc_bank */

typedef struct
{
  ia_drc_gain_dec_struct dec;
  ia_drc_instructions_struct ins;
  FLOAT32 *band_ptr[MAX_CHANNEL_COUNT * 4];
  FLOAT32 *bands;
  int n_ch, n_seq;
} ref_drctd_handle;

static void set_f(ia_iir_filter_struct *d, const h_biquad *s)
{
  d->a0 = 1.0f, d->a1 = s->a1, d->a2 = s->a2, d->b0 = s->b0, d->b1 = s->b1, d->b2 = s->b2;
}

/* banks [n_groups + 1]: one per channel group, the last one for the channels outside every group */
void *ref_drctd_create(int n_ch, const int *group_of_ch, int n_groups, const void *banks_v)
{
  const h_bank *banks = (const h_bank *)banks_v;
  ref_drctd_handle *h = (ref_drctd_handle *)calloc(1, sizeof(*h));
  ia_drc_params_struct *p = &h->dec.ia_drc_params_struct;
  h->n_ch = n_ch;
  h->bands = (FLOAT32 *)calloc((size_t)MAX_CHANNEL_COUNT * 4 * 1024, sizeof(FLOAT32));
  for (int i = 0; i < MAX_CHANNEL_COUNT * 4; ++i) h->band_ptr[i] = h->bands + (size_t)i * 1024;
  h->dec.audio_band_buffer.non_interleaved_audio = h->band_ptr;
  p->drc_frame_size = 1024;
  p->channel_offset = 0;
  p->num_ch_process = n_ch;
  p->multiband_sel_drc_idx = 0;
  p->sel_drc_array[0].drc_instrns_idx = 0;
  h->ins.drc_set_id = 1;
  h->ins.audio_num_chan = n_ch;
  h->ins.num_drc_ch_groups = n_groups;
  for (int c = 0; c < n_ch; ++c) h->ins.channel_group_of_ch[c] = group_of_ch[c];
  for (int g = 0; g <= n_groups; ++g)
  {
    ia_drc_filter_bank_struct *b = &h->dec.ia_filter_banks_struct.str_drc_filter_bank[g];
    const h_bank *s = &banks[g];
    if (g < n_groups)
    {
      h->ins.band_count_of_ch_group[g] = s->n_bands;
      h->n_seq += s->n_bands;
    }
    b->num_bands = s->n_bands;
    b->str_all_pass_cascade.num_filter = s->n_cascade > 0 ? s->n_cascade - 1 : 0;
    for (int i = 0; i < s->n_cascade && i < CASCADE_ALLPASS_COUNT_MAX; ++i)
      set_f(&b->str_all_pass_cascade.str_ap_cascade_filt[i].str_ap_stage, &s->ap[i]);
    set_f(&b->str_two_band_filt_bank.str_lp, &s->f[0]);
    set_f(&b->str_two_band_filt_bank.str_hp, &s->f[1]);
    set_f(&b->str_three_band_filt_bank.str_lp_stage_1, &s->f[0]);
    set_f(&b->str_three_band_filt_bank.str_hp_stage_1, &s->f[1]);
    set_f(&b->str_three_band_filt_bank.str_ap_stage_2, &s->f[2]);
    set_f(&b->str_three_band_filt_bank.str_lp_stage_2, &s->f[3]);
    set_f(&b->str_three_band_filt_bank.str_hp_stage_2, &s->f[4]);
    set_f(&b->str_four_band_filt_bank.str_lp_stage_1, &s->f[0]);
    set_f(&b->str_four_band_filt_bank.str_hp_stage_1, &s->f[1]);
    set_f(&b->str_four_band_filt_bank.str_ap_stage_2_low, &s->f[2]);
    set_f(&b->str_four_band_filt_bank.str_ap_stage_2_high, &s->f[3]);
    set_f(&b->str_four_band_filt_bank.str_lp_stage_3_low, &s->f[4]);
    set_f(&b->str_four_band_filt_bank.str_hp_stage_3_low, &s->f[5]);
    set_f(&b->str_four_band_filt_bank.str_lp_stage_3_high, &s->f[6]);
    set_f(&b->str_four_band_filt_bank.str_hp_stage_3_high, &s->f[7]);
  }
  h->dec.drc_gain_buffers.pstr_gain_buf[0].pstr_buf_interp =
      (ia_drc_interp_buf_struct *)calloc(h->n_seq ? h->n_seq : 1, sizeof(ia_drc_interp_buf_struct));
  h->dec.drc_gain_buffers.pstr_gain_buf[0].buf_interp_cnt = h->n_seq;
  return h;
}
void ref_drctd_destroy(void *q)
{
  ref_drctd_handle *h = (ref_drctd_handle *)q;
  free(h->dec.drc_gain_buffers.pstr_gain_buf[0].pstr_buf_interp);
  free(h->bands);
  free(h);
}
/* x [n_ch][1024] in place; gains [n_seq][1024] (sequence order: groups, bands within a group) */
int ref_drctd_process(void *q, float *x, const float *gains)
{
  ref_drctd_handle *h = (ref_drctd_handle *)q;
  FLOAT32 *rows[MAX_CHANNEL_COUNT];
  int err;
  for (int c = 0; c < h->n_ch; ++c) rows[c] = x + (size_t)c * 1024;
  for (int k = 0; k < h->n_seq; ++k)
    memcpy(h->dec.drc_gain_buffers.pstr_gain_buf[0].pstr_buf_interp[k].lpcm_gains + MAX_SIGNAL_DELAY,
           gains + (size_t)k * 1024, 4096);
  err = (int)impd_drc_filter_banks_process(rows, &h->ins, &h->dec, 0);
  if (err) return err;
  impd_drc_apply_gains_and_add(&h->dec, rows, &h->ins, 1, 0);
  return 0;
}
