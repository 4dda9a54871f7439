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
