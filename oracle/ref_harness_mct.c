/* ref_harness_mct.c -- TEST INFRASTRUCTURE: function-level entry points into the UNMODIFIED reference build for the
 * multichannel coding tool (decoder/impeghd_multichannel.c).  Compiled into oracle/_ref/libmpeghref.so only. */
#include <stdlib.h>
#include <string.h>

#include <impeghd_type_def.h>
#include "impd_drc_common.h"
#include "impd_drc_extr_delta_coded_info.h"
#include "impd_drc_struct.h"
#include "impeghd_intrinsics_flt.h"
#include "ia_core_coder_bitbuffer.h"
#include "ia_core_coder_cnst.h"
#include "ia_core_coder_constants.h"
#include "ia_core_coder_defines.h"
#include "ia_core_coder_acelp_info.h"
#include "ia_core_coder_interface.h"
#include "ia_core_coder_info.h"
#include "ia_core_coder_tns_usac.h"
#include "ia_core_coder_rom.h"
#include "impeghd_tbe_dec.h"
#include "ia_core_coder_igf_data_struct.h"
#include "ia_core_coder_stereo_lpd.h"
#include "impeghd_error_codes.h"
#include "impeghd_hoa_common_values.h"
#include "impeghd_hoa_matrix_struct.h"
#include "impeghd_hoa_matrix.h"
#include "ia_core_coder_config.h"
#include "impeghd_multichannel.h"
#include "impeghd_multichannel_rom.h"

void ref_mct_rom(float *sin_alpha /*[65]*/, float *cos_alpha /*[65]*/)
{
  memcpy(sin_alpha, impeghd_mc_index_to_sin_alpha, 65 * sizeof(float));
  memcpy(cos_alpha, impeghd_mc_index_to_cos_alpha, 65 * sizeof(float));
}

/* impeghd_mc_process (impeghd_multichannel.c:899) for pair 0 of a tool instance filled from the arguments */
void ref_mct_process(float *dmx, float *res, const int *coef_idx, const int *mask, int bands_per_window,
                     int total_sfb, const short *sfb_offset, int num_samples, int signal_type, int pred_dir,
                     int fullband)
{
  ia_multichannel_data *mc = (ia_multichannel_data *)calloc(1, sizeof(*mc));
  WORD32 c[MAX_NUM_MC_BANDS] = {0}, m[MAX_NUM_MC_BANDS] = {0};
  for (int i = 0; i < bands_per_window && i < MAX_NUM_MC_BANDS; ++i) c[i] = coef_idx[i], m[i] = mask[i];
  if (bands_per_window == 0) c[0] = coef_idx[0];
  mc->use_tool = 1;
  mc->signal_type = signal_type;
  mc->pred_dir[0] = pred_dir;
  mc->bandwise_coeff_flag[0] = fullband ? 0 : 1;
  mc->mask_flag[0] = 0;
  impeghd_mc_process(mc, dmx, res, c, m, bands_per_window, total_sfb, 0, sfb_offset, num_samples);
  free(mc);
}
