/* oracle_drc.c -- TEST INFRASTRUCTURE (CPU restatement, never shipped, never on the product path).
 *
 * Sub-band application of DRC gains: restates the arithmetic of impd_drc_apply_gains_subband
 * (decoder/impd_drc_process.c:243-385) on flat arrays.  Gain decoding, DRC set selection and the interpolation into
 * lpcm_gains are host work; what enters here is, per gain sequence and time slot, the value the reference picks from
 * lpcm_gains at the slot centre (:328, :345).  Pinned against the reference build through oracle/ref_harness_drc.c
 * (tests/test_oracle_drc.py) and tests/golden/drc_golden.npz. */
#include "oracle.h"

/* re / im: [n_ch][n_slots * n_sb] in place; group_of_ch [n_ch] (< 0: channel without DRC); band_count [n_groups];
 * gains [n_seq][n_slots], n_seq = sum of band_count; weights [n_groups][4][256] (overlap_weight of each band);
 * stft != 0: STFT256 packing, im[0] of a slot is the Nyquist bin. */
void oracle_drc_gains_subband(float *re, float *im, int n_ch, const int *group_of_ch, const int *band_count,
                              int n_groups, const float *gains, const float *weights, int n_slots, int n_sb, int stft)
{
  int first[64];
  first[0] = 0;
  for (int g = 0; g + 1 < n_groups && g + 1 < 64; ++g) first[g + 1] = first[g] + band_count[g]; /* :296-304 */
  for (int ch = 0; ch < n_ch; ++ch)
  {
    const int g = group_of_ch[ch];
    float *r = re + (long)ch * n_slots * n_sb, *q = im + (long)ch * n_slots * n_sb;
    if (g < 0) continue;
    for (int s = n_slots - 1; s >= 0; --s)
    {
      if (band_count[g] <= 1)
      { /* single band (:322-335) */
        const float gain = gains[first[g] * n_slots + s];
        for (int k = 0; k < n_sb; ++k)
        {
          r[s * n_sb + k] *= gain;
          q[s * n_sb + k] *= gain;
        }
      }
      else
      { /* :338-371 */
        for (int k = n_sb - 1; k >= 0; --k)
        {
          float gain = 0;
          for (int b = 0; b < band_count[g]; ++b)
            gain = gain + weights[(g * 4 + b) * 256 + k] * gains[(first[g] + b) * n_slots + s];
          r[s * n_sb + k] *= gain;
          if (!stft)
            q[s * n_sb + k] *= gain;
          else
          {
            if (k != 0) q[s * n_sb + k] *= gain;
            if (k == n_sb - 1) q[s * n_sb] *= gain;
          }
        }
      }
    }
  }
}
