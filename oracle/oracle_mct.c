/* oracle_mct.c -- TEST INFRASTRUCTURE (CPU restatement, never shipped, never on the product path).
 *
 * Multichannel coding tool, the signal half: inverse rotation / prediction "boxes" applied to a pair of channel
 * spectra.  Restates impeghd_mc_process (decoder/impeghd_multichannel.c:899) with its two static helpers
 * impeghd_apply_mc_rotation (:781) and impeghd_apply_mc_prediction (:817), and the angle tables
 * impeghd_mc_index_to_sin_alpha / _cos_alpha (decoder/impeghd_multichannel_rom.c:129,141).  Bitstream parsing of the
 * boxes (impeghd_mc_parse, :386) and stereo filling (:104, :698) are host work and not restated.
 * Pinned against the reference build through oracle/ref_harness_mct.c (tests/test_oracle_mct.py) and
 * tests/golden/mct_golden.npz. */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>

#include "oracle.h"

/* The ROM holds sin / cos of (idx - 32) * pi / 64 printed with six decimals: rebuilt the same way. */
void oracle_mct_rom(float *sin_alpha /*[65]*/, float *cos_alpha /*[65]*/)
{
  for (int i = 0; i < 65; ++i)
  {
    char buf[32];
    const double a = (i - 32) * 3.14159265358979323846 / 64.0;
    double s = sin(a), c = cos(a);
    if (fabs(s) < 5e-7) s = 0.0; /* "-0.000000" / "0.000000": the ROM has +0 */
    if (fabs(c) < 5e-7) c = 0.0;
    snprintf(buf, sizeof buf, "%.6f", s);
    sin_alpha[i] = strtof(buf, NULL);
    snprintf(buf, sizeof buf, "%.6f", c);
    cos_alpha[i] = strtof(buf, NULL);
  }
}

static void rotation(float *dmx, float *res, int alpha_idx, int n, const float *sn, const float *cs)
{
  for (int i = 0; i < n; ++i)
  {
    const float d = dmx[i], r = res[i];
    const float l = d * cs[alpha_idx] - r * sn[alpha_idx]; /* :792 */
    const float rr = d * sn[alpha_idx] + r * cs[alpha_idx]; /* :794 */
    dmx[i] = l, res[i] = rr;
  }
}

static void prediction(float *dmx, float *res, int alpha_q, int n, int pred_dir)
{
  const float alpha = alpha_q * 0.1f; /* :824 */
  for (int i = 0; i < n; ++i)
  {
    const float d = dmx[i], r = res[i];
    const float l = (d + alpha * d) + r;
    const float rr = (pred_dir == 0) ? (d - alpha * d) - r : ((-d) + alpha * d) + r;
    dmx[i] = l, res[i] = rr;
  }
}

/* One window of one box: impeghd_mc_process with use_tool = 1.  fullband = !bandwise_coeff_flag && !mask_flag. */
void oracle_mct_process(float *dmx, float *res, const int *coef_idx, const int *mask, int bands_per_window,
                        int total_sfb, const short *sfb_offset, int num_samples, int signal_type, int pred_dir,
                        int fullband)
{
  float sn[65], cs[65];
  oracle_mct_rom(sn, cs);
  if (fullband)
  {
    if (signal_type == 0)
      prediction(dmx, res, coef_idx[0], num_samples, pred_dir);
    else
      rotation(dmx, res, coef_idx[0], num_samples, sn, cs);
    return;
  }
  int sfb = -1;
  for (int band = 0; band < bands_per_window; ++band)
  {
    if (mask[band] == 1)
    {
      const int start = (sfb < 0) ? 0 : sfb_offset[sfb];
      const int stop = (sfb + 2 < total_sfb) ? sfb_offset[sfb + 2] : sfb_offset[sfb + 1];
      if (signal_type == 0)
        prediction(dmx + start, res + start, coef_idx[band], stop - start, pred_dir);
      else
        rotation(dmx + start, res + start, coef_idx[band], stop - start, sn, cs);
    }
    sfb += 2;
    if (sfb >= total_sfb - 1) break;
  }
}
