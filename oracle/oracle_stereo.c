/* oracle/oracle_stereo.c -- TEST INFRASTRUCTURE ONLY (see oracle.h).
 *
 * Joint-stereo tools of a channel pair, on the dequantised spectra, in the reference's order of operations
 * (decoder/ia_core_coder_ext_ch_ele.c):
 *   previous-frame downmix   ia_core_coder_cplx_prev_mdct_dmx   :703   (parse phase there; same inputs here)
 *   M/S by band mask         impeghd_ms_stereo                  :1427
 *   complex prediction       ia_core_coder_cplx_pred_upmixing   :561   with the MDST estimate of the downmix
 *                            ia_core_coder_estimate_dmx_im :495 / ia_core_coder_filter_and_add :412
 *   spectrum save            ia_core_coder_usac_cplx_save_prev  :93
 * Every product and sum is a separately rounded binary32 operation (compiled with -ffp-contract=off).
 * Groups of short windows are expanded to windows by the caller (all windows of a group share their side info).
 */
#include <string.h>

#include "oracle.h"

static float g_mdst_curr[4][2][2][7];
static float g_mdst_prev[2][2][7];

void oracle_stereo_set_rom(const float *mdst_curr, const float *mdst_prev)
{
  memcpy(g_mdst_curr, mdst_curr, sizeof(g_mdst_curr));
  memcpy(g_mdst_prev, mdst_prev, sizeof(g_mdst_prev));
}

/* window class of the MDST filter pair (:507-531): 0 ONLY_LONG / EIGHT_SHORT, 1 LONG_START, 2 LONG_STOP,
 * 3 STOP_START (and anything else); previous-frame filter class: 0 for classes 0-1, 1 for classes 2-3 */
static int mdst_class(int window_sequence)
{
  switch (window_sequence)
  {
  case 0:
  case 2:
    return 0;
  case 1:
    return 1;
  case 3:
    return 2;
  default:
    return 3;
  }
}

/* tap j of the 7-tap filter around bin i reads in[reflect(i + j - 3)]: the edge forms written out at :422-438 and
 * :461-483 are mirror extensions, index -k -> k - 1 and len - 1 + k -> len - k */
static int reflect(int j, int len)
{
  if (j < 0)
    return -j - 1;
  if (j >= len)
    return 2 * len - 1 - j;
  return j;
}

static float mdst_tap_sum(const float *in, int len, const float *f, int i)
{
  const float a = f[6] * in[reflect(i - 3, len)] + f[5] * in[reflect(i - 2, len)];
  const float b = f[4] * in[reflect(i - 1, len)] + f[3] * in[reflect(i, len)];
  const float c = f[2] * in[reflect(i + 1, len)] + f[1] * in[reflect(i + 2, len)];
  const float d = f[0] * in[reflect(i + 3, len)];
  return (a + b) + (c + d);
}

/* out[i] += s on odd bins, out[i] += s * factor_even on even bins (:421-483) */
static void filter_and_add(const float *in, int len, const float *f, float *out, int factor_even)
{
  int i;
  for (i = 0; i < len; i++)
  {
    const float s = mdst_tap_sum(in, len, f, i);
    out[i] = (i & 1) ? out[i] + s : out[i] + s * (float)factor_even;
  }
}

void oracle_stereo_apply(const oracle_stereo_side *sd, const short *sfb_top, int n_sfb, const oracle_stereo_state *st,
                         float *left, float *right)
{
  const int n_win = sd->is_short ? 8 : 1;
  const int bins = sd->is_short ? 128 : 1024;
  const int stride = sd->is_short ? 16 : 0;
  int w, sfb, b;

  if (sd->tool == 1)
  { /* M/S where the mask says so, bands [first_band, last_band) */
    for (w = 0; w < n_win; w++)
      for (sfb = sd->first_band; sfb < sd->last_band; sfb++)
        if (sd->used[w * stride + sfb])
          for (b = sfb ? sfb_top[sfb - 1] : 0; b < sfb_top[sfb]; b++)
          {
            const float l = left[w * bins + b], r = right[w * bins + b];
            left[w * bins + b] = l + r;
            right[w * bins + b] = l - r;
          }
    return;
  }
  if (sd->tool != 3)
    return;

  {
    const float factor = sd->pred_dir ? -1.0f : 1.0f;
    float dmx_re[1024], dmx_im[1024], dmx_prev[1024];
    /* downmix: prediction bands carry it in the left spectrum, the others are L/R coded (:593-615) */
    for (w = 0; w < n_win; w++)
      for (sfb = 0; sfb < n_sfb; sfb++)
        for (b = sfb ? sfb_top[sfb - 1] : 0; b < sfb_top[sfb]; b++)
        {
          const int i = w * bins + b;
          dmx_re[i] = (sd->used[w * stride + sfb] == 1) ? left[i] : 0.5f * (left[i] + factor * right[i]);
        }
    memset(dmx_im, 0, sizeof(dmx_im));
    if (sd->complex_coef)
    {
      const int cls = mdst_class(sd->window_sequence);
      const float *fc = g_mdst_curr[cls][sd->window_shape_prev][sd->window_shape];
      const float *fp = g_mdst_prev[cls >> 1][sd->window_shape_prev];
      const float *prev = 0;
      if (sd->use_prev_frame)
      { /* ia_core_coder_cplx_prev_mdct_dmx: last window of the saved spectra, or zeros */
        const int off = 1024 - bins;
        for (b = 0; b < bins; b++)
          dmx_prev[b] = sd->prev_dmx ? 0.5f * (st->save_l[off + b] + factor * st->save_r[off + b]) : 0.0f;
        prev = dmx_prev;
      }
      for (w = 0; w < n_win; w++)
      {
        filter_and_add(dmx_re + w * bins, bins, fc, dmx_im + w * bins, 1);
        if (prev)
          filter_and_add(prev, bins, fp, dmx_im + w * bins, -1);
        prev = dmx_re + w * bins;
      }
    }
    for (w = 0; w < n_win; w++)
      for (sfb = 0; sfb < n_sfb; sfb++)
      {
        const float a_re = (float)sd->alpha_re[w * stride + sfb] * 0.1f;
        const float a_im = (float)sd->alpha_im[w * stride + sfb] * 0.1f;
        int on = sd->used[w * stride + sfb] != 0;
        if (!sd->complex_coef)
          on = on && sfb >= sd->first_band && sfb < sd->last_band;
        if (!on)
          continue;
        for (b = sfb ? sfb_top[sfb - 1] : 0; b < sfb_top[sfb]; b++)
        {
          const int i = w * bins + b;
          float mix = right[i] - a_re * left[i];
          if (sd->complex_coef)
            mix = mix - a_im * dmx_im[i];
          right[i] = factor * (left[i] - mix);
          left[i] = left[i] + mix;
        }
      }
  }
}

void oracle_stereo_save(const oracle_stereo_side *sd, oracle_stereo_state *st, const float *left, const float *right)
{
  const int bins = sd->is_short ? 128 : 1024;
  const int off = 1024 - bins;
  if (sd->save == 1)
  {
    memcpy(st->save_l + off, left + off, bins * sizeof(float));
    memcpy(st->save_r + off, right + off, bins * sizeof(float));
  }
  else if (sd->save == 2)
  {
    memset(st->save_l + off, 0, bins * sizeof(float));
    memset(st->save_r + off, 0, bins * sizeof(float));
  }
}
