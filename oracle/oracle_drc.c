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

/* ---- domain switcher: the STFT pair around the sub-band DRC (impeghd_domain_switcher_process,
 * decoder/ia_core_coder_process.c:505-627; called at decoder/ia_core_coder_decode_main.c:2128 and :2140) ---------
 * win = ia_core_coder_sine_window256.  Layout per channel as the reference leaves it in time_sample_stft: re / im
 * planes of [4 slots][256]; im[slot][0] carries the Nyquist bin (:575-583). */
void oracle_ds_t2f(const float *win, float *in_hist /*[n_ch][256]*/, const float *x /*[n_ch][1024]*/, float *re_out,
                   float *im_out, int n_ch)
{
  float re[512], im[512];
  for (int slot = 0; slot < 4; ++slot)
    for (int ch = 0; ch < n_ch; ++ch)
    {
      float *h = in_hist + ch * 256;
      const float *cur = x + ch * 1024 + slot * 256;
      for (int i = 0; i < 256; ++i)
      {
        const float v = cur[i];
        re[i] = h[i] * win[i];
        re[256 + i] = win[255 - i] * v;
        h[i] = v;
        im[i] = im[256 + i] = 0.0f;
      }
      oracle_cplx_fft_pow2(re, im, 512); /* impeghd_rad2_cplx_fft (:570) */
      float *ro = re_out + ch * 1024 + slot * 256, *io = im_out + ch * 1024 + slot * 256;
      ro[0] = re[0];
      io[0] = re[256];
      for (int k = 1; k < 256; ++k) ro[k] = re[k], io[k] = im[k];
    }
}

void oracle_ds_f2t(const float *win, float *out_hist /*[n_ch][256]*/, const float *re_in, const float *im_in,
                   float *y /*[n_ch][1024]*/, int n_ch)
{
  float re[512], im[512];
  for (int slot = 0; slot < 4; ++slot)
    for (int ch = 0; ch < n_ch; ++ch)
    {
      const float *ri = re_in + ch * 1024 + slot * 256, *ii = im_in + ch * 1024 + slot * 256;
      float *h = out_hist + ch * 256, *o = y + ch * 1024 + slot * 256;
      re[0] = ri[0], re[256] = ii[0], im[0] = 0, im[256] = 0;
      for (int k = 1; k < 256; ++k)
      {
        re[k] = ri[k], re[512 - k] = re[k];
        im[k] = ii[k], im[512 - k] = -im[k];
      }
      oracle_cplx_ifft_pow2(re, im, 512); /* impeghd_rad2_cplx_ifft (:608) */
      for (int i = 0; i < 256; ++i)
      {
        o[i] = (re[i] / 512.0f) * win[i] + h[i] * win[255 - i]; /* :613-616 */
        h[i] = re[256 + i] / 512.0f;
      }
    }
}

/* ---- time-domain multi-band DRC: filter banks, per-band gain curves, band sum -------------------------------
 * impd_drc_filter_banks_process (decoder/impd_drc_process.c:399), the Linkwitz-Riley banks of
 * decoder/impd_drc_filter_bank.c (low/high pair :410, all-pass :367, 2/3/4-band trees :486,:511,:550, phase
 * alignment cascade :601) and the gain/sum stage of impd_drc_apply_gains_and_add (impd_drc_process.c:68-225) for one
 * channel.  Reference quirks kept: the all-pass sections run BACKWARDS over the frame (:380) with carried state;
 * the cascade runs sections num_filter .. 0 (:604), i.e. n_cascade = num_filter + 1 of them. */
static void lowhigh(const oracle_drc_biquad *lp, const oracle_drc_biquad *hp, float *sl, float *sh, float *lo,
                    float *hi, const float *in, int n)
{
  float s1l = sl[0], s2l = sl[1], s3l = sl[2], s4l = sl[3], s1h = sh[0], s2h = sh[1], s3h = sh[2], s4h = sh[3];
  for (int i = 0; i < n; ++i)
  {
    const float x = in[i];
    float t = s1l + lp->b0 * x;
    s1l = (lp->b1 * x - lp->a1 * t) + s2l;
    s2l = lp->b2 * x - lp->a2 * t;
    const float ol = s3l + lp->b0 * t;
    s3l = (lp->b1 * t - lp->a1 * ol) + s4l;
    s4l = lp->b2 * t - lp->a2 * ol;
    t = s1h + hp->b0 * x;
    s1h = (hp->b1 * x - hp->a1 * t) + s2h;
    s2h = hp->b2 * x - hp->a2 * t;
    const float oh = s3h + hp->b0 * t;
    s3h = (hp->b1 * t - hp->a1 * oh) + s4h;
    s4h = hp->b2 * t - hp->a2 * oh;
    lo[i] = ol, hi[i] = oh;
  }
  sl[0] = s1l, sl[1] = s2l, sl[2] = s3l, sl[3] = s4l, sh[0] = s1h, sh[1] = s2h, sh[2] = s3h, sh[3] = s4h;
}

static void allpass(const oracle_drc_biquad *f, float *st, float *out, const float *in, int n)
{
  float s1 = st[0], s2 = st[2]; /* x_p[2 ch], y_p[2 ch] */
  for (int i = n - 1; i >= 0; --i)
  {
    const float x = in[i];
    const float y = f->b0 * x + s1;
    s1 = (f->b1 * x - f->a1 * y) + s2;
    s2 = f->b2 * x - f->a2 * y;
    out[i] = y;
  }
  st[0] = s1, st[2] = s2;
}

/* x [1024] in place; state [17][4] (bank filters 0..7, cascade sections 8..16); gains [n_bands][1024] or NULL for a
 * channel outside every DRC channel group (then only the cascade touches it) */
void oracle_drc_td_channel(const oracle_drc_bank *bank, float *state, float *x, const float *gains)
{
  static __thread float b[4][1024];
  const int n = 1024;
  for (int i = bank->n_cascade - 1; i >= 0; --i) allpass(&bank->ap[i], state + 4 * (8 + i), x, x, n);
  if (!gains) return;
  const oracle_drc_biquad *f = bank->f;
  switch (bank->n_bands)
  {
  case 2:
    lowhigh(&f[0], &f[1], state + 0, state + 4, b[0], b[1], x, n);
    break;
  case 3:
    lowhigh(&f[0], &f[1], state + 0, state + 4, b[0], b[1], x, n);
    allpass(&f[2], state + 8, b[2], b[1], n);
    lowhigh(&f[3], &f[4], state + 12, state + 16, b[0], b[1], b[0], n);
    break;
  case 4:
    lowhigh(&f[0], &f[1], state + 0, state + 4, b[0], b[1], x, n);
    allpass(&f[2], state + 8, b[0], b[0], n);
    allpass(&f[3], state + 12, b[2], b[1], n);
    lowhigh(&f[4], &f[5], state + 16, state + 20, b[0], b[1], b[0], n);
    lowhigh(&f[6], &f[7], state + 24, state + 28, b[2], b[3], b[2], n);
    break;
  default:
    for (int i = 0; i < n; ++i) b[0][i] = x[i];
    break;
  }
  const int nb = bank->n_bands < 1 ? 1 : bank->n_bands;
  for (int k = 0; k < nb; ++k)
    for (int i = 0; i < n; ++i) b[k][i] = b[k][i] * gains[k * 1024 + i];
  for (int i = 0; i < n; ++i)
  {
    float sum = 0.0f;
    for (int k = nb - 1; k >= 0; --k) sum = sum + b[k][i];
    x[i] = sum;
  }
}
