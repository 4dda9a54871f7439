/* oracle/oracle_tail.c -- TEST INFRASTRUCTURE ONLY (see oracle.h).
 *
 * Plain-C restatement of the end of the reference's chain: loudness gain, peak limiter, PCM packer.
 *   impegh_dec_ln_pl_process        decoder/ia_core_coder_decode_main.c:1496
 *   impegh_dec_peak_limiter         decoder/ia_core_coder_decode_main.c:1429   (transposes only)
 *   impeghd_peak_limiter_init       decoder/impeghd_peak_limiter.c:146
 *   impeghd_peak_limiter_process    decoder/impeghd_peak_limiter.c:190
 *   ia_core_coder_samples_sat       decoder/ia_core_coder_decode_main.c:209
 *
 * State is kept in TIME ORDER (oldest first) instead of the reference's circular buffers, which is the
 * layout the CUDA kernels use:
 *   max_hist[A]     the per-sample channel maxima of the last A input samples
 *   delay[ch][A]    the last A input samples of each channel
 * with A = attack_time_samples.  The reference's max_idx bookkeeping (:213-236) keeps the invariant
 * max_buf[max_idx] == max(max_buf[0..A]) (rescan when the slot holding the maximum is overwritten, replace
 * when the new value is >= the maximum, else keep), so `maximum` is simply the largest of the last A+1
 * channel maxima; that is what is computed here.  tests/test_oracle_tail.py pins this against the reference.
 */
#include <math.h>
#include <stdint.h>
#include <string.h>

#include "oracle.h"

void oracle_limiter_init(oracle_limiter *st, int n_ch, int sample_rate, int limiter_on)
{
  /* impeghd_peak_limiter.c:152-177 */
  int attack = (int)(5.0f * sample_rate / 1000);
  memset(st, 0, sizeof(*st));
  st->n_ch = n_ch;
  st->attack_samples = attack;
  st->limiter_on = limiter_on;
  st->threshold = 0.89125094f * 32768.0f;
  st->gain_modified = 1.0f;
  st->pre_smoothed_gain = 1.0f;
  st->attack_const = (float)pow(0.1, 1.0 / (attack + 1));
  st->release_const = (float)pow(0.1, 1.0 / (50.0f * sample_rate / 1000 + 1));
}

/* planar [n_ch][n] in place */
void oracle_limiter_process(oracle_limiter *st, float *planar, int n)
{
  const int A = st->attack_samples, nch = st->n_ch;
  float gm = st->gain_modified, psg = st->pre_smoothed_gain;
  const float thr = st->threshold, ac = st->attack_const, rc = st->release_const;
  float m[ORACLE_LIM_MAX_ATTACK + 4096];
  float g[4096];
  int i, j, c;

  if (!st->limiter_on && psg >= 1.0f)
  {
    /* pure delay (impeghd_cpy_delay_samples, :66) */
    for (i = 0; i < n; i++)
      g[i] = -1.0f; /* marker: copy through */
  }
  else
  {
    memcpy(m, st->max_hist, sizeof(float) * A);
    for (i = 0; i < n; i++)
    {
      float tmp = 0.0f;
      for (c = 0; c < nch; c++)
        tmp = (float)fmax(tmp, (float)fabs(planar[c * n + i])); /* :209-212 */
      m[A + i] = tmp;
    }
    for (i = 0; i < n; i++)
    {
      float maximum = m[i], gain;
      for (j = 1; j <= A; j++)
        if (maximum < m[i + j])
          maximum = m[i + j];
      gain = (thr >= maximum) ? 1.0f : thr / maximum; /* :240-247 */
      if (gain < psg)
        gm = (float)fmin(gm, (gain - 0.1f * psg) * 1.11111111f); /* :248-253 */
      else
        gm = gain;
      if (gm >= psg)
        psg = gm + rc * (psg - gm); /* :258-262 */
      else
      {
        psg = gm + ac * (psg - gm);
        psg = (float)fmax(psg, gain);
      }
      g[i] = psg;
    }
    memcpy(st->max_hist, m + n, sizeof(float) * A);
  }
  for (c = 0; c < nch; c++)
  {
    float x[ORACLE_LIM_MAX_ATTACK + 4096];
    memcpy(x, st->delay[c], sizeof(float) * A);
    memcpy(x + A, planar + c * n, sizeof(float) * n);
    for (i = 0; i < n; i++)
    {
      float t = x[i];
      if (g[i] >= 0.0f)
      {
        t = t * g[i]; /* impeghd_cpy_delay_samples_with_gain, :113-121 */
        if (thr < t)
          t = thr;
        else if (t < -thr)
          t = -thr;
      }
      planar[c * n + i] = t;
    }
    memcpy(st->delay[c], x + n, sizeof(float) * A);
  }
  st->gain_modified = gm;
  st->pre_smoothed_gain = psg;
}

/* loudness normalisation gain (decode_main.c:1533-1544): gain = (float)pow(10, db / 20), x *= gain */
float oracle_loudness_gain(float gain_db) { return (float)pow(10.0, gain_db / 20.0); }

/* ia_core_coder_samples_sat.  planar [n_ch][n]; out_ch_map[n_ch] (-1 = identity); returns bytes written.
 * 32-bit quirk, pinned against the reference build: MAX_FLT_VAL_32 (2147483647.0f) is 2^31 in binary32.
 * In the clamp-high branch gcc folds (WORD32)MAX_FLT_VAL_32 to INT32_MAX at compile time, so values above
 * 2^31 come out as 0x7FFFFFFF; a product that is exactly 2^31 (x == 32768.0f) is not "greater than" the
 * bound, reaches the run-time cvttss2si out of range and comes out as 0x80000000. */
static int32_t cast_i32(float v) { return (v >= 2147483648.0f || v < -2147483648.0f || v != v) ? INT32_MIN : (int32_t)v; }

int oracle_samples_sat(const float *planar, int n, int n_ch, int pcmsize, const int *out_ch_map,
                       int *delay_samples, unsigned char *out)
{
  int pos[64];
  int ch, s, num, delay = *delay_samples;
  unsigned char *o = out;
  if (delay <= n)
  {
    num = n_ch * (n - delay);
    *delay_samples = 0;
  }
  else
  {
    num = 0;
    *delay_samples = delay - n;
  }
  for (ch = 0; ch < n_ch; ch++)
    pos[ch] = ch;
  for (ch = 0; ch < n_ch; ch++)
  {
    int m = out_ch_map[ch] == -1 ? ch : out_ch_map[ch];
    pos[m] = ch;
  }
  for (s = delay; s < n; s++)
    for (ch = 0; ch < n_ch; ch++)
    {
      float v = planar[(size_t)pos[ch] * n + s];
      int32_t w;
      if (pcmsize == 24)
      {
        v = v * 256;
        if (v < -8388608.0f)
          v = -8388608.0f;
        else if (8388607.0f < v)
          v = 8388607.0f;
        w = cast_i32(v);
        *o++ = w & 0xff;
        *o++ = (w >> 8) & 0xff;
        *o++ = (w >> 16) & 0xff;
      }
      else if (pcmsize == 16)
      {
        if (v < -32768.0f)
          v = -32768.0f;
        else if (32767.0f < v)
          v = 32767.0f;
        w = (int16_t)v;
        *o++ = w & 0xff;
        *o++ = (w >> 8) & 0xff;
      }
      else
      {
        v = v * 65536;
        if (v < -2147483647.0f)
          v = -2147483647.0f;
        if (2147483647.0f < v)
          w = INT32_MAX;
        else
          w = cast_i32(v);
        *o++ = w & 0xff;
        *o++ = (w >> 8) & 0xff;
        *o++ = (w >> 16) & 0xff;
        *o++ = (w >> 24) & 0xff;
      }
    }
  return num * (pcmsize >> 3);
}
