/* oracle/oracle_ltpf.c -- TEST INFRASTRUCTURE ONLY (see oracle.h).
 *
 * Long-term post-filter of an FD channel, ia_core_coder_usac_ltpf_process (decoder/ia_core_coder_ltpf.c:395):
 * the frame is filtered in three stretches -- [0, 512) with the previous frame's parameters, a 128-sample
 * transition (same parameters: plain; fade-in; fade-out; or, when both frames filter with different parameters,
 * the new filter minus the zero-input response of the LPC-shaped difference, :247-335), and [640, 1024) with the
 * current parameters.  Filter (:134): out[n] = in[n] + g * sum_k out[n - pitch + k - 4] c1[fr][k]
 *                                             - 0.95 * (g * sum_k in[n - k] c2[gain_idx][k]).
 * State: the last 6 inputs, the last pitch_max + 4 outputs, the previous parameters.
 */
#include <string.h>

#include "oracle.h"

#define FRAME 1024
#define DELAY 512
#define TRANS 128
#define LPC_ORDER 24

static const float k_c1[2][8] = {
    {0.0000000f, 0.0304386f, 0.1162701f, 0.2195613f, 0.2674597f, 0.2195613f, 0.1162701f, 0.0304386f},
    {0.0076226f, 0.0676508f, 0.1700032f, 0.2547232f, 0.2547232f, 0.1700032f, 0.0676508f, 0.0076226f}};
static const float k_c2[4][7] = {
    {0.27150189f, 0.44286013f, 0.23027992f, 0.05759155f, -0.00172290f, -0.00045168f, -0.00005891f},
    {0.27581838f, 0.44682277f, 0.22783915f, 0.05410054f, -0.00353758f, -0.00092331f, -0.00011995f},
    {0.28044685f, 0.45103979f, 0.22519192f, 0.05037740f, -0.00545541f, -0.00141719f, -0.00018336f},
    {0.28543320f, 0.45554676f, 0.22230634f, 0.04638935f, -0.00749011f, -0.00193612f, -0.00024943f}};

void oracle_ltpf_cfg_init(oracle_ltpf_cfg *c, int sample_rate)
{
  c->pitch_min = (int)((34.f * (((float)sample_rate / 2.f) / 12800.f)) + 0.5f) * 2;
  c->pitch_fr1 = 320;
  c->pitch_fr2 = 324 - c->pitch_min;
  c->pitch_max = 54 + 6 * c->pitch_min;
  c->pitch_res = 2;
}

void oracle_ltpf_state_init(oracle_ltpf_state *st) { memset(st, 0, sizeof(*st)); }

static void decode_params(const oracle_ltpf_cfg *c, int present, int lag_idx, int gain_idx, int *pitch, int *fr,
                          float *gain)
{
  *pitch = 0, *fr = 0, *gain = 0.0f;
  if (!present)
    return;
  {
    const int a = (c->pitch_fr2 - c->pitch_min) * c->pitch_res, b = c->pitch_fr1 - c->pitch_fr2;
    if (lag_idx < a)
    {
      *pitch = c->pitch_min + lag_idx / c->pitch_res;
      *fr = lag_idx - (*pitch - c->pitch_min) * c->pitch_res;
    }
    else if (lag_idx < a + b * (c->pitch_res >> 1))
      *pitch = c->pitch_fr2 + lag_idx - a;
    else
      *pitch = (lag_idx - a - b) * c->pitch_res + c->pitch_fr1;
    *gain = (float)(gain_idx + 1) * 0.0625f;
  }
}

/* in / out point at the first sample of the stretch; history lies at negative indices */
static void filter(const float *in, float *out, int len, int pitch, int fr, float gain, int gain_idx, int mode,
                   const float *zir)
{
  float alpha = 0.0f, step = 0.0f;
  int n, k;
  if (gain == 0)
  {
    memcpy(out, in, len * sizeof(float));
    return;
  }
  if (mode == 2)
    alpha = 0.f, step = 1.f / (float)len;
  else if (mode == 3)
    alpha = 1.f, step = -1.f / (float)len;
  for (n = 0; n < len; n++)
  {
    float s1 = 0, s2 = 0;
    for (k = 0; k < 8; k++)
      s1 = s1 + out[n - pitch + k - 4] * k_c1[fr][k];
    for (k = 0; k < 7; k++)
      s2 = s2 + in[n - k] * k_c2[gain_idx][k];
    if (mode == 0)
      out[n] = (in[n] + gain * s1) - 0.95f * (gain * s2);
    else if (mode == 1)
      out[n] = ((in[n] + gain * s1) - 0.95f * (gain * s2)) - zir[n];
    else
      out[n] = in[n] + alpha * (gain * s1 - 0.95f * (gain * s2));
    if (mode > 1)
      alpha = alpha + step;
  }
}

/* LPC of the 256 outputs before `at` (autocorrelation + Levinson-Durbin, :210-245) */
static void lpc_of(const float *at, int len, float *a)
{
  float r[LPC_ORDER + 1], rc[LPC_ORDER], sigma2;
  int i, j;
  for (i = 0; i <= LPC_ORDER; i++)
  {
    float s = 0.0f;
    for (j = 0; j < len - i; j++)
      s = s + at[j - len] * at[j - len + i];
    r[i] = s;
  }
  r[0] = r[0] > 100.0f ? r[0] : 100.0f;
  r[0] = r[0] * 1.0001f;
  memset(a, 0, (LPC_ORDER + 1) * sizeof(float));
  a[0] = 1.0f;
  rc[0] = -r[1] / r[0];
  a[1] = rc[0];
  sigma2 = r[0] + r[1] * rc[0];
  for (i = 2; i <= LPC_ORDER; i++)
  {
    float sum = 0.0f;
    for (j = 0; j < i; j++)
      sum = sum + r[i - j] * a[j];
    rc[i - 1] = -sum / sigma2;
    sigma2 = sigma2 * (1.0f - rc[i - 1] * rc[i - 1]);
    if (sigma2 <= 1.0E-09f)
    {
      for (j = i; j <= LPC_ORDER; j++)
        a[j] = 0.0f;
      return;
    }
    for (j = 1; j <= i / 2; j++)
    {
      const float v = a[j] + rc[i - 1] * a[i - j];
      a[i - j] = a[i - j] + rc[i - 1] * a[j];
      a[j] = v;
    }
    a[i] = rc[i - 1];
  }
}

/* zero-input response over the transition (:247-335) */
static void zir_of(const float *in, const float *out, float *zir, const float *a, float gain, int gain_idx, int pitch,
                   int fr)
{
  float buf[LPC_ORDER + TRANS];
  float alpha, step;
  int n, k;
  memset(buf, 0, sizeof(buf));
  for (n = 0; n < LPC_ORDER; n++)
  {
    float s1 = 0, s2 = 0;
    for (k = 0; k < 8; k++)
      s1 = s1 + out[n - LPC_ORDER - pitch + k - 4] * k_c1[fr][k];
    for (k = 0; k < 7; k++)
      s2 = s2 + in[n - LPC_ORDER - k] * k_c2[gain_idx][k];
    buf[n] = (in[n - LPC_ORDER] - 0.95f * (gain * s2)) - (out[n - LPC_ORDER] - gain * s1);
  }
  for (n = 0; n < TRANS; n++)
    for (k = 1; k <= LPC_ORDER; k++)
      buf[LPC_ORDER + n] = buf[LPC_ORDER + n] - a[k] * buf[LPC_ORDER + n - k];
  memcpy(zir, buf + LPC_ORDER, (TRANS / 2) * sizeof(float));
  alpha = 1.f;
  step = 1.f / (float)(TRANS / 2);
  for (n = TRANS / 2; n < TRANS; n++)
  {
    zir[n] = buf[LPC_ORDER + n] * alpha;
    alpha = alpha - step;
  }
}

void oracle_ltpf_process(const oracle_ltpf_cfg *c, oracle_ltpf_state *st, int data_present, int pitch_lag_idx,
                         int gain_idx, float *x)
{
  float bin[6 + FRAME], bout[1600 + FRAME], zir[TRANS], a[LPC_ORDER + 1];
  const int hist = c->pitch_max + 4;
  float *in = bin + 6, *out = bout + hist;
  int pitch, fr, mode, t_pitch, t_fr, t_gidx;
  float gain, t_gain;
  const float *zp = 0;

  memcpy(bin, st->prev_in, 6 * sizeof(float));
  memcpy(in, x, FRAME * sizeof(float));
  memcpy(bout, st->prev_out, hist * sizeof(float));
  memset(out, 0, FRAME * sizeof(float));
  decode_params(c, data_present, pitch_lag_idx, gain_idx, &pitch, &fr, &gain);

  filter(in, out, DELAY, st->prev_pitch, st->prev_pitch_fr, st->prev_gain, st->prev_gain_idx, 0, 0);
  if (gain == st->prev_gain && pitch == st->prev_pitch && fr == st->prev_pitch_fr)
    t_gidx = gain_idx, t_pitch = pitch, t_fr = fr, t_gain = gain, mode = 0;
  else if (st->prev_gain == 0.f)
    t_gidx = gain_idx, t_pitch = pitch, t_fr = fr, t_gain = gain, mode = 2;
  else if (gain == 0.f)
    t_gidx = st->prev_gain_idx, t_pitch = st->prev_pitch, t_fr = st->prev_pitch_fr, t_gain = st->prev_gain, mode = 3;
  else
  {
    lpc_of(out + DELAY, FRAME / 4, a);
    zir_of(in + DELAY, out + DELAY, zir, a, gain, gain_idx, pitch, fr);
    t_gidx = gain_idx, t_pitch = pitch, t_fr = fr, t_gain = gain, zp = zir, mode = 1;
  }
  filter(in + DELAY, out + DELAY, TRANS, t_pitch, t_fr, t_gain, t_gidx, mode, zp);
  filter(in + DELAY + TRANS, out + DELAY + TRANS, FRAME - DELAY - TRANS, pitch, fr, gain, gain_idx, 0, 0);

  memcpy(x, out, FRAME * sizeof(float));
  st->prev_pitch = pitch, st->prev_pitch_fr = fr, st->prev_gain = gain, st->prev_gain_idx = gain_idx;
  memcpy(st->prev_in, in + FRAME - 6, 6 * sizeof(float));
  memcpy(st->prev_out, bout + FRAME, hist * sizeof(float));
}
