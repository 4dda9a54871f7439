/* oracle/oracle_tns_rs.c -- TEST INFRASTRUCTURE ONLY (see oracle.h).
 *
 * TNS: impeghd_tns_decode_coeff_ar_filter (decoder/ia_core_coder_tns.c:83): coefficient decode (host-side
 * double sin, then the step-up recursion in float) and the all-pole filter over a bin range, forwards or
 * backwards.  ia_core_coder_tns_apply (:204) resolves band numbers to bin ranges; tests do that with the same
 * clamps (oracle_tns_range).
 * Resampler: impeghd_resample (decoder/impeghd_resampler.c:93): 60-tap polyphase x2 / x3 up, optional 60-tap
 * decimation by 2; state = the last input samples, the last up-sampled samples and the polyphase memory.
 */
#include <math.h>
#include <string.h>

#include "oracle.h"

#define PI_F 3.14159265358979323846264338327950288

/* coef[order] (transmitted indices), coef_res 3 or 4 -> lpc[order] = the reference's lpc[1..order] */
void oracle_tns_lpc(int order, int coef_res, const int *coef, float *lpc_out)
{
  float tmp[16], b[16], lpc[16];
  const float pi2 = (float)PI_F / 2;
  const float iqfac = (float)(((float)(1 << (coef_res - 1)) - 0.5) / (pi2));
  const float iqfac_m = (float)(((float)(1 << (coef_res - 1)) + 0.5) / (pi2));
  int i, m;
  for (i = 0; i < order; i++)
    tmp[i + 1] = (float)sin((short)coef[i] / (((short)coef[i] >= 0) ? iqfac : iqfac_m));
  lpc[0] = 1;
  for (m = 1; m <= order; m++)
  {
    b[0] = lpc[0];
    for (i = 1; i < m; i++)
      b[i] = lpc[i] + tmp[m] * lpc[m - i];
    b[m] = tmp[m];
    for (i = 0; i <= m; i++)
      lpc[i] = b[i];
  }
  for (i = 0; i < order; i++)
    lpc_out[i] = lpc[i + 1];
}

void oracle_tns_filter(float *spec, int size, int inc, int order, const float *lpc)
{
  float state[16];
  int i, j;
  memset(state, 0, sizeof(state));
  if (inc == -1)
    spec += size - 1;
  for (i = 0; i < size; i++)
  {
    float y = *spec;
    for (j = 0; j < order; j++)
      y = y - lpc[j] * state[j];
    for (j = order - 1; j > 0; j--)
      state[j] = state[j - 1];
    state[0] = y;
    *spec = y;
    spec += inc;
  }
}

/* band numbers -> [start, stop) bins of one window, ia_core_coder_tns_apply :262-284 with IGF inactive */
int oracle_tns_range(int start_band, int stop_band, int tns_max_band, int nbands, const short *sfb_tbl,
                     int sfb_per_sbk, int *start_bin, int *stop_bin)
{
  int start = start_band, stop = stop_band;
  start = start < tns_max_band ? start : tns_max_band;
  start = start < nbands ? start : nbands;
  if (start > sfb_per_sbk)
    return -1;
  stop = stop < tns_max_band ? stop : tns_max_band;
  stop = stop < nbands ? stop : nbands;
  if (stop > sfb_per_sbk)
    return -1;
  *start_bin = start > 0 ? sfb_tbl[start - 1] : 0;
  *stop_bin = stop > 0 ? sfb_tbl[stop - 1] : 0;
  return 0;
}

void oracle_rs_state_init(oracle_rs_state *st) { memset(st, 0, sizeof(*st)); }

/* in [1024] -> out [1024 * up / down]; f_up = the 60-tap table for this up factor, f_down = the x2 table */
void oracle_rs_process(oracle_rs_state *st, int up, int down, const float *f_up, const float *f_down, const float *in,
                       float *out)
{
  static __thread float y[3072 + 64];
  const int hist = 59 / up, taps = 60 / up, n_in = 1024;
  float *o = (down != 1) ? y : out;
  int n, k, p;
  for (n = 0; n < n_in; n++)
  {
    float poly[3] = {0, 0, 0};
    for (k = 0; k < taps; k++)
    {
      const int idx = n + k - hist;
      const float x = idx < 0 ? st->in_hist[hist + idx] : in[idx];
      for (p = 0; p < up; p++)
        poly[p] += x * f_up[k * up + p];
    }
    for (p = 0; p < up; p++)
    {
      float v = 0;
      int q;
      for (q = p; q < up - 1; q++)
        v = v + st->poly_mem[q];
      for (q = 0; q <= p; q++)
        v = v + poly[q];
      o[n * up + p] = v;
    }
    for (p = 1; p < up; p++)
      st->poly_mem[p - 1] = poly[p];
  }
  for (k = 0; k < hist; k++)
    st->in_hist[k] = in[n_in - hist + k];
  if (down == 2)
  {
    const int n_out = n_in * up / 2;
    for (n = 0; n < n_out; n++)
    {
      float mac = 0;
      for (k = 0; k < 60; k++)
      {
        const int idx = 2 * n + k - 59;
        mac += (idx < 0 ? st->up_hist[59 + idx] : y[idx]) * f_down[k];
      }
      out[n] = mac;
    }
    for (k = 0; k < 59; k++)
      st->up_hist[k] = y[n_in * up - 59 + k];
  }
}
