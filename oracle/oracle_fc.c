/* oracle/oracle_fc.c -- TEST INFRASTRUCTURE ONLY (see oracle.h).
 *
 * Plain-C restatement of the reference format converter's per-frame work:
 *   impeghd_format_converter_process             decoder/ia_core_coder_process.c:641  (4 STFT slots per frame:
 *       sine-windowed 512-point transform of [previous 256 | current 256] samples, inverse transform, /512,
 *       windowed overlap-add with the tail kept from the previous slot)
 *   impeghd_format_conv_active_dmx_stft_process  decoder/ia_core_coder_process.c:127  (frequency-dependent downmix,
 *       target vs realised energy in 58 ERB bands, one-pole smoothing, EQ = sqrt(target / realised) clipped to
 *       [-10 dB, +8 dB])
 * The downmix tensor D[out][in][257] comes from the reference's init and is an input.
 * Reference details kept: an input contributes to an output only when D[out][in][0] != 0; the DC and Nyquist
 * bins are energy-counted in ERB bands 0 and 57; the forward direction uses the e^{+j} transform and keeps
 * Im as it comes (no conjugation), the inverse direction uses the e^{-j} transform on the Hermitian extension.
 */
#include <math.h>
#include <string.h>

#include "oracle.h"

#define FC_ALPHA (0.0435f)
#define FC_BETA (1 - FC_ALPHA)
#define FC_EPS 1e-35f
#define FC_EQ_MAX 2.51188636f
#define FC_EQ_MIN 0.316227764f

static const int erb_stop[58] = {1,  2,  3,  4,  5,  6,  7,   8,   9,   10,  11,  12,  13,  14,  15,
                                 16, 17, 18, 19, 20, 21, 22,  23,  24,  25,  26,  27,  28,  29,  30,
                                 31, 32, 33, 36, 39, 43, 47,  51,  55,  60,  65,  70,  77,  84,  91,
                                 98, 106, 115, 125, 136, 148, 160, 174, 189, 205, 222, 241, 256};

void oracle_fc_state_init(oracle_fc_state *st) { memset(st, 0, sizeof(*st)); }

/* D [n_out][n_in][257], win = ia_core_coder_sine_window256, in [n_in][1024], out [n_out][1024] */
void oracle_fc_process(const float *D, int n_in, int n_out, const float *win, oracle_fc_state *st, const float *in,
                       float *out)
{
  static __thread float X[ORACLE_FC_MAX_CH][512], Y[ORACLE_FC_MAX_CH][512];
  float re[512], im[512];
  int slot, ci, co, i, k, erb;
  for (slot = 0; slot < 4; slot++)
  {
    for (ci = 0; ci < n_in; ci++)
    {
      const float *cur = in + ci * 1024 + slot * 256;
      for (i = 0; i < 256; i++)
      {
        re[i] = st->in_prev[ci][i] * win[i];
        re[256 + i] = win[255 - i] * cur[i];
        st->in_prev[ci][i] = cur[i];
      }
      memset(im, 0, sizeof(im));
      oracle_cplx_ifft_pow2(re, im, 512);
      X[ci][0] = re[0];
      X[ci][1] = re[256];
      for (k = 1; k < 256; k++)
        X[ci][2 * k] = re[k], X[ci][2 * k + 1] = im[k];
    }
    for (co = 0; co < n_out; co++)
    {
      float T[58], R[58], *y = Y[co];
      memset(y, 0, sizeof(float) * 512);
      memset(T, 0, sizeof(T));
      memset(R, 0, sizeof(R));
      for (ci = 0; ci < n_in; ci++)
      {
        const float *d = D + ((size_t)co * n_in + ci) * 257, *x = X[ci];
        float tmp;
        if (!d[0])
          continue;
        tmp = d[0] * x[0];
        y[0] = y[0] + tmp;
        T[0] = T[0] + tmp * tmp;
        tmp = d[256] * x[1];
        y[1] = y[1] + tmp;
        T[57] = T[57] + tmp * tmp;
        k = 1;
        for (erb = 1; erb < 58; erb++)
        {
          float t = T[erb];
          for (; k < erb_stop[erb]; k++)
          {
            tmp = d[k] * x[2 * k];
            y[2 * k] = tmp + y[2 * k];
            t = t + tmp * tmp;
            tmp = d[k] * x[2 * k + 1];
            y[2 * k + 1] = tmp + y[2 * k + 1];
            t = t + tmp * tmp;
          }
          T[erb] = t;
        }
      }
      R[0] = y[0] * y[0];
      R[57] = y[1] * y[1];
      k = 1;
      for (erb = 0; erb < 58; erb++)
      {
        float r = R[erb];
        for (; k < erb_stop[erb]; k++)
        {
          r = r + y[2 * k] * y[2 * k];
          r = r + y[2 * k + 1] * y[2 * k + 1];
        }
        T[erb] = (FC_ALPHA * T[erb]) + (FC_BETA * st->t_prev[co][erb]);
        R[erb] = (FC_ALPHA * r) + (FC_BETA * st->r_prev[co][erb]);
        st->t_prev[co][erb] = T[erb];
        st->r_prev[co][erb] = R[erb];
      }
      {
        float eq = (float)sqrt(T[0] / (FC_EPS + R[0]));
        eq = eq > FC_EQ_MAX ? FC_EQ_MAX : (eq < FC_EQ_MIN ? FC_EQ_MIN : eq);
        y[0] *= eq;
        eq = (float)sqrt(T[57] / (FC_EPS + R[57]));
        eq = eq > FC_EQ_MAX ? FC_EQ_MAX : (eq < FC_EQ_MIN ? FC_EQ_MIN : eq);
        y[1] *= eq;
        k = 1;
        for (erb = 1; erb < 58; erb++)
        {
          eq = (float)sqrt(T[erb] / (FC_EPS + R[erb]));
          eq = eq > FC_EQ_MAX ? FC_EQ_MAX : (eq < FC_EQ_MIN ? FC_EQ_MIN : eq);
          for (; k < erb_stop[erb]; k++)
          {
            y[2 * k] = y[2 * k] * eq;
            y[2 * k + 1] = eq * y[2 * k + 1];
          }
        }
      }
      re[0] = y[0], im[0] = 0;
      re[256] = y[1], im[256] = 0;
      for (k = 1; k < 256; k++)
      {
        re[k] = y[2 * k];
        re[512 - k] = re[k];
        im[k] = y[2 * k + 1];
        im[512 - k] = -im[k];
      }
      oracle_cplx_fft_pow2(re, im, 512);
      for (i = 0; i < 256; i++)
      {
        out[co * 1024 + slot * 256 + i] = ((re[i] / ((float)512)) * win[i]) + st->out_prev[co][i] * win[255 - i];
        st->out_prev[co][i] = re[256 + i] / ((float)512);
      }
    }
  }
}
