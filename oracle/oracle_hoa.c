/* oracle/oracle_hoa.c -- TEST INFRASTRUCTURE ONLY (see oracle.h).
 *
 * Plain-C restatement of the reference HOA renderer's per-frame work:
 *   impeghd_hoa_ren_block_process      decoder/impeghd_hoa_renderer.c:181   (NFC, render-matrix apply, clip)
 *   impeghd_hoa_ren_nfc_filter_apply   decoder/impeghd_hoa_nfc_filtering.c:201
 *   impeghd_hoa_ren_iir_ord_2_filter   :64,  impeghd_hoa_ren_iir_ord_1_filter :99
 * The render matrix (impeghd_hoa_ren_renderer_init, :399) and the NFC coefficients
 * (impeghd_hoa_ren_nfc_set_filter_params, nfc_filtering.c:127) are designed by the reference's host init and
 * passed in.  nfc[mn][17] = {global_gain, n_biquads, n_first_order, (a1,a2,b1,b2) x 3, (a1,b1)};
 * nfc_state[mn][14] = {(x1,x2,y1,y2) x 3, (x1,y1)}.
 */
#include <string.h>

#include "oracle.h"

void oracle_hoa_render(const float *M, int n_out, int n_coef, const float *nfc, float *nfc_state,
                       const float *in, int n, int clip, float *out)
{
  static __thread float buf[64 * 4096];
  int mn, och, t, k;
  memcpy(buf, in, sizeof(float) * (size_t)n_coef * n);
  if (nfc)
  {
    for (mn = 0; mn < n_coef; mn++)
    {
      const float *c = nfc + mn * 17;
      float *st = nfc_state + mn * 14;
      float *y = buf + (size_t)mn * n;
      int n2 = (int)c[1], n1 = (int)c[2];
      if (c[0] != 1.0f)
        for (t = n - 1; t >= 0; t--)
          y[t] = y[t] * c[0];
      for (k = 0; k < n2; k++)
      {
        const float a1 = c[3 + 4 * k], a2 = c[4 + 4 * k], b1 = c[5 + 4 * k], b2 = c[6 + 4 * k];
        float x1 = st[4 * k], x2 = st[4 * k + 1], y1 = st[4 * k + 2], y2 = st[4 * k + 3];
        for (t = 0; t < n; t++)
        {
          float x0 = y[t];
          float v = (((x2 * b2) - (y1 * a1)) - (y2 * a2)) + ((x1 * b1) + (x0 * 1.0f)); /* b0 = 1 */
          y[t] = v;
          y2 = y1, y1 = v, x2 = x1, x1 = x0;
        }
        st[4 * k] = x1, st[4 * k + 1] = x2, st[4 * k + 2] = y1, st[4 * k + 3] = y2;
      }
      if (n1)
      {
        const float a1 = c[15], b1 = c[16];
        float x1 = st[12], y1 = st[13];
        for (t = 0; t < n; t++)
        {
          float x0 = y[t];
          float v = ((x1 * b1) - (y1 * a1)) + (x0 * 1.0f);
          y[t] = v;
          y1 = v, x1 = x0;
        }
        st[12] = x1, st[13] = y1;
      }
    }
  }
  for (och = 0; och < n_out; och++)
    for (t = 0; t < n; t++)
    {
      float val = 0;
      for (mn = 0; mn < n_coef; mn++)
        val = val + M[och * n_coef + mn] * buf[(size_t)mn * n + t];
      if (clip)
      {
        val = (val < 32767.0f) ? val : 32767.0f;
        val = (val > -32768.0f) ? val : -32768.0f;
      }
      out[(size_t)och * n + t] = val;
    }
}
