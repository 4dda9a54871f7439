/* oracle/oracle_fd.c -- TEST INFRASTRUCTURE ONLY (see oracle.h).
 *
 * CPU restatement of the FD synthesis stage: power-of-two complex transform, IMDCT via pre/post twiddle,
 * windowing + overlap-add for the five window sequences, two window shapes and four transform kernel
 * types.  Written in "gather" form (each output sample is one expression of the IMDCT output, the old
 * overlap and the window tables) instead of the reference's pointer-walking scatter form; the rounding
 * sequence of every sample is the reference's.
 */
#include <math.h>
#include <string.h>

#include "oracle.h"

static oracle_rom g_rom;
void oracle_set_rom(const oracle_rom *rom) { g_rom = *rom; }
const oracle_rom *oracle_get_rom(void) { return &g_rom; }

/* ------------------------------------------------------------------------------------------------
 * Complex transform, n = 2^lg.  Follows impeghd_rad2_cplx_ifft (impeghd_ifft.c:297-843):
 *   pass 0   radix-4, no twiddles, input in base-4 digit-reversed order (DIG_REV, :326-379)
 *   pass p   radix-4 with twiddles, butterfly span del = 4^p (:384-777); twiddle index j = m * 256 / del
 *            into the 1024-point quarter-wave table; W^2 and W^3 use folded entries in four j regions
 *   last     one radix-2 pass when lg is odd (:779-837)
 * Butterflies use the "a += b; b = a - 2b" form (:353-369): one extra rounding versus a +- b.
 * This is the e^{+j} transform ("ifft"); the e^{-j} one (impeghd_fft.c:297) has its own rounding
 * sequence and is restated separately in oracle_fft.c.
 * ---------------------------------------------------------------------------------------------- */
static unsigned rev_base4_16(unsigned v)
{
  unsigned r = 0;
  int k;
  for (k = 0; k < 8; k++)
  {
    r = (r << 2) | (v & 3u);
    v >>= 2;
  }
  return r;
}

typedef struct
{
  float r, i;
} cpx;

/* out = x * (wc - j*ws) written as the reference writes it: (xr*wc + xi*ws, -(xr*ws) + xi*wc). */
static cpx rot_std(cpx x, float wc, float ws)
{
  cpx o;
  o.r = x.r * wc + x.i * ws;
  o.i = -(x.r * ws) + x.i * wc;
  return o;
}
/* folded form used when the table index passed 256: (xr*b - xi*a, xr*a + xi*b) */
static cpx rot_fold(cpx x, float a, float b)
{
  cpx o;
  o.r = x.r * b - x.i * a;
  o.i = x.r * a + x.i * b;
  return o;
}

static void bfly4_inv(cpx *p0, cpx *p1, cpx *p2, cpx *p3, int region4)
{
  cpx x0 = *p0, x1 = *p1, x2 = *p2, x3 = *p3;
  x0.r = x0.r + x2.r;
  x0.i = x0.i + x2.i;
  x2.r = x0.r - x2.r * 2;
  x2.i = x0.i - x2.i * 2;
  x1.r = x1.r + x3.r;
  x3.r = x1.r - x3.r * 2;
  if (!region4)
  {
    x1.i = x1.i + x3.i;
    x3.i = x1.i - x3.i * 2;
  }
  else
  { /* x3.i arrives negated from the region-4 twiddle form (:730-744) */
    x1.i = x1.i - x3.i;
    x3.i = x1.i + x3.i * 2;
  }
  x0.r = x0.r + x1.r;
  x0.i = x0.i + x1.i;
  x1.r = x0.r - x1.r * 2;
  x1.i = x0.i - x1.i * 2;
  x2.r = x2.r - x3.i;
  x2.i = x2.i + x3.r;
  {
    float d_r = x2.r + x3.i * 2;
    float d_i = x2.i - x3.r * 2;
    p0->r = x0.r, p0->i = x0.i;
    p1->r = x2.r, p1->i = x2.i;
    p2->r = x1.r, p2->i = x1.i;
    p3->r = d_r, p3->i = d_i;
  }
}

void oracle_cplx_ifft_pow2(float *re, float *im, int n)
{
  cpx y[1024];
  const float *w = g_rom.fft_twiddle;
  int lg = 0, odd, npass, b, p, m, g, del, ns, groups, i;
  while ((1 << lg) < n)
    lg++;
  odd = lg & 1;
  npass = lg >> 1;

  /* pass 0 */
  for (b = 0; b < n / 4; b++)
  {
    unsigned h = rev_base4_16((unsigned)(4 * b)) >> (15 - lg);
    int c, q = n / 4;
    cpx x[4];
    if (odd)
      h = (h + 1) & ~1u;
    c = (int)(h >> 1);
    for (i = 0; i < 4; i++)
      x[i].r = re[c + i * q], x[i].i = im[c + i * q];
    bfly4_inv(&x[0], &x[1], &x[2], &x[3], 0);
    for (i = 0; i < 4; i++)
      y[4 * b + i] = x[i];
  }

  /* twiddled radix-4 passes */
  del = 4;
  ns = 64;
  groups = n >> 4;
  for (p = 1; p < npass; p++)
  {
    const int full = ns * del;                                     /* always 256 */
    const int lim1 = full / 4 + full / 8 - full / 16 + full / 32 - full / 64 + full / 128 - full / 256;
    for (m = 0; m < del; m++)
    {
      const int j = m * ns;
      for (g = 0; g < groups; g++)
      {
        cpx *d = &y[g * 4 * del + m];
        cpx x0 = d[0], x1 = d[del], x2 = d[2 * del], x3 = d[3 * del];
        int region4 = 0;
        if (m > 0)
        {
          x1 = rot_std(x1, w[j], w[j + 257]);
          if (j <= full / 2)
            x2 = rot_std(x2, w[2 * j], w[2 * j + 257]);
          else
            x2 = rot_fold(x2, w[2 * j - 256], w[2 * j + 1]);
          if (j <= lim1)
            x3 = rot_std(x3, w[3 * j], w[3 * j + 257]);
          else if (j <= 2 * lim1)
            x3 = rot_fold(x3, w[3 * j - 256], w[3 * j + 1]);
          else
          {
            const float a = w[3 * j - 512], bq = w[3 * j - 512 + 257];
            cpx t;
            t.r = -(x3.r * a) - x3.i * bq;
            t.i = -(x3.r * bq) + x3.i * a;
            x3 = t;
            region4 = 1;
          }
        }
        bfly4_inv(&x0, &x1, &x2, &x3, region4);
        d[0] = x0, d[del] = x1, d[2 * del] = x2, d[3 * del] = x3;
      }
    }
    ns >>= 2;
    del <<= 2;
    groups >>= 2;
  }

  if (odd)
  {
    const int half = del / 2, step = ns << 1;
    for (i = 0; i < del; i++)
    {
      const int k = (i < half) ? i : i - half;
      const float wc = w[k * step], ws = w[k * step + 257];
      cpx x0 = y[i], x1 = y[i + del], t;
      if (i < half)
        t = rot_std(x1, wc, ws);
      else
      {
        t.r = x1.r * ws - x1.i * wc;
        t.i = x1.r * wc + x1.i * ws;
      }
      y[i + del].r = x0.r - t.r;
      y[i + del].i = x0.i - t.i;
      y[i].r = x0.r + t.r;
      y[i].i = x0.i + t.i;
    }
  }
  for (i = 0; i < n; i++)
  {
    re[i] = y[i].r;
    im[i] = y[i].i;
  }
}

/* ------------------------------------------------------------------------------------------------
 * IMDCT of one block: spec[0..N-1] -> aliased time signal x[0..N-1] in place, N = 1024 or 128.
 * ia_core_coder_acelp_imdct (ia_core_coder_imdct.c:366) + ia_core_coder_fft_based_imdct (:294) with
 * ia_core_coder_calc_pre_twid_dec (:87) / ia_core_coder_calc_post_twid_dec (:203).
 * ---------------------------------------------------------------------------------------------- */
static void imdct_block(float *x, int N, int tkt)
{
  float r[512], q[512];
  const int nl = N / 2;
  const float *c = (nl == 512) ? g_rom.twid_cos_512 : g_rom.twid_cos_64;
  const float *s = (nl == 512) ? g_rom.twid_sin_512 : g_rom.twid_sin_64;
  const float norm = 2.0f / (float)(2 * N); /* npoints = 2N (:371) */
  int i;

  for (i = 0; i < N; i++)
    x[i] = x[i] * norm;

  if (tkt == 1 || tkt == 2)
  {
    const float *C = g_rom.dctdst_cos_1024, *S = g_rom.dctdst_sin_1024;
    const float s4 = (float)sin(3.14159265358979323846264338327950288 * 0.25f);
    const float half = 0.5f;
    float tmp = 0.0f;
    if (tkt == 1)
      for (i = 0; i < nl; i++)
      {
        float t = x[i];
        x[i] = x[N - 1 - i];
        x[N - 1 - i] = t;
      }
    r[0] = x[0] + x[nl] * s4;
    q[0] = x[0] - x[nl] * s4;
    for (i = 1; i <= nl / 2; i++)
    {
      float re1 = (x[i] * C[i - 1] + x[N - i] * S[i - 1]) * half;
      float re2 = (x[nl - i] * S[nl + i - 1] + x[nl + i] * C[nl + i - 1]) * half;
      float im1 = (x[i] * S[i - 1] - x[N - i] * C[i - 1]) * half;
      float im2 = (x[nl - i] * C[nl + i - 1] - x[nl + i] * S[nl + i - 1]) * half;
      tmp = im1 - im2;
      q[i] = (re1 - re2) * C[4 * i - 1] - (im1 + im2) * S[4 * i - 1];
      q[nl - i] = q[i] - tmp;
      q[i] = q[i] + tmp;
      tmp = re1 + re2;
      r[i] = (im1 + im2) * C[4 * i - 1] + (re1 - re2) * S[4 * i - 1];
      r[nl - i] = tmp + r[i];
      r[i] = tmp - r[i];
    }
    r[nl / 2] = tmp - r[nl / 2];
  }
  else if (tkt == 3)
  {
    for (i = 0; i < N / 4; i++)
    {
      float t = x[2 * i + 1];
      x[2 * i + 1] = -x[N - 1 - 2 * i];
      x[N - 1 - 2 * i] = -t;
    }
    for (i = 0; i < nl; i++)
    {
      r[i] = x[2 * i + 1] * c[i] - x[2 * i] * s[i];
      q[i] = x[2 * i] * c[i] + x[2 * i + 1] * s[i];
    }
  }
  else
  {
    for (i = 0; i < nl; i++)
    {
      const float a = x[2 * i], b = x[N - 1 - 2 * i];
      r[i] = (-(a * c[i])) - b * s[i];
      q[i] = b * c[i] - a * s[i];
    }
  }

  oracle_cplx_ifft_pow2(r, q, nl);

  if (tkt == 1 || tkt == 2)
  {
    for (i = 0; i < nl / 2; i++)
    {
      x[4 * i] = r[i];
      x[4 * i + 1] = q[nl - 1 - i];
      x[4 * i + 2] = q[i];
      x[4 * i + 3] = r[nl - 1 - i];
    }
    if (tkt == 1)
      for (i = 1; i < N; i += 2)
        x[i] = -x[i];
  }
  else if (tkt == 3)
  {
    for (i = 0; i < nl; i++)
    {
      x[2 * i + 1] = (-q[i]) * c[i] - r[i] * s[i];
      x[2 * i] = r[i] * c[i] - q[i] * s[i];
    }
    for (i = 0; i < N / 4; i++)
    {
      float t;
      x[2 * i] = -x[2 * i];
      x[2 * i + N / 2] = -x[2 * i + N / 2];
      t = x[2 * i + 1];
      x[2 * i + 1] = -x[N - 1 - 2 * i];
      x[N - 1 - 2 * i] = -t;
    }
  }
  else
  {
    for (i = 0; i < nl; i++)
    {
      x[2 * i] = -(r[i] * c[i] - q[i] * s[i]);
      x[N - 1 - 2 * i] = -(q[i] * c[i] + r[i] * s[i]);
    }
  }
}

static const float *window_table(int len, int shape)
{
  if (len == 1024)
    return shape ? g_rom.kbd_1024 : g_rom.sine_1024;
  return shape ? g_rom.kbd_128 : g_rom.sine_128;
}

/* ------------------------------------------------------------------------------------------------
 * Long block: windowing_long1 (ia_core_coder_basic_ops.c:124) for ONLY_LONG / LONG_START,
 * windowing_long3 (:230) for LONG_STOP / STOP_START, then the overlap save of
 * ia_core_coder_fd_imdct_long (ia_core_coder_imdct.c:762-815).
 * x = IMDCT output (1024).  s(i) is the unfolded second half of x.
 * ---------------------------------------------------------------------------------------------- */
static void long_block(const float *x, float *ov, float *out, int seq, int shape_prev, int tkt)
{
  int i;
  if (seq == ORACLE_ONLY_LONG || seq == ORACLE_LONG_START)
  {
    const float *w = window_table(1024, shape_prev);
    const int pos = (tkt == 2 || tkt == 3);
    for (i = 0; i < 512; i++)
    {
      const float a = x[512 + i];
      out[i] = a * w[i] + ov[i] * w[1023 - i];
      out[1023 - i] = (pos ? a : -a) * w[1023 - i] + ov[1023 - i] * w[i];
    }
  }
  else
  {
    const float *ws = window_table(128, shape_prev);
    const int pos = (tkt == 3);
    for (i = 0; i < 448; i++)
      out[i] = ov[i];
    for (i = 448; i < 512; i++)
      out[i] = x[512 + i] * ws[i - 448] + ov[i] * ws[127 - (i - 448)];
    for (i = 512; i < 576; i++)
    {
      const float a = x[512 + (1023 - i)];
      out[i] = (pos ? a : -a) * ws[i - 448] + ov[i] * ws[127 - (i - 448)];
    }
    for (i = 576; i < 1024; i++)
    {
      const float a = x[512 + (1023 - i)];
      out[i] = pos ? a : -a;
    }
  }
  for (i = 0; i < 512; i++)
  {
    const float a = x[i];
    ov[512 + i] = (tkt == 1 || tkt == 2) ? a : -a;
    ov[511 - i] = (tkt == 2 || tkt == 3) ? a : -a;
  }
}

/* ------------------------------------------------------------------------------------------------
 * Eight short blocks: ia_core_coder_fd_imdct_short (ia_core_coder_imdct.c:397) with
 * windowing_short2/3/4 (ia_core_coder_basic_ops.c:335,383,429), no FAC.
 * Work buffer positions p in [0, 2048): [0,1024) becomes the frame output, [1024,2048) the new overlap.
 * Block k occupies [448 + 128k, 448 + 128k + 256).  Each position receives at most the falling half of
 * one block followed by the rising half of the next, accumulated in that order onto 0.
 * ---------------------------------------------------------------------------------------------- */
static float short_rise(const float *xk, const float *w, int t, int tkt)
{ /* rising half of a block's window, sample t in [0,128) */
  if (t < 64)
    return xk[64 + t] * w[t];
  return (tkt == 3 ? xk[191 - t] : -xk[191 - t]) * w[t];
}
static float short_fall(const float *xk, const float *w, int t, int tkt, int windowed)
{ /* falling half, sample t in [0,128): window value w[127 - t]; the last block is left unwindowed */
  float v;
  if (t < 64)
    v = (tkt == 3) ? xk[63 - t] : -xk[63 - t];
  else
    v = -xk[t - 64];
  return windowed ? v * w[127 - t] : v;
}

static void short_blocks(const float *x, float *ov, float *out, int shape, int shape_prev, int tkt)
{
  float buf[2048];
  const float *wc = window_table(128, shape), *wp = window_table(128, shape_prev);
  int p, k, t;
  memset(buf, 0, sizeof(buf));
  for (p = 0; p < 448; p++)
    buf[p] = ov[p];
  /* block 0 rising half against the old overlap (windowing_short2) */
  for (t = 0; t < 64; t++)
  {
    const float a = x[64 + t];
    buf[448 + t] = a * wp[t] + ov[448 + t] * wp[127 - t];
    buf[448 + 127 - t] = (tkt == 3 ? a : -a) * wp[127 - t] + ov[448 + 127 - t] * wp[t];
  }
  /* block 0 falling half (windowing_short3) */
  for (t = 0; t < 128; t++)
  {
    float *d = &buf[576 + t];
    if (tkt == 3 && t < 64)
      *d = x[63 - t] * (wc[127 - t] + *d); /* reference quirk, basic_ops.c:404 */
    else
      *d += short_fall(x, wc, t, tkt, 1);
  }
  for (k = 1; k < 8; k++)
  {
    const float *xk = x + 128 * k;
    const int base = 448 + 128 * k;
    for (t = 0; t < 128; t++)
      buf[base + t] += short_rise(xk, wc, t, tkt);
    for (t = 0; t < 128; t++)
      buf[base + 128 + t] += short_fall(xk, wc, t, tkt, k != 7);
  }
  for (p = 0; p < 1024; p++)
  {
    out[p] = buf[p];
    ov[p] = buf[1024 + p];
  }
}

int oracle_fd_frm_dec(float *coef, float *overlap, float *out, int window_sequence, int window_shape,
                      int window_shape_prev, int prev_aliasing_symmetry, int curr_aliasing_symmetry)
{
  const int tkt = ((prev_aliasing_symmetry & 1) << 1) + (curr_aliasing_symmetry & 1);
  int k;
  if (window_sequence < 0 || window_sequence > 4 || (window_shape | window_shape_prev) & ~1)
    return -1;
  if (window_sequence == ORACLE_EIGHT_SHORT)
  {
    for (k = 0; k < 8; k++)
      imdct_block(coef + 128 * k, 128, tkt);
    short_blocks(coef, overlap, out, window_shape, window_shape_prev, tkt);
  }
  else
  {
    imdct_block(coef, 1024, tkt);
    long_block(coef, overlap, out, window_sequence, window_shape_prev, tkt);
  }
  return 0;
}

int oracle_fd_frm_dec_batch(const float *coef, float *overlap, float *out, const int *side, int n_frames,
                            int n_ch)
{
  int f, c, err = 0;
  float tmp[1024];
  for (f = 0; f < n_frames; f++)
    for (c = 0; c < n_ch; c++)
    {
      const int *s = side + ((size_t)f * n_ch + c) * 5;
      const size_t o = ((size_t)f * n_ch + c) * 1024;
      memcpy(tmp, coef + o, sizeof(tmp));
      err |= oracle_fd_frm_dec(tmp, overlap + (size_t)c * 1024, out + o, s[0], s[1], s[2], s[3], s[4]);
    }
  return err;
}
