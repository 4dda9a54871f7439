/* oracle/oracle_fft.c -- TEST INFRASTRUCTURE ONLY (see oracle.h).
 *
 * Plain-C restatement of
 *   impeghd_rad2_cplx_fft    decoder/impeghd_fft.c:297    e^{-j} transform, n = 2^k <= 1024
 *   impeghd_cplx_ifft_8k     decoder/impeghd_ifft.c:1264  n = 1024 x {2, 8, 16} (see the notes on 4096 and 16384)
 *   impeghd_cplx_fft_16k     decoder/impeghd_fft.c:1415   n = 1024 x {2, 4, 8, 16}
 * The e^{-j} transform is the conjugate mirror of oracle_cplx_ifft_pow2 (oracle_fd.c) node for node, with
 * one rounding difference: its first pass forms a - 2b as (a - b) - b (impeghd_fft.c:352-386), two
 * roundings instead of one.  It is written out natively rather than as conj(ifft(conj(x))) because the
 * mirror does not commute with the sign of exact-cancellation zeros.
 *
 * Reference behaviour kept bug for bug (the binaural renderer's results depend on it):
 *   - impeghd_cplx_ifft_8 (impeghd_ifft.c:1118) writes bins 2 and 6 as (n21 - x4, n20 + x5) / (n21 + x4,
 *     n20 - x5): real and imaginary parts of the even half are swapped relative to an 8-point e^{+j} DFT.
 *   - impeghd_cplx_ifft_8k has no column transform for n = 16384 (impeghd_ifft.c:1318-1359): what comes back is the
 *     twiddled row transforms, transposed.  Restated as it is (a pure function of the input; pinned by
 *     tests/test_oracle_binaural.py).
 *   - its 4-point column transform (impeghd_cplx_ifft_4, :1212) reads elements 4 and 6 of 4-element rows: the
 *     imaginary row, the next row's (not yet transformed) real part and, for row 1023, two floats BEHIND the transform's
 *     interim buffer -- scratch memory the transform never wrote.  n = 4096 is therefore not a function of its input
 *     (tests/test_oracle_binaural.py::test_ifft_4096_depends_on_stale_scratch shows two different results for one input)
 *     and is not restated: oracle_cplx_ifft_big returns -1 for it.
 */
#include <string.h>

#include "oracle.h"

extern const oracle_rom *oracle_get_rom(void);

typedef struct
{
  float r, i;
} cpx;

static unsigned rev4_16(unsigned v)
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

/* x * (wc + j*ws) in the reference's product order (ws is the table's -sin entry) */
static cpx rotf_std(cpx x, float wc, float ws)
{
  cpx o;
  o.r = x.r * wc - x.i * ws;
  o.i = x.r * ws + x.i * wc;
  return o;
}
static cpx rotf_fold(cpx x, float a, float b)
{
  cpx o;
  o.r = x.r * b + x.i * a;
  o.i = -(x.r * a) + x.i * b;
  return o;
}

/* forward butterfly; two_step selects the first-pass (a - b) - b form */
static void bfly4_fwd(cpx *p0, cpx *p1, cpx *p2, cpx *p3, int region4, int two_step)
{
  cpx x0 = *p0, x1 = *p1, x2 = *p2, x3 = *p3;
  float dr, di;
  x0.r = x0.r + x2.r;
  x0.i = x0.i + x2.i;
  if (two_step)
  {
    float t = x0.r - x2.r;
    x2.r = t - x2.r;
    t = x0.i - x2.i;
    x2.i = t - x2.i;
  }
  else
  {
    x2.r = x0.r - x2.r * 2;
    x2.i = x0.i - x2.i * 2;
  }
  x1.r = x1.r + x3.r;
  if (two_step)
  {
    float t;
    x1.i = x1.i + x3.i;
    t = x1.r - x3.r;
    x3.r = t - x3.r;
    t = x1.i - x3.i;
    x3.i = t - x3.i;
  }
  else
  {
    x3.r = x1.r - x3.r * 2;
    if (!region4)
    {
      x1.i = x1.i + x3.i;
      x3.i = x1.i - x3.i * 2;
    }
    else
    {
      x1.i = x1.i - x3.i;
      x3.i = x1.i + x3.i * 2;
    }
  }
  x0.r = x0.r + x1.r;
  x0.i = x0.i + x1.i;
  if (two_step)
  {
    float t = x0.r - x1.r;
    x1.r = t - x1.r;
    t = x0.i - x1.i;
    x1.i = t - x1.i;
  }
  else
  {
    x1.r = x0.r - x1.r * 2;
    x1.i = x0.i - x1.i * 2;
  }
  x2.r = x2.r + x3.i;
  x2.i = x2.i - x3.r;
  if (two_step)
  {
    float t = x2.r - x3.i;
    dr = t - x3.i;
    t = x2.i + x3.r;
    di = t + x3.r;
  }
  else
  {
    dr = x2.r - x3.i * 2;
    di = x2.i + x3.r * 2;
  }
  p0->r = x0.r, p0->i = x0.i;
  p1->r = x2.r, p1->i = x2.i;
  p2->r = x1.r, p2->i = x1.i;
  p3->r = dr, p3->i = di;
}

void oracle_cplx_fft_pow2(float *re, float *im, int n)
{
  cpx y[1024];
  const float *w = oracle_get_rom()->fft_twiddle;
  int lg = 0, odd, npass, b, p, m, g, del, ns, groups, i;
  while ((1 << lg) < n)
    lg++;
  odd = lg & 1;
  npass = lg >> 1;

  for (b = 0; b < n / 4; b++)
  {
    unsigned h = rev4_16((unsigned)(4 * b)) >> (15 - lg);
    int c, q = n / 4;
    cpx x[4];
    if (odd)
      h = (h + 1) & ~1u;
    c = (int)(h >> 1);
    for (i = 0; i < 4; i++)
      x[i].r = re[c + i * q], x[i].i = im[c + i * q];
    bfly4_fwd(&x[0], &x[1], &x[2], &x[3], 0, 1);
    for (i = 0; i < 4; i++)
      y[4 * b + i] = x[i];
  }

  del = 4;
  ns = 64;
  groups = n >> 4;
  for (p = 1; p < npass; p++)
  {
    const int full = ns * del;
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
          x1 = rotf_std(x1, w[j], w[j + 257]);
          if (j <= full / 2)
            x2 = rotf_std(x2, w[2 * j], w[2 * j + 257]);
          else
            x2 = rotf_fold(x2, w[2 * j - 256], w[2 * j + 1]);
          if (j <= lim1)
            x3 = rotf_std(x3, w[3 * j], w[3 * j + 257]);
          else if (j <= 2 * lim1)
            x3 = rotf_fold(x3, w[3 * j - 256], w[3 * j + 1]);
          else
          {
            const float a = w[3 * j - 512], bq = w[3 * j - 512 + 257];
            cpx t;
            t.r = -(x3.r * a) + x3.i * bq;
            t.i = x3.r * bq + x3.i * a;
            x3 = t;
            region4 = 1;
          }
        }
        bfly4_fwd(&x0, &x1, &x2, &x3, region4, 0);
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
        t = rotf_std(x1, wc, ws);
      else
      {
        t.r = x1.r * ws + x1.i * wc;
        t.i = -(x1.r * wc) + x1.i * ws;
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

/* ---- small column transforms --------------------------------------------------------------------------- */
/* impeghd_cplx_ifft_8 (impeghd_ifft.c:1118), including its swapped bins 2 and 6 */
static void col_ifft8(float *xr, float *xi)
{
  const float s = 0.7071067811f;
  float ar, ai, br, bi, cr, ci, dr, di;
  float n00, n01, n10, n11, n20, n21, n30, n31;
  float e0, e1, e2, e3, e4, e5, e6, e7, t0, t1;
  ar = xr[0] + xr[4], ai = xi[0] + xi[4], br = xr[0] - xr[4], bi = xi[0] - xi[4];
  cr = xr[2] + xr[6], ci = xi[2] + xi[6], dr = xr[2] - xr[6], di = xi[2] - xi[6];
  n00 = ar + cr, n01 = ai + ci, n10 = br - di, n11 = bi + dr;
  n20 = ar - cr, n21 = ai - ci, n30 = br + di, n31 = bi - dr;
  ar = xr[1] + xr[5], ai = xi[1] + xi[5], br = xr[1] - xr[5], bi = xi[1] - xi[5];
  cr = xr[3] + xr[7], ci = xi[3] + xi[7], dr = xr[3] - xr[7], di = xi[3] - xi[7];
  e0 = ar + cr, e1 = ai + ci, e2 = br - di, e3 = bi + dr;
  e4 = ar - cr, e5 = ai - ci, e6 = br + di, e7 = bi - dr;
  t0 = (e2 - e3) * s, t1 = (e3 + e2) * s;
  e2 = t0, e3 = t1;
  t0 = (e6 + e7) * -s, t1 = (e6 - e7) * s;
  e6 = t0, e7 = t1;
  xr[0] = n00 + e0, xi[0] = n01 + e1;
  xr[1] = n10 + e2, xi[1] = n11 + e3;
  xr[2] = n21 - e4, xi[2] = n20 + e5;
  xr[3] = n30 + e6, xi[3] = n31 + e7;
  xr[4] = n00 - e0, xi[4] = n01 - e1;
  xr[5] = n10 - e2, xi[5] = n11 - e3;
  xr[6] = n21 + e4, xi[6] = n20 - e5;
  xr[7] = n30 - e6, xi[7] = n31 - e7;
}

/* impeghd_cplx_fft_8 (impeghd_fft.c:1304) */
static void col_fft8(float *xr, float *xi)
{
  const float s = 0.7071067811f;
  float ar, ai, br, bi, cr, ci, dr, di;
  float n00, n01, n10, n11, n20, n21, n30, n31;
  float e0, e1, e2, e3, e4, e5, e6, e7, t0, t1;
  ar = xr[0] + xr[4], ai = xi[0] + xi[4], br = xr[0] - xr[4], bi = xi[0] - xi[4];
  cr = xr[2] + xr[6], ci = xi[2] + xi[6], dr = xr[2] - xr[6], di = xi[2] - xi[6];
  n00 = ar + cr, n01 = ai + ci, n20 = ar - cr, n21 = ai - ci;
  n10 = br + di, n11 = bi - dr, n30 = br - di, n31 = bi + dr;
  ar = xr[1] + xr[5], ai = xi[1] + xi[5], br = xr[1] - xr[5], bi = xi[1] - xi[5];
  cr = xr[3] + xr[7], ci = xi[3] + xi[7], dr = xr[3] - xr[7], di = xi[3] - xi[7];
  e0 = ar + cr, e1 = ai + ci, e4 = ar - cr, e5 = ai - ci;
  e2 = br + di, e3 = bi - dr, e6 = br - di, e7 = bi + dr;
  t0 = (e2 + e3) * s, t1 = (e3 - e2) * s;
  e2 = t0, e3 = t1;
  t0 = (e7 - e6) * s, t1 = (e6 + e7) * -s;
  e6 = t0, e7 = t1;
  t0 = e4, e4 = e5, e5 = -t0;
  xr[0] = n00 + e0, xi[0] = n01 + e1;
  xr[1] = n10 + e2, xi[1] = n11 + e3;
  xr[2] = n20 + e4, xi[2] = n21 + e5;
  xr[3] = n30 + e6, xi[3] = n31 + e7;
  xr[4] = n00 - e0, xi[4] = n01 - e1;
  xr[5] = n10 - e2, xi[5] = n11 - e3;
  xr[6] = n20 - e4, xi[6] = n21 - e5;
  xr[7] = n30 - e6, xi[7] = n31 - e7;
}

/* impeghd_cplx_fft_4 (impeghd_fft.c:1256) */
static void col_fft4(float *xr, float *xi)
{
  float ar = xr[0] + xr[2], ai = xi[0] + xi[2], br = xr[0] - xr[2], bi = xi[0] - xi[2];
  float cr = xr[1] + xr[3], ci = xi[1] + xi[3], dr = xr[1] - xr[3], di = xi[1] - xi[3];
  xr[0] = ar + cr, xi[0] = ai + ci;
  xr[2] = ar - cr, xi[2] = ai - ci;
  xr[1] = br + di, xi[1] = bi - dr;
  xr[3] = br - di, xi[3] = bi + dr;
}

/* n = 1024 * dim2: dim2 row transforms of 1024 points over the decimated input, inter-stage twiddle
 * from the [1024][16] table, dim2-point column transform, transposed write-back. */
static int big_transform(float *re, float *im, int n, int inverse)
{
  static __thread float rr[16][1024], ri[16][1024];
  const oracle_rom *rom = oracle_get_rom();
  const int dim2 = n >> 10, dim1 = 1024;
  const int fac = 16 / dim2;
  int i, j;
  if (n <= 1024)
  {
    if (inverse)
      oracle_cplx_ifft_pow2(re, im, n);
    else
      oracle_cplx_fft_pow2(re, im, n);
    return 0;
  }
  if (dim2 * 1024 != n || (dim2 != 2 && dim2 != 4 && dim2 != 8 && dim2 != 16))
    return -1;
  if (inverse && dim2 == 4)
    return -1; /* the reference's 4-point e^{+j} column transform reads outside the transform's buffers: see header */
  for (i = 0; i < dim2; i++)
  {
    for (j = 0; j < dim1; j++)
      rr[i][j] = re[dim2 * j + i], ri[i][j] = im[dim2 * j + i];
    if (inverse)
      oracle_cplx_ifft_pow2(rr[i], ri[i], dim1);
    else
      oracle_cplx_fft_pow2(rr[i], ri[i], dim1);
  }
  for (i = 0; i < dim1; i++)
  {
    float cr[16], ci[16];
    for (j = 0; j < dim2; j++)
    {
      const float c = rom->mixed_rad_cos[(i * dim2 + j) * fac], s = rom->mixed_rad_sin[(i * dim2 + j) * fac];
      const float a = rr[j][i], b = ri[j][i];
      if (inverse)
        cr[j] = a * c - b * s, ci[j] = a * s + b * c;
      else
        cr[j] = a * c + b * s, ci[j] = b * c - a * s;
    }
    if (dim2 == 8)
    {
      if (inverse)
        col_ifft8(cr, ci);
      else
        col_fft8(cr, ci);
    }
    else if (dim2 == 4)
      col_fft4(cr, ci);
    else if (dim2 == 2)
    {
      float a = cr[0] + cr[1], b = cr[0] - cr[1], c = ci[0] + ci[1], d = ci[0] - ci[1];
      cr[0] = a, cr[1] = b, ci[0] = c, ci[1] = d;
    }
    else if (!inverse)
      oracle_cplx_fft_pow2(cr, ci, 16);
    /* inverse, dim2 == 16: impeghd_cplx_ifft_8k has no column transform for this size (impeghd_ifft.c:1318-1359 tests
     * 8192, 4096 and 2048 only): the twiddled row outputs are written back as they are.  Not a Fourier transform of the
     * input, but a pure function of it, and what the reference's binaural renderer runs on for direct filters of 4097 ..
     * 8192 taps. */
    for (j = 0; j < dim2; j++)
      re[j * dim1 + i] = cr[j], im[j * dim1 + i] = ci[j];
  }
  return 0;
}

int oracle_cplx_ifft_big(float *re, float *im, int n) { return big_transform(re, im, n, 1); }
int oracle_cplx_fft_big(float *re, float *im, int n) { return big_transform(re, im, n, 0); }
