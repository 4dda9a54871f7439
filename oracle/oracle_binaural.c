/* oracle/oracle_binaural.c -- TEST INFRASTRUCTURE ONLY (see oracle.h).
 *
 * Plain-C restatement of the reference's time-domain binaural renderer (FFT overlap-add fast convolution):
 *   filter spectra      impeghd_binaural_struct_set             decoder/impeghd_binaural_renderer.c:244
 *   one block           impeghd_binaural_struct_process_ld_ola  :543  (the body of its while loop)
 *   spectral MAC        impeghd_binaural_mul_add_perm           :498
 *   packed formats      impeghd_fft_complex_to_perm_fmt_i :112, impeghd_fft_perm_fmt_to_complex_i :466
 * A "perm" spectrum of an N-point transform of real data is N floats: [0] = Re X[0], [1] = Re X[N/2],
 * [2k] = Re X[k], [2k+1] = -Im X[k] for 0 < k < N/2, X being the output of impeghd_cplx_ifft_8k.
 *
 * Reference behaviours kept on purpose:
 *   - the forward direction uses the e^{+j} transform (with its 8-point column quirk, see oracle_fft.c) and
 *     the inverse direction the e^{-j} one; 1/N is applied after the inverse transform;
 *   - impeghd_fft_perm_fmt_to_complex_i never writes Im Y[0] and Im Y[N/2]: they keep whatever the scratch
 *     array holds, i.e. Im X[0] and Im X[N/2] of the LAST input channel's forward transform of this block;
 *   - the diffuse downmix spectrum of the current block is stored but first used by the NEXT block.
 * NOT reproducible: with 2 diffuse blocks the reference cycles write_idx over 3 slots of prev_in_buf[2][..]
 * (impeghd_binaural_renderer.h:74, renderer.c:653): the third slot aliases diffuse_weight[] and whatever
 * follows the struct.  This restatement keeps a real 3-slot ring (the evident intent); only 0 and 1 diffuse
 * blocks are pinned against the reference.
 */
#include <stdlib.h>
#include <string.h>

#include "oracle.h"

/* taps[n_taps] -> perm spectrum out[fft] (n_taps <= fft/2) */
int oracle_binaural_spectrum(const float *taps, int n_taps, int fft, float *out)
{
  float *re = (float *)calloc((size_t)2 * fft, sizeof(float)), *im = re + fft;
  int i, rc;
  memcpy(re, taps, sizeof(float) * n_taps);
  rc = oracle_cplx_ifft_big(re, im, fft);
  out[0] = re[0];
  out[1] = re[fft / 2];
  for (i = 1; i < fft / 2; i++)
  {
    out[2 * i] = re[i];
    out[2 * i + 1] = -1 * im[i];
  }
  free(re);
  return rc;
}

static void mul_add_perm(const float *a, const float *b, float *d, int len)
{
  int i;
  d[0] = d[0] + a[0] * b[0];
  d[1] = d[1] + a[1] * b[1];
  for (i = 1; i < len; i++)
  {
    float v = d[2 * i] + a[2 * i] * b[2 * i];
    d[2 * i] = v - a[2 * i + 1] * b[2 * i + 1];
    v = d[2 * i + 1] + a[2 * i] * b[2 * i + 1];
    d[2 * i + 1] = v + a[2 * i + 1] * b[2 * i];
  }
}

/* One block of B = fft/2 samples per input.
 *   h_direct [2][n_in][fft], h_diffuse [n_blocks][2][fft]  perm spectra
 *   direct_len [2][n_in], diffuse_len [n_blocks][2]         MAC lengths in bins
 *   weight [n_in]                                           diffuse downmix weights
 *   prev_in [n_blocks + 1][fft], *write_idx, tail [2][B]    state, advanced
 *   in [n_in][B] -> out [2][B] */
int oracle_binaural_block(int fft, int n_in, int n_blocks, const float *h_direct, const float *h_diffuse,
                          const int *direct_len, const int *diffuse_len, const float *weight, float *prev_in,
                          int *write_idx, float *tail, const float *in, float *out)
{
  const int B = fft / 2;
  float *spec = (float *)calloc((size_t)n_in * fft + 3 * (size_t)fft, sizeof(float));
  float *re = spec + (size_t)n_in * fft, *im = re + fft, *acc = im + fft;
  float *dmx = prev_in + (size_t)(*write_idx) * fft;
  int inp, o, j, k, rc = 0;
  for (inp = 0; inp < n_in; inp++)
  {
    float *x = spec + (size_t)inp * fft;
    const float w = weight[inp];
    memset(re, 0, sizeof(float) * fft);
    memset(im, 0, sizeof(float) * fft);
    memcpy(re, in + (size_t)inp * B, sizeof(float) * B);
    rc |= oracle_cplx_ifft_big(re, im, fft);
    x[0] = re[0];
    x[1] = re[B];
    for (j = 1; j < B; j++)
    {
      x[2 * j] = re[j];
      x[2 * j + 1] = -1 * im[j];
    }
    if (inp == 0)
      for (j = 0; j < fft; j++)
        dmx[j] = x[j] * w;
    else
      for (j = 0; j < fft; j++)
        dmx[j] += x[j] * w;
  }
  /* im[] now holds the last input's imaginary output: im[0] and im[B] leak into the inverse transform */
  for (o = 0; o < 2; o++)
  {
    const float im0 = im[0], imn = im[B];
    const float scale = 1.f / fft;
    memset(acc, 0, sizeof(float) * fft);
    for (inp = 0; inp < n_in; inp++)
      mul_add_perm(h_direct + ((size_t)o * n_in + inp) * fft, spec + (size_t)inp * fft, acc,
                   direct_len[o * n_in + inp]);
    k = (*write_idx + n_blocks) % (n_blocks + 1);
    for (j = 0; j < n_blocks; j++)
    {
      mul_add_perm(h_diffuse + ((size_t)j * 2 + o) * fft, prev_in + (size_t)k * fft, acc, diffuse_len[j * 2 + o]);
      k = (k + n_blocks) % (n_blocks + 1);
    }
    {
      float *yr = (float *)calloc((size_t)2 * fft, sizeof(float)), *yi = yr + fft;
      yr[0] = acc[0];
      yr[B] = acc[1];
      yi[0] = im0;
      yi[B] = imn;
      for (j = 1; j < B; j++)
      {
        yi[j] = -1 * acc[2 * j + 1];
        yi[fft - j] = -1 * yi[j];
        yr[j] = acc[2 * j];
        yr[fft - j] = yr[j];
      }
      rc |= oracle_cplx_fft_big(yr, yi, fft);
      for (j = 0; j < fft; j++)
        yr[j] = yr[j] * scale;
      for (j = 0; j < B; j++)
        yr[j] = yr[j] + tail[o * B + j];
      memcpy(out + (size_t)o * B, yr, sizeof(float) * B);
      memcpy(tail + (size_t)o * B, yr + B, sizeof(float) * B);
      /* the inverse transform overwrote the shared imaginary scratch: the second ear sees ITS leftovers */
      im[0] = yi[0];
      im[B] = yi[B];
      free(yr);
    }
  }
  *write_idx = (*write_idx + 1) % (n_blocks + 1);
  free(spec);
  return rc;
}
