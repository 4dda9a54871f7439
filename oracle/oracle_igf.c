/* oracle/oracle_igf.c -- TEST INFRASTRUCTURE ONLY (see oracle.h).
 *
 * Intelligent gap filling of one FD channel-frame, mono path, in the reference's order of operations
 * (decoder/ia_core_coder_igf_dec.c): ia_core_coder_igf_mono (:2479) = source-tile whitening (:913), per-band
 * energies of the coded and of the mapped source spectrum (:1083, tile mapping :726-880), group normalisation
 * (:1002), gain and fill of the zero bins (:1321) with random-sign noise for whitening level 2 (:882); followed by
 * ia_core_coder_igf_tnf (:2297): temporal noise flattening of the filled range (detect :503, LPC :330-416,
 * filter :418).  Long blocks (igf_win_type 0) and eight-short blocks (type 1); the TCX window types 2/3 belong
 * to the LPD path and stay on the host.  Every product and sum is a separately rounded binary32 operation.
 */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "oracle.h"

#define TILE_STRIDE 1024

static const float k_acf_win[8] = {0.997803f, 0.991211f, 0.980225f, 0.964844f,
                                   0.945068f, 0.920898f, 0.892334f, 0.859375f};

/* 2^((42 - n) / 2) printed with seven decimals (entry 41 truncated, not rounded), n < 64:
 * ia_core_coder_pow_igf_whitening (decoder/ia_core_coder_rom.c:1287); 2^21 * 2^(n / 2): ..._pow_neg_... (:1302) */
static float g_whiten[2][64];
static int g_whiten_ready;

static void whitening_init(void)
{
  int n;
  if (g_whiten_ready)
    return;
  for (n = 0; n < 64; n++)
  { /* the ROM prints seven decimals */
    char buf[64];
    snprintf(buf, sizeof(buf), "%.7f", pow(2.0, (42 - n) / 2.0));
    g_whiten[0][n] = strtof(buf, 0);
    g_whiten[1][n] = (float)(2097152.0 * pow(2.0, n / 2.0));
  }
  g_whiten[0][41] = 1.4142135f;
  g_whiten_ready = 1;
}

static float whitening_gain(int n, int neg)
{
  whitening_init();
  return g_whiten[neg][n & 63];
}

/* out [2][64]: the two gain tables, for pinning against the reference ROM */
void oracle_igf_whitening_table(float *out)
{
  whitening_init();
  memcpy(out, g_whiten, sizeof(g_whiten));
}

int oracle_igf_tiles(int sfb_start, int sfb_stop, int igf_min, int use_high_res, const short *swb_offset, int *tile)
{
  const int igf_bgn = swb_offset[sfb_start], igf_end = swb_offset[sfb_stop];
  int range = igf_end - igf_bgn, mem = sfb_start, sbs = igf_bgn, n = 0, i, j;
  if (range < 8)
    range = 8;
  for (i = 0; i < 4; i++)
    tile[i] = range >> 2;
  if (igf_bgn > igf_min && igf_bgn < igf_end)
  {
    for (i = 0; i < 4; i++)
    {
      int lim = sbs + tile[i] < igf_end ? sbs + tile[i] : igf_end, t;
      j = 1;
      while (lim > swb_offset[j])
        j++;
      if (!use_high_res)
      {
        if (i == 3)
          j = sfb_stop;
        else if ((((j - sfb_start) & 1) == 1) && ((j - mem) > 1) && (i < 3))
          j--;
        mem = j;
      }
      t = swb_offset[j] - sbs < igf_end - sbs ? swb_offset[j] - sbs : igf_end - sbs;
      tile[i] = t > 2 ? t : 2;
      n++;
      sbs += tile[i];
      if (sbs == igf_end)
        break;
    }
  }
  for (i = n; i < 4; i++)
    tile[i] = 0;
  return n;
}

static int tile_of(int tb, int igf_bgn, int n_tiles, const int *tile)
{
  int t, sum = igf_bgn;
  for (t = 0; t < n_tiles; t++)
  {
    sum += tile[t];
    if (tb < sum)
      break;
  }
  return t;
}

static int source_bin(int igf_min, int igf_bgn, int tb, int n_tiles, const int *tile, const unsigned char *tile_num)
{
  const int src = igf_bgn - igf_min;
  int sbs = igf_bgn, offset = 0, sb = tb, i;
  if (src > 0)
  {
    for (i = 0; i < n_tiles; i++)
    {
      if (tb >= sbs && tb < sbs + tile[i])
      {
        const float sel = (float)(signed char)tile_num[i];
        offset = (int)((2.5f + (-0.5f * sel)) * tile[i] + sbs - igf_bgn);
        offset += offset % 2;
        break;
      }
      sbs += tile[i];
    }
    sb = tb - offset;
    if (sb < igf_min)
      sb = igf_min + tb % src;
  }
  return sb;
}

static float random_sign(unsigned int *seed)
{
  *seed = (unsigned int)((*seed) * 69069ull + 5);
  return ((*seed) & 0x10000) ? -1.0f : 1.0f;
}

static void whiten(const oracle_igf_side *sd, int n_tiles, int stop, float *tiles)
{
  float flat[2048];
  float fac = 0.0f; /* like n, carried from tile to tile (:927-928) */
  int ix, i, j, n = 0;
  for (ix = 0; ix < n_tiles; ix++)
  {
    float *x = tiles + ix * TILE_STRIDE;
    if (sd->whitening_level[ix] != 0)
      continue;
    for (i = sd->igf_min; i < stop - 3; i++)
    {
      float env = 1e-3f;
      for (j = i - 3; j <= i + 3; j++)
        env = env + x[j] * x[j];
      if (1.0f <= env)
      { /* n = floor(log2 env): first power of two above env, minus one */
        int e;
        for (e = 0; e < 64; e++)
          if (env < ldexpf(1.0f, e))
          {
            n = e - 1;
            break;
          }
        fac = whitening_gain(n, 0);
      }
      else
      {
        int e;
        for (e = 0; e < 36; e++)
          if (ldexpf(1.0f, -e) <= env)
          {
            n = e;
            break;
          }
        fac = whitening_gain(n, 1);
      }
      flat[i] = x[i] * fac;
    }
    for (; i < stop; i++)
      flat[i] = x[i] * fac;
    if (stop > sd->igf_min)
      memcpy(x + sd->igf_min, flat + sd->igf_min, (stop - sd->igf_min) * sizeof(float));
  }
}

/* temporal noise flattening of tnf[bgn, end) (:503-584); order stays 0 when a third of the range is silent */
static void tnf_flatten(float *tnf, int bgn, int end)
{
  const int range = end - bgn;
  float norms[3], acf[9] = {0}, par[16] = {0}, pred[17] = {0}, buf[1024];
  float mem[32] = {0};
  int i, j, lag, order = 0, lpc_order;
  for (i = 0; i < 3; i++)
  {
    const int a = bgn + range * i / 3, b = bgn + range * (i + 1) / 3;
    float acc = 0.0f;
    for (j = a; j < b; j++)
      acc = acc + tnf[j] * tnf[j];
    norms[i] = acc;
  }
  for (i = 0; i < 3 && norms[i] > 0.0000000037f; i++)
  {
    const int a = bgn + range * i / 3, b = bgn + range * (i + 1) / 3, n = b - a;
    const float fac = 1.0f / norms[i];
    for (lag = 1; lag <= 8; lag++)
    {
      float acc = 0.0f;
      if (n - lag)
        acc = tnf[a] * tnf[a + lag];
      for (j = 1; j < n - lag; j++)
        acc = acc + tnf[a + j] * tnf[a + j + lag];
      acf[lag] = acf[lag] + fac * (k_acf_win[lag - 1] * acc);
    }
  }
  if (i != 3)
    return;
  order = 8;
  acf[0] = 3.0f;
  lpc_order = range >> 2 < 8 ? range >> 2 : 8;
  /* reflection coefficients (:330) */
  {
    float *pm = mem + lpc_order;
    memcpy(mem, acf, lpc_order * sizeof(float));
    memcpy(pm, acf + 1, lpc_order * sizeof(float));
    for (i = 0; i < lpc_order; i++)
    {
      float t = 0.0f;
      if (mem[0] >= 0.0000152f)
        t = -pm[i] / mem[0];
      t = t > -0.999f ? t : -0.999f;
      t = t < 0.999f ? t : 0.999f;
      par[i] = t;
      for (j = i; j < lpc_order; j++)
      {
        const float t2 = pm[j] + t * mem[j - i];
        mem[j - i] = mem[j - i] + t * pm[j];
        pm[j] = t2;
      }
    }
  }
  /* to prediction coefficients (:375) */
  pred[0] = 1.0f;
  pred[1] = par[0];
  for (i = 1; i < lpc_order; i++)
  {
    for (j = 0; j < (i >> 1); j++)
    {
      const float t = pred[j + 1];
      pred[j + 1] = pred[j + 1] + par[i] * pred[i - 1 - j + 1];
      pred[i - 1 - j + 1] = pred[i - 1 - j + 1] + par[i] * t;
    }
    if (i & 1)
      pred[j + 1] = pred[j + 1] + par[i] * pred[j + 1];
    pred[i + 1] = par[i];
  }
  /* FIR over the unfiltered values (:418) */
  for (j = 0; j < range; j++)
  {
    buf[j] = tnf[bgn + j];
    for (i = 1; i < order && j >= i; i++)
      tnf[bgn + j] = tnf[bgn + j] + pred[i] * buf[j - i];
  }
}

unsigned int oracle_igf_mono(const oracle_igf_side *sd, const short *sfb_top, int n_sfb, float *coef, float *tiles,
                             float *tnf_state)
{
  short swb[72];
  int tile[4], n_tiles, sfb, bin, w = 0, g, k;
  const int bins = sd->win_type == 1 ? 128 : 1024;
  const int lvl_stride = sd->win_type == 1 ? 16 : 0;
  const int step = (sd->use_high_res || sd->win_type >= 1) ? 1 : 2;
  const int noise_ok = sd->win_type != 1 && sd->use_whitening;
  const int tnf_on = sd->use_tnf && sd->win_type != 1;
  unsigned int seed_calc = sd->seed, seed_apply = sd->seed;
  /* igf_tnf_spec_float / igf_tnf_mask persist in the decoder handle: a frame that signals TNF without having
   * filled anything (igf_all_zero) re-flattens and re-applies what the last filled long block left there */
  float *tnf_spec = tnf_state;
  float *tnf_mask = tnf_state + 1024; /* 1 = coded bin, 0 = filled bin */
  int igf_bgn, igf_end;

  swb[0] = 0;
  for (sfb = 0; sfb < n_sfb; sfb++)
    swb[sfb + 1] = sfb_top[sfb];
  igf_bgn = swb[sd->sfb_start], igf_end = swb[sd->sfb_stop];
  if (sd->all_zero)
    goto tnf;
  n_tiles = oracle_igf_tiles(sd->sfb_start, sd->sfb_stop, sd->igf_min, sd->use_high_res, swb, tile);
  if (noise_ok)
    whiten(sd, n_tiles, igf_bgn, tiles);

  for (g = 0; g < sd->num_groups; g++)
  {
    float sn[64] = {0}, pn[64] = {0};
    const int w0 = w;
    for (k = 0; k < sd->group_len[g]; k++, w++)
    { /* energies of the coded spectrum and of the mapped source on the zero bins (:1083) */
      const float *c = coef + w * bins;
      for (sfb = sd->sfb_start; sfb < sd->sfb_stop; sfb += step)
      {
        const int hi = sfb + step < sd->sfb_stop ? sfb + step : sd->sfb_stop;
        float e = 0.0f;
        for (bin = swb[sfb]; bin < swb[hi]; bin++)
          e = e + c[bin] * c[bin];
        sn[sfb] = sn[sfb] + e;
        e = 0.0f;
        for (bin = swb[sfb]; bin < swb[hi]; bin++)
          if (c[bin] == 0)
          {
            const int ix = tile_of(bin, igf_bgn, n_tiles, tile);
            const int sb = source_bin(sd->igf_min, igf_bgn, bin, n_tiles, tile, sd->tile_num);
            float v = tiles[ix * TILE_STRIDE + w * bins + sb];
            if (noise_ok && sd->whitening_level[ix] == 2)
              v = random_sign(&seed_calc) * 2097152.0f;
            e = e + v * v;
          }
        pn[sfb] = pn[sfb] + e;
      }
    }
    for (sfb = sd->sfb_stop - 1; sfb >= sd->sfb_start; sfb--)
    {
      sn[sfb] /= sd->group_len[g];
      pn[sfb] /= sd->group_len[g];
    }
    for (w = w0, k = 0; k < sd->group_len[g]; k++, w++)
    { /* gains and fill (:1321) */
      float *c = coef + w * bins;
      if (tnf_on)
      {
        for (bin = 0; bin < 1024; bin++)
          tnf_spec[bin] = 0.0f, tnf_mask[bin] = 1.0f;
      }
      for (sfb = sd->sfb_start; sfb < sd->sfb_stop; sfb += step)
      {
        const int hi = sfb + step < sd->sfb_stop ? sfb + step : sd->sfb_stop;
        const int width = swb[hi] - swb[sfb];
        const float lev = sd->levels[lvl_stride * g + sfb];
        const float mn = (lev * lev) * (float)width - sn[sfb];
        float gn = 0.0f;
        if (0 < mn && 0 < pn[sfb])
        {
          gn = (float)sqrt((double)mn / pn[sfb]);
          gn = gn < 10.f ? gn : 10.f;
        }
        for (bin = swb[sfb]; bin < swb[hi]; bin++)
        {
          const int ix = tile_of(bin, igf_bgn, n_tiles, tile);
          const int sb = source_bin(sd->igf_min, igf_bgn, bin, n_tiles, tile, sd->tile_num);
          float v = tiles[ix * TILE_STRIDE + w * bins + sb];
          const float cb = c[bin];
          if (cb == 0 && noise_ok && sd->whitening_level[ix] == 2)
            v = random_sign(&seed_apply) * 2097152.0f;
          if (tnf_on)
          {
            tnf_spec[bin] = gn * v;
            if (cb == 0)
              tnf_mask[bin] = 0;
          }
          if (cb == 0)
            c[bin] = gn * v;
        }
      }
    }
  }
tnf:
  if (tnf_on && sd->is_tnf == 1)
  { /* ia_core_coder_igf_tnf (:2297): flatten, then take the flattened values on the filled bins */
    tnf_flatten(tnf_spec, igf_bgn, igf_end);
    for (bin = igf_bgn; bin < igf_end; bin++)
      if (tnf_mask[bin] == 0)
        coef[bin] = tnf_spec[bin];
  }
  return sd->all_zero ? sd->seed : seed_calc;
}

void oracle_igf_tnf(const oracle_igf_side *sd, const short *sfb_top, int n_sfb, float *coef, float *tnf_state)
{
  short swb[72];
  int sfb, bin, igf_bgn, igf_end;
  float *tnf_spec = tnf_state, *tnf_mask = tnf_state + 1024;
  if (!(sd->use_tnf && sd->win_type != 1 && sd->is_tnf == 1))
    return;
  swb[0] = 0;
  for (sfb = 0; sfb < n_sfb; sfb++)
    swb[sfb + 1] = sfb_top[sfb];
  igf_bgn = swb[sd->sfb_start], igf_end = swb[sd->sfb_stop];
  tnf_flatten(tnf_spec, igf_bgn, igf_end);
  for (bin = igf_bgn; bin < igf_end; bin++)
    if (tnf_mask[bin] == 0)
      coef[bin] = tnf_spec[bin];
}

void oracle_igf_stereo(const oracle_igf_side *sl, const oracle_igf_side *sr, const unsigned char *mask,
                       const short *sfb_top, int n_sfb, float *coef_l, float *coef_r, float *tiles_l, float *tiles_r,
                       float *tnf_l, float *tnf_r, unsigned int *seeds_out)
{
  const oracle_igf_side *sd[2] = {sl, sr};
  float *coef[2] = {coef_l, coef_r}, *tiles[2] = {tiles_l, tiles_r}, *tnf[2] = {tnf_l, tnf_r};
  short swb[72];
  int tile[4], n_tiles, sfb, bin, w = 0, g, k, ch;
  const int bins = sl->win_type == 1 ? 128 : 1024;
  const int stride = sl->win_type == 1 ? 16 : 0;
  const int step = (sl->use_high_res || sl->win_type >= 1) ? 1 : 2;
  const int noise_ok = sl->win_type != 1 && sl->use_whitening;
  const int tnf_on = sl->use_tnf && sl->win_type != 1;
  unsigned int seed_calc[2] = {sl->seed, sr->seed}, seed_apply[2] = {sl->seed, sr->seed};
  int igf_bgn;

  seeds_out[0] = sl->seed, seeds_out[1] = sr->seed;
  if (sl->all_zero && sr->all_zero)
    return;
  swb[0] = 0;
  for (sfb = 0; sfb < n_sfb; sfb++)
    swb[sfb + 1] = sfb_top[sfb];
  igf_bgn = swb[sl->sfb_start];
  n_tiles = oracle_igf_tiles(sl->sfb_start, sl->sfb_stop, sl->igf_min, sl->use_high_res, swb, tile);
  if (noise_ok)
  {
    whiten(sl, n_tiles, igf_bgn, tiles_l);
    whiten(sr, n_tiles, igf_bgn, tiles_r);
  }
  for (g = 0; g < sl->num_groups; g++)
  {
    float sn[2][64] = {{0}}, pn[2][64] = {{0}};
    const unsigned char *m = mask + stride * g;
    const int w0 = w;
    for (k = 0; k < sl->group_len[g]; k++, w++)
    { /* ia_core_coder_igf_calc_stereo (:1169) */
      for (sfb = sl->sfb_start; sfb < sl->sfb_stop; sfb += step)
      {
        const int hi = sfb + step < sl->sfb_stop ? sfb + step : sl->sfb_stop;
        float e[2] = {0.0f, 0.0f};
        for (bin = swb[sfb]; bin < swb[hi]; bin++)
          for (ch = 1; ch >= 0; ch--)
            e[ch] = e[ch] + coef[ch][w * bins + bin] * coef[ch][w * bins + bin];
        for (ch = 1; ch >= 0; ch--)
          sn[ch][sfb] = sn[ch][sfb] + e[ch];
        e[0] = e[1] = 0.0f;
        for (bin = swb[sfb]; bin < swb[hi]; bin++)
        {
          const int ix = tile_of(bin, igf_bgn, n_tiles, tile);
          float v[2];
          for (ch = 1; ch >= 0; ch--)
          {
            const int sb = source_bin(sl->igf_min, igf_bgn, bin, n_tiles, tile, sd[ch]->tile_num);
            v[ch] = tiles[ch][ix * TILE_STRIDE + w * bins + sb];
            if (noise_ok && sd[ch]->whitening_level[ix] == 2)
              v[ch] = 2097152.f * random_sign(&seed_calc[ch]);
          }
          if (m[sfb])
          {
            const float t = v[0];
            v[0] = 0.5f * (t + v[1]);
            v[1] = 0.5f * (t - v[1]);
          }
          for (ch = 1; ch >= 0; ch--)
            if (coef[ch][w * bins + bin] == 0.0f)
              e[ch] = e[ch] + v[ch] * v[ch];
        }
        for (ch = 1; ch >= 0; ch--)
          pn[ch][sfb] = pn[ch][sfb] + e[ch];
      }
    }
    if (sl->group_len[g] != 1)
      for (ch = 0; ch < 2; ch++)
        for (sfb = sl->sfb_stop - 1; sfb >= sl->sfb_start; sfb--)
        {
          sn[ch][sfb] /= sl->group_len[g];
          pn[ch][sfb] /= sl->group_len[g];
        }
    for (w = w0, k = 0; k < sl->group_len[g]; k++, w++)
    { /* ia_core_coder_igf_stereo_proc (:1431) */
      if (tnf_on)
        for (ch = 0; ch < 2; ch++)
          for (bin = 0; bin < 1024; bin++)
            tnf[ch][bin] = 0.0f, tnf[ch][1024 + bin] = 1.0f;
      for (sfb = sl->sfb_start; sfb < sl->sfb_stop; sfb += step)
      {
        const int hi = sfb + step < sl->sfb_stop ? sfb + step : sl->sfb_stop;
        const int width = swb[hi] - swb[sfb];
        float gn[2];
        for (ch = 1; ch >= 0; ch--)
        {
          const float lev = sd[ch]->levels[stride * g + sfb];
          const float mn = (lev * lev) * (float)width - sn[ch][sfb];
          gn[ch] = 0.0f;
          if (0.0f < mn && 0.0f < pn[ch][sfb])
          {
            const float q = (float)sqrt((double)(mn / pn[ch][sfb])); /* float quotient, then double sqrt */
            gn[ch] = q < 10.f ? q : 10.f;
          }
        }
        for (bin = swb[sfb]; bin < swb[hi]; bin++)
        {
          const int ix = tile_of(bin, igf_bgn, n_tiles, tile);
          const int i = w * bins + bin;
          float v[2], cb[2];
          for (ch = 1; ch >= 0; ch--)
          {
            const int sb = source_bin(sl->igf_min, igf_bgn, bin, n_tiles, tile, sd[ch]->tile_num);
            v[ch] = tiles[ch][ix * TILE_STRIDE + w * bins + sb];
            if (noise_ok && sd[ch]->whitening_level[ix] == 2)
              v[ch] = 2097152.f * random_sign(&seed_apply[ch]);
          }
          if (m[sfb])
          {
            const float t = v[0];
            v[0] = 0.5f * (t + v[1]);
            v[1] = 0.5f * (t - v[1]);
          }
          cb[0] = coef[0][i], cb[1] = coef[1][i];
          if (tnf_on)
          {
            for (ch = 1; ch >= 0; ch--)
            {
              tnf[ch][bin] = gn[ch] * v[ch];
              if (cb[ch] == 0.0f)
                tnf[ch][1024 + bin] = 0.0f;
            }
            if (m[sfb])
            {
              const float any = (tnf[1][1024 + bin] != 0.0f || tnf[0][1024 + bin] != 0.0f) ? 1.0f : 0.0f;
              tnf[1][1024 + bin] = any;
              tnf[0][1024 + bin] = any;
            }
          }
          for (ch = 1; ch >= 0; ch--)
            if (cb[ch] == 0.0f)
              coef[ch][i] = gn[ch] * v[ch];
        }
      }
    }
  }
  seeds_out[0] = seed_calc[0], seeds_out[1] = seed_calc[1];
}
