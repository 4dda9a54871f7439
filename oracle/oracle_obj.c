/* oracle/oracle_obj.c -- TEST INFRASTRUCTURE ONLY (see oracle.h).
 *
 * Plain-C restatement of the reference object renderer's per-frame work:
 *   impeghd_obj_renderer_dec              decoder/impeghd_obj_ren_dec.c:1319   (ramped mix)
 *   impegh_obj_ren_vbap_process           decoder/impeghd_obj_ren_vbap.c:485   (gains -> imaginary-speaker
 *                                                                               downmix -> power normalise)
 *   impeghd_calc_vbap                     :437
 *   impeghd_calc_one_source_point         :156   (triangle search)
 *   impeghd_spreading_calc_mdap_vectors   :336, impeghd_vector_process :262, impeghd_calc_spread_gains :113
 *   impeghd_nrm_values                    :82
 *   impeghd_pol_2_cart_degree             decoder/impeghd_3d_vec_basic_ops.c:86
 * The layout-dependent tables (triangles, inverse matrices, speaker vectors, downmix matrix) come from the
 * reference's init (impeghd_obj_renderer_dec_init, :1057), which stays host code.
 * Every float operation below is a separately rounded binary32 operation in the reference's order; sqrt /
 * sin / cos / tan go through double like the reference's ia_sqrt_flt / libm calls.
 */
#include <math.h>
#include <string.h>

#include "oracle.h"

#define DEG2RAD ((float)(M_PI / 180.0f))

static void pol2cart(float phi, float theta, float radius, float *c /* x,y,z */)
{
  float sin_phi = (float)sin(phi * DEG2RAD), cos_phi = (float)cos(phi * DEG2RAD);
  float sin_theta = (float)sin(theta * DEG2RAD), cos_theta = (float)cos(theta * DEG2RAD);
  c[2] = radius * sin_phi;
  c[1] = (radius * sin_theta) * cos_phi;
  c[0] = (radius * cos_theta) * cos_phi;
}

/* host-side trigonometry of one object: what mpeghb200_obj_side_from_polar also produces */
void oracle_obj_trig(const float *q /* az, el, radius, gain, sw, sh, sd */, oracle_obj_side *s)
{
  float phi = q[1], theta = q[0];
  memset(s, 0, sizeof(*s));
  pol2cart(phi, theta, q[2], s->vec_c);
  s->gain = q[3];
  s->spread_w = q[4];
  s->spread_h = q[5];
  if (!(q[4] <= 0.0f && q[5] <= 0.0f))
  {
    float alpha = q[4], phi2;
    if (alpha > 90.0f)
      alpha = 90.0f;
    else if (alpha < 0.001f)
      alpha = 0.001f;
    s->tan_alpha = (float)tan(alpha * DEG2RAD);
    pol2cart(phi, theta, 1.0f, s->p0);
    phi2 = (phi < 0) ? phi + 90 : phi - 90;
    pol2cart(phi2, theta, 1.0f, s->v);
  }
}

static void one_source_point(const oracle_obj_layout *L, const float *vec, float *ls_gains)
{
  int i, j, winner = 0, hi_lo = 3;
  float highest_least = -100000.f, g[3] = {0.f, 0.f, 0.f}, power;
  for (i = 0; i < L->n_tri; i++)
  {
    float gains[3], min_gain = 10000000.f;
    int neg = 3;
    for (j = 2; j >= 0; j--)
    {
      const float *r = &L->inv[i][j * 3];
      gains[j] = vec[0] * r[0];
      gains[j] = gains[j] + vec[1] * r[1];
      gains[j] = gains[j] + vec[2] * r[2];
      if (gains[j] >= -0.01f)
        neg--;
      if (gains[j] < min_gain)
        min_gain = gains[j];
    }
    if ((neg < hi_lo) || (neg <= hi_lo && min_gain > highest_least))
    {
      g[0] = gains[0], g[1] = gains[1], g[2] = gains[2];
      highest_least = min_gain;
      hi_lo = neg;
      winner = i;
    }
  }
  for (i = 2; i >= 0; i--)
    if (g[i] < -0.01f)
      g[i] = 0.f; /* the source vector is also replaced (:209-222) but never read again */
  for (i = 2; i >= 0; i--)
    if (g[i] < 0.f)
      g[i] = 0.f;
  power = (float)sqrt(g[0] * g[0] + g[1] * g[1] + g[2] * g[2]);
  if (power != 0.f)
    for (i = 0; i < 3; i++)
      g[i] = g[i] / power;
  ls_gains[L->tri[winner][2]] = g[2] + ls_gains[L->tri[winner][2]];
  ls_gains[L->tri[winner][1]] = g[1] + ls_gains[L->tri[winner][1]];
  ls_gains[L->tri[winner][0]] = g[0] + ls_gains[L->tri[winner][0]];
}

static void sva(const float *a, float ga, const float *b, float gb, float *r)
{
  /* impeghd_scaled_vec_addition_cart; r may alias a */
  float z = ga * a[2] + gb * b[2], y = ga * a[1] + gb * b[1], x = ga * a[0] + gb * b[0];
  r[0] = x, r[1] = y, r[2] = z;
}

/* final gains of one object into fg[n_non_lfe] */
void oracle_obj_gains(const oracle_obj_layout *L, const oracle_obj_side *s, float *fg)
{
  float ls[44], m[19][3], norm, len, gain_sum;
  int tot = L->n_non_lfe + L->n_imag, i, j;
  memset(ls, 0, sizeof(ls));
  if (s->spread_w <= 0.0f && s->spread_h <= 0.0f)
  {
    one_source_point(L, s->vec_c, ls);
  }
  else
  {
    float u[3], v[3], p0[3], alpha, unit;
    memcpy(p0, s->p0, sizeof(p0));
    memcpy(v, s->v, sizeof(v));
    u[2] = v[0] * p0[1] - v[1] * p0[0];
    u[1] = v[2] * p0[0] - v[0] * p0[2];
    u[0] = v[1] * p0[2] - v[2] * p0[1];
    if (!L->height_present)
      v[0] = v[1] = v[2] = 0.0f;
    memcpy(m[0], s->vec_c, sizeof(m[0]));
    memcpy(m[1], u, sizeof(u));
    sva(u, 0.75f, p0, 0.25f, m[2]);
    sva(u, 0.375f, p0, 0.625f, m[3]);
    m[4][0] = -u[0], m[4][1] = -u[1], m[4][2] = -u[2];
    sva(u, -0.75f, p0, 0.25f, m[5]);
    sva(u, -0.375f, p0, 0.625f, m[6]);
    sva(u, 0.5f, v, 0.866f, m[7]);
    sva(m[7], 1.0f, p0, 0.333f, m[7]);
    sva(m[7], 0.5f, p0, 0.5f, m[8]);
    sva(m[7], 0.25f, p0, 0.75f, m[9]);
    sva(u, -0.5f, v, 0.866f, m[10]);
    sva(m[10], 1.0f, p0, 0.333f, m[10]);
    sva(m[10], 0.5f, p0, 0.5f, m[11]);
    sva(m[10], 0.25f, p0, 0.75f, m[12]);
    sva(u, -0.5f, v, -0.866f, m[13]);
    sva(m[13], 1.0f, p0, 0.333f, m[13]);
    sva(m[13], 0.5f, p0, 0.5f, m[14]);
    sva(m[13], 0.25f, p0, 0.75f, m[15]);
    sva(u, 0.5f, v, -0.866f, m[16]);
    sva(m[16], 1.0f, p0, 0.333f, m[16]);
    sva(m[16], 0.5f, p0, 0.5f, m[17]);
    sva(m[16], 0.25f, p0, 0.75f, m[18]);
    p0[0] = p0[0] / s->tan_alpha;
    p0[1] = p0[1] / s->tan_alpha;
    p0[2] = p0[2] / s->tan_alpha;
    for (i = 0; i < 19; i++)
    {
      float vl;
      m[i][0] = m[i][0] + p0[0];
      m[i][1] = m[i][1] + p0[1];
      m[i][2] = m[i][2] + p0[2];
      vl = (float)sqrt(m[i][0] * m[i][0] + (m[i][1] * m[i][1] + m[i][2] * m[i][2]));
      if (vl != 0.0f)
      {
        float inv = 1.0f / vl;
        m[i][0] = m[i][0] * inv, m[i][1] = m[i][1] * inv, m[i][2] = m[i][2] * inv;
      }
    }
    for (i = 0; i < 19; i++)
      one_source_point(L, m[i], ls);
    /* impeghd_calc_spread_gains with the raw width */
    norm = 0.0f;
    unit = (float)(1.f / sqrt((float)tot));
    alpha = (float)((s->spread_w * DEG2RAD) / (M_PI / 2.f) - 1.f);
    for (i = 0; i < tot; i++)
      norm = norm + ls[i] * ls[i];
    norm = (float)sqrt(norm);
    if (alpha > 1.0f)
      alpha = 1.0f;
    else if (alpha < 0.0f)
      alpha = 0.0f;
    for (i = 0; i < tot; i++)
      ls[i] = (1.f - alpha) * ls[i] / norm + alpha * unit;
  }
  /* impeghd_nrm_values */
  len = 0.0f;
  for (i = 0; i < tot; i++)
    len = len + ls[i] * ls[i];
  len = (float)sqrt(len);
  len = 1.0f / len;
  if (len != 0.0f)
    for (i = 0; i < tot; i++)
      ls[i] = ls[i] * len;
  /* imaginary-speaker downmix + power normalisation (impeghd_obj_ren_vbap.c:563-586) */
  gain_sum = 0.0f;
  for (i = 0; i < L->n_non_lfe; i++)
  {
    float value = 0;
    for (j = 0; j < tot; j++)
      value = value + L->dmx[i][j] * ls[j];
    fg[i] = value;
    gain_sum = gain_sum + value * value;
  }
  gain_sum = (float)sqrt(gain_sum);
  if (gain_sum != 0.0f)
    for (i = 0; i < L->n_non_lfe; i++)
      fg[i] = (fg[i] * s->gain) / gain_sum;
}

/* one call of impeghd_obj_renderer_dec (decoder/impeghd_obj_ren_dec.c:1319): frame_len samples from `offset` on of
 * 1024-sample rows, the gain ramp spread over frame_len samples; accumulates into out (the caller zeroed it once per
 * core frame, decoder/ia_core_coder_decode_main.c:1374) */
void oracle_obj_render_sub(const oracle_obj_layout *L, oracle_obj_state *st, const oracle_obj_side *side, int n_obj,
                           const float *in, float *out, int frame_len, int offset, int metadata_set)
{
  int obj, ch, j;
  for (obj = 0; obj < n_obj; obj++)
  {
    float fg[44];
    int lfe = 0;
    oracle_obj_gains(L, &side[obj], fg);
    /* impeghd_obj_ren_vbap.c:497 tests first_frame_flag[sub_frame_offset + obj] and writes initial_gains of THAT row:
     * only a call rendering with metadata set 0 starts the ramp at the final gains; :1367 clears the object's own flag
     * after every call */
    if (st->first_frame[obj] && metadata_set == 0)
      memcpy(st->initial_gains[obj], fg, sizeof(float) * L->n_non_lfe);
    for (ch = 0; ch < L->n_spk; ch++)
    {
      if (L->lfe[ch])
      {
        lfe++;
      }
      else
      {
        int ci = ch - lfe;
        float step = (fg[ci] - st->initial_gains[obj][ci]) / frame_len;
        float scale = st->initial_gains[obj][ci] + step;
        float *o = out + (size_t)ch * 1024 + offset;
        const float *x = in + (size_t)obj * 1024 + offset;
        for (j = 0; j < frame_len; j++)
        {
          o[j] = o[j] + x[j] * scale;
          scale = scale + step;
        }
      }
    }
    memcpy(st->initial_gains[obj], fg, sizeof(float) * L->n_non_lfe);
    st->first_frame[obj] = 0;
  }
}

/* one frame: in [n_obj][1024], out [n_spk][1024] (overwritten: zero + sum over objects, LFE rows zero) */
void oracle_obj_render(const oracle_obj_layout *L, oracle_obj_state *st, const oracle_obj_side *side, int n_obj,
                       const float *in, float *out)
{
  memset(out, 0, sizeof(float) * 1024 * L->n_spk);
  oracle_obj_render_sub(L, st, side, n_obj, in, out, 1024, 0, 0);
}
