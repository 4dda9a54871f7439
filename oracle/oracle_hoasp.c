/* oracle/oracle_hoasp.c -- TEST INFRASTRUCTURE ONLY (see oracle.h).
 *
 * Plain-C restatement of the signal half of the reference HOA spatial decoder, i.e. what
 * impeghd_hoa_spatial_process (decoder/impeghd_hoa_spatial_decoder.c:663) does after the side-info decode:
 *   inverse dynamic correction      decoder/impeghd_hoa_inverse_dyn_correction.c:75
 *   ambience synthesis              decoder/impeghd_hoa_amb_syn.c:76
 *   channel reassignment (clearing) decoder/impeghd_hoa_ch_reassignment.c:71
 *   vector-based synthesis          decoder/impeghd_hoa_vector_based_predom_sound_syn.c:76
 *   direction-based synthesis       decoder/impeghd_hoa_dir_based_pre_dom_sound_syn.c:145 (direction signals),
 *                                   :362 + :523 (spatial prediction), :304 (sum)
 * written per output coefficient in "gather" form: nothing but the T transport signals of the current frame
 * enters the sample arithmetic, every other dependence on the past is side information kept in
 * oracle_hoasp_state.  Every product and sum keeps the reference's order and rounding.
 *
 * Reference behaviours kept on purpose:
 *   - negative gain exponents all use row 6 of the window-power table (offset = min(|e| + 5, 6));
 *   - the signal-index set starts with the vector channels, direction channels only survive beyond
 *     position n_vec (impeghd_hoa_pre_dom_sound_syn.c:83-91);
 *   - "not found" tests compare against 49, not against the list length, so the entry one past the end of
 *     the direction lists (stale from earlier frames) decides the cross-fade window (dir_based...c:186,253);
 *   - the stale tail of vectors[].index takes part in the interpolation-length decision (vector_based...c:112).
 */
#include <math.h>
#include <string.h>

#include "oracle.h"

void oracle_hoasp_state_init(oracle_hoasp_state *st)
{
  int i;
  memset(st, 0, sizeof(*st));
  for (i = 0; i < HOASP_MAX_CH; i++)
  {
    st->prev_gain[i] = 1.0f;
    st->prev_ps_set[i] = -1;
  }
}

static float pow2f(const oracle_hoasp_tables *tb, int e) { return e >= 0 ? tb->pow2[e] : tb->pow2[64 - e]; }

int oracle_hoasp_process(const oracle_hoasp_tables *tb, oracle_hoasp_state *st, const hoasp_side *sd, const float *in,
                         float *out)
{
  static __thread float x[HOASP_MAX_CH][1024], xa[HOASP_MAX_CH][1024];
  static __thread float tmpc[HOASP_MAX_COEF][1024], tmpl[HOASP_MAX_COEF][1024];
  const int nc = tb->n_coef, T = tb->n_transport, N = 1024;
  const float *w_in_vec = tb->wins, *w_out_vec = tb->wins + 1024, *w_in_dir = tb->wins + 2048,
              *w_out_dir = tb->wins + 3072;
  int reassign[HOASP_MAX_COEF], keep[HOASP_MAX_CH], set[HOASP_MAX_CH];
  float fac[HOASP_MAX_PRED][HOASP_MAX_COEF];
  int ch, c, t, i, k, g;
  int cur_flag = 0, last_flag = 0;

  if (tb->use_phase_shift_decorr)
    return -1;
  /* ---- inverse dynamic correction ---------------------------------------------------------------- */
  for (ch = 0; ch < T; ch++)
  {
    const int e = sd->exponents[ch];
    int offset = (e < 0) ? (-e + 5) : (e - 1);
    float last, reci;
    if (sd->prev_gain_reset[ch] > 0.0f)
      st->prev_gain[ch] = sd->prev_gain_reset[ch];
    last = st->prev_gain[ch];
    reci = (float)(1.0 / (last));
    if (offset > 6)
      offset = 6;
    if (e == 0)
    {
      for (t = 0; t < N; t++)
        x[ch][t] = reci * in[ch * N + t];
      continue;
    }
    if (sd->is_exception[ch])
    {
      const float m = pow2f(tb, e) * reci;
      for (t = 0; t < N; t++)
        x[ch][t] = m * in[ch * N + t];
      last = pow2f(tb, -e) * last;
    }
    else if (e > 6 || e < -6)
    {
      for (t = 0; t < N; t++)
      {
        const float v = reci * in[ch * N + t];
        x[ch][t] = (float)(pow(tb->dyn_win[t], e)) * v;
      }
      last = last * (float)(pow(tb->dyn_win[N - 1], e));
    }
    else
    {
      for (t = 0; t < N; t++)
        x[ch][t] = (in[ch * N + t] * reci) * tb->dyn_pow[offset * N + t];
      last = last / tb->dyn_pow[offset * N + N - 1];
    }
    st->prev_gain[ch] = last;
  }
  /* ---- ambience: uses the signals before clearing -------------------------------------------------- */
  memcpy(xa, x, sizeof(float) * T * N);
  for (c = 0; c < HOASP_MAX_COEF; c++)
    reassign[c] = -1;
  for (ch = 0; ch < T; ch++)
    if (sd->amb_hoa_assign[ch] != 0)
      reassign[sd->amb_hoa_assign[ch] - 1] = ch;
  /* ---- clearing of the additional coders that carry neither a vector nor a direction signal ------------ */
  for (ch = 0; ch < HOASP_MAX_CH; ch++)
    keep[ch] = 0;
  for (i = 0; i < sd->n_vec; i++)
    keep[sd->vec_index[i] - 1] = 1;
  for (i = 0; i < sd->n_dir; i++)
    keep[sd->dir_ch[i] - 1] = 1;
  for (ch = 0; ch < tb->n_addnl && ch < T; ch++)
    if (!keep[ch])
      memset(x[ch], 0, sizeof(float) * N);
  for (i = 0; i < HOASP_MAX_CH; i++)
    set[i] = 0;
  for (i = 0; i < sd->n_dir; i++)
    set[i] = sd->dir_ch[i];
  for (i = 0; i < sd->n_vec; i++)
    set[i] = sd->vec_index[i];

  /* ---- spatial prediction: faded predicted grid signals, current and last parameters ------------------- */
  {
    const float scale = pow2f(tb, 1 - tb->num_bits_per_sf);
    int r;
    for (r = 0; r < tb->pred_dir_sigs; r++)
      for (g = 0; g < nc; g++)
        fac[r][g] = (float)(scale * ((float)(sd->pred_quant[r][g]) + 0.5));
    if (tb->n_addnl != 0)
      for (k = 0; k < 2; k++)
      {
        const int *type = k ? sd->pred_type : st->last_pred_type;
        const int(*idx)[HOASP_MAX_COEF] = k ? sd->pred_idx : st->last_pred_idx;
        const float(*f)[HOASP_MAX_COEF] = k ? fac : st->last_pred_fac;
        const float *fade = k ? w_in_dir : w_out_dir;
        float(*tmp)[1024] = k ? tmpc : tmpl;
        for (g = 0; g < nc; g++)
        {
          int n = 0, s;
          if (type[g] == 0)
            continue;
          if (k)
            cur_flag = 1;
          else
            last_flag = 1;
          while (n < tb->pred_dir_sigs && idx[n][g] != 0)
            n++;
          for (t = 0; t < N; t++)
          {
            float a = 0.0f;
            for (s = 0; s < n; s++)
              a += f[s][g] * x[idx[s][g] - 1][t];
            tmp[g][t] = a * fade[t];
          }
        }
      }
  }

  for (c = 0; c < nc; c++)
  {
    float *o = out + c * N;
    int in_non = 0, in_en = 0, in_dis = 0;
    for (i = 0; i < HOASP_MAX_COEF; i++)
      if (sd->non_en_dis[i] == c + 1)
        in_non = 1;
    for (i = 0; i < sd->n_enable; i++)
      if (sd->enable[i] - 1 == c)
        in_en++;
    for (i = 0; i < sd->n_disable; i++)
      if (sd->disable[i] - 1 == c)
        in_dis++;
    for (t = 0; t < N; t++)
    {
      float v = 0.0f, amb, pred, dir = 0.0f;
      /* vector-based part */
      if (c >= tb->vec_start)
      {
        for (i = 0; i < st->num_prev_vec; i++)
        {
          const int vi = st->prev_vec_index[i];
          int it, mp, ns = N;
          const float *win = w_out_dir;
          for (it = 0; it < HOASP_MAX_CH; it++)
            if (set[it] == vi)
              break;
          if (it == HOASP_MAX_CH)
            continue;
          for (mp = 0; mp < HOASP_MAX_CH; mp++)
            if (sd->vec_index[mp] == vi)
              break;
          if (mp != HOASP_MAX_CH)
          {
            ns = tb->spat_interp_time;
            win = w_out_vec;
          }
          if (t < ns)
            v += (st->prev_vec[i][c - tb->vec_start] * win[t]) * x[vi - 1][t];
        }
        for (i = 0; i < sd->n_vec; i++)
        {
          const int vi = sd->vec_index[i];
          int it;
          float iv = sd->vec[i][c - tb->vec_start];
          for (it = 0; it < HOASP_MAX_CH; it++)
            if (st->prev_ps_set[it] == vi)
              break;
          if (it != HOASP_MAX_CH)
            iv = iv * w_in_vec[t];
          v += iv * x[vi - 1][t];
        }
      }
      if (tb->coded_v_vec_length == 1)
      {
        for (i = 0; i < in_en; i++)
          v = v * w_out_dir[t];
        for (i = 0; i < in_dis; i++)
          v = v * w_in_dir[t];
      }
      /* ambience */
      if (c >= tb->low_order)
        amb = (reassign[c] == -1) ? 0.0f : xa[reassign[c]][t];
      else
      {
        int m;
        amb = 0.0f;
        for (m = 0; m < tb->low_order; m++)
          amb = amb + tb->order_matrix[c * tb->low_order + m] * (reassign[m] < 0 ? 0.0f : xa[reassign[m]][t]);
      }
      /* spatial prediction */
      {
        float pc = 0.0f, pl = 0.0f;
        if (tb->n_addnl != 0)
          for (g = 0; g < nc; g++)
          {
            if (sd->pred_type[g] != 0)
              pc += tmpc[g][t] * tb->coarse[c * nc + g];
            if (st->last_pred_type[g] != 0)
              pl += tmpl[g][t] * tb->coarse[c * nc + g];
          }
        pred = in_non ? 0.0f : (pc + pl);
        if (cur_flag)
          for (i = 0; i < in_en; i++)
            pred = pred * w_out_dir[t];
        if (last_flag)
          for (i = 0; i < in_dis; i++)
            pred = pred * w_in_dir[t];
      }
      /* direction signals */
      for (i = 0; i < st->num_prev_dir; i++)
      {
        const int grid = st->last_dir_id[i], dch = st->last_dir_ch[i];
        int di;
        const float *win = 0;
        float cc;
        if (grid == 0)
          continue;
        for (di = 0; di < sd->n_dir; di++)
          if (sd->dir_ch[di] == dch)
            break;
        if (di == HOASP_MAX_COEF)
        {
          int q;
          for (q = 0; q < HOASP_MAX_CH; q++)
            if (set[q] == dch)
              break;
          if (q != HOASP_MAX_CH)
            win = w_out_vec;
        }
        else if (sd->dir_id[di] != 0)
          win = w_out_dir;
        cc = x[dch - 1][t] * tb->fine[(grid - 1) * HOASP_MAX_COEF + c];
        if (win)
          cc *= win[t];
        dir += cc;
      }
      for (i = 0; i < sd->n_dir; i++)
      {
        const int grid = sd->dir_id[i], dch = sd->dir_ch[i];
        int di;
        const float *win = 0;
        float cc;
        if (grid == 0)
          continue;
        for (di = 0; di < st->num_prev_dir; di++)
          if (dch == st->last_dir_ch[di])
            break;
        if (di == HOASP_MAX_COEF)
        {
          int q;
          for (q = 0; q < st->num_prev_ps; q++)
            if (dch == st->old_ps_set[q])
              break;
          if (q != HOASP_MAX_COEF)
            win = w_in_dir;
        }
        else if (st->last_dir_id[di] != 0)
          win = w_in_dir;
        cc = x[dch - 1][t] * tb->fine[(grid - 1) * HOASP_MAX_COEF + c];
        if (win)
          cc *= win[t];
        dir += cc;
      }
      o[t] = v + ((amb + pred) + dir);
    }
  }

  /* ---- state updates ------------------------------------------------------------------------------- */
  for (i = 0; i < sd->n_vec; i++)
  {
    st->prev_vec_index[i] = sd->vec_index[i];
    memcpy(st->prev_vec[i], sd->vec[i], sizeof(float) * HOASP_MAX_COEF);
    if (tb->coded_v_vec_length == 1)
      for (k = 0; k < sd->n_enable; k++)
        st->prev_vec[i][sd->enable[k] - 1 - tb->vec_start] = 0;
  }
  st->num_prev_vec = sd->n_vec;
  memcpy(st->prev_ps_set, set, sizeof(set));
  st->num_prev_ps = sd->n_vec;
  for (i = 0; i < sd->n_vec; i++)
    st->old_ps_set[i] = set[i];
  st->num_prev_dir = sd->n_dir;
  for (i = 0; i < sd->n_dir; i++)
  {
    st->last_dir_ch[i] = sd->dir_ch[i];
    st->last_dir_id[i] = sd->dir_id[i];
  }
  for (g = 0; g < nc; g++)
    st->last_pred_type[g] = sd->pred_type[g];
  {
    /* the reference copies the first n_coef * pred_dir_sigs entries of matrices with row stride n_coef */
    int r;
    for (r = 0; r < tb->pred_dir_sigs; r++)
      for (g = 0; g < nc; g++)
      {
        st->last_pred_fac[r][g] = fac[r][g];
        st->last_pred_idx[r][g] = sd->pred_idx[r][g];
      }
  }
  return 0;
}
