/* oracle/ref_chain.c -- TEST INFRASTRUCTURE ONLY (linked into oracle/_ref/libmpeghref_ns.so, never into the product).
 *
 * The UNMODIFIED reference's stage functions composed into the chains of BASELINE.json's configurations, in the order
 * of decoder/ia_core_coder_decode_main.c:2736-2893, over many independent streams on a pool of persistent, pinned
 * worker threads ("pinned persistent C workers", one disjoint block of streams each):
 *
 *   kind 0  FD synthesis of n_in channels (ia_core_coder_fd_frm_dec)                       [+ limiter + PCM packer]
 *   kind 1  HOA: impeghd_hoa_spatial_process stages -> impeghd_hoa_ren_block_process (NFC)   + limiter + PCM packer
 *   kind 2  binaural: impeghd_binaural_struct_process_ld_ola (x 1/32768 in, x 32768 out)     + PCM packer
 *   kind 3  objects: impeghd_obj_renderer_dec                                                + limiter + PCM packer
 *
 * Two uses: (1) the checker of the chain parity tests (tests/test_oracle_chain.py pins tests/chain_cases.py to it,
 * the GPU chains are then compared with the PCM bytes); (2) bench.py's cpu_baseline / --impl reference legs, timed
 * inside ref_chain_run with clock_gettime around the start / done barriers (thread start-up is outside).
 * Contains no reference code: only calls into the harness entry points (ref_harness*.c), which call the reference.
 */
#define _GNU_SOURCE
#include <pthread.h>
#include <sched.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

/* harness entry points (ref_harness.c, _hoasp.c, _hoa.c, _obj.c, _bin.c, _ns.c) */
void *ref_fd_create(void);
void ref_fd_destroy(void *p);
int ref_fd_frm_dec(void *p, const float *coef, float *overlap, float *out, int window_sequence, int window_shape,
                   int window_shape_prev, int prev_aliasing_symmetry, int curr_aliasing_symmetry);
void *ref_limiter_create(int n_ch, int sample_rate, int limiter_on);
void ref_limiter_destroy(void *p);
void ref_limiter_process(void *p, float *planar, int n);
void *ref_hoasp_create(const int *c, int *err_out);
void ref_hoasp_destroy(void *p);
int ref_hoasp_process(void *p, const void *sd, const float *in, float *out);
void *ref_hoa_ren_create(int cicp, int order, int lfe_enabled, float nfc_radius, int sample_rate, int *err_out);
void ref_hoa_ren_destroy(void *p);
int ref_hoa_ren_process(void *p, const float *in, int rows, int n, int clip, float *out);
void *ref_obj_create(int cicp_idx, int *err_out);
void ref_obj_destroy(void *p);
int ref_obj_render(void *p, const float *params, int n_obj, const float *in, float *out, float *gains_out);
void *ref_bin_create(int n_in, int len, int n_blocks, const float *taps_direct, const float *taps_diffuse,
                     const float *direct_fc, const float *diffuse_fc, const float *inv_diffuse_weight, int via_init,
                     int *err_out);
void ref_bin_destroy(void *p);
int ref_bin_process(void *p, const float *in, int n, float *out, int out_stride);
int ref_samples_sat(const float *planar, int n, int n_ch, int pcmsize, const int *out_ch_map, int *delay_samples,
                    unsigned char *out);

#define FRAME 1024
#define MAX_CH 64

typedef struct ref_chain ref_chain;
typedef struct
{
  ref_chain *c;
  int idx;
  float *buf_a; /* [MAX_CH][4096] */
  float *buf_b;
  unsigned char *pcm;
} ref_worker;

struct ref_chain
{
  int kind, S, n_in, n_out, pcm_bits, tail, side_bytes, fpr, rec_bytes, n_coef, n_threads;
  void **a, **b, **lim;
  int *delay;
  float *overlap; /* kind 0: [S][n_in][1024] */
  /* job */
  const float *in;
  const unsigned char *side;
  unsigned char *pcm;
  int *pcm_bytes;
  int n_frames, quit;
  volatile int err;
  pthread_t *th;
  ref_worker *w;
  pthread_barrier_t start, done;
};

static void one_stream(ref_chain *c, ref_worker *w, int s)
{
  const int S = c->S, n_in = c->n_in;
  int map[MAX_CH], f, ch, i;
  for (ch = 0; ch < MAX_CH; ch++)
    map[ch] = -1;
  for (f = 0; f < c->n_frames; f++)
  {
    const float *x = c->in + ((size_t)f * S + s) * n_in * FRAME;
    const unsigned char *sd = c->side ? c->side + ((size_t)f * S + s) * c->side_bytes : 0;
    float *y = w->buf_a;
    int n = FRAME, n_ch = c->n_out, nb, rec = f / c->fpr;
    switch (c->kind)
    {
    case 0:
      for (ch = 0; ch < n_in; ch++)
      {
        const int *q = (const int *)sd + ch * 5;
        if (ref_fd_frm_dec(c->a[s], x + (size_t)ch * FRAME, c->overlap + ((size_t)s * n_in + ch) * FRAME,
                           y + (size_t)ch * FRAME, q[0], q[1], q[2], q[3], q[4]))
          c->err = 1;
      }
      break;
    case 1:
      if (ref_hoasp_process(c->a[s], sd, x, w->buf_b))
        c->err = 1;
      if (ref_hoa_ren_process(c->b[s], w->buf_b, c->n_coef, FRAME, 1, y))
        c->err = 1;
      break;
    case 2:
      for (i = 0; i < n_in * FRAME; i++) /* ia_div_q15_flt (decode_main.c:2833) */
        w->buf_b[i] = x[i] / 32768.0f;
      n = ref_bin_process(c->a[s], w->buf_b, FRAME, y, 4096 * 2);
      for (ch = 0; ch < 2 && n > 0; ch++)
      {
        float *p = y + (size_t)ch * 4096 * 2;
        for (i = 0; i < n; i++)
          p[i] = p[i] * 32768.0f;
        if (ch) /* pack rows contiguously: [2][n] */
          memmove(y + n, p, sizeof(float) * n);
      }
      break;
    default:
      if (ref_obj_render(c->a[s], (const float *)sd, n_in, x, y, 0))
        c->err = 1;
      break;
    }
    if (!c->tail || n <= 0)
      continue;
    if (c->lim)
      ref_limiter_process(c->lim[s], y, n);
    {
      unsigned char *dst = c->pcm ? c->pcm + ((size_t)rec * S + s) * c->rec_bytes : w->pcm;
      nb = ref_samples_sat(y, n, n_ch, c->pcm_bits, map, &c->delay[s], dst);
      if (c->pcm_bytes)
        c->pcm_bytes[(size_t)rec * S + s] = nb;
    }
  }
}

static void *worker_main(void *arg)
{
  ref_worker *w = (ref_worker *)arg;
  ref_chain *c = w->c;
  for (;;)
  {
    int lo, hi, s;
    pthread_barrier_wait(&c->start);
    if (c->quit)
      break;
    lo = (int)((long long)c->S * w->idx / c->n_threads);
    hi = (int)((long long)c->S * (w->idx + 1) / c->n_threads);
    for (s = lo; s < hi; s++)
      one_stream(c, w, s);
    pthread_barrier_wait(&c->done);
  }
  return 0;
}

static int start_pool(ref_chain *c, int n_threads)
{
  cpu_set_t allowed;
  int cpus[1024], n_cpu = 0, i;
  CPU_ZERO(&allowed);
  if (sched_getaffinity(0, sizeof(allowed), &allowed) == 0)
    for (i = 0; i < 1024 && i < CPU_SETSIZE; i++)
      if (CPU_ISSET(i, &allowed))
        cpus[n_cpu++] = i;
  if (n_threads < 1)
    n_threads = 1;
  if (n_threads > c->S)
    n_threads = c->S;
  c->n_threads = n_threads;
  c->th = (pthread_t *)calloc(n_threads, sizeof(pthread_t));
  c->w = (ref_worker *)calloc(n_threads, sizeof(ref_worker));
  pthread_barrier_init(&c->start, 0, n_threads + 1);
  pthread_barrier_init(&c->done, 0, n_threads + 1);
  for (i = 0; i < n_threads; i++)
  {
    pthread_attr_t at;
    c->w[i].c = c;
    c->w[i].idx = i;
    c->w[i].buf_a = (float *)calloc((size_t)MAX_CH * 4096, sizeof(float));
    c->w[i].buf_b = (float *)calloc((size_t)MAX_CH * 4096, sizeof(float));
    c->w[i].pcm = (unsigned char *)calloc((size_t)MAX_CH * 4096, 4);
    pthread_attr_init(&at);
    pthread_attr_setstacksize(&at, 64 << 20); /* the reference keeps large frames on the stack */
    if (n_cpu > 0)
    {
      cpu_set_t one;
      CPU_ZERO(&one);
      CPU_SET(cpus[i % n_cpu], &one);
      pthread_attr_setaffinity_np(&at, sizeof(one), &one);
    }
    if (pthread_create(&c->th[i], &at, worker_main, &c->w[i]))
      return -1;
    pthread_attr_destroy(&at);
  }
  return 0;
}

static ref_chain *chain_new(int kind, int S, int n_in, int n_out, int pcm_bits, int tail, int limiter, int side_bytes)
{
  ref_chain *c = (ref_chain *)calloc(1, sizeof(ref_chain));
  int s;
  c->kind = kind, c->S = S, c->n_in = n_in, c->n_out = n_out, c->pcm_bits = pcm_bits, c->tail = tail;
  c->side_bytes = side_bytes, c->fpr = 1;
  c->rec_bytes = FRAME * n_out * (pcm_bits / 8);
  c->a = (void **)calloc(S, sizeof(void *));
  c->b = (void **)calloc(S, sizeof(void *));
  c->delay = (int *)calloc(S, sizeof(int));
  if (tail && limiter)
  {
    c->lim = (void **)calloc(S, sizeof(void *));
    for (s = 0; s < S; s++)
      c->lim[s] = ref_limiter_create(n_out, 48000, limiter == 1);
  }
  return c;
}

/* ---- constructors: one per configuration; every stream gets its own reference state ---------------------------- */
void *ref_chain_create_fd(int S, int n_ch, int tail, int pcm_bits, int start_delay, int n_threads)
{
  ref_chain *c = chain_new(0, S, n_ch, n_ch, pcm_bits, tail, 1, n_ch * 5 * (int)sizeof(int));
  int s;
  c->overlap = (float *)calloc((size_t)S * n_ch * FRAME, sizeof(float));
  for (s = 0; s < S; s++)
    c->a[s] = ref_fd_create(), c->delay[s] = start_delay;
  return start_pool(c, n_threads) ? 0 : c;
}

void *ref_chain_create_hoa(int S, const int *hoasp_cfg12, int side_bytes, int cicp, int lfe_enabled, float nfc_radius,
                           int pcm_bits, int start_delay, int n_threads, int *err_out)
{
  int order = hoasp_cfg12[0], s, err = 0, e;
  ref_chain *c = chain_new(1, S, hoasp_cfg12[1], 0, pcm_bits, 1, 1, side_bytes);
  c->n_coef = (order + 1) * (order + 1);
  for (s = 0; s < S; s++)
  {
    c->a[s] = ref_hoasp_create(hoasp_cfg12, &e), err |= e;
    c->b[s] = ref_hoa_ren_create(cicp, order, lfe_enabled, nfc_radius, 48000, &e), err |= e;
    c->delay[s] = start_delay;
  }
  if (err_out)
    *err_out = err;
  return c; /* n_out is set by ref_chain_set_out (the render matrix decides it) before the first run */
}

void *ref_chain_create_obj(int S, int n_obj, int cicp, int n_spk, int pcm_bits, int start_delay, int n_threads,
                           int *err_out)
{
  ref_chain *c = chain_new(3, S, n_obj, n_spk, pcm_bits, 1, 1, n_obj * 7 * (int)sizeof(float));
  int s, err = 0, e;
  for (s = 0; s < S; s++)
    c->a[s] = ref_obj_create(cicp, &e), err |= e, c->delay[s] = start_delay;
  if (err_out)
    *err_out = err;
  return start_pool(c, n_threads) ? 0 : c;
}

/* filters of stream s: set_of_stream[s] (or 0) into taps_* laid out [set][...] as ref_bin_create takes one set */
void *ref_chain_create_bin(int S, int n_in, int len, int n_blocks, int n_sets, const int *set_of_stream,
                           const float *taps_direct, const float *taps_diffuse, const float *direct_fc,
                           const float *diffuse_fc, const float *weight, int pcm_bits, int start_delay, int n_threads,
                           int *err_out)
{
  ref_chain *c = chain_new(2, S, n_in, 2, pcm_bits, 1, 0, 0);
  int s, err = 0, e, fft = 2;
  while (fft < 2 * len)
    fft <<= 1;
  c->fpr = fft / 2 / FRAME;
  if (c->fpr < 1)
    c->fpr = 1;
  c->rec_bytes = (fft / 2 < FRAME ? FRAME : fft / 2) * 2 * (pcm_bits / 8);
  for (s = 0; s < S; s++)
  {
    const int k = set_of_stream ? set_of_stream[s] % n_sets : 0;
    c->a[s] = ref_bin_create(n_in, len, n_blocks, taps_direct + (size_t)k * 2 * n_in * len,
                             taps_diffuse + (size_t)k * n_blocks * 2 * len, direct_fc + (size_t)k * 2 * n_in,
                             diffuse_fc + (size_t)k * 2 * n_blocks, weight + (size_t)k * n_in, 0, &e);
    err |= e;
    c->delay[s] = start_delay;
  }
  if (err_out)
    *err_out = err;
  return start_pool(c, n_threads) ? 0 : c;
}

/* HOA: the loudspeaker count comes from the designed matrix; call once after ref_chain_create_hoa */
int ref_chain_set_out(void *p, int n_out, int n_threads)
{
  ref_chain *c = (ref_chain *)p;
  int s;
  c->n_out = n_out;
  c->rec_bytes = FRAME * n_out * (c->pcm_bits / 8);
  if (c->lim)
    for (s = 0; s < c->S; s++)
    {
      ref_limiter_destroy(c->lim[s]);
      c->lim[s] = ref_limiter_create(n_out, 48000, 1);
    }
  return start_pool(c, n_threads);
}

void *ref_chain_stream_handle(void *p, int s, int which) { return which ? ((ref_chain *)p)->b[s] : ((ref_chain *)p)->a[s]; }
int ref_chain_threads(void *p) { return ((ref_chain *)p)->n_threads; }
int ref_chain_rec_bytes(void *p) { return ((ref_chain *)p)->rec_bytes; }
int ref_chain_frames_per_record(void *p) { return ((ref_chain *)p)->fpr; }

/* n_frames consecutive frames for every stream: in [frame][stream][n_in][1024], side [frame][stream][side_bytes] (or
 * NULL), pcm [record][stream][rec_bytes] and pcm_bytes [record][stream] (either may be NULL: the PCM then goes to
 * per-thread scratch).  Returns the wall time of the parallel region in seconds, < 0 on a reference error. */
double ref_chain_run(void *p, const float *in, const void *side, int n_frames, unsigned char *pcm, int *pcm_bytes)
{
  ref_chain *c = (ref_chain *)p;
  struct timespec t0, t1;
  c->in = in, c->side = (const unsigned char *)side, c->pcm = pcm, c->pcm_bytes = pcm_bytes, c->n_frames = n_frames;
  c->err = 0;
  clock_gettime(CLOCK_MONOTONIC, &t0);
  pthread_barrier_wait(&c->start);
  pthread_barrier_wait(&c->done);
  clock_gettime(CLOCK_MONOTONIC, &t1);
  if (c->err)
    return -1.0;
  return (double)(t1.tv_sec - t0.tv_sec) + 1e-9 * (double)(t1.tv_nsec - t0.tv_nsec);
}

void ref_chain_destroy(void *p)
{
  ref_chain *c = (ref_chain *)p;
  int i, s;
  if (!c)
    return;
  if (c->th)
  {
    c->quit = 1;
    pthread_barrier_wait(&c->start);
    for (i = 0; i < c->n_threads; i++)
    {
      pthread_join(c->th[i], 0);
      free(c->w[i].buf_a), free(c->w[i].buf_b), free(c->w[i].pcm);
    }
    pthread_barrier_destroy(&c->start), pthread_barrier_destroy(&c->done);
    free(c->th), free(c->w);
  }
  for (s = 0; s < c->S; s++)
  {
    switch (c->kind)
    {
    case 0: ref_fd_destroy(c->a[s]); break;
    case 1: ref_hoasp_destroy(c->a[s]), ref_hoa_ren_destroy(c->b[s]); break;
    case 2: ref_bin_destroy(c->a[s]); break;
    default: ref_obj_destroy(c->a[s]); break;
    }
    if (c->lim)
      ref_limiter_destroy(c->lim[s]);
  }
  free(c->a), free(c->b), free(c->lim), free(c->delay), free(c->overlap), free(c);
}
