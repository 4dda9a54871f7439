/* seam_core.c -- libmpeghgpu.so: the reference's decoder API in front of the B200 signal path.
 *
 * Exports, unchanged, the five entry points of decoder/impeghd_api.h:47-51 (ia_mpegh_dec_get_lib_id_strings, _create,
 * _init, _execute, _delete) and, additively,
 *
 *   IA_ERRORCODE mpeghb200_dec_execute_batch(pVOID *objs, pVOID *ins, pVOID *outs, WORD32 n, IA_ERRORCODE *errs);
 *
 * which advances n decoder handles by one ia_mpegh_dec_execute each.  The bitstream half of every handle (MHAS / USAC
 * parsing, arithmetic decoding, OAM / HOA side info) is the reference's own code, linked as a shared library behind
 * this one; the post-parse signal functions it calls resolve, by ELF symbol precedence, to the definitions in
 * seam_interpose*.c, which run on the GPU through include/mpeghb200.h.
 *
 * Batching without touching the reference source: each handle's ia_mpegh_dec_execute runs on its own coroutine
 * (ucontext, 4 MB lazily committed stack).  An interposed signal function does not compute; it describes its call in a seam_req and
 * yields.  When every coroutine of the batch is parked the scheduler gives all pending requests of a kind to that
 * kind's back-end -- one gather, one upload, ONE kernel launch over the whole batch, one download, one scatter -- and
 * resumes the handles.  n handles x k interposed calls per frame thus cost k launches, not n x k.  The coroutines of a
 * batch are spread over MPEGHB200_PARSE_THREADS worker threads (default: the host's cores, at most 16), so the
 * reference's parse halves of different handles run in parallel between two GPU rounds: the parse pool of SURVEY.md
 * section 8(f) rank 4.  ia_mpegh_dec_execute itself is a batch of one.
 *
 * The reference library is found through the dynamic linker only (RTLD_NEXT): nothing here names a path.
 */
#define _GNU_SOURCE
#include <dlfcn.h>
#include <pthread.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/mman.h>
#include <time.h>
#include <ucontext.h>
#include <unistd.h>

#include <impeghd_type_def.h>
#include "impeghd_error_codes.h"
#include "impeghd_memory_standards.h"
#include "impeghd_api.h"

#include "seam_internal.h"

seam_counters seam_n;
static mpeghb200_ctx *g_ctx;
static seam_backend g_backend[RQ_KINDS];
static const char *g_backend_name[RQ_KINDS];
static unsigned long g_blocks_made, g_blocks_freed; /* device state blocks created / freed with their handles */
static double g_t_parse, g_t_backend[RQ_KINDS]; /* wall time of the batch scheduler: handles running / back-end rounds */
static pthread_mutex_t g_init_mu = PTHREAD_MUTEX_INITIALIZER;

void seam_die(const char *what, mpeghb200_status st)
{
  fprintf(stderr, "mpeghb200 seam: %s failed (0x%08x): %s -- there is no CPU fallback for the GPU signal path\n", what,
          (unsigned)st, mpeghb200_last_error(g_ctx));
  abort();
}

static void seam_stats(void)
{
  if (getenv("MPEGHB200_SEAM_STATS"))
    fprintf(stderr,
            "mpeghb200 seam: fd_frm_dec gpu=%lu host(LPD transition)=%lu tns_filters gpu=%lu limiter_frames gpu=%lu "
            "ltpf_frames gpu=%lu format_conv_frames gpu=%lu mct_boxes gpu=%lu kernel_launches=%llu "
            "obj_render gpu=%lu host(sub-frames)=%lu pcm_pack gpu=%lu batch_calls=%lu backend_rounds=%lu max_batch=%lu "
            "hoa_render gpu=%lu host=%lu resample gpu=%lu hoa_spatial gpu=%lu host=%lu "
            "binaural_blocks gpu=%lu host(calls)=%lu "
            "stereo_tools gpu=%lu host=%lu igf mono=%lu tnf=%lu stereo=%lu host=%lu "
            "drc_channels subband=%lu time_domain=%lu host(calls)=%lu stft=%lu\n",
            seam_n.fd_gpu, seam_n.fd_host, seam_n.tns, seam_n.lim, seam_n.ltpf, seam_n.fc, seam_n.mct,
            g_ctx ? (unsigned long long)mpeghb200_launch_count(g_ctx) : 0ull, seam_n.obj, seam_n.obj_host, seam_n.pcm,
            seam_n.batch_calls, seam_n.batches, seam_n.max_batch, seam_n.hoa_ren, seam_n.hoa_ren_host, seam_n.resample, seam_n.hoa_spatial,
            seam_n.hoa_spatial_host, seam_n.binaural, seam_n.binaural_host, seam_n.stereo, seam_n.stereo_host, seam_n.igf,
            seam_n.igf_tnf, seam_n.igf_stereo, seam_n.igf_host, seam_n.drc_sb, seam_n.drc_td, seam_n.drc_host,
            seam_n.drc_stft);
  if (getenv("MPEGHB200_SEAM_STATS"))
    fprintf(stderr, "mpeghb200 seam: device state blocks created=%lu freed with their handles=%lu\n", g_blocks_made,
            g_blocks_freed);
  if (getenv("MPEGHB200_SEAM_STATS") && g_t_parse > 0.0)
  {
    int k;
    fprintf(stderr, "mpeghb200 seam: batch scheduler wall time: reference code on the parse threads %.3f s", g_t_parse);
    for (k = 0; k < RQ_KINDS; k++)
      if (g_t_backend[k] > 0.0)
        fprintf(stderr, ", %s %.3f s", g_backend_name[k] ? g_backend_name[k] : "?", g_t_backend[k]);
    fprintf(stderr, "\n");
  }
}

void *seam_real(const char *name)
{
  void *f = dlsym(RTLD_NEXT, name);
  if (!f)
  {
    fprintf(stderr, "mpeghb200 seam: the reference library's %s is not visible behind this library (load the reference "
                    "decoder with global symbol scope: link against it, or dlopen it with RTLD_GLOBAL)\n", name);
    abort();
  }
  return f;
}

mpeghb200_ctx *seam_ctx(void)
{
  if (!g_ctx)
  {
    pthread_mutex_lock(&g_init_mu);
    if (!g_ctx)
    {
      const char *dev = getenv("MPEGHB200_DEVICE");
      mpeghb200_ctx *c = NULL;
      mpeghb200_status st = mpeghb200_create(&c, dev ? atoi(dev) : 0);
      if (st != MPEGHB200_OK)
      {
        fprintf(stderr, "mpeghb200 seam: mpeghb200_create failed (0x%08x): %s -- no CPU fallback\n", (unsigned)st,
                mpeghb200_last_error(NULL));
        abort();
      }
      atexit(seam_stats);
      g_ctx = c;
    }
    pthread_mutex_unlock(&g_init_mu);
  }
  return g_ctx;
}

void seam_need(seam_buf *b, size_t bytes)
{
  if (bytes <= b->cap)
    return;
  bytes = (bytes + (bytes >> 1) + 4095) & ~(size_t)4095;
  if (b->h)
    CK(mpeghb200_host_free(seam_ctx(), b->h));
  if (b->d)
    CK(mpeghb200_dev_free(seam_ctx(), b->d));
  CK(mpeghb200_host_alloc(seam_ctx(), bytes, &b->h));
  CK(mpeghb200_dev_alloc(seam_ctx(), bytes, &b->d));
  b->cap = bytes;
}

/* ---- per-handle device state ------------------------------------------------------------------------------------ */
static seam_block *pool_find(seam_pool *p, const void *key, int *idx)
{
  seam_block *b;
  int i;
  for (b = p->blocks; b; b = b->next)
    for (i = 0; i < b->n; i++)
      if (b->keys[i] == key)
      {
        *idx = i;
        return b;
      }
  return NULL;
}

/* every pool that owns slots, so that ia_mpegh_dec_delete can drop the ones of the handle that goes away */
#define MAX_POOLS 512
static seam_pool *g_pools[MAX_POOLS];
static int g_n_pools;

static seam_block *pool_new_block(seam_pool *p, const void **keys, int n)
{
  seam_block *b = (seam_block *)calloc(1, sizeof(*b));
  if (!p->registered && g_n_pools < MAX_POOLS)
    g_pools[g_n_pools++] = p, p->registered = 1;
  g_blocks_made++;
  b->n = n;
  b->keys = (const void **)malloc(sizeof(void *) * (size_t)n);
  memcpy(b->keys, keys, sizeof(void *) * (size_t)n);
  CK(mpeghb200_dev_alloc(seam_ctx(), p->stride * (size_t)n, (void **)&b->d));
  CK(mpeghb200_dev_memset(seam_ctx(), b->d, 0, p->stride * (size_t)n));
  b->next = p->blocks;
  p->blocks = b;
  return b;
}

void *seam_pool_range(seam_pool *p, const void **keys, int n, int *fresh)
{
  int idx = 0, i, known = 0;
  seam_block *b = pool_find(p, keys[0], &idx);
  *fresh = 0;
  if (b)
  {
    if (idx + n > b->n)
      return NULL;
    for (i = 1; i < n; i++)
      if (b->keys[idx + i] != keys[i])
        return NULL;
    return b->d + p->stride * (size_t)idx;
  }
  for (i = 1; i < n && !known; i++)
  {
    int j;
    known = pool_find(p, keys[i], &j) != NULL;
  }
  if (known)
    return NULL;
  *fresh = 1;
  return pool_new_block(p, keys, n)->d;
}

void *seam_pool_one(seam_pool *p, const void *key, int *fresh)
{
  int idx = 0;
  seam_block *b = pool_find(p, key, &idx);
  *fresh = 0;
  if (b)
    return b->d + p->stride * (size_t)idx;
  *fresh = 1;
  return pool_new_block(p, &key, 1)->d;
}

/* A decoder handle is deleted: the device state that was keyed by addresses inside its memory [lo, hi) is dead.  Its
 * slots lose their key (a new handle the allocator puts at the same address starts from fresh state, as the reference's
 * own create + init gives it); a block without a live slot is freed. */
static void pools_release_range(const char *lo, const char *hi)
{
  int k, i;
  for (k = 0; k < g_n_pools; k++)
  {
    seam_block **link = &g_pools[k]->blocks, *b;
    while ((b = *link) != NULL)
    {
      int live = 0;
      for (i = 0; i < b->n; i++)
      {
        const char *key = (const char *)b->keys[i];
        if (key && key >= lo && key < hi)
          b->keys[i] = NULL;
        live += b->keys[i] != NULL;
      }
      if (!live)
      {
        *link = b->next;
        if (g_ctx)
          mpeghb200_dev_free(g_ctx, b->d);
        free(b->keys);
        free(b);
        g_blocks_freed++;
      }
      else
        link = &b->next;
    }
  }
}

/* ---- coroutines and the batch scheduler ------------------------------------------------------------------------- */
#define CORO_STACK (4u << 20) /* the reference needs 0.5 - 1 MB of stack per execute call (measured with ulimit -s) */
typedef struct coro
{
  ucontext_t ctx;
  char *stack;
  int done, started;
  seam_req *pending;
  pVOID obj, in, out;
  IA_ERRORCODE ret;
} coro;

static pthread_mutex_t g_batch_mu = PTHREAD_MUTEX_INITIALIZER;  /* one batch (or inline request) at a time */
static pthread_mutex_t g_inline_mu = PTHREAD_MUTEX_INITIALIZER; /* stages that still compute inside the interposer */
static __thread coro *tl_cur;          /* coroutine running on this thread, NULL outside a batch */
static __thread ucontext_t tl_sched;   /* where a parked coroutine returns to */

void seam_register(int kind, seam_backend fn, const char *name)
{
  g_backend[kind] = fn;
  g_backend_name[kind] = name;
}

void seam_submit(seam_req *r)
{
  r->status = 0;
  if (!tl_cur)
  { /* not inside a batch (e.g. a frame decoded by ia_mpegh_dec_init): a batch of one, right here */
    seam_req *one = r;
    pthread_mutex_lock(&g_batch_mu);
    g_backend[r->kind](&one, 1);
    seam_n.batches++;
    pthread_mutex_unlock(&g_batch_mu);
    return;
  }
  tl_cur->pending = r;
  swapcontext(&tl_cur->ctx, &tl_sched);
}

typedef IA_ERRORCODE (*api3_fn)(pVOID, pVOID, pVOID);
typedef IA_ERRORCODE (*api2_fn)(pVOID, pVOID);
typedef IA_ERRORCODE (*api1_fn)(pVOID);
static void *next_sym(const char *name)
{
  void *p = dlsym(RTLD_NEXT, name);
  if (!p)
  {
    fprintf(stderr, "mpeghb200 seam: the reference decoder library behind libmpeghgpu.so does not export %s\n", name);
    abort();
  }
  return p;
}

static void coro_entry(unsigned lo, unsigned hi)
{
  coro *c = (coro *)(((uintptr_t)hi << 32) | (uintptr_t)lo);
  static api3_fn real;
  if (!real)
    real = (api3_fn)next_sym("ia_mpegh_dec_execute");
  c->ret = real(c->obj, c->in, c->out);
  c->done = 1;
  /* back to the scheduler of the thread this coroutine runs on (a coroutine never changes threads); makecontext froze
   * uc_link at creation time on the creating thread, so the hand-over is explicit and this function never returns */
  swapcontext(&c->ctx, &tl_sched);
  abort();
}

/* worker pool: thread w owns coroutines w, w + W, w + 2W, ... of the current batch */
typedef struct
{
  pthread_t th;
  int idx;
} worker;
static struct
{
  int n_workers, started;
  worker *w;
  pthread_barrier_t go, done;
  coro *batch;
  int n;
  volatile int quit;
  seam_par_fn job; /* non-NULL: the round is a parallel loop of a back-end, not handles to run */
  void *job_arg;
  int job_n;
} g_pool;
static int g_in_round; /* the scheduler is serving back-ends: the workers sit at the `go` barrier */
static char **g_stacks;
static int g_n_stacks;

static double now_s(void)
{
  struct timespec ts;
  clock_gettime(CLOCK_MONOTONIC, &ts);
  return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec;
}

static void run_slice(coro *batch, int n, int first, int step)
{
  int i;
  for (i = first; i < n; i += step)
  {
    coro *c = &batch[i];
    if (c->done || c->pending)
      continue;
    tl_cur = c;
    swapcontext(&tl_sched, &c->ctx);
    tl_cur = NULL;
  }
}

static void *worker_main(void *arg)
{
  worker *w = (worker *)arg;
  for (;;)
  {
    pthread_barrier_wait(&g_pool.go);
    if (g_pool.quit)
      break;
    if (g_pool.job)
    {
      int i;
      for (i = w->idx + 1; i < g_pool.job_n; i += g_pool.n_workers + 1)
        g_pool.job(i, g_pool.job_arg);
    }
    else
      run_slice(g_pool.batch, g_pool.n, w->idx + 1, g_pool.n_workers + 1);
    pthread_barrier_wait(&g_pool.done);
  }
  return NULL;
}

static void pool_start(void)
{
  const char *e = getenv("MPEGHB200_PARSE_THREADS");
  long n = e ? atol(e) : sysconf(_SC_NPROCESSORS_ONLN);
  int i;
  if (n > 16)
    n = 16;
  if (n < 1)
    n = 1;
  g_pool.n_workers = (int)n - 1; /* the calling thread is worker 0 */
  g_pool.started = 1;
  if (g_pool.n_workers == 0)
    return;
  g_pool.w = (worker *)calloc((size_t)g_pool.n_workers, sizeof(worker));
  pthread_barrier_init(&g_pool.go, NULL, (unsigned)g_pool.n_workers + 1);
  pthread_barrier_init(&g_pool.done, NULL, (unsigned)g_pool.n_workers + 1);
  for (i = 0; i < g_pool.n_workers; i++)
  {
    g_pool.w[i].idx = i;
    pthread_create(&g_pool.w[i].th, NULL, worker_main, &g_pool.w[i]);
  }
}

/* The gather / scatter loops of a back-end, spread over the parse threads (idle while the scheduler serves a round).
 * Outside a batch round (a single handle, an interposed call made from outside ia_mpegh_dec_execute) it is a plain loop. */
void seam_parallel_for(int n, seam_par_fn fn, void *arg)
{
  int i;
  if (!g_in_round || g_pool.n_workers == 0 || n < 2 * (g_pool.n_workers + 1))
  {
    for (i = 0; i < n; i++)
      fn(i, arg);
    return;
  }
  g_pool.job = fn, g_pool.job_arg = arg, g_pool.job_n = n;
  pthread_barrier_wait(&g_pool.go);
  for (i = 0; i < n; i += g_pool.n_workers + 1)
    fn(i, arg);
  pthread_barrier_wait(&g_pool.done);
  g_pool.job = NULL;
}

/* The format converter and the MCT boxes are not request-ised yet: they compute inside the interposer with shared
 * staging buffers, possibly on several parse threads at once -> one at a time. */
void seam_inline_lock(void) { pthread_mutex_lock(&g_inline_mu); }
void seam_inline_unlock(void) { pthread_mutex_unlock(&g_inline_mu); }

IA_ERRORCODE mpeghb200_dec_execute_batch(pVOID *objs, pVOID *ins, pVOID *outs, WORD32 n, IA_ERRORCODE *errs)
{
  coro *batch;
  seam_req **reqs;
  IA_ERRORCODE worst = IA_MPEGH_DEC_NO_ERROR;
  int i, k, remaining;
  if (n <= 0 || !objs || !ins || !outs)
    return IA_MPEGH_DEC_NO_ERROR;
  if (tl_cur)
  {
    fprintf(stderr, "mpeghb200 seam: mpeghb200_dec_execute_batch called from inside a decoder callback\n");
    abort();
  }
  pthread_mutex_lock(&g_batch_mu); /* one batch at a time per process: the back-ends share staging buffers */
  if (!g_pool.started)
    pool_start();
  if (n > g_n_stacks)
  {
    g_stacks = (char **)realloc(g_stacks, sizeof(char *) * (size_t)n);
    for (i = g_n_stacks; i < n; i++)
    {
      g_stacks[i] = (char *)mmap(NULL, CORO_STACK, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS | MAP_STACK, -1, 0);
      if (g_stacks[i] == MAP_FAILED)
      {
        perror("mpeghb200 seam: mmap of a coroutine stack");
        abort();
      }
      mprotect(g_stacks[i], 4096, PROT_NONE); /* guard page: an overflow faults instead of corrupting a neighbour */
    }
    g_n_stacks = n;
  }
  batch = (coro *)calloc((size_t)n, sizeof(coro));
  reqs = (seam_req **)malloc(sizeof(seam_req *) * (size_t)n);
  for (i = 0; i < n; i++)
  {
    coro *c = &batch[i];
    const uintptr_t p = (uintptr_t)c;
    c->obj = objs[i], c->in = ins[i], c->out = outs[i];
    getcontext(&c->ctx);
    c->ctx.uc_stack.ss_sp = g_stacks[i];
    c->ctx.uc_stack.ss_size = CORO_STACK;
    c->ctx.uc_link = NULL;
    makecontext(&c->ctx, (void (*)(void))coro_entry, 2, (unsigned)(p & 0xffffffffu), (unsigned)(p >> 32));
  }
  seam_n.batch_calls++;
  if ((unsigned long)n > seam_n.max_batch)
    seam_n.max_batch = (unsigned long)n;
  for (;;)
  {
    const int use_workers = g_pool.n_workers > 0 && n > 1;
    double t0 = now_s(), t1;
    /* run every coroutine that is neither finished nor parked until it parks or finishes */
    if (use_workers)
    {
      g_pool.batch = batch, g_pool.n = n;
      pthread_barrier_wait(&g_pool.go);
      run_slice(batch, n, 0, g_pool.n_workers + 1);
      pthread_barrier_wait(&g_pool.done);
    }
    else
    {
      run_slice(batch, n, 0, 1);
    }
    g_t_parse += now_s() - t0;
    remaining = 0;
    for (i = 0; i < n; i++)
      remaining += !batch[i].done;
    if (!remaining)
      break;
    /* every unfinished handle is parked on a request: serve them kind by kind, one back-end call per kind */
    for (k = 0; k < RQ_KINDS; k++)
    {
      int m = 0;
      for (i = 0; i < n; i++)
        if (!batch[i].done && batch[i].pending && batch[i].pending->kind == k)
          reqs[m++] = batch[i].pending;
      if (!m)
        continue;
      t1 = now_s();
      g_in_round = use_workers;
      g_backend[k](reqs, m);
      g_in_round = 0;
      g_t_backend[k] += now_s() - t1;
      seam_n.batches++;
      for (i = 0; i < n; i++)
        if (!batch[i].done && batch[i].pending && batch[i].pending->kind == k)
          batch[i].pending = NULL;
    }
  }
  for (i = 0; i < n; i++)
  {
    if (errs)
      errs[i] = batch[i].ret;
    if (batch[i].ret != IA_MPEGH_DEC_NO_ERROR && worst == IA_MPEGH_DEC_NO_ERROR)
      worst = batch[i].ret;
  }
  free(reqs);
  free(batch);
  pthread_mutex_unlock(&g_batch_mu);
  return worst;
}

/* ---- the reference API, unchanged signatures (decoder/impeghd_api.h:47-51) -------------------------------------- */
IA_ERRORCODE ia_mpegh_dec_get_lib_id_strings(pVOID pv_output)
{
  static api1_fn real;
  if (!real)
    real = (api1_fn)next_sym("ia_mpegh_dec_get_lib_id_strings");
  return real(pv_output);
}

IA_ERRORCODE ia_mpegh_dec_create(pVOID pv_input, pVOID pv_output)
{
  static api2_fn real;
  if (!real)
    real = (api2_fn)next_sym("ia_mpegh_dec_create");
  seam_ctx(); /* fail at create time, loudly, when there is no GPU */
  return real(pv_input, pv_output);
}

void seam_release_range(const void *lo, const void *hi)
{ /* back-ends (the other users of the pools) only run while every handle of the batch is parked; the handles of a
     batch run side by side, hence the lock */
  seam_inline_lock();
  pools_release_range((const char *)lo, (const char *)hi);
  seam_inline_unlock();
}

static void release_handle_state(const ia_output_config *oc)
{
  int i;
  pthread_mutex_lock(&g_batch_mu);
  for (i = 0; oc && i < NUM_MEMTABS; i++)
    if (oc->mem_info_table[i].mem_ptr && oc->mem_info_table[i].ui_size)
      pools_release_range((const char *)oc->mem_info_table[i].mem_ptr,
                          (const char *)oc->mem_info_table[i].mem_ptr + oc->mem_info_table[i].ui_size);
  pthread_mutex_unlock(&g_batch_mu);
}

IA_ERRORCODE ia_mpegh_dec_init(pVOID p_ia_module_obj, pVOID pv_input, pVOID pv_output)
{
  static api3_fn real;
  if (!real)
    real = (api3_fn)next_sym("ia_mpegh_dec_init");
  /* (re-)initialisation: the reference sets up its persistent structs again (limiter, renderers, converter ...), so
   * what the device holds for this handle is the past; the frames init itself decodes start from fresh state */
  release_handle_state((const ia_output_config *)pv_output);
  return real(p_ia_module_obj, pv_input, pv_output);
}

IA_ERRORCODE ia_mpegh_dec_execute(pVOID p_ia_module_obj, pVOID pv_input, pVOID pv_output)
{
  IA_ERRORCODE err = IA_MPEGH_DEC_NO_ERROR;
  mpeghb200_dec_execute_batch(&p_ia_module_obj, &pv_input, &pv_output, 1, &err);
  return err;
}

IA_ERRORCODE ia_mpegh_dec_delete(pVOID pv_output)
{
  static api1_fn real;
  const ia_output_config *oc = (const ia_output_config *)pv_output;
  if (!real)
    real = (api1_fn)next_sym("ia_mpegh_dec_delete");
  /* the handle's memory blocks (decoder/impeghd_api.c:341-392): every persistent struct of the handle lives in them;
   * the table itself sits in the caller's ia_output_config and survives the delete */
  release_handle_state(oc);
  return real(pv_output);
}
