/* seam_internal.h -- what the translation units of libmpeghgpu.so / libmpeghb200_seam.so share.
 *
 * Every interposed signal function turns its call into a REQUEST and hands it to seam_submit().  Outside a batch
 * the request executes at once (a batch of one).  Inside mpeghb200_dec_execute_batch every decoder handle runs the
 * reference's ia_mpegh_dec_execute on its own coroutine; seam_submit() parks the coroutine, and once every handle
 * of the batch is parked (or finished) the scheduler hands all pending requests of one kind to that kind's
 * BACK-END in one call: gather from the handles -> one upload -> ONE kernel launch over all of them -> one download
 * -> scatter.  The handles then continue to their next interposed call.  (seam_core.c)
 */
#ifndef MPEGHB200_SEAM_INTERNAL_H
#define MPEGHB200_SEAM_INTERNAL_H
#include <stddef.h>
#include "mpeghb200.h"

mpeghb200_ctx *seam_ctx(void); /* creates the context on first use; never returns without one */
void seam_die(const char *what, mpeghb200_status st);
#define CK(call)                                                                                   \
  do                                                                                               \
  {                                                                                                \
    mpeghb200_status st_ = (call);                                                                 \
    if (st_ != MPEGHB200_OK)                                                                       \
      seam_die(#call, st_);                                                                        \
  } while (0)

/* ---- requests ---- */
typedef enum
{
  RQ_FD = 0, /* ia_core_coder_fd_frm_dec */
  RQ_TNS,    /* ia_core_coder_tns_apply */
  RQ_LTPF,   /* ia_core_coder_usac_ltpf_process */
  RQ_LIM,    /* impeghd_peak_limiter_process */
  RQ_OBJ,    /* impeghd_obj_renderer_dec */
  RQ_PCM,    /* ia_core_coder_samples_sat */
  RQ_HOAREN, /* impeghd_hoa_ren_block_process */
  RQ_RESAMPLE, /* impeghd_resample */
  RQ_HOASP,    /* the four signal stages of impeghd_hoa_spatial_process */
  RQ_BINAURAL, /* impeghd_binaural_struct_process_ld_ola, one request per completed block */
  RQ_STEREO,   /* impeghd_ms_stereo, ia_core_coder_cplx_pred_upmixing */
  RQ_IGF,      /* ia_core_coder_igf_mono (+ its TNF step), ia_core_coder_igf_tnf */
  RQ_IGFST,    /* ia_core_coder_igf_apply_stereo */
  RQ_DS,       /* impeghd_domain_switcher_process (time <-> STFT256 around sub-band DRC) */
  RQ_DRCSB,    /* impd_drc_apply_gains_subband */
  RQ_DRCTD,    /* impd_drc_filter_banks_process + impd_drc_apply_gains_and_add */
  RQ_KINDS
} seam_kind;

typedef struct seam_req
{
  int kind;
  void *p[6];  /* pointers into the calling handle (meaning per kind) */
  long v[6];   /* scalars */
  int status;  /* back-end result handed back to the interposer (reference error code) */
} seam_req;

typedef void (*seam_backend)(seam_req **reqs, int n);
void seam_register(int kind, seam_backend fn, const char *name);
void seam_submit(seam_req *r);
void seam_inline_lock(void);
void seam_inline_unlock(void);

/* ---- counters (MPEGHB200_SEAM_STATS=1 prints them at exit) ---- */
typedef struct
{
  unsigned long fd_gpu, fd_host, tns, lim, ltpf, fc, mct, obj, obj_host, pcm, batches, batch_calls, max_batch, hoa_ren,
      hoa_ren_host, resample, hoa_spatial, hoa_spatial_host, binaural, binaural_host, stereo, stereo_host, igf, igf_tnf, igf_stereo, igf_host, drc_sb, drc_td, drc_host, drc_stft;
} seam_counters;
extern seam_counters seam_n;

/* the reference's own definition of an interposed function (dlsym RTLD_NEXT); aborts with a message when the reference
 * library is not in the global symbol scope (a dlopen without RTLD_GLOBAL) instead of calling through NULL */
void *seam_real(const char *name);

/* gather / scatter loops of a back-end on all parse threads: fn(i, arg) for i in [0, n), any order, any thread */
typedef void (*seam_par_fn)(int i, void *arg);
void seam_parallel_for(int n, seam_par_fn fn, void *arg);

/* ---- staging buffers that grow on demand: page-locked host + device of the same size ---- */
typedef struct
{
  void *h, *d;
  size_t cap;
} seam_buf;
void seam_need(seam_buf *b, size_t bytes);

/* ---- device-resident per-handle state, keyed by a pointer inside the reference handle ----
 * Slots live in blocks of consecutive records so that a batch whose keys sit in consecutive slots (the normal case:
 * the same handle array batch after batch) is ONE launch over `n` streams at seam_pool_run's base pointer.  Keys
 * never seen before get a fresh block in batch order; a batch that mixes blocks falls back to one launch per key. */
typedef struct seam_block
{
  struct seam_block *next;
  char *d;
  int n;
  const void **keys;
} seam_block;
typedef struct
{
  seam_block *blocks;
  size_t stride; /* bytes per slot */
  int registered; /* known to the core, which drops a deleted handle's slots (ia_mpegh_dec_delete) */
} seam_pool;
/* Finds (or creates: *fresh = 1) consecutive slots for keys[0..n).  Returns the device pointer of slot 0 when the n
 * keys are consecutive in one block, NULL otherwise (use seam_pool_one per key then). */
void *seam_pool_range(seam_pool *p, const void **keys, int n, int *fresh);
void *seam_pool_one(seam_pool *p, const void *key, int *fresh);
/* The reference has re-initialised everything inside [lo, hi) in place (ia_core_coder_decoder_flush_api, from inside an
 * execute call): slots keyed by addresses in that range start over.  Callable from a handle's coroutine. */
void seam_release_range(const void *lo, const void *hi);
#endif
