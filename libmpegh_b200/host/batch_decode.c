/* batch_decode.c -- mpeghb200_batch_decode: the host feed path around libmpeghgpu.so.
 *
 * Decodes N MPEG-H streams (MHAS or MP4 files, each opened `-n` times) side by side through the reference's own API
 * (decoder/impeghd_api.h:47-51; call sequence of test/impeghd_main.c:552-823: create -> init until ui_init_done ->
 * execute until the input is exhausted -> delete), one decoder handle per stream, every frame of all handles advanced
 * by ONE mpeghb200_dec_execute_batch call.  The demultiplexer is the reference's (test/mp4/impeghd_mp4_file_wrapper.c,
 * compiled in place by the Makefile; raw MHAS and MP4 both go through impeghd_mp4_fw_read).  Per handle the PCM is
 * hashed (FNV-1a 64) and, for the first `-write:K` handles, written as a WAV file with the reference's header layout
 * (test/impeghd_main.c:227-248), so that a test can compare every handle with the reference's own output without
 * gigabytes of files.
 *
 *   mpeghb200_batch_decode [-n:<copies>] [-batch:0|1|2] [-cycles:<K>] [-write:<K>] [-cicp:<idx>] [-pcmsz:<bits>] -o:<prefix> in1 [in2 ..]
 *
 * -batch:0 advances the handles one ia_mpegh_dec_execute at a time instead (same GPU stages, a launch per handle).
 * Prints one JSON line: handles, frames, wall seconds, audio-seconds per wall-second, per-handle hashes.
 */
#include <pthread.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#include <impeghd_type_def.h>
#include "impeghd_memory_standards.h"
#include "impeghd_api.h"
#include "impeghd_error_codes.h"
#include "impeghd_config_params.h"
#include "impeghd_mp4_parser.h"
#include "impeghd_mp4_file_wrapper.h"

IA_ERRORCODE mpeghb200_dec_execute_batch(pVOID *objs, pVOID *ins, pVOID *outs, WORD32 n, IA_ERRORCODE *errs);
IA_ERRORCODE impeghd_mp4_parse_mae_boxes(ia_file_wrapper *g_pf_inp, pVOID ptr_dec_api);

typedef struct
{
  ia_mpeghd_api_struct api;
  ia_file_wrapper *in;
  FILE *out;
  pVOID obj;
  pWORD8 inp_buf, out_buf;
  WORD32 buff_size, total_bytes, frames, active;
  unsigned long long hash;
  const char *name;
} stream_t;

static pVOID malloc_global(UWORD32 size, UWORD32 alignment) { return malloc(size + alignment); }
static VOID free_global(pVOID ptr) { free(ptr); }

static void w16(FILE *fp, unsigned v) { fputc(v & 0xff, fp), fputc((v >> 8) & 0xff, fp); }
static void w32(FILE *fp, unsigned v) { w16(fp, v & 0xffff), w16(fp, v >> 16); }
static void wav_header(FILE *fp, WORD32 pcmbytes, WORD32 freq, WORD32 channels, WORD32 bits)
{
  const WORD32 bytes = (bits + 7) / 8;
  fwrite("RIFF", 1, 4, fp);
  w32(fp, (unsigned)(pcmbytes + 44 - 8));
  fwrite("WAVEfmt ", 2, 4, fp);
  w32(fp, 16), w16(fp, 1), w16(fp, (unsigned)channels), w32(fp, (unsigned)freq);
  w32(fp, (unsigned)(freq * channels * bytes)), w16(fp, (unsigned)(channels * bytes)), w16(fp, (unsigned)bits);
  fwrite("data", 1, 4, fp);
  w32(fp, (unsigned)pcmbytes);
}

/* test/impeghd_main.c:577-594 / :698-709: slide the unconsumed bytes down, top the buffer up from the file */
static void refill(stream_t *s)
{
  ia_output_config *o = &s->api.output_config;
  const WORD32 cap = (WORD32)o->ui_inp_buf_size, left = s->buff_size - o->i_bytes_consumed;
  WORD32 got = 0;
  if (cap - left > 0)
  {
    memmove(s->inp_buf, s->inp_buf + o->i_bytes_consumed, (size_t)(left > 0 ? left : 0));
    impeghd_mp4_fw_read(s->in, (pUWORD8)(s->inp_buf + left), cap - left, (pUWORD32)&got);
    s->buff_size = s->buff_size - (o->i_bytes_consumed - got);
  }
  s->api.input_config.num_inp_bytes = s->buff_size;
}

static double now(void)
{
  struct timespec t;
  clock_gettime(CLOCK_MONOTONIC, &t);
  return (double)t.tv_sec + 1e-9 * (double)t.tv_nsec;
}

/* -batch:2: the reference's threading contract (one handle per thread is safe): every handle is decoded to the end by a
 * thread of its own through the unchanged ia_mpegh_dec_execute */
static void *decode_thread(void *arg)
{
  stream_t *s = (stream_t *)arg;
  while (s->active)
  {
    const ia_output_config *oc = &s->api.output_config;
    const unsigned char *p = (const unsigned char *)s->out_buf;
    IA_ERRORCODE err;
    WORD32 b;
    refill(s);
    if (s->api.input_config.num_inp_bytes <= 0)
      break;
    err = ia_mpegh_dec_execute(s->obj, &s->api.input_config, &s->api.output_config);
    if (err == IA_MPEGH_DEC_EXE_NONFATAL_INSUFFICIENT_SD_BYTES || err == IA_MPEGH_DEC_EXE_NONFATAL_INSUFFICIENT_HOA_BYTES ||
        err == IA_MPEGH_DEC_EXE_NONFATAL_INSUFFICIENT_METADATA_BYTES || (err & 0x80000000))
      break;
    for (b = 0; b < oc->num_out_bytes; b++)
      s->hash = (s->hash ^ p[b]) * 1099511628211ull;
    s->total_bytes += oc->num_out_bytes;
    if (oc->num_out_bytes > 0)
    {
      s->frames++;
      if (s->out)
        fwrite(s->out_buf, 1, (size_t)oc->num_out_bytes, s->out);
    }
    if (s->buff_size <= 0)
      break;
  }
  s->active = 0;
  return NULL;
}

int main(int argc, char **argv)
{
  int copies = 1, use_batch = 1, n_write = 1, cicp = IMPEGHD_CONFIG_PARAM_CICP_IDX_DFLT_VAL, cycles = 1, cycle;
  int pcmsz = IMPEGHD_CONFIG_PARAM_PCM_WD_SZ_DFLT_VAL, n_files = 0, n, i, k, n_active;
  const char *prefix = NULL, *files[256];
  stream_t *st;
  pVOID *objs, *ins, *outs;
  IA_ERRORCODE *errs;
  int *map;
  double t0, t1, t_init, t_run, t_dec = 0.0;
  long long total_frames = 0;
  for (i = 1; i < argc; i++)
  {
    if (!strncmp(argv[i], "-n:", 3))
      copies = atoi(argv[i] + 3);
    else if (!strncmp(argv[i], "-batch:", 7))
      use_batch = atoi(argv[i] + 7);
    else if (!strncmp(argv[i], "-write:", 7))
      n_write = atoi(argv[i] + 7);
    else if (!strncmp(argv[i], "-cicp:", 6))
      cicp = atoi(argv[i] + 6);
    else if (!strncmp(argv[i], "-pcmsz:", 7))
      pcmsz = atoi(argv[i] + 7);
    else if (!strncmp(argv[i], "-cycles:", 8))
      cycles = atoi(argv[i] + 8);
    else if (!strncmp(argv[i], "-o:", 3))
      prefix = argv[i] + 3;
    else if (n_files < 256)
      files[n_files++] = argv[i];
  }
  if (!n_files || copies < 1 || !prefix)
  {
    fprintf(stderr, "usage: %s [-n:<copies>] [-batch:0|1|2] [-cycles:<K>] [-write:<K>] [-cicp:<idx>] [-pcmsz:<bits>] -o:<prefix> in1 [in2 ..]\n",
            argv[0]);
    return 2;
  }
  n = n_files * copies;
  /* -cycles:K: create, decode and delete all handles K times in this process (a server's handles come and go); one
   * JSON line per cycle */
  for (cycle = 0; cycle < cycles; cycle++)
  {
  t_dec = 0.0, total_frames = 0;
  st = (stream_t *)calloc((size_t)n, sizeof(stream_t));
  objs = (pVOID *)calloc((size_t)n, sizeof(pVOID)), ins = (pVOID *)calloc((size_t)n, sizeof(pVOID));
  outs = (pVOID *)calloc((size_t)n, sizeof(pVOID));
  errs = (IA_ERRORCODE *)calloc((size_t)n, sizeof(IA_ERRORCODE));
  map = (int *)calloc((size_t)n, sizeof(int));
  t0 = now();
  /* ---- create + init, handle by handle (once per stream; the parse of the configuration is not the hot loop) ---- */
  for (i = 0; i < n; i++)
  {
    stream_t *s = &st[i];
    ia_input_config *ic = &s->api.input_config;
    ia_output_config *oc = &s->api.output_config;
    IA_ERRORCODE err, err_reinit = IA_MPEGH_DEC_NO_ERROR;
    s->name = files[i / copies];
    s->hash = 1469598103934665603ull;
    s->in = impeghd_mp4_fw_open((WORD8 *)s->name);
    if (!s->in)
    {
      fprintf(stderr, "cannot open %s\n", s->name);
      return 1;
    }
    ic->ui_mhas_flag = IMPEGHD_CONFIG_PARAM_MHAS_FLAG_DFLT_VAL;
    ic->ui_pcm_wd_sz = (UWORD32)pcmsz;
    ic->ui_cicp_layout_idx = cicp;
    ic->ui_effect = IMPEGHD_CONFIG_PARAM_EFFECT_TYPE_DFLT_VAL;
    ic->i_preset_id = IMPEGHD_CONFIG_PARAM_PRESET_ID_DFLT_VAL;
    oc->malloc_mpegh = &malloc_global;
    oc->free_mpegh = &free_global;
    if (s->in->is_mp4_file)
      impeghd_mp4_parse_mae_boxes(s->in, &s->api);
    err = ia_mpegh_dec_create((pVOID)ic, (pVOID)oc);
    if (err)
    {
      fprintf(stderr, "ia_mpegh_dec_create: 0x%x\n", (unsigned)err);
      return 1;
    }
    s->obj = oc->pv_ia_process_api_obj;
    s->inp_buf = (pWORD8)oc->mem_info_table[IA_MEMTYPE_INPUT].mem_ptr;
    s->out_buf = (pWORD8)oc->mem_info_table[IA_MEMTYPE_OUTPUT].mem_ptr;
    if (s->in->is_mp4_file && !s->in->is_mp4_mhm1)
      ic->ui_raw_flag = 1;
    oc->i_bytes_consumed = 0;
    s->buff_size = 0;
    do
    {
      refill(s);
      if (s->buff_size <= 0)
      {
        fprintf(stderr, "%s: input ended during initialisation (0x%x)\n", s->name, (unsigned)err_reinit);
        return 1;
      }
      err = ia_mpegh_dec_init(s->obj, (pVOID)ic, (pVOID)oc);
      err_reinit = err;
      if (err & 0x80000000)
      {
        fprintf(stderr, "%s: ia_mpegh_dec_init fatal 0x%x\n", s->name, (unsigned)err);
        return 1;
      }
    } while (!oc->ui_init_done);
    if (i < n_write)
    {
      char path[1024];
      snprintf(path, sizeof path, "%s_%d.wav", prefix, i);
      s->out = fopen(path, "wb");
      if (!s->out)
      {
        perror(path);
        return 1;
      }
      wav_header(s->out, 0, oc->i_samp_freq, oc->i_num_chan, oc->i_pcm_wd_sz);
    }
    s->active = 1;
  }
  t_init = now() - t0;
  /* ---- the hot loop: one batch call advances every stream that still has input by one frame ---- */
  t0 = now();
  if (use_batch == 2)
  {
    pthread_t *th = (pthread_t *)calloc((size_t)n, sizeof(pthread_t));
    for (i = 0; i < n; i++)
      pthread_create(&th[i], NULL, decode_thread, &st[i]);
    for (i = 0; i < n; i++)
      pthread_join(th[i], NULL), total_frames += st[i].frames;
    free(th);
    t_dec = now() - t0;
  }
  for (; use_batch != 2;)
  {
    n_active = 0;
    for (i = 0; i < n; i++)
    {
      stream_t *s = &st[i];
      if (!s->active)
        continue;
      refill(s);
      if (s->api.input_config.num_inp_bytes <= 0)
      {
        s->active = 0;
        continue;
      }
      map[n_active] = i;
      objs[n_active] = s->obj, ins[n_active] = &s->api.input_config, outs[n_active] = &s->api.output_config;
      n_active++;
    }
    if (!n_active)
      break;
    t1 = now();
    if (use_batch)
      mpeghb200_dec_execute_batch(objs, ins, outs, n_active, errs);
    else
      for (k = 0; k < n_active; k++)
        errs[k] = ia_mpegh_dec_execute(objs[k], ins[k], outs[k]);
    t_dec += now() - t1;
    for (k = 0; k < n_active; k++)
    {
      stream_t *s = &st[map[k]];
      const ia_output_config *oc = &s->api.output_config;
      const unsigned char *p = (const unsigned char *)s->out_buf;
      WORD32 b;
      if (errs[k] == IA_MPEGH_DEC_EXE_NONFATAL_INSUFFICIENT_SD_BYTES ||
          errs[k] == IA_MPEGH_DEC_EXE_NONFATAL_INSUFFICIENT_HOA_BYTES ||
          errs[k] == IA_MPEGH_DEC_EXE_NONFATAL_INSUFFICIENT_METADATA_BYTES || (errs[k] & 0x80000000))
      {
        if (errs[k] & 0x80000000)
          fprintf(stderr, "%s (handle %d): execute fatal 0x%x\n", s->name, map[k], (unsigned)errs[k]);
        s->active = 0;
        continue;
      }
      for (b = 0; b < oc->num_out_bytes; b++)
        s->hash = (s->hash ^ p[b]) * 1099511628211ull;
      s->total_bytes += oc->num_out_bytes;
      if (oc->num_out_bytes > 0)
      {
        s->frames++;
        total_frames++;
        if (s->out)
          fwrite(s->out_buf, 1, (size_t)oc->num_out_bytes, s->out);
      }
      if (s->buff_size <= 0)
        s->active = 0;
    }
  }
  t_run = now() - t0;
  /* run_s: the whole loop of this tool (input refill, the decode calls, the FNV hash of every handle's PCM, WAV writes);
   * decode_s: the time inside the library's execute calls alone */
  printf("{\"handles\": %d, \"batch\": %d, \"frames\": %lld, \"init_s\": %.4f, \"run_s\": %.4f, \"decode_s\": %.4f, "
         "\"audio_s_per_wall_s\": %.2f, \"audio_s_per_decode_s\": %.2f, \"streams\": [",
         n, use_batch, total_frames, t_init, t_run, t_dec, (double)total_frames * 1024.0 / 48000.0 / (t_run > 0 ? t_run : 1),
         (double)total_frames * 1024.0 / 48000.0 / (t_dec > 0 ? t_dec : 1));
  for (i = 0; i < n; i++)
  {
    stream_t *s = &st[i];
    ia_output_config *oc = &s->api.output_config;
    printf("%s{\"file\": \"%s\", \"bytes\": %d, \"frames\": %d, \"channels\": %d, \"fnv1a64\": \"%016llx\"}", i ? ", " : "",
           s->name, s->total_bytes, s->frames, oc->i_num_chan, s->hash);
    if (s->out)
    {
      if (!fseek(s->out, 0, SEEK_SET))
        wav_header(s->out, s->total_bytes, oc->i_samp_freq, oc->i_num_chan, oc->i_pcm_wd_sz);
      fclose(s->out);
    }
    ia_mpegh_dec_delete((pVOID)oc);
    impeghd_mp4_fw_close(s->in);
  }
  printf("]}\n");
  fflush(stdout);
  free(st), free(objs), free(ins), free(outs), free(errs), free(map);
  }
  return 0;
}
