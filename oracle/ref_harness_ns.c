/* oracle/ref_harness_ns.c -- TEST INFRASTRUCTURE ONLY.
 *
 * Entry points into `static` functions of the unmodified reference.  oracle/Makefile compiles the reference
 * files that hold them a second time with -Dstatic= (linkage only; the code is unchanged) and links this
 * harness with them into oracle/_ref/libmpeghref_ns.so.  Contains no reference code, only calls into it.
 */
#include <stdlib.h>
#include <string.h>
#include <impeghd_type_def.h>

/* ia_core_coder_samples_sat (ia_core_coder_decode_main.c:209) */
VOID ia_core_coder_samples_sat(WORD8 *outbuffer, WORD32 num_samples_out, WORD32 pcmsize,
                               FLOAT32 **out_samples, WORD32 *out_ch_map, WORD32 *ptr_delay_samples,
                               WORD32 *out_bytes, WORD32 num_channel_out);

/* planar [n_ch][n] floats -> interleaved PCM bytes.  out_ch_map[n_ch] may hold -1 (identity), is not
 * modified; *delay_samples is updated like the reference does.  Returns the byte count reported. */
int ref_samples_sat(const float *planar, int n, int n_ch, int pcmsize, const int *out_ch_map,
                    int *delay_samples, unsigned char *out)
{
  FLOAT32 *rows[64];
  WORD32 map[64];
  WORD32 out_bytes = 0;
  int c;
  for (c = 0; c < n_ch; c++)
  {
    rows[c] = (FLOAT32 *)(planar + (size_t)c * n);
    map[c] = out_ch_map[c];
  }
  ia_core_coder_samples_sat((WORD8 *)out, n, pcmsize, rows, map, delay_samples, &out_bytes, n_ch);
  return out_bytes;
}
