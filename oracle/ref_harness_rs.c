/* oracle/ref_harness_rs.c -- TEST INFRASTRUCTURE ONLY.
 * Function-level entry point into the UNMODIFIED reference resampler impeghd_resample
 * (decoder/impeghd_resampler.c:93) and its filter tables.  No reference code, only calls. */
#include <stdlib.h>
#include <string.h>
#include <impeghd_type_def.h>
#include "impeghd_resampler.h"
#include "impeghd_resampler_rom.h"

void *ref_rs_create(int fac_up, int fac_down)
{
  ia_resampler_struct *r = (ia_resampler_struct *)calloc(1, sizeof(ia_resampler_struct));
  r->fac_up = fac_up;
  r->fac_down = fac_down;
  r->input_length = 1024;
  r->output_length = 1024 * fac_up / fac_down;
  return r;
}
void ref_rs_destroy(void *p) { free(p); }
void ref_rs_filters(float *f2, float *f3)
{
  memcpy(f2, impeghd_resampler_filter, sizeof(float) * RESAMPLE_FILTER_LEN);
  memcpy(f3, impeghd_resampler_filter_3, sizeof(float) * RESAMPLE_FILTER_LEN);
}
/* in [1024] -> out [1024 * up / down] */
void ref_rs_process(void *p, const float *in, float *out)
{
  static __thread float scratch[4096], inbuf[1024];
  memcpy(inbuf, in, sizeof(inbuf));
  impeghd_resample((ia_resampler_struct *)p, inbuf, out, scratch);
}
