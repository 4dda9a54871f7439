/* seam_internal.h -- what the translation units of libmpeghb200_seam.so share. */
#ifndef MPEGHB200_SEAM_INTERNAL_H
#define MPEGHB200_SEAM_INTERNAL_H
#include "mpeghb200.h"

mpeghb200_ctx *seam_ctx(void); /* creates the context on first use; never returns without one */
void seam_die(const char *what, mpeghb200_status st);
void seam_count_fc(void);
#define CK(call)                                                                                   \
  do                                                                                               \
  {                                                                                                \
    mpeghb200_status st_ = (call);                                                                 \
    if (st_ != MPEGHB200_OK)                                                                       \
      seam_die(#call, st_);                                                                        \
  } while (0)
#endif
