// bands.h -- scale-factor-band tables of one sampling rate on the device, shared by the spectral tools
// (stereo.cu, igf.cu): upper band edges of a long window and of a short window, and the bin -> band maps.
#pragma once
#include <cstdint>

#include "common.h"

struct mpeghb200_bands {
  mpeghb200_ctx* ctx = nullptr;
  uint8_t* d_blob = nullptr;            // one allocation: the four tables below
  const uint8_t* d_bin2sfb_long = nullptr;   // [1024]
  const uint8_t* d_bin2sfb_short = nullptr;  // [128]
  const int16_t* d_sfb_long = nullptr;       // [n_long] upper edges
  const int16_t* d_sfb_short = nullptr;      // [n_short]
  int n_long = 0, n_short = 0;
};
