// rom_tables.cpp -- formula + patch-list generation of the run-time constant tables (see rom_tables.h).
#include "rom_tables.h"

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>

namespace mpeghb200 {
namespace {

constexpr double kPi = 3.14159265358979323846264338327950288;

// The reference's core-coder tables are literals printed with six decimals ("0.999954f"): reproduce the
// decimal rounding, then let strtof do the decimal -> binary32 conversion the C compiler did.
float dec6(double v) {
  char buf[48];
  std::snprintf(buf, sizeof(buf), "%.6f", v);
  return std::strtof(buf, nullptr);
}

// Modified Bessel function of the first kind, order 0 (power series, double precision).
double bessel_i0(double x) {
  double sum = 1.0, term = 1.0;
  const double q = x * x / 4.0;
  for (int k = 1; k < 200; ++k) {
    term *= q / (double(k) * double(k));
    sum += term;
    if (term < sum * 1e-18) break;
  }
  return sum;
}

// Kaiser-Bessel-derived window, rising slope of length n (ISO/IEC 14496-3 4.6.11.3.2): alpha = 4 for the
// long window, 6 for the short one.
std::vector<float> kbd_window(int n, double alpha) {
  std::vector<double> w(n + 1), cs(n + 1);
  double acc = 0.0;
  for (int i = 0; i <= n; ++i) {
    const double r = 2.0 * i / n - 1.0;
    w[i] = bessel_i0(kPi * alpha * std::sqrt(1.0 - r * r));
    acc += w[i];
    cs[i] = acc;
  }
  std::vector<float> out(n);
  for (int i = 0; i < n; ++i) out[i] = dec6(std::sqrt(cs[i] / cs[n]));
  return out;
}

std::vector<float> sine_window(int n) {
  std::vector<float> out(n);
  for (int i = 0; i < n; ++i) out[i] = dec6(std::sin(kPi * (i + 0.5) / (2.0 * n)));
  return out;
}

// MDCT pre/post twiddle, angle 2*pi*(i + 1/8)/(4n).
std::vector<float> mdct_twiddle(int n, bool is_sin) {
  std::vector<float> out(n);
  for (int i = 0; i < n; ++i) {
    const double a = 2.0 * kPi * (i + 0.125) / (4.0 * n);
    out[i] = dec6(is_sin ? std::sin(a) : std::cos(a));
  }
  return out;
}

std::vector<float> generate(RomId id) {
  std::vector<float> t;
  switch (id) {
    case ROM_FFT_TWIDDLE: {
      // [0..256] = cos(2 pi k / 1024), [257..513] = -sin(2 pi k / 1024)
      t.resize(514);
      for (int k = 0; k <= 256; ++k) {
        t[k] = float(std::cos(2.0 * kPi * k / 1024.0));
        t[257 + k] = (k == 0) ? 0.0f : float(-std::sin(2.0 * kPi * k / 1024.0));
      }
      break;
    }
    case ROM_TWID_COS_512: t = mdct_twiddle(512, false); break;
    case ROM_TWID_SIN_512: t = mdct_twiddle(512, true); break;
    case ROM_TWID_COS_64: t = mdct_twiddle(64, false); break;
    case ROM_TWID_SIN_64: t = mdct_twiddle(64, true); break;
    case ROM_DCTDST_COS_1024:
    case ROM_DCTDST_SIN_1024: {
      t.resize(1024);
      for (int i = 0; i < 1024; ++i) {
        const double a = kPi * (i + 1) / 2048.0;
        t[i] = dec6(id == ROM_DCTDST_SIN_1024 ? std::sin(a) : std::cos(a));
      }
      break;
    }
    case ROM_SINE_1024: t = sine_window(1024); break;
    case ROM_SINE_128: t = sine_window(128); break;
    case ROM_SINE_256: {  // this one is printed with eight decimals (decoder/ia_core_coder_avq_rom.c:539)
      t.resize(256);
      for (int i = 0; i < 256; ++i) {
        char buf[48];
        std::snprintf(buf, sizeof(buf), "%.8f", std::sin(kPi * (i + 0.5) / 512.0));
        t[i] = std::strtof(buf, nullptr);
      }
      break;
    }
    case ROM_KBD_1024: t = kbd_window(1024, 4.0); break;
    case ROM_KBD_128: t = kbd_window(128, 6.0); break;
    case ROM_MIXED_RAD_COS:
    case ROM_MIXED_RAD_SIN: {
      // entry [16*i + j] = cos / sin(2 pi i j / 16384): the inter-stage twiddle of the 1024 x {2..16}
      // decomposition used for transforms longer than 1024 (impeghd_ifft.c:1264, impeghd_fft.c:1415).
      t.resize(16384);
      for (int idx = 0; idx < 16384; ++idx) {
        const double a = 2.0 * kPi * double(idx >> 4) * double(idx & 15) / 16384.0;
        t[idx] = float(id == ROM_MIXED_RAD_SIN ? std::sin(a) : std::cos(a));
      }
      break;
    }
    default: break;
  }
  return t;
}

struct Patch {
  int id;
  int index;
  uint32_t bits;
};
const Patch kPatches[] = {
#include "rom_patches.inc"
    {-1, 0, 0u}};

std::vector<float> g_tables[ROM_COUNT][2];
std::once_flag g_once;

void build_all() {
  for (int id = 0; id < ROM_COUNT; ++id) {
    g_tables[id][0] = generate(RomId(id));
    g_tables[id][1] = g_tables[id][0];
  }
  for (const Patch& p : kPatches) {
    if (p.id < 0 || p.id >= ROM_COUNT) continue;
    if (p.index < 0 || size_t(p.index) >= g_tables[p.id][1].size()) continue;
    float v;
    std::memcpy(&v, &p.bits, sizeof(v));
    g_tables[p.id][1][p.index] = v;
  }
}

}  // namespace

const std::vector<float>& rom_table(RomId id, bool patched) {
  std::call_once(g_once, build_all);
  return g_tables[id][patched ? 1 : 0];
}

const char* rom_name(RomId id) {
  static const char* names[ROM_COUNT] = {"fft_twiddle",    "twid_cos_512",   "twid_sin_512", "twid_cos_64",
                                         "twid_sin_64",    "dctdst_cos_1024", "dctdst_sin_1024",
                                         "sine_1024",      "sine_128",       "kbd_1024",     "kbd_128",
                                         "mixed_rad_cos",  "mixed_rad_sin",  "sine_256"};
  return (id >= 0 && id < ROM_COUNT) ? names[id] : "?";
}

}  // namespace mpeghb200
