// rom_tables.h -- constant tables the signal path indexes at run time.
//
// The reference keeps these as literal arrays (decoder/ia_core_coder_rom.c, decoder/impeghd_fft_ifft_rom.c).
// Parity needs the *same float values*, but the product must stand alone, so every table is generated
// here from its defining formula with the reference's literal precision (6 decimals for the core-coder
// tables, full float precision for the FFT twiddles) and then corrected by a short list of entries whose
// last printed digit was rounded differently by the reference's (unknown) generator
// (rom_patches.inc, produced by tools/gen_rom_patches.py from tests/golden/rom_tables.npz).
// tests/test_rom_tables.py checks every table bit-for-bit against the fixture.
#pragma once
#include <cstdint>
#include <vector>

namespace mpeghb200 {

enum RomId : int {
  ROM_FFT_TWIDDLE = 0,   // ia_fft_twiddle_table_float[514]        impeghd_fft_ifft_rom.c:45
  ROM_TWID_COS_512 = 1,  // ia_core_coder_pre_post_twid_cos_512    ia_core_coder_rom.c:1705
  ROM_TWID_SIN_512 = 2,  // ia_core_coder_pre_post_twid_sin_512    ia_core_coder_rom.c:1771
  ROM_TWID_COS_64 = 3,   // ia_core_coder_pre_post_twid_cos_64     ia_core_coder_rom.c:1594
  ROM_TWID_SIN_64 = 4,   // ia_core_coder_pre_post_twid_sin_64     ia_core_coder_rom.c:1583
  ROM_DCTDST_COS_1024 = 5,  // ia_core_coder_pre_twid_type2_dct_dst_cos_1024  ia_core_coder_rom.c:1967
  ROM_DCTDST_SIN_1024 = 6,  // ia_core_coder_pre_twid_type2_dct_dst_sin_1024  ia_core_coder_rom.c:1837
  ROM_SINE_1024 = 7,     // ia_core_coder_sine_window1024
  ROM_SINE_128 = 8,      // ia_core_coder_sine_window128
  ROM_KBD_1024 = 9,      // ia_core_coder_kbd_win1024              ia_core_coder_rom.c:2103
  ROM_KBD_128 = 10,      // ia_core_coder_kbd_window128
  ROM_MIXED_RAD_COS = 11,  // ia_mixed_rad_twiddle_cos[16384]      impeghd_fft_ifft_rom.c:1122
  ROM_MIXED_RAD_SIN = 12,  // ia_mixed_rad_twiddle_sin[16384]      impeghd_fft_ifft_rom.c:4401
  ROM_SINE_256 = 13,       // ia_core_coder_sine_window256 (format converter STFT window)
  ROM_COUNT
};

// Generated (and patched unless patched == false) table; built once, thread-safe.
const std::vector<float>& rom_table(RomId id, bool patched = true);
const char* rom_name(RomId id);

}  // namespace mpeghb200
