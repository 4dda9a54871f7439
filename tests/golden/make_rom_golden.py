#!/usr/bin/env python3
"""Dump the reference ROM tables the hot path indexes at run time into tests/golden/rom_tables.npz.

TEST INFRASTRUCTURE.  Runs only where oracle/_ref/libmpeghref.so exists (built by oracle/Makefile from
/root/reference).  The fixture pins (a) the tables the product library generates by formula + patch list
(libmpegh_b200/csrc/rom_tables.cpp) and (b) feeds the plain-C oracle restatement, so that neither needs
the reference tree on the GPU box.

Sources (reference file:line of each array):
  ia_fft_twiddle_table_float[514]                  decoder/impeghd_fft_ifft_rom.c:45
  ia_mixed_rad_twiddle_cos/sin[16384]              decoder/impeghd_fft_ifft_rom.c:1122,4401
  ia_core_coder_pre_post_twid_{cos,sin}_{512,64}   decoder/ia_core_coder_rom.c:1705,1771,1594,1583
  ia_core_coder_pre_twid_type2_dct_dst_{cos,sin}_1024  decoder/ia_core_coder_rom.c:1967,1837
  ia_core_coder_sine_window{1024,128}, kbd_win1024, kbd_window128   decoder/ia_core_coder_rom.c:2103ff
"""
import ctypes
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))

TABLES = {
    # name in the fixture: (reference symbol, ctype, count)
    "fft_twiddle": ("ia_fft_twiddle_table_float", ctypes.c_float, 514),
    "mixed_rad_cos": ("ia_mixed_rad_twiddle_cos", ctypes.c_float, 16384),
    "mixed_rad_sin": ("ia_mixed_rad_twiddle_sin", ctypes.c_float, 16384),
    "twid_cos_512": ("ia_core_coder_pre_post_twid_cos_512", ctypes.c_float, 512),
    "twid_sin_512": ("ia_core_coder_pre_post_twid_sin_512", ctypes.c_float, 512),
    "twid_cos_64": ("ia_core_coder_pre_post_twid_cos_64", ctypes.c_float, 64),
    "twid_sin_64": ("ia_core_coder_pre_post_twid_sin_64", ctypes.c_float, 64),
    "dctdst_cos_1024": ("ia_core_coder_pre_twid_type2_dct_dst_cos_1024", ctypes.c_float, 1024),
    "dctdst_sin_1024": ("ia_core_coder_pre_twid_type2_dct_dst_sin_1024", ctypes.c_float, 1024),
    "sine_1024": ("ia_core_coder_sine_window1024", ctypes.c_float, 1024),
    "sine_128": ("ia_core_coder_sine_window128", ctypes.c_float, 128),
    "kbd_1024": ("ia_core_coder_kbd_win1024", ctypes.c_float, 1024),
    "kbd_128": ("ia_core_coder_kbd_window128", ctypes.c_float, 128),
    "sine_256": ("ia_core_coder_sine_window256", ctypes.c_float, 256),
}


def main():
    so = os.path.join(ROOT, "oracle", "_ref", "libmpeghref.so")
    if not os.path.exists(so):
        sys.exit("oracle/_ref/libmpeghref.so missing: run `make -C oracle ref` where /root/reference exists")
    lib = ctypes.CDLL(so)
    out = {}
    for key, (sym, ct, n) in TABLES.items():
        arr = (ct * n).in_dll(lib, sym)
        out[key] = np.ctypeslib.as_array(arr).copy()
    np.savez_compressed(os.path.join(HERE, "rom_tables.npz"), **out)
    print("wrote rom_tables.npz:", {k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
