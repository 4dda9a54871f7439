"""ctypes loaders for the parity checkers -- TEST INFRASTRUCTURE ONLY.

`load_oracle()`  -> oracle/liboracle.so   (our plain-C restatement, oracle_*.c)
`load_ref()`     -> oracle/_ref/libmpeghref.so (the unmodified reference sources + ref_harness.c), or None
                    when it has not been built (it is built in the dev container by `make -C oracle ref`
                    and travels to the GPU box as a prebuilt file).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import this.
"""
import ctypes
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)

_f32p = np.ctypeslib.ndpointer(dtype=np.float32, flags="C_CONTIGUOUS")
_i32p = np.ctypeslib.ndpointer(dtype=np.int32, flags="C_CONTIGUOUS")
_u8p = np.ctypeslib.ndpointer(dtype=np.uint8, flags="C_CONTIGUOUS")

ROM_EXTRA = ["sine_256"]  # in the fixture, not part of oracle_rom
ROM_FIELDS = [
    "fft_twiddle", "twid_cos_512", "twid_sin_512", "twid_cos_64", "twid_sin_64", "dctdst_cos_1024",
    "dctdst_sin_1024", "sine_1024", "sine_128", "kbd_1024", "kbd_128", "mixed_rad_cos", "mixed_rad_sin",
]


class OracleRom(ctypes.Structure):
    _fields_ = [(n, ctypes.POINTER(ctypes.c_float)) for n in ROM_FIELDS]


def build(target="oracle"):
    subprocess.run(["make", "-s", "-C", HERE, target], check=True)


def golden_rom():
    return np.load(os.path.join(ROOT, "tests", "golden", "rom_tables.npz"))


_oracle = None
_ref = None
_keep = []


def load_oracle():
    global _oracle
    if _oracle is not None:
        return _oracle
    so = os.path.join(HERE, "liboracle.so")
    if not os.path.exists(so):
        build("oracle")
    lib = ctypes.CDLL(so)
    rom_np = golden_rom()
    rom = OracleRom()
    for n in ROM_FIELDS:
        a = np.ascontiguousarray(rom_np[n], dtype=np.float32)
        _keep.append(a)
        setattr(rom, n, a.ctypes.data_as(ctypes.POINTER(ctypes.c_float)))
    lib.oracle_set_rom.argtypes = [ctypes.POINTER(OracleRom)]
    lib.oracle_set_rom(ctypes.byref(rom))
    _keep.append(rom)

    lib.oracle_cplx_ifft_pow2.argtypes = [_f32p, _f32p, ctypes.c_int]
    lib.oracle_fd_frm_dec.argtypes = [_f32p, _f32p, _f32p] + [ctypes.c_int] * 5
    lib.oracle_fd_frm_dec.restype = ctypes.c_int
    lib.oracle_fd_frm_dec_batch.argtypes = [_f32p, _f32p, _f32p, _i32p, ctypes.c_int, ctypes.c_int]
    lib.oracle_fd_frm_dec_batch.restype = ctypes.c_int
    lib.oracle_limiter_init.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int]
    lib.oracle_limiter_process.argtypes = [ctypes.c_void_p, _f32p, ctypes.c_int]
    lib.oracle_loudness_gain.argtypes = [ctypes.c_float]
    lib.oracle_loudness_gain.restype = ctypes.c_float
    lib.oracle_samples_sat.argtypes = [_f32p, ctypes.c_int, ctypes.c_int, ctypes.c_int, _i32p, _i32p, _u8p]
    lib.oracle_samples_sat.restype = ctypes.c_int
    lib.oracle_obj_trig.argtypes = [ctypes.c_void_p, ctypes.c_void_p]
    lib.oracle_obj_gains.argtypes = [ctypes.c_void_p, ctypes.c_void_p, _f32p]
    lib.oracle_obj_render.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, _f32p, _f32p]
    lib.oracle_obj_render_sub.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, _f32p, _f32p,
                                          ctypes.c_int, ctypes.c_int, ctypes.c_int]
    lib.oracle_hoa_render.argtypes = [_f32p, ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, _f32p,
                                      ctypes.c_int, ctypes.c_int, _f32p]
    lib.oracle_cplx_fft_pow2.argtypes = [_f32p, _f32p, ctypes.c_int]
    for f in (lib.oracle_cplx_ifft_big, lib.oracle_cplx_fft_big):
        f.argtypes = [_f32p, _f32p, ctypes.c_int]
        f.restype = ctypes.c_int
    lib.oracle_binaural_spectrum.argtypes = [_f32p, ctypes.c_int, ctypes.c_int, _f32p]
    lib.oracle_binaural_spectrum.restype = ctypes.c_int
    lib.oracle_binaural_block.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_int, _f32p, _f32p, _i32p, _i32p, _f32p,
                                          _f32p, _i32p, _f32p, _f32p, _f32p]
    lib.oracle_binaural_block.restype = ctypes.c_int
    lib.oracle_hoasp_state_init.argtypes = [ctypes.c_void_p]
    lib.oracle_hoasp_process.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, _f32p, _f32p]
    lib.oracle_hoasp_process.restype = ctypes.c_int
    lib.oracle_fc_state_init.argtypes = [ctypes.c_void_p]
    lib.oracle_fc_process.argtypes = [_f32p, ctypes.c_int, ctypes.c_int, _f32p, ctypes.c_void_p, _f32p, _f32p]
    _i16p = np.ctypeslib.ndpointer(dtype=np.int16, flags="C_CONTIGUOUS")
    lib.oracle_tns_lpc.argtypes = [ctypes.c_int, ctypes.c_int, _i32p, _f32p]
    lib.oracle_tns_filter.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int, _f32p]
    lib.oracle_tns_range.argtypes = [ctypes.c_int] * 4 + [_i16p, ctypes.c_int, _i32p, _i32p]
    lib.oracle_tns_range.restype = ctypes.c_int
    lib.oracle_rs_state_init.argtypes = [ctypes.c_void_p]
    lib.oracle_rs_process.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, _f32p, _f32p, _f32p, _f32p]
    lib.oracle_stereo_set_rom.argtypes = [_f32p, _f32p]
    lib.oracle_stereo_apply.argtypes = [ctypes.c_char_p, _i16p, ctypes.c_int, _f32p, _f32p, _f32p]
    lib.oracle_stereo_save.argtypes = [ctypes.c_char_p, _f32p, _f32p, _f32p]
    lib.oracle_igf_mono.argtypes = [ctypes.c_char_p, _i16p, ctypes.c_int, _f32p, _f32p, _f32p]
    lib.oracle_igf_mono.restype = ctypes.c_uint
    lib.oracle_igf_stereo.argtypes = [ctypes.c_char_p, ctypes.c_char_p, _u8p, _i16p, ctypes.c_int, _f32p, _f32p, _f32p,
                                      _f32p, _f32p, _f32p, np.ctypeslib.ndpointer(dtype=np.uint32, flags="C_CONTIGUOUS")]
    lib.oracle_igf_tnf.argtypes = [ctypes.c_char_p, _i16p, ctypes.c_int, _f32p, _f32p]
    lib.oracle_ltpf_cfg_init.argtypes = [ctypes.c_void_p, ctypes.c_int]
    lib.oracle_ltpf_state_init.argtypes = [ctypes.c_void_p]
    lib.oracle_ltpf_process.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int, _f32p]
    lib.oracle_drc_gains_subband.argtypes = [_f32p, _f32p, ctypes.c_int, _i32p, _i32p, ctypes.c_int, _f32p, _f32p,
                                             ctypes.c_int, ctypes.c_int, ctypes.c_int]
    lib.oracle_drc_td_channel.argtypes = [ctypes.c_void_p, _f32p, _f32p, ctypes.c_void_p]
    lib.oracle_ds_t2f.argtypes = [_f32p, _f32p, _f32p, _f32p, _f32p, ctypes.c_int]
    lib.oracle_ds_f2t.argtypes = [_f32p, _f32p, _f32p, _f32p, _f32p, ctypes.c_int]
    _mct_args = [_f32p, _f32p, _i32p, _i32p, ctypes.c_int, ctypes.c_int, _i16p] + [ctypes.c_int] * 4
    lib.oracle_mct_process.argtypes = _mct_args
    lib.oracle_mct_rom.argtypes = [_f32p, _f32p]
    lib.oracle_igf_tiles.argtypes = [ctypes.c_int] * 4 + [_i16p, _i32p]
    lib.oracle_igf_tiles.restype = ctypes.c_int
    lib.oracle_igf_whitening_table.argtypes = [_f32p]
    sg_path = os.path.join(ROOT, "tests", "golden", "stereo_golden.npz")
    if os.path.exists(sg_path):  # MDST filter ROM of the complex-prediction tool (absent only while it is being made)
        sg = np.load(sg_path)
        lib.oracle_stereo_set_rom(np.ascontiguousarray(sg["mdst_curr"], np.float32).reshape(-1),
                                  np.ascontiguousarray(sg["mdst_prev"], np.float32).reshape(-1))
    _oracle = lib
    return lib


LIM_MAX_ATTACK, LIM_MAX_CH = 240, 32


class OracleLimiter(ctypes.Structure):
    """Mirror of oracle.h's oracle_limiter (state in time order: the layout the CUDA kernels use too)."""
    _fields_ = [("n_ch", ctypes.c_int), ("attack_samples", ctypes.c_int), ("limiter_on", ctypes.c_int),
                ("threshold", ctypes.c_float), ("attack_const", ctypes.c_float), ("release_const", ctypes.c_float),
                ("gain_modified", ctypes.c_float), ("pre_smoothed_gain", ctypes.c_float),
                ("max_hist", ctypes.c_float * LIM_MAX_ATTACK),
                ("delay", (ctypes.c_float * LIM_MAX_ATTACK) * LIM_MAX_CH)]


def oracle_limiter_new(n_ch, sample_rate=48000, limiter_on=1):
    st = OracleLimiter()
    load_oracle().oracle_limiter_init(ctypes.byref(st), n_ch, sample_rate, limiter_on)
    return st


def oracle_limiter_run(st, planar):
    """planar [n_ch, n] float32 -> limited copy; st is advanced."""
    x = np.ascontiguousarray(planar, dtype=np.float32).copy()
    load_oracle().oracle_limiter_process(ctypes.byref(st), x, x.shape[1])
    return x


def oracle_samples_sat(planar, pcmsize, ch_map=None, delay=0):
    """-> (bytes ndarray, new_delay)"""
    planar = np.ascontiguousarray(planar, dtype=np.float32)
    nch, n = planar.shape
    m = np.full(nch, -1, np.int32) if ch_map is None else np.ascontiguousarray(ch_map, dtype=np.int32)
    d = np.array([delay], np.int32)
    out = np.zeros(nch * n * (pcmsize // 8), np.uint8)
    nb = load_oracle().oracle_samples_sat(planar, n, nch, pcmsize, m, d, out)
    return out[:nb].copy(), int(d[0])


def _bind_harnesses(lib):
    """ctypes signatures of the function-level harnesses (oracle/ref_harness_*.c): both reference builds carry them."""
    lib.ref_fd_create.restype = ctypes.c_void_p
    lib.ref_fd_destroy.argtypes = [ctypes.c_void_p]
    lib.ref_fd_frm_dec.argtypes = [ctypes.c_void_p, _f32p, _f32p, _f32p] + [ctypes.c_int] * 5
    lib.ref_fd_frm_dec.restype = ctypes.c_int
    lib.ref_fd_frm_dec_batch.argtypes = [ctypes.c_void_p, _f32p, _f32p, _f32p, _i32p, ctypes.c_int,
                                         ctypes.c_int]
    lib.ref_fd_frm_dec_batch.restype = ctypes.c_int
    lib.ref_rad2_cplx_ifft.argtypes = [_f32p, _f32p, ctypes.c_int]
    lib.ref_rad2_cplx_fft.argtypes = [_f32p, _f32p, ctypes.c_int]
    lib.ref_limiter_create.restype = ctypes.c_void_p
    lib.ref_limiter_create.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_int]
    lib.ref_limiter_destroy.argtypes = [ctypes.c_void_p]
    lib.ref_limiter_process.argtypes = [ctypes.c_void_p, _f32p, ctypes.c_int]
    lib.ref_limiter_constants.argtypes = [ctypes.c_void_p, _f32p]
    lib.ref_obj_create.restype = ctypes.c_void_p
    lib.ref_obj_create.argtypes = [ctypes.c_int, ctypes.POINTER(ctypes.c_int)]
    lib.ref_obj_destroy.argtypes = [ctypes.c_void_p]
    lib.ref_obj_layout.argtypes = [ctypes.c_void_p, _i32p, _i32p, _i32p, _f32p, _f32p, _f32p]
    lib.ref_obj_first_flags.argtypes = [ctypes.c_void_p, _i32p, ctypes.c_int]
    lib.ref_obj_render.argtypes = [ctypes.c_void_p, _f32p, ctypes.c_int, _f32p, _f32p, _f32p]
    lib.ref_obj_render.restype = ctypes.c_int
    lib.ref_obj_render_subframes.argtypes = [ctypes.c_void_p, _f32p, _i32p, ctypes.c_int, ctypes.c_int, _f32p, _f32p,
                                             _i32p]
    lib.ref_obj_render_subframes.restype = ctypes.c_int
    lib.ref_hoa_ren_create.restype = ctypes.c_void_p
    lib.ref_hoa_ren_create.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_float, ctypes.c_int,
                                       ctypes.POINTER(ctypes.c_int)]
    lib.ref_hoa_ren_destroy.argtypes = [ctypes.c_void_p]
    lib.ref_hoa_ren_matrix.argtypes = [ctypes.c_void_p, _i32p, _f32p]
    lib.ref_hoa_ren_nfc_coeffs.argtypes = [ctypes.c_void_p, ctypes.c_int, _f32p]
    lib.ref_hoa_ren_process.argtypes = [ctypes.c_void_p, _f32p, ctypes.c_int, ctypes.c_int, ctypes.c_int, _f32p]
    lib.ref_hoa_ren_process.restype = ctypes.c_int
    lib.ref_cplx_ifft_8k.argtypes = [_f32p, _f32p, ctypes.c_int]
    lib.ref_cplx_ifft_8k_fill.argtypes = [_f32p, _f32p, ctypes.c_int, ctypes.c_float]
    lib.ref_cplx_fft_16k.argtypes = [_f32p, _f32p, ctypes.c_int]
    lib.ref_bin_create.restype = ctypes.c_void_p
    lib.ref_bin_create.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_int, _f32p, _f32p, _f32p, _f32p, _f32p,
                                   ctypes.c_int, ctypes.POINTER(ctypes.c_int)]
    lib.ref_bin_destroy.argtypes = [ctypes.c_void_p]
    lib.ref_bin_info.argtypes = [ctypes.c_void_p, _i32p]
    lib.ref_bin_spectra.argtypes = [ctypes.c_void_p, _f32p, _f32p, _i32p, _i32p]
    lib.ref_bin_process.argtypes = [ctypes.c_void_p, _f32p, ctypes.c_int, _f32p, ctypes.c_int]
    lib.ref_bin_process.restype = ctypes.c_int
    lib.ref_hoasp_create.restype = ctypes.c_void_p
    lib.ref_hoasp_create.argtypes = [_i32p, ctypes.POINTER(ctypes.c_int)]
    lib.ref_hoasp_destroy.argtypes = [ctypes.c_void_p]
    lib.ref_hoasp_tables.argtypes = [ctypes.c_void_p, _i32p] + [_f32p] * 7
    lib.ref_hoasp_process.argtypes = [ctypes.c_void_p, ctypes.c_void_p, _f32p, _f32p]
    lib.ref_hoasp_process.restype = ctypes.c_int
    lib.ref_fc_create.restype = ctypes.c_void_p
    lib.ref_fc_create.argtypes = [ctypes.c_int] * 3 + [ctypes.POINTER(ctypes.c_int)]
    lib.ref_fc_destroy.argtypes = [ctypes.c_void_p]
    lib.ref_fc_info.argtypes = [ctypes.c_void_p, _i32p]
    lib.ref_fc_get_matrix.argtypes = [ctypes.c_void_p, _f32p]
    lib.ref_fc_set_matrix.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, _f32p]
    lib.ref_fc_process.argtypes = [ctypes.c_void_p, _f32p, _f32p]
    lib.ref_fc_process.restype = ctypes.c_int
    _i16p = np.ctypeslib.ndpointer(dtype=np.int16, flags="C_CONTIGUOUS")
    lib.ref_tns_apply.argtypes = [_f32p, ctypes.c_int, _i32p, _i32p, _i16p, ctypes.c_int, ctypes.c_int, ctypes.c_int]
    lib.ref_tns_apply.restype = ctypes.c_int
    lib.ref_cpe_create.restype = ctypes.c_void_p
    lib.ref_cpe_create.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_int]
    lib.ref_cpe_destroy.argtypes = [ctypes.c_void_p]
    lib.ref_cpe_sfb_table.argtypes = [ctypes.c_int, ctypes.c_int, _i16p, ctypes.c_int]
    lib.ref_cpe_sfb_table.restype = ctypes.c_int
    lib.ref_cpe_process.argtypes = [ctypes.c_void_p, ctypes.c_char_p, _i32p, ctypes.c_int, _i32p, _i32p, _i32p, _f32p,
                                    _f32p, _f32p, _f32p, _f32p, _f32p]
    lib.ref_cpe_process.restype = ctypes.c_int
    lib.ref_mdst_rom.argtypes = [_f32p, _f32p]
    _i16p = np.ctypeslib.ndpointer(dtype=np.int16, flags="C_CONTIGUOUS")
    lib.ref_mct_process.argtypes = [_f32p, _f32p, _i32p, _i32p, ctypes.c_int, ctypes.c_int, _i16p] + [ctypes.c_int] * 4
    lib.ref_mct_rom.argtypes = [_f32p, _f32p]
    lib.ref_drc_gains_subband.argtypes = [_f32p, _f32p, ctypes.c_int, _i32p, _i32p, ctypes.c_int, _f32p, _f32p] + \
        [ctypes.c_int] * 5
    lib.ref_drc_gains_subband.restype = ctypes.c_int
    lib.ref_drctd_whole_create.restype = ctypes.c_void_p
    lib.ref_drctd_whole_create.argtypes = [ctypes.c_int, np.ctypeslib.ndpointer(dtype=np.int32, flags="C_CONTIGUOUS"),
                                           ctypes.c_int, ctypes.c_void_p]
    lib.ref_drctd_whole_destroy.argtypes = [ctypes.c_void_p]
    lib.ref_drctd_whole_process.argtypes = [ctypes.c_void_p, _f32p, _f32p]
    lib.ref_drctd_whole_process.restype = ctypes.c_int
    lib.ref_ds_create.restype = ctypes.c_void_p
    lib.ref_ds_create.argtypes = [ctypes.c_int]
    lib.ref_ds_destroy.argtypes = [ctypes.c_void_p]
    lib.ref_ds_t2f.argtypes = [ctypes.c_void_p, _f32p, _f32p, _f32p]
    lib.ref_ds_f2t.argtypes = [ctypes.c_void_p, _f32p, _f32p, _f32p]
    lib.ref_ltpf_create.restype = ctypes.c_void_p
    lib.ref_ltpf_create.argtypes = [ctypes.c_int]
    lib.ref_ltpf_destroy.argtypes = [ctypes.c_void_p]
    lib.ref_ltpf_cfg.argtypes = [ctypes.c_void_p, _i32p]
    lib.ref_ltpf_process.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int, _f32p]
    lib.ref_igf_create.restype = ctypes.c_void_p
    lib.ref_igf_create.argtypes = [ctypes.c_int] * 7
    lib.ref_igf_destroy.argtypes = [ctypes.c_void_p]
    lib.ref_igf_grid.argtypes = [ctypes.c_void_p, _i32p]
    lib.ref_igf_process.argtypes = [ctypes.c_void_p, ctypes.c_char_p, _f32p, _f32p]
    lib.ref_igf_process.restype = ctypes.c_uint
    lib.ref_igf_stereo_process.argtypes = [ctypes.c_void_p, ctypes.c_char_p, ctypes.c_char_p, _u8p, ctypes.c_int, _f32p,
                                           _f32p, _f32p, _f32p,
                                           np.ctypeslib.ndpointer(dtype=np.uint32, flags="C_CONTIGUOUS")]
    lib.ref_rs_create.restype = ctypes.c_void_p
    lib.ref_rs_create.argtypes = [ctypes.c_int, ctypes.c_int]
    lib.ref_rs_destroy.argtypes = [ctypes.c_void_p]
    lib.ref_rs_filters.argtypes = [_f32p, _f32p]
    lib.ref_rs_process.argtypes = [ctypes.c_void_p, _f32p, _f32p]


def load_ref():
    """The reference build, or None if oracle/_ref has not been built."""
    global _ref
    if _ref is not None:
        return _ref
    so = os.path.join(HERE, "_ref", "libmpeghref.so")
    if not os.path.exists(so):
        if os.path.isdir("/root/reference/decoder"):
            build("ref")
        else:
            return None
    lib = ctypes.CDLL(so)
    _bind_harnesses(lib)
    _ref = lib
    return lib


# ---- object renderer ------------------------------------------------------------------------------
OBJ_MAX_LS, OBJ_MAX_OBJ = 44, 96


class ObjLayout(ctypes.Structure):
    """oracle.h oracle_obj_layout"""
    _fields_ = [("n_spk", ctypes.c_int), ("n_non_lfe", ctypes.c_int), ("n_imag", ctypes.c_int),
                ("n_tri", ctypes.c_int), ("height_present", ctypes.c_int),
                ("lfe", ctypes.c_int * OBJ_MAX_LS), ("tri", (ctypes.c_int * 3) * OBJ_MAX_LS),
                ("inv", (ctypes.c_float * 9) * OBJ_MAX_LS), ("cart", (ctypes.c_float * 3) * OBJ_MAX_LS),
                ("dmx", (ctypes.c_float * OBJ_MAX_LS) * OBJ_MAX_LS)]


class ObjState(ctypes.Structure):
    _fields_ = [("initial_gains", (ctypes.c_float * OBJ_MAX_LS) * OBJ_MAX_OBJ),
                ("first_frame", ctypes.c_int * OBJ_MAX_OBJ)]


OBJ_SIDE_DTYPE = np.dtype([("vec_c", np.float32, 3), ("gain", np.float32), ("spread_w", np.float32),
                           ("spread_h", np.float32), ("tan_alpha", np.float32), ("pad0", np.float32),
                           ("p0", np.float32, 3), ("pad1", np.float32), ("v", np.float32, 3), ("pad2", np.float32)])
assert OBJ_SIDE_DTYPE.itemsize == 64


def obj_layout_from_arrays(a):
    """dict of arrays (counts, lfe, tri, inv, cart, dmx as ref_obj_layout returns them) -> ObjLayout"""
    L = ObjLayout()
    c = a["counts"]
    L.n_spk, L.n_non_lfe, L.n_imag, L.n_tri, L.height_present = (int(c[i]) for i in range(5))
    ctypes.memmove(L.lfe, np.ascontiguousarray(a["lfe"], np.int32).ctypes.data, 4 * OBJ_MAX_LS)
    ctypes.memmove(L.tri, np.ascontiguousarray(a["tri"], np.int32).ctypes.data, 12 * OBJ_MAX_LS)
    ctypes.memmove(L.inv, np.ascontiguousarray(a["inv"], np.float32).ctypes.data, 36 * OBJ_MAX_LS)
    ctypes.memmove(L.cart, np.ascontiguousarray(a["cart"], np.float32).ctypes.data, 12 * OBJ_MAX_LS)
    ctypes.memmove(L.dmx, np.ascontiguousarray(a["dmx"], np.float32).ctypes.data, 4 * OBJ_MAX_LS * OBJ_MAX_LS)
    return L


def ref_obj_layout_arrays(ref, h):
    a = {"counts": np.zeros(8, np.int32), "lfe": np.zeros(OBJ_MAX_LS, np.int32),
         "tri": np.zeros((OBJ_MAX_LS, 3), np.int32), "inv": np.zeros((OBJ_MAX_LS, 9), np.float32),
         "cart": np.zeros((OBJ_MAX_LS, 3), np.float32), "dmx": np.zeros((OBJ_MAX_LS, OBJ_MAX_LS), np.float32)}
    ref.ref_obj_layout(h, a["counts"], a["lfe"], a["tri"].reshape(-1), a["inv"].reshape(-1), a["cart"].reshape(-1),
                       a["dmx"].reshape(-1))
    return a


def obj_layout_golden(cicp):
    """Layout tables dumped from the reference init (tests/golden/obj_layouts.npz)."""
    g = np.load(os.path.join(ROOT, "tests", "golden", "obj_layouts.npz"))
    return {k: g[f"cicp{cicp}_{k}"] for k in ("counts", "lfe", "tri", "inv", "cart", "dmx")}


def oracle_obj_side(params):
    """params [n_obj, 7] (az, el, radius, gain, spread w/h/d) -> structured side array (host trig)."""
    params = np.ascontiguousarray(params, dtype=np.float32)
    side = np.zeros(params.shape[0], OBJ_SIDE_DTYPE)
    lib = load_oracle()
    for o in range(params.shape[0]):
        lib.oracle_obj_trig(params[o].ctypes.data_as(ctypes.c_void_p),
                            ctypes.c_void_p(side.ctypes.data + 64 * o))
    return side


def oracle_obj_state_new(n_first=24):
    st = ObjState()
    for i in range(min(n_first, OBJ_MAX_OBJ)):
        st.first_frame[i] = 1
    return st


def oracle_obj_render(L, st, side, x):
    """x [n_obj, 1024] -> out [n_spk, 1024]"""
    x = np.ascontiguousarray(x, dtype=np.float32)
    out = np.zeros((L.n_spk, 1024), np.float32)
    load_oracle().oracle_obj_render(ctypes.byref(L), ctypes.byref(st), ctypes.c_void_p(side.ctypes.data), x.shape[0],
                                    x, out)
    return out


def subframe_sets(present, num_iter):
    """Which metadata set each sub-frame call renders with: the sub_frame_number bookkeeping of
    impeghd_obj_md_dec_ren_process (decoder/ia_core_coder_decode_main.c:1388-1404), minus one."""
    n, used = 0, []
    for i, p in enumerate(present):
        if p:
            n += 1
            if n > num_iter:
                n %= num_iter
        elif i == 0:
            n = num_iter
        used.append(n - 1)
    return used


def oracle_obj_render_subframes(L, st, params, present, frame_len, x):
    """params [n_sub, n_obj, 7]; x [n_obj, 1024] -> out [n_spk, 1024] (the frame rendered in n_sub calls)"""
    x = np.ascontiguousarray(x, dtype=np.float32)
    out = np.zeros((L.n_spk, 1024), np.float32)
    n_sub = 1024 // frame_len
    for i, k in enumerate(subframe_sets(present, n_sub)):
        side = oracle_obj_side(params[k])
        load_oracle().oracle_obj_render_sub(ctypes.byref(L), ctypes.byref(st), ctypes.c_void_p(side.ctypes.data),
                                            x.shape[0], x, out, frame_len, i * frame_len, k)
    return out


def oracle_obj_gains(L, side_one):
    fg = np.zeros(OBJ_MAX_LS, np.float32)
    load_oracle().oracle_obj_gains(ctypes.byref(L), ctypes.c_void_p(side_one.ctypes.data), fg)
    return fg[:L.n_non_lfe].copy()


# ---- HOA renderer -----------------------------------------------------------------------------------
def oracle_hoa_render(M, x, nfc=None, nfc_state=None, clip=1):
    """M [n_out, n_coef], x [n_coef, n]; nfc [n_coef, 17] + nfc_state [n_coef, 14] (advanced in place) or None."""
    M = np.ascontiguousarray(M, dtype=np.float32)
    x = np.ascontiguousarray(x, dtype=np.float32)
    out = np.zeros((M.shape[0], x.shape[1]), np.float32)
    load_oracle().oracle_hoa_render(M, M.shape[0], M.shape[1], None if nfc is None else nfc.ctypes.data,
                                    None if nfc is None else nfc_state.ctypes.data, x, x.shape[1], clip, out)
    return out


def ref_hoa_ren_tables(ref, h, order):
    info = np.zeros(4, np.int32)
    m = np.zeros(64 * 64, np.float32)
    ref.ref_hoa_ren_matrix(h, info, m)
    n_out, rows, cols = int(info[0]), int(info[1]), int(info[2])
    nc = (order + 1) ** 2
    nfc = np.zeros((nc, 17), np.float32)
    ref.ref_hoa_ren_nfc_coeffs(h, nc, nfc.reshape(-1))
    return m[:rows * cols].reshape(rows, cols).copy(), nfc, n_out


_ref_ns = None


def load_ref_ns():
    """Reference build with the needed `static` functions exported (libmpeghref_ns.so), or None."""
    global _ref_ns
    if _ref_ns is not None:
        return _ref_ns
    so = os.path.join(HERE, "_ref", "libmpeghref_ns.so")
    if not os.path.exists(so):
        if os.path.isdir("/root/reference/decoder"):
            build("ref")
        else:
            return None
    lib = ctypes.CDLL(so)
    _bind_harnesses(lib)
    lib.ref_samples_sat.argtypes = [_f32p, ctypes.c_int, ctypes.c_int, ctypes.c_int, _i32p, _i32p, _u8p]
    lib.ref_samples_sat.restype = ctypes.c_int
    _i32q = np.ctypeslib.ndpointer(dtype=np.int32, flags="C_CONTIGUOUS")
    lib.ref_drctd_create.restype = ctypes.c_void_p
    lib.ref_drctd_create.argtypes = [ctypes.c_int, _i32q, ctypes.c_int, ctypes.c_void_p]
    lib.ref_drctd_destroy.argtypes = [ctypes.c_void_p]
    lib.ref_drctd_process.argtypes = [ctypes.c_void_p, _f32p, _f32p]
    lib.ref_drctd_process.restype = ctypes.c_int
    vp, ci = ctypes.c_void_p, ctypes.c_int
    lib.ref_chain_create_fd.restype = vp
    lib.ref_chain_create_fd.argtypes = [ci] * 6
    lib.ref_chain_create_hoa.restype = vp
    lib.ref_chain_create_hoa.argtypes = [ci, _i32q, ci, ci, ci, ctypes.c_float, ci, ci, ci, ctypes.POINTER(ci)]
    lib.ref_chain_create_obj.restype = vp
    lib.ref_chain_create_obj.argtypes = [ci] * 7 + [ctypes.POINTER(ci)]
    lib.ref_chain_create_bin.restype = vp
    lib.ref_chain_create_bin.argtypes = [ci] * 5 + [vp] * 6 + [ci, ci, ci, ctypes.POINTER(ci)]
    lib.ref_chain_set_out.argtypes = [vp, ci, ci]
    lib.ref_chain_stream_handle.restype = vp
    lib.ref_chain_stream_handle.argtypes = [vp, ci, ci]
    for n in ("ref_chain_threads", "ref_chain_rec_bytes", "ref_chain_frames_per_record"):
        getattr(lib, n).argtypes = [vp]
    lib.ref_chain_run.restype = ctypes.c_double
    lib.ref_chain_run.argtypes = [vp, vp, vp, ci, vp, vp]
    lib.ref_chain_destroy.argtypes = [vp]
    lib.ref_hoa_ren_matrix.argtypes = [vp, _i32q, _f32p]
    _ref_ns = lib
    return lib


class RefChain:
    """oracle/ref_chain.c: the unmodified reference's stage functions chained per configuration over S streams on a
    pool of pinned worker threads.  kind: "fd", "hoa", "bin", "obj"."""

    def __init__(self, kind, S, n_threads=1, pcm_bits=24, start_delay=0, **kw):
        self.ns = load_ref_ns()
        assert self.ns is not None, "oracle/_ref/libmpeghref_ns.so not built"
        err = ctypes.c_int(0)
        self.kind, self.S = kind, S
        if kind == "fd":
            self.n_in = kw["n_ch"]
            self.h = self.ns.ref_chain_create_fd(S, self.n_in, int(kw.get("tail", 1)), pcm_bits, start_delay, n_threads)
        elif kind == "hoa":
            cfg = np.array(list(kw["cfg"]) + [0, 0], np.int32)
            cicp, order, lfe, radius = kw["ren"]
            self.n_in = int(cfg[1])
            self.h = self.ns.ref_chain_create_hoa(S, cfg, kw["side_bytes"], cicp, lfe, radius, pcm_bits, start_delay,
                                                  n_threads, ctypes.byref(err))
            info, m = np.zeros(4, np.int32), np.zeros(64 * 64, np.float32)
            self.ns.ref_hoa_ren_matrix(self.ns.ref_chain_stream_handle(self.h, 0, 1), info, m)
            self.n_out = int(info[0])
            assert self.ns.ref_chain_set_out(self.h, self.n_out, n_threads) == 0
        elif kind == "obj":
            self.n_in = kw["n_obj"]
            self.h = self.ns.ref_chain_create_obj(S, self.n_in, kw["cicp"], kw["n_spk"], pcm_bits, start_delay,
                                                  n_threads, ctypes.byref(err))
        elif kind == "bin":
            f = kw["filters"]  # list of dicts (sets)
            self._keep = [np.ascontiguousarray(np.stack([d[k] for d in f]), np.float32)
                          for k in ("taps_direct", "taps_diffuse", "direct_fc", "diffuse_fc", "weight")]
            td = self._keep[0]
            self.n_in = td.shape[2]
            sets = kw.get("sets")
            self._sets = None if sets is None else np.ascontiguousarray(sets, np.int32)
            self.h = self.ns.ref_chain_create_bin(S, self.n_in, td.shape[3], self._keep[1].shape[1], td.shape[0],
                                                  None if self._sets is None else self._sets.ctypes.data,
                                                  *[a.ctypes.data for a in self._keep], pcm_bits, start_delay,
                                                  n_threads, ctypes.byref(err))
        else:
            raise ValueError(kind)
        assert self.h and err.value == 0, (kind, hex(err.value & 0xFFFFFFFF))
        self.rec_bytes = self.ns.ref_chain_rec_bytes(self.h)
        self.fpr = self.ns.ref_chain_frames_per_record(self.h)
        self.n_threads = self.ns.ref_chain_threads(self.h)

    def run(self, x, side=None, want_pcm=True):
        """x [F, S, n_in, 1024]; side [F, S, ...] or None -> (seconds, pcm [F/fpr, S, rec_bytes], bytes [F/fpr, S])"""
        x = np.ascontiguousarray(x, np.float32)
        F = x.shape[0]
        assert x.shape[1:] == (self.S, self.n_in, 1024), x.shape
        side = None if side is None else np.ascontiguousarray(side)
        pcm = np.zeros((F // self.fpr, self.S, self.rec_bytes), np.uint8) if want_pcm else None
        nb = np.zeros((F // self.fpr, self.S), np.int32) if want_pcm else None
        t = self.ns.ref_chain_run(self.h, x.ctypes.data, None if side is None else side.ctypes.data, F,
                                  None if pcm is None else pcm.ctypes.data, None if nb is None else nb.ctypes.data)
        assert t >= 0, "reference stage reported an error"
        return t, pcm, nb

    def close(self):
        if self.h:
            self.ns.ref_chain_destroy(self.h)
            self.h = None


def bits_equal(a, b):
    """Bit-for-bit equality of two float32 arrays (distinguishes -0.0 from +0.0, NaN payloads)."""
    a = np.ascontiguousarray(a, dtype=np.float32)
    b = np.ascontiguousarray(b, dtype=np.float32)
    return a.shape == b.shape and bool(np.array_equal(a.view(np.uint32), b.view(np.uint32)))


# ---- binaural renderer ------------------------------------------------------------------------------
class OracleBinaural:
    """Per-stream state + filter spectra of the oracle's binaural renderer (oracle_binaural.c)."""

    def __init__(self, taps_direct, taps_diffuse, direct_fc, diffuse_fc, weight):
        """taps_direct [2, n_in, len], taps_diffuse [n_blocks, 2, len], direct_fc [2, n_in], diffuse_fc [2, n_blocks]"""
        lib = load_oracle()
        self.n_in, self.len = taps_direct.shape[1], taps_direct.shape[2]
        self.n_blocks = taps_diffuse.shape[0]
        fft = 2
        while fft < 2 * self.len:
            fft <<= 1
        self.fft, self.B = fft, fft // 2
        self.h_direct = np.zeros((2, self.n_in, fft), np.float32)
        self.h_diffuse = np.zeros((max(self.n_blocks, 1), 2, fft), np.float32)
        for o in range(2):
            for i in range(self.n_in):
                assert lib.oracle_binaural_spectrum(np.ascontiguousarray(taps_direct[o, i], np.float32),
                                                    min(self.len, self.B), fft, self.h_direct[o, i]) == 0
        for b in range(self.n_blocks):
            for o in range(2):
                assert lib.oracle_binaural_spectrum(np.ascontiguousarray(taps_diffuse[b, o], np.float32),
                                                    min(self.len, self.B), fft, self.h_diffuse[b, o]) == 0
        self.direct_len = (np.minimum(np.asarray(direct_fc, np.float32), np.float32(1.0)) *
                           np.float32(self.B)).astype(np.int32).reshape(2, self.n_in).copy()
        self.diffuse_len = np.zeros((max(self.n_blocks, 1), 2), np.int32)
        if self.n_blocks:
            dl = (np.minimum(np.asarray(diffuse_fc, np.float32), np.float32(1.0)) * np.float32(self.B)).astype(np.int32)
            self.diffuse_len[:self.n_blocks] = dl.reshape(2, self.n_blocks).T
        self.weight = np.ascontiguousarray(weight, np.float32)
        self.reset()

    def reset(self):
        self.prev_in = np.zeros((self.n_blocks + 1, self.fft), np.float32)
        self.tail = np.zeros((2, self.B), np.float32)
        self.write_idx = np.zeros(1, np.int32)

    def clone_state(self):
        import copy
        c = copy.copy(self)
        c.reset()
        return c

    def block(self, x):
        """x [n_in, B] -> [2, B]"""
        x = np.ascontiguousarray(x, np.float32)
        out = np.zeros((2, self.B), np.float32)
        rc = load_oracle().oracle_binaural_block(self.fft, self.n_in, self.n_blocks, self.h_direct.reshape(-1),
                                                 self.h_diffuse.reshape(-1), self.direct_len.reshape(-1),
                                                 self.diffuse_len.reshape(-1), self.weight, self.prev_in.reshape(-1),
                                                 self.write_idx, self.tail.reshape(-1), x.reshape(-1), out.reshape(-1))
        assert rc == 0
        return out


# ---- HOA spatial decoder ------------------------------------------------------------------------------
HOASP_TABLE_KEYS = ("order_matrix", "fine", "coarse", "wins", "dyn_pow", "dyn_win", "pow2")


def ref_hoasp_tables(ref, h):
    """Tables designed by the reference init -> dict (info = low_order_mat_sz, vec_start, n_coef, 0)."""
    t = {"info": np.zeros(4, np.int32), "order_matrix": np.zeros(49 * 49, np.float32),
         "fine": np.zeros(900 * 49, np.float32), "coarse": np.zeros(49 * 49, np.float32),
         "wins": np.zeros(4 * 1024, np.float32), "dyn_pow": np.zeros(7 * 1024, np.float32),
         "dyn_win": np.zeros(1024, np.float32), "pow2": np.zeros(128, np.float32)}
    ref.ref_hoasp_tables(h, t["info"], *[t[k] for k in HOASP_TABLE_KEYS])
    return t


class OracleHoaspTables(ctypes.Structure):
    _fields_ = [(n, ctypes.c_int) for n in ("n_coef", "n_transport", "n_addnl", "low_order", "vec_start",
                                            "coded_v_vec_length", "spat_interp_time", "pred_dir_sigs",
                                            "num_bits_per_sf", "use_phase_shift_decorr")] + \
               [(n, ctypes.POINTER(ctypes.c_float)) for n in HOASP_TABLE_KEYS]


class OracleHoasp:
    """oracle_hoasp.c: one stream's spatial decoder (tables from the reference init or the golden fixture)."""
    STATE_BYTES = 4 * (16 + 16 + 4 + 16 + 16 + 17 + 17 + 49 + 4 * 49 + 4 * 49 + 16 * 49)

    def __init__(self, cfg, tables):
        self.cfg, self.t = cfg, {k: np.ascontiguousarray(v) for k, v in tables.items()}
        tb = OracleHoaspTables()
        tb.n_coef = (cfg[0] + 1) ** 2
        tb.n_transport, tb.n_addnl = cfg[1], cfg[2]
        tb.low_order, tb.vec_start = int(self.t["info"][0]), int(self.t["info"][1])
        tb.coded_v_vec_length, tb.spat_interp_time, tb.pred_dir_sigs, tb.num_bits_per_sf = cfg[4], cfg[5], cfg[7], cfg[8]
        tb.use_phase_shift_decorr = cfg[9]
        for k in HOASP_TABLE_KEYS:
            setattr(tb, k, self.t[k].ctypes.data_as(ctypes.POINTER(ctypes.c_float)))
        self.tb = tb
        self.n_coef = tb.n_coef
        self.state = ctypes.create_string_buffer(self.STATE_BYTES)
        load_oracle().oracle_hoasp_state_init(self.state)

    def process(self, side, x):
        x = np.ascontiguousarray(x, np.float32)
        out = np.zeros((self.n_coef, 1024), np.float32)
        side = np.ascontiguousarray(side)
        rc = load_oracle().oracle_hoasp_process(ctypes.byref(self.tb), self.state, side.ctypes.data, x.reshape(-1),
                                                out.reshape(-1))
        assert rc == 0
        return out


# ---- format converter ---------------------------------------------------------------------------------
FC_STATE_BYTES = 4 * 24 * (256 + 256 + 58 + 58)


class OracleFc:
    """oracle_fc.c: one stream's format converter; D [n_out, n_in, 257] from the reference init / golden fixture."""

    def __init__(self, D):
        self.D = np.ascontiguousarray(D, np.float32)
        self.n_out, self.n_in = self.D.shape[:2]
        self.win = np.ascontiguousarray(golden_rom()["sine_256"], np.float32)
        self.state = ctypes.create_string_buffer(FC_STATE_BYTES)
        load_oracle().oracle_fc_state_init(self.state)

    def process(self, x):
        x = np.ascontiguousarray(x, np.float32)
        out = np.zeros((self.n_out, 1024), np.float32)
        load_oracle().oracle_fc_process(self.D.reshape(-1), self.n_in, self.n_out, self.win, self.state, x.reshape(-1),
                                        out.reshape(-1))
        return out


def ref_fc_new(ref, cicp_in, cicp_out, sample_rate=48000):
    """-> (handle, D [n_out, n_in, 257]) designed by the reference init"""
    err = ctypes.c_int()
    h = ref.ref_fc_create(cicp_in, cicp_out, sample_rate, ctypes.byref(err))
    assert err.value == 0, hex(err.value & 0xFFFFFFFF)
    info = np.zeros(2, np.int32)
    ref.ref_fc_info(h, info)
    D = np.zeros((int(info[1]), int(info[0]), 257), np.float32)
    ref.ref_fc_get_matrix(h, D.reshape(-1))
    return h, D


# ---- TNS ------------------------------------------------------------------------------------------------
def oracle_tns_apply(spec, islong, n_filt, filt, sfb_tbl, tns_max_band, nbands):
    """Restatement-level TNS of one channel: returns (filtered copy, list of resolved filters).
    filt [n_win, 3, 20] ints as ref_tns_apply takes them; each resolved filter = (start, size, inc, order, lpc[16])
    with start relative to the 1024-bin frame."""
    lib = load_oracle()
    spec = np.ascontiguousarray(spec, np.float32).copy()
    n_win, bins = (1, 1024) if islong else (8, 128)
    sfb_tbl = np.ascontiguousarray(sfb_tbl, np.int16)
    resolved = []
    for w in range(n_win):
        for f in range(int(n_filt[w])):
            order, sb, eb, direction, res = (int(v) for v in filt[w, f, :5])
            if order == 0:
                continue
            a, b = np.zeros(1, np.int32), np.zeros(1, np.int32)
            assert lib.oracle_tns_range(sb, eb, tns_max_band, nbands, sfb_tbl, len(sfb_tbl), a, b) == 0
            size = int(b[0] - a[0])
            if size <= 0:
                continue
            lpc = np.zeros(16, np.float32)
            lib.oracle_tns_lpc(order, res, np.ascontiguousarray(filt[w, f, 5:20], np.int32), lpc)
            start = w * bins + int(a[0])
            inc = -1 if direction else 1
            seg = spec[start:start + size]
            lib.oracle_tns_filter(seg.ctypes.data, size, inc, order, lpc)
            resolved.append((start, size, inc, order, lpc))
    return spec, resolved


# ---- resampler ----------------------------------------------------------------------------------------
class OracleResampler:
    def __init__(self, up, down, f2, f3):
        self.up, self.down = up, down
        self.f_up = np.ascontiguousarray(f3 if up == 3 else f2, np.float32)
        self.f_down = np.ascontiguousarray(f2, np.float32)
        self.state = ctypes.create_string_buffer(4 * (32 + 64 + 4))
        load_oracle().oracle_rs_state_init(self.state)

    def process(self, x):
        x = np.ascontiguousarray(x, np.float32)
        out = np.zeros(1024 * self.up // self.down, np.float32)
        load_oracle().oracle_rs_process(self.state, self.up, self.down, self.f_up, self.f_down, x, out)
        return out


# ---- time-domain multi-band DRC -------------------------------------------------------------------------
DRC_BANK_DTYPE = np.dtype([("n_bands", np.int32), ("n_cascade", np.int32), ("f", np.float32, (8, 5)),
                           ("ap", np.float32, (9, 5))])  # oracle_drc_bank / mpeghb200_drc_td_bank


# ---- LTPF -----------------------------------------------------------------------------------------------
class OracleLtpf:
    """One channel's LTPF (oracle/oracle_ltpf.c) with carried state."""

    def __init__(self, sample_rate=48000):
        self.lib = load_oracle()
        self.cfg = (ctypes.c_int * 5)()
        self.lib.oracle_ltpf_cfg_init(self.cfg, sample_rate)
        self.state = ctypes.create_string_buffer(4 * (4 + 6 + 1600))
        self.lib.oracle_ltpf_state_init(self.state)

    def process(self, data_present, pitch_lag_idx, gain_idx, x):
        y = np.ascontiguousarray(x, np.float32).copy()
        self.lib.oracle_ltpf_process(self.cfg, self.state, int(data_present), int(pitch_lag_idx), int(gain_idx), y)
        return y
