"""ctypes binding of libmpeghb200.so (see include/mpeghb200.h).  No fallbacks: a missing library raises."""
import ctypes
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
SO_PATH = os.path.join(HERE, "libmpeghb200.so")

OK = 0
ERR_NO_DEVICE = -0x7FFF  # 0xFFFF8001 as int32
FRAME_LEN = 1024

ONLY_LONG, LONG_START, EIGHT_SHORT, LONG_STOP, STOP_START = range(5)


def _i32(x):
    return ctypes.c_int32(x).value


class MpeghError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"mpeghb200 status 0x{code & 0xFFFFFFFF:08X}: {msg}")
        self.code = code


def build(verbose=False):
    """Compile libmpeghb200.so in-tree for sm_100a (nvcc cross-compiles without a GPU)."""
    subprocess.run(["make", "-s" if not verbose else "-k", "-C", os.path.join(HERE, "csrc"), "-j8"], check=True)
    return SO_PATH


def fd_side(seq, shape=0, shape_prev=0, prev_sym=0, curr_sym=0):
    """MPEGHB200_FD_SIDE packing (numpy-broadcastable)."""
    seq, shape, shape_prev, prev_sym, curr_sym = (np.asarray(v, dtype=np.uint32) for v in
                                                  (seq, shape, shape_prev, prev_sym, curr_sym))
    return ((seq & 7) | ((shape & 1) << 3) | ((shape_prev & 1) << 4) | ((prev_sym & 1) << 5) |
            ((curr_sym & 1) << 6)).astype(np.uint32)


class LimiterParams(ctypes.Structure):
    """mpeghb200_limiter_params"""
    _fields_ = [("n_ch", ctypes.c_int32), ("attack_samples", ctypes.c_int32), ("limiter_on", ctypes.c_int32),
                ("threshold", ctypes.c_float), ("attack_const", ctypes.c_float), ("release_const", ctypes.c_float)]


OBJ_MAX_LS = 44


class ObjLayoutDesc(ctypes.Structure):
    """mpeghb200_obj_layout_desc"""
    _fields_ = [("n_spk", ctypes.c_int32), ("n_non_lfe", ctypes.c_int32), ("n_imag", ctypes.c_int32),
                ("n_tri", ctypes.c_int32), ("height_present", ctypes.c_int32),
                ("lfe", ctypes.c_int32 * OBJ_MAX_LS), ("tri", (ctypes.c_int32 * 3) * OBJ_MAX_LS),
                ("inv", (ctypes.c_float * 9) * OBJ_MAX_LS), ("cart", (ctypes.c_float * 3) * OBJ_MAX_LS),
                ("dmx", (ctypes.c_float * OBJ_MAX_LS) * OBJ_MAX_LS)]

    @classmethod
    def from_arrays(cls, a):
        """a: dict with counts[>=5], lfe[44], tri[44,3], inv[44,9], cart[44,3], dmx[44,44]"""
        d = cls()
        c = a["counts"]
        d.n_spk, d.n_non_lfe, d.n_imag, d.n_tri, d.height_present = (int(c[i]) for i in range(5))
        for name, dt in (("lfe", np.int32), ("tri", np.int32), ("inv", np.float32), ("cart", np.float32),
                         ("dmx", np.float32)):
            src = np.ascontiguousarray(a[name], dtype=dt)
            ctypes.memmove(getattr(d, name), src.ctypes.data, src.nbytes)
        return d


class HoaSpatialDesc(ctypes.Structure):
    """mpeghb200_hoa_spatial_desc"""
    _fields_ = [(n, ctypes.c_int32) for n in ("order", "n_transport", "n_addnl_coders", "low_order_mat_sz",
                                              "vec_start", "coded_v_vec_length", "spat_interp_time", "pred_dir_sigs",
                                              "num_bits_per_sf", "use_phase_shift_decorr")] + \
               [(n, ctypes.c_void_p) for n in ("order_matrix", "fine_grid_mode_matrix", "coarse_grid_mode_matrix",
                                               "fade_windows", "dyn_win_pow_table", "dyn_win_func", "pow2_table")]


class ChainDesc(ctypes.Structure):
    """mpeghb200_chain_desc"""
    _fields_ = [(n, ctypes.c_int32) for n in ("n_streams", "core_is_spectra", "n_bed_ch", "n_obj", "n_hoa_transport",
                                              "pcm_bits", "limiter", "start_delay")] + \
               [("loudness_gain", ctypes.c_float), ("max_frames_per_call", ctypes.c_int32)] + \
               [(n, ctypes.c_void_p) for n in ("obj_layout", "hoa_spatial", "hoa_renderer", "format_conv", "binaural",
                                               "out_ch_map", "binaural_set_of_stream")]


class ChainFrame(ctypes.Structure):
    """mpeghb200_chain_frame (host or device pointers)"""
    _fields_ = [(n, ctypes.c_void_p) for n in ("core", "fd_side", "obj_side", "hoa_side", "pcm", "pcm_bytes")]


LTPF_SIDE_DTYPE = np.dtype([("unit", "<i4"), ("pitch_lag_idx", "<u2"), ("data_present", "u1"), ("gain_idx", "u1")])
DRC_SB_CHANNEL_DTYPE = np.dtype([("first_seq", np.int32), ("band_count", np.int32), ("weight_row", np.int32),
                                 ("reserved", np.int32)])

DRC_TD_BANK_DTYPE = np.dtype([("n_bands", np.int32), ("n_cascade", np.int32), ("f", np.float32, (8, 5)),
                              ("ap", np.float32, (9, 5))])
DRC_TD_CHANNEL_DTYPE = np.dtype([("bank", np.int32), ("first_seq", np.int32)])

MCT_BOX_DTYPE = np.dtype([("unit_l", np.int32), ("unit_r", np.int32), ("op", np.int32), ("n_seg", np.int32),
                          ("start", np.int16, 64), ("len", np.int16, 64), ("coef", np.int16, 64)])

TNS_CHANNEL_DTYPE = np.dtype([
    ("unit", "<i4"), ("is_long", "<i4"), ("max_sfb", "<i4"), ("igf_active", "<i4"), ("igf_after_tns_synth", "<i4"),
    ("igf_sfb_start", "<i4"), ("igf_sfb_stop", "<i4"), ("n_filt", "<i4", 8), ("coef_res", "<i4", 8),
    ("filt", [("order", "<i4"), ("start_band", "<i4"), ("stop_band", "<i4"), ("direction", "<i4"),
              ("coef", "<i2", 16)], (8, 3))])
TNS_FILTER_DTYPE = np.dtype([("unit", np.int32), ("start", np.int32), ("size", np.int32), ("inc", np.int32),
                             ("order", np.int32), ("pad", np.int32, 3), ("lpc", np.float32, 16)])
OBJ_SIDE_DTYPE = np.dtype([("vec_c", np.float32, 3), ("gain", np.float32), ("spread_w", np.float32),
                           ("spread_h", np.float32), ("tan_alpha", np.float32), ("pad0", np.float32),
                           ("p0", np.float32, 3), ("pad1", np.float32), ("v", np.float32, 3), ("pad2", np.float32)])


class Lib:
    """Thin object wrapper: one instance = one loaded library + (lazily) one context per device."""

    def __init__(self, path=SO_PATH):
        if not os.path.exists(path):
            raise FileNotFoundError(
                f"{path} not built: run `python -c 'import __graft_entry__ as g; g.build()'` "
                "(there is no CPU fallback)")
        self.c = ctypes.CDLL(path)
        c = self.c
        vp, i32, u64, sz = ctypes.c_void_p, ctypes.c_int32, ctypes.c_uint64, ctypes.c_size_t
        c.mpeghb200_version.restype = ctypes.c_char_p
        c.mpeghb200_create.argtypes = [ctypes.POINTER(vp), ctypes.c_int]
        c.mpeghb200_create.restype = i32
        c.mpeghb200_destroy.argtypes = [vp]
        c.mpeghb200_last_error.argtypes = [vp]
        c.mpeghb200_last_error.restype = ctypes.c_char_p
        c.mpeghb200_launch_count.argtypes = [vp]
        c.mpeghb200_launch_count.restype = u64
        c.mpeghb200_rom_table.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.POINTER(ctypes.c_int)]
        c.mpeghb200_rom_table.restype = ctypes.POINTER(ctypes.c_float)
        c.mpeghb200_rom_count.restype = ctypes.c_int
        c.mpeghb200_rom_name.argtypes = [ctypes.c_int]
        c.mpeghb200_rom_name.restype = ctypes.c_char_p
        c.mpeghb200_dev_alloc.argtypes = [vp, sz, ctypes.POINTER(vp)]
        c.mpeghb200_dev_alloc.restype = i32
        c.mpeghb200_dev_free.argtypes = [vp, vp]
        c.mpeghb200_dev_free.restype = i32
        c.mpeghb200_dev_memset.argtypes = [vp, vp, ctypes.c_int, sz]
        c.mpeghb200_dev_memset.restype = i32
        c.mpeghb200_h2d.argtypes = [vp, vp, vp, sz]
        c.mpeghb200_h2d.restype = i32
        c.mpeghb200_d2h.argtypes = [vp, vp, vp, sz]
        c.mpeghb200_d2h.restype = i32
        c.mpeghb200_sync.argtypes = [vp]
        c.mpeghb200_sync.restype = i32
        c.mpeghb200_host_alloc.argtypes = [vp, sz, ctypes.POINTER(vp)]
        c.mpeghb200_host_alloc.restype = i32
        c.mpeghb200_host_free.argtypes = [vp, vp]
        c.mpeghb200_host_free.restype = i32
        c.mpeghb200_fd_synth_dev.argtypes = [vp, vp, vp, vp, vp, ctypes.c_int, vp]
        c.mpeghb200_fd_synth_dev.restype = i32
        c.mpeghb200_fd_synth_frames_dev.argtypes = [vp, vp, vp, vp, vp, ctypes.c_int, ctypes.c_int, vp]
        c.mpeghb200_fd_synth_frames_dev.restype = i32
        c.mpeghb200_fd_open.argtypes = [vp, ctypes.c_int, ctypes.c_int, ctypes.POINTER(vp)]
        c.mpeghb200_fd_open.restype = i32
        c.mpeghb200_fd_execute.argtypes = [vp, vp, vp, vp]
        c.mpeghb200_fd_execute.restype = i32
        c.mpeghb200_fd_submit.argtypes = [vp, vp, vp, vp]
        c.mpeghb200_fd_submit.restype = i32
        c.mpeghb200_fd_wait.argtypes = [vp]
        c.mpeghb200_fd_wait.restype = i32
        c.mpeghb200_fd_reset_stream.argtypes = [vp, ctypes.c_int]
        c.mpeghb200_fd_reset_stream.restype = i32
        c.mpeghb200_fd_close.argtypes = [vp]
        c.mpeghb200_limiter_params_init.argtypes = [ctypes.POINTER(LimiterParams), ctypes.c_int, ctypes.c_int,
                                                    ctypes.c_int]
        c.mpeghb200_limiter_params_init.restype = i32
        c.mpeghb200_limiter_state_floats.argtypes = [ctypes.POINTER(LimiterParams)]
        c.mpeghb200_limiter_state_floats.restype = sz
        c.mpeghb200_limiter_state_init_dev.argtypes = [vp, ctypes.POINTER(LimiterParams), vp, ctypes.c_int, vp]
        c.mpeghb200_limiter_state_init_dev.restype = i32
        c.mpeghb200_limiter_dev.argtypes = [vp, ctypes.POINTER(LimiterParams), vp, vp, vp, vp, ctypes.c_int, vp]
        c.mpeghb200_limiter_dev.restype = i32
        c.mpeghb200_pcm_pack_dev.argtypes = [vp, vp, ctypes.c_int, ctypes.c_int, vp, vp, vp, vp, ctypes.c_int, vp]
        c.mpeghb200_pcm_pack_dev.restype = i32
        c.mpeghb200_obj_layout_create.argtypes = [vp, ctypes.POINTER(ObjLayoutDesc), ctypes.POINTER(vp)]
        c.mpeghb200_obj_layout_create.restype = i32
        c.mpeghb200_obj_layout_destroy.argtypes = [vp]
        c.mpeghb200_obj_side_from_polar.argtypes = [vp, ctypes.c_int, vp]
        c.mpeghb200_obj_state_floats.argtypes = [vp, ctypes.c_int]
        c.mpeghb200_obj_state_floats.restype = sz
        c.mpeghb200_obj_scratch_floats.argtypes = [vp, ctypes.c_int]
        c.mpeghb200_obj_scratch_floats.restype = sz
        c.mpeghb200_obj_state_init_dev.argtypes = [vp, vp, ctypes.c_int, vp, vp, ctypes.c_int, vp]
        c.mpeghb200_obj_state_init_dev.restype = i32
        c.mpeghb200_obj_render_dev.argtypes = [vp, vp, ctypes.c_int, vp, vp, vp, vp, vp, ctypes.c_int, vp]
        c.mpeghb200_obj_render_dev.restype = i32
        c.mpeghb200_obj_render_sub_dev.argtypes = [vp, vp, ctypes.c_int, vp, vp, vp, vp, vp, ctypes.c_int, ctypes.c_int,
                                                   ctypes.c_int, vp, vp]
        c.mpeghb200_obj_render_sub_dev.restype = i32
        c.mpeghb200_hoa_renderer_create.argtypes = [vp, vp, ctypes.c_int, ctypes.c_int, vp, ctypes.c_int,
                                                    ctypes.POINTER(vp)]
        c.mpeghb200_hoa_renderer_create.restype = i32
        c.mpeghb200_hoa_renderer_destroy.argtypes = [vp]
        c.mpeghb200_hoa_nfc_state_floats.argtypes = [vp]
        c.mpeghb200_hoa_nfc_state_floats.restype = sz
        c.mpeghb200_hoa_renderer_set_tensor_mode.argtypes = [vp, ctypes.c_int]
        c.mpeghb200_hoa_renderer_set_tensor_mode.restype = i32
        c.mpeghb200_hoa_render_dev.argtypes = [vp, vp, vp, vp, vp, ctypes.c_int, vp]
        c.mpeghb200_hoa_render_dev.restype = i32
        c.mpeghb200_binaural_create.argtypes = [vp, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, vp, vp, vp,
                                                vp, vp, ctypes.POINTER(vp)]
        c.mpeghb200_binaural_create.restype = i32
        c.mpeghb200_binaural_destroy.argtypes = [vp]
        c.mpeghb200_binaural_block_len.argtypes = [vp]
        c.mpeghb200_binaural_block_len.restype = ctypes.c_int
        c.mpeghb200_binaural_state_floats.argtypes = [vp]
        c.mpeghb200_binaural_state_floats.restype = sz
        c.mpeghb200_binaural_spectra.argtypes = [vp, vp, vp, vp]
        c.mpeghb200_binaural_spectra.restype = i32
        c.mpeghb200_binaural_block_dev.argtypes = [vp, vp, vp, ctypes.c_int, ctypes.c_float, ctypes.c_float, vp, vp,
                                                   vp, ctypes.c_int, vp]
        c.mpeghb200_binaural_block_dev.restype = i32
        c.mpeghb200_hoa_spatial_create.argtypes = [vp, ctypes.POINTER(HoaSpatialDesc), ctypes.POINTER(vp)]
        c.mpeghb200_hoa_spatial_create.restype = i32
        c.mpeghb200_hoa_spatial_destroy.argtypes = [vp]
        c.mpeghb200_hoa_spatial_num_coeffs.argtypes = [vp]
        c.mpeghb200_hoa_spatial_num_coeffs.restype = ctypes.c_int
        c.mpeghb200_hoa_spatial_state_bytes.argtypes = [vp]
        c.mpeghb200_hoa_spatial_state_bytes.restype = sz
        c.mpeghb200_hoa_spatial_state_init_dev.argtypes = [vp, vp, vp, ctypes.c_int, vp]
        c.mpeghb200_hoa_spatial_state_init_dev.restype = i32
        c.mpeghb200_hoa_spatial_dev.argtypes = [vp, vp, vp, vp, vp, vp, ctypes.c_int, vp]
        c.mpeghb200_hoa_spatial_dev.restype = i32
        c.mpeghb200_format_conv_create.argtypes = [vp, ctypes.c_int, ctypes.c_int, vp, ctypes.POINTER(vp)]
        c.mpeghb200_format_conv_create.restype = i32
        c.mpeghb200_format_conv_destroy.argtypes = [vp]
        c.mpeghb200_format_conv_state_floats.argtypes = [vp]
        c.mpeghb200_format_conv_state_floats.restype = sz
        c.mpeghb200_format_conv_dev.argtypes = [vp, vp, vp, vp, vp, ctypes.c_int, vp]
        c.mpeghb200_format_conv_dev.restype = i32
        c.mpeghb200_tns_lpc_from_coef.argtypes = [ctypes.c_int, ctypes.c_int, vp, vp]
        c.mpeghb200_tns_lpc_from_coef.restype = i32
        c.mpeghb200_tns_resolve.argtypes = [vp, vp, ctypes.c_int, ctypes.c_int, vp, ctypes.c_int, vp, ctypes.c_int, vp, vp]
        c.mpeghb200_tns_resolve.restype = i32
        c.mpeghb200_tns_dev.argtypes = [vp, vp, vp, vp, ctypes.c_int, vp]
        c.mpeghb200_tns_dev.restype = i32
        c.mpeghb200_resampler_create.argtypes = [vp, ctypes.c_int, ctypes.c_int, vp, vp, ctypes.POINTER(vp)]
        c.mpeghb200_resampler_create.restype = i32
        c.mpeghb200_resampler_destroy.argtypes = [vp]
        c.mpeghb200_resampler_state_floats.argtypes = [vp]
        c.mpeghb200_resampler_state_floats.restype = sz
        c.mpeghb200_resampler_out_len.argtypes = [vp]
        c.mpeghb200_resampler_out_len.restype = ctypes.c_int
        c.mpeghb200_resample_dev.argtypes = [vp, vp, vp, vp, vp, ctypes.c_int, vp]
        c.mpeghb200_resample_dev.restype = i32
        c.mpeghb200_bands_create.argtypes = [vp, vp, ctypes.c_int, vp, ctypes.c_int, ctypes.POINTER(vp)]
        c.mpeghb200_bands_create.restype = i32
        c.mpeghb200_bands_destroy.argtypes = [vp]
        c.mpeghb200_igf_dev.argtypes = [vp, vp, vp, vp, vp, vp, vp, ctypes.c_int, vp]
        c.mpeghb200_igf_dev.restype = i32
        c.mpeghb200_gain_subband_dev.argtypes = [vp, vp, vp, vp, vp, vp, ctypes.c_int, vp]
        c.mpeghb200_gain_subband_dev.restype = i32
        c.mpeghb200_drc_stft_dev.argtypes = [vp, vp, vp, vp, ctypes.c_int, vp, vp, ctypes.c_int, vp]
        c.mpeghb200_stft_dev.argtypes = [vp, vp, vp, vp, ctypes.c_int, ctypes.c_int, vp]
        c.mpeghb200_stft_dev.restype = i32
        c.mpeghb200_drc_stft_dev.restype = i32
        c.mpeghb200_drc_td_dev.argtypes = [vp, vp, vp, vp, vp, vp, ctypes.c_int, vp]
        c.mpeghb200_drc_td_dev.restype = i32
        c.mpeghb200_drc_td_state_floats.restype = sz
        c.mpeghb200_mct_resolve.argtypes = [ctypes.c_int] * 7 + [vp, vp, vp, ctypes.c_int, ctypes.c_int, vp]
        c.mpeghb200_mct_resolve.restype = i32
        c.mpeghb200_mct_dev.argtypes = [vp, vp, vp, vp, ctypes.c_int, vp]
        c.mpeghb200_mct_dev.restype = i32
        c.mpeghb200_mct_rom.argtypes = [vp, vp]
        c.mpeghb200_mct_rom.restype = None
        c.mpeghb200_bus_add_dev.argtypes = [vp, vp, vp, vp, sz, vp]
        c.mpeghb200_bus_add_dev.restype = i32
        c.mpeghb200_interleave_dev.argtypes = [vp, vp, vp, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, vp]
        c.mpeghb200_interleave_dev.restype = i32
        c.mpeghb200_channel_mask_dev.argtypes = [vp, vp, vp, ctypes.c_int, vp]
        c.mpeghb200_channel_mask_dev.restype = i32
        c.mpeghb200_ltpf_state_floats.argtypes = [ctypes.c_int]
        c.mpeghb200_ltpf_state_floats.restype = sz
        c.mpeghb200_ltpf_dev.argtypes = [vp, ctypes.c_int, vp, vp, vp, ctypes.c_int, vp]
        c.mpeghb200_ltpf_dev.restype = i32
        c.mpeghb200_gain_apply_dev.argtypes = [vp, vp, vp, vp, ctypes.c_int, vp]
        c.mpeghb200_gain_apply_dev.restype = i32
        c.mpeghb200_igf_stereo_dev.argtypes = [vp, vp, vp, vp, vp, vp, vp, vp, ctypes.c_int, vp]
        c.mpeghb200_igf_stereo_dev.restype = i32
        c.mpeghb200_igf_tnf_dev.argtypes = [vp, vp, vp, vp, vp, ctypes.c_int, vp]
        c.mpeghb200_igf_tnf_dev.restype = i32
        c.mpeghb200_igf_whitening_rom.argtypes = [vp]
        c.mpeghb200_igf_whitening_rom.restype = None
        c.mpeghb200_stereo_state_floats.argtypes = []
        c.mpeghb200_stereo_state_floats.restype = sz
        c.mpeghb200_stereo_dev.argtypes = [vp, vp, vp, vp, vp, ctypes.c_int, vp]
        c.mpeghb200_stereo_dev.restype = i32
        c.mpeghb200_stereo_save_dev.argtypes = [vp, vp, vp, vp, vp, ctypes.c_int, vp]
        c.mpeghb200_stereo_save_dev.restype = i32
        c.mpeghb200_stereo_mdst_rom.argtypes = [vp, vp]
        c.mpeghb200_stereo_mdst_rom.restype = None
        c.mpeghb200_limiter_pack_dev.argtypes = [vp, ctypes.POINTER(LimiterParams), vp, vp, vp, vp, ctypes.c_int, vp, vp,
                                                 vp, vp, ctypes.c_int, vp]
        c.mpeghb200_limiter_pack_scratch_floats.argtypes = [ctypes.POINTER(LimiterParams)]
        c.mpeghb200_limiter_pack_scratch_floats.restype = sz
        c.mpeghb200_limiter_pack_dev.restype = i32
        c.mpeghb200_pcm_pack_block_dev.argtypes = [vp, vp, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, vp, vp,
                                                   vp, vp, ctypes.c_int, vp]
        c.mpeghb200_pcm_pack_block_dev.restype = i32
        for name in ("obj_layout_num_speakers", "hoa_renderer_num_out", "hoa_renderer_num_coeffs", "binaural_num_in",
                     "hoa_spatial_num_transport", "format_conv_num_in", "format_conv_num_out", "chain_out_channels",
                     "chain_frames_per_record"):
            f = getattr(c, "mpeghb200_" + name)
            f.argtypes, f.restype = [vp], ctypes.c_int
        c.mpeghb200_chain_open.argtypes = [vp, ctypes.POINTER(ChainDesc), ctypes.POINTER(vp)]
        c.mpeghb200_chain_open.restype = i32
        c.mpeghb200_chain_close.argtypes = [vp]
        c.mpeghb200_chain_close.restype = None
        c.mpeghb200_chain_core_floats.argtypes = [vp]
        c.mpeghb200_chain_core_floats.restype = sz
        c.mpeghb200_chain_pcm_record_bytes.argtypes = [vp]
        c.mpeghb200_chain_pcm_record_bytes.restype = sz
        c.mpeghb200_chain_submit.argtypes = [vp, ctypes.POINTER(ChainFrame), ctypes.c_int]
        c.mpeghb200_chain_submit.restype = i32
        c.mpeghb200_chain_wait.argtypes = [vp]
        c.mpeghb200_chain_wait.restype = i32
        c.mpeghb200_chain_execute.argtypes = [vp, ctypes.POINTER(ChainFrame), ctypes.c_int]
        c.mpeghb200_chain_execute.restype = i32
        c.mpeghb200_chain_run_dev.argtypes = [vp, ctypes.POINTER(ChainFrame), ctypes.c_int, vp]
        c.mpeghb200_chain_run_dev.restype = i32
        self.ctx = None

    # ---- context ------------------------------------------------------------------------------
    def version(self):
        return self.c.mpeghb200_version().decode()

    def create(self, device=0):
        h = ctypes.c_void_p()
        st = self.c.mpeghb200_create(ctypes.byref(h), device)
        if st != OK:
            raise MpeghError(st, self.c.mpeghb200_last_error(None).decode())
        self.ctx = h
        return self

    def close(self):
        if self.ctx:
            self.c.mpeghb200_destroy(self.ctx)
            self.ctx = None

    def check(self, st):
        if st != OK:
            raise MpeghError(st, self.c.mpeghb200_last_error(self.ctx).decode())

    def launch_count(self):
        return int(self.c.mpeghb200_launch_count(self.ctx))

    def host_array(self, shape, dtype=np.float32):
        """numpy view of page-locked memory from mpeghb200_host_alloc (freed with host_free(arr))."""
        n = int(np.prod(shape)) * np.dtype(dtype).itemsize
        p = ctypes.c_void_p()
        self.check(self.c.mpeghb200_host_alloc(self.ctx, n, ctypes.byref(p)))
        buf = (ctypes.c_byte * n).from_address(p.value)
        arr = np.frombuffer(buf, dtype=dtype).reshape(shape)
        self._pinned = getattr(self, "_pinned", {})
        self._pinned[arr.ctypes.data] = p
        return arr

    def host_free(self, arr):
        p = self._pinned.pop(arr.ctypes.data)
        self.check(self.c.mpeghb200_host_free(self.ctx, p))

    # ---- constant tables (host side, no GPU needed) ------------------------------------------------
    def rom_tables(self, patched=True):
        out = {}
        for i in range(self.c.mpeghb200_rom_count()):
            n = ctypes.c_int()
            p = self.c.mpeghb200_rom_table(i, 1 if patched else 0, ctypes.byref(n))
            out[self.c.mpeghb200_rom_name(i).decode()] = np.ctypeslib.as_array(p, shape=(n.value,)).copy()
        return out

    # ---- FD synthesis -------------------------------------------------------------------------
    def fd_synth_dev(self, d_coef, d_side, d_overlap, d_out, n_units, stream=0):
        """Raw device pointers (ints), e.g. torch tensor .data_ptr(); stream = cudaStream_t as int."""
        self.check(self.c.mpeghb200_fd_synth_dev(self.ctx, d_coef, d_side, d_overlap, d_out, n_units, stream))

    def fd_synth_frames_dev(self, d_coef, d_side, d_overlap, d_out, n_units, n_frames, stream=0):
        """coef / side / out [n_frames][n_units][...]: n_frames consecutive frames of every unit in one launch."""
        self.check(self.c.mpeghb200_fd_synth_frames_dev(self.ctx, d_coef, d_side, d_overlap, d_out, n_units, n_frames,
                                                        stream))

    def fd_open(self, n_streams, n_ch):
        return FdSession(self, n_streams, n_ch)

    # ---- HOA renderer ---------------------------------------------------------------------------
    def hoa_renderer_create(self, matrix, nfc=None, clip=1):
        m = np.ascontiguousarray(matrix, dtype=np.float32)
        n = None if nfc is None else np.ascontiguousarray(nfc, dtype=np.float32)
        h = ctypes.c_void_p()
        self.check(self.c.mpeghb200_hoa_renderer_create(self.ctx, m.ctypes.data, m.shape[0], m.shape[1],
                                                        None if n is None else n.ctypes.data, clip, ctypes.byref(h)))
        return h

    def hoa_renderer_set_tensor_mode(self, h, on):
        self.check(self.c.mpeghb200_hoa_renderer_set_tensor_mode(h, int(on)))

    def hoa_renderer_destroy(self, h):
        self.c.mpeghb200_hoa_renderer_destroy(h)

    def hoa_nfc_state_floats(self, h):
        return int(self.c.mpeghb200_hoa_nfc_state_floats(h))

    def hoa_render_dev(self, h, d_in, d_out, d_state, n_streams, stream=0):
        self.check(self.c.mpeghb200_hoa_render_dev(self.ctx, h, d_in, d_out, d_state, n_streams, stream))

    # ---- TNS / resampler ---------------------------------------------------------------------------
    def tns_lpc_from_coef(self, order, coef_res, coef):
        """Host-side coefficient decode (no GPU needed) -> lpc[16]"""
        c16 = np.zeros(16, np.int16)
        c16[:order] = np.asarray(coef[:order], np.int16)
        lpc = np.zeros(16, np.float32)
        st = self.c.mpeghb200_tns_lpc_from_coef(order, coef_res, c16.ctypes.data, lpc.ctypes.data)
        if st != OK:
            raise MpeghError(st, "tns_lpc_from_coef: bad order or coefficient resolution")
        return lpc

    def tns_resolve(self, ch, sfb_top, tns_max_band):
        """ch: TNS_CHANNEL_DTYPE record -> list of chains (lists of TNS_FILTER_DTYPE records); raises MpeghError"""
        sfb = np.ascontiguousarray(sfb_top, np.int16)
        recs = np.zeros(24, TNS_FILTER_DTYPE)
        lens = np.zeros(8, np.int32)
        nf, nc = ctypes.c_int(), ctypes.c_int()
        buf = ch.tobytes()
        self.check(self.c.mpeghb200_tns_resolve(buf, sfb.ctypes.data, len(sfb), tns_max_band, recs.ctypes.data, 24,
                                                lens.ctypes.data, 8, ctypes.byref(nf), ctypes.byref(nc)))
        out, k = [], 0
        for i in range(nc.value):
            out.append([recs[k + j].copy() for j in range(int(lens[i]))])
            k += int(lens[i])
        return out

    def tns_dev(self, d_coef, d_filters, d_chain_first, n_chains, stream=0):
        self.check(self.c.mpeghb200_tns_dev(self.ctx, d_coef, d_filters, d_chain_first, n_chains, stream))

    # ---- multichannel coding tool -------------------------------------------------------------------
    def mct_resolve(self, signal_type, pred_dir, fullband, n_windows, bins_per_window, bands_per_window, sfb_offset,
                    coef_idx, mask, unit_l, unit_r):
        """-> one MCT_BOX_DTYPE record; raises MpeghError on out-of-range side info"""
        sfb = np.ascontiguousarray(sfb_offset, np.int16)
        c = np.ascontiguousarray(coef_idx, np.int32)
        m = np.ascontiguousarray(mask, np.int32)
        box = np.zeros(1, MCT_BOX_DTYPE)
        self.check(self.c.mpeghb200_mct_resolve(signal_type, pred_dir, fullband, n_windows, bins_per_window,
                                                bands_per_window, len(sfb), sfb.ctypes.data, c.ctypes.data,
                                                m.ctypes.data, unit_l, unit_r, box.ctypes.data))
        return box[0]

    def mct_dev(self, d_spec, d_boxes, d_group_first, n_groups, stream=0):
        self.check(self.c.mpeghb200_mct_dev(self.ctx, d_spec, d_boxes, d_group_first, n_groups, stream))

    def mct_rom(self):
        s, c = np.zeros(65, np.float32), np.zeros(65, np.float32)
        self.c.mpeghb200_mct_rom(s.ctypes.data, c.ctypes.data)
        return s, c

    # ---- joint-stereo tools -------------------------------------------------------------------------
    def bands_create(self, sfb_long, sfb_short):
        a = np.ascontiguousarray(sfb_long, np.int16)
        b = np.ascontiguousarray(sfb_short, np.int16)
        h = ctypes.c_void_p()
        self.check(self.c.mpeghb200_bands_create(self.ctx, a.ctypes.data, len(a), b.ctypes.data, len(b),
                                                 ctypes.byref(h)))
        return h

    def bands_destroy(self, h):
        self.c.mpeghb200_bands_destroy(h)

    def igf_dev(self, bands, d_coef, d_tiles, d_side, d_tnf_state, d_seed_out, n_units, stream=0):
        self.check(self.c.mpeghb200_igf_dev(self.ctx, bands, d_coef, d_tiles, d_side, d_tnf_state, d_seed_out,
                                            n_units, stream))

    def igf_stereo_dev(self, bands, d_coef, d_tiles, d_side, d_mask, d_tnf_state, d_seed_out, n_pairs, stream=0):
        self.check(self.c.mpeghb200_igf_stereo_dev(self.ctx, bands, d_coef, d_tiles, d_side, d_mask, d_tnf_state,
                                                   d_seed_out, n_pairs, stream))

    def igf_tnf_dev(self, bands, d_coef, d_side, d_tnf_state, n_units, stream=0):
        self.check(self.c.mpeghb200_igf_tnf_dev(self.ctx, bands, d_coef, d_side, d_tnf_state, n_units, stream))

    def gain_subband_dev(self, d_re, d_im, d_chan, d_gains, d_weights, n_channels, stream=0):
        self.check(self.c.mpeghb200_gain_subband_dev(self.ctx, d_re, d_im, d_chan, d_gains, d_weights, n_channels,
                                                     stream))

    def drc_stft_dev(self, d_audio, d_state, d_chan, n_sets, d_gains, d_weights, n_channels, stream=0):
        self.check(self.c.mpeghb200_drc_stft_dev(self.ctx, d_audio, d_state, d_chan, n_sets, d_gains, d_weights,
                                                 n_channels, stream))

    def stft_dev(self, d_audio, d_spec, d_state, to_frequency, n_channels, stream=0):
        """One half of the domain switcher: d_audio [n][1024] <-> d_spec [n][2048] (re | im planes)."""
        self.check(self.c.mpeghb200_stft_dev(self.ctx, d_audio, d_spec, d_state, int(to_frequency), n_channels, stream))

    def drc_td_dev(self, d_audio, d_state, d_chan, d_banks, d_gains, n_channels, stream=0):
        self.check(self.c.mpeghb200_drc_td_dev(self.ctx, d_audio, d_state, d_chan, d_banks, d_gains, n_channels, stream))

    def drc_td_state_floats(self):
        return int(self.c.mpeghb200_drc_td_state_floats())

    def bus_add_dev(self, d_a, d_b, d_out, n_floats, stream=0):
        self.check(self.c.mpeghb200_bus_add_dev(self.ctx, d_a, d_b, d_out, n_floats, stream))

    def interleave_dev(self, d_in, d_out, n_ch, n_samples, to_planar, n_streams, stream=0):
        """to_planar: [stream][n_samples][n_ch] -> [stream][n_ch][n_samples]; else the reverse."""
        self.check(self.c.mpeghb200_interleave_dev(self.ctx, d_in, d_out, n_ch, n_samples, int(to_planar), n_streams,
                                                   stream))

    def channel_mask_dev(self, d_audio, d_on, n_channels, stream=0):
        self.check(self.c.mpeghb200_channel_mask_dev(self.ctx, d_audio, d_on, n_channels, stream))

    def gain_apply_dev(self, d_audio, d_gains, d_seq_of_ch, n_channels, stream=0):
        self.check(self.c.mpeghb200_gain_apply_dev(self.ctx, d_audio, d_gains, d_seq_of_ch, n_channels, stream))

    def ltpf_state_floats(self, sample_rate=48000):
        return int(self.c.mpeghb200_ltpf_state_floats(sample_rate))

    def ltpf_dev(self, sample_rate, d_audio, d_side, d_state, n_units, stream=0):
        self.check(self.c.mpeghb200_ltpf_dev(self.ctx, sample_rate, d_audio, d_side, d_state, n_units, stream))

    def igf_whitening_rom(self):
        t = np.zeros((2, 64), np.float32)
        self.c.mpeghb200_igf_whitening_rom(t.ctypes.data)
        return t

    def stereo_state_floats(self):
        return int(self.c.mpeghb200_stereo_state_floats())

    def stereo_dev(self, h, d_coef, d_pairs, d_state, n_pairs, stream=0):
        self.check(self.c.mpeghb200_stereo_dev(self.ctx, h, d_coef, d_pairs, d_state, n_pairs, stream))

    def stereo_save_dev(self, h, d_coef, d_pairs, d_state, n_pairs, stream=0):
        self.check(self.c.mpeghb200_stereo_save_dev(self.ctx, h, d_coef, d_pairs, d_state, n_pairs, stream))

    def stereo_mdst_rom(self):
        curr, prev = np.zeros((4, 2, 2, 7), np.float32), np.zeros((2, 2, 7), np.float32)
        self.c.mpeghb200_stereo_mdst_rom(curr.ctypes.data, prev.ctypes.data)
        return curr, prev

    def resampler_create(self, up, down, f_up, f_down=None):
        fu = np.ascontiguousarray(f_up, np.float32)
        fd = None if f_down is None else np.ascontiguousarray(f_down, np.float32)
        h = ctypes.c_void_p()
        self.check(self.c.mpeghb200_resampler_create(self.ctx, up, down, fu.ctypes.data,
                                                     None if fd is None else fd.ctypes.data, ctypes.byref(h)))
        return h

    def resampler_destroy(self, h):
        self.c.mpeghb200_resampler_destroy(h)

    def resampler_state_floats(self, h):
        return int(self.c.mpeghb200_resampler_state_floats(h))

    def resampler_out_len(self, h):
        return int(self.c.mpeghb200_resampler_out_len(h))

    def resample_dev(self, h, d_in, d_out, d_state, n_units, stream=0):
        self.check(self.c.mpeghb200_resample_dev(self.ctx, h, d_in, d_out, d_state, n_units, stream))

    # ---- format converter ------------------------------------------------------------------------
    def format_conv_create(self, dmx):
        """dmx [n_out, n_in, 257] float32"""
        d = np.ascontiguousarray(dmx, dtype=np.float32)
        h = ctypes.c_void_p()
        self.check(self.c.mpeghb200_format_conv_create(self.ctx, d.shape[1], d.shape[0], d.ctypes.data, ctypes.byref(h)))
        return h

    def format_conv_destroy(self, h):
        self.c.mpeghb200_format_conv_destroy(h)

    def format_conv_state_floats(self, h):
        return int(self.c.mpeghb200_format_conv_state_floats(h))

    def format_conv_dev(self, h, d_in, d_out, d_state, n_streams, stream=0):
        self.check(self.c.mpeghb200_format_conv_dev(self.ctx, h, d_in, d_out, d_state, n_streams, stream))

    # ---- HOA spatial decoder ---------------------------------------------------------------------
    def hoa_spatial_create(self, cfg, tables):
        """cfg = (order, n_transport, n_addnl, min_amb_order, coded_v_vec_length, spat_interp_time, interp_method,
        pred_dir_sigs, num_bits_per_sf, use_phase_shift_decorr); tables: dict with info = (low_order_mat_sz,
        vec_start, ..) and the float tables named as in mpeghb200_hoa_spatial_desc's comment."""
        d = HoaSpatialDesc()
        d.order, d.n_transport, d.n_addnl_coders = cfg[0], cfg[1], cfg[2]
        d.low_order_mat_sz, d.vec_start = int(tables["info"][0]), int(tables["info"][1])
        d.coded_v_vec_length, d.spat_interp_time, d.pred_dir_sigs, d.num_bits_per_sf = cfg[4], cfg[5], cfg[7], cfg[8]
        d.use_phase_shift_decorr = cfg[9]
        keep = []
        for field, key in (("order_matrix", "order_matrix"), ("fine_grid_mode_matrix", "fine"),
                           ("coarse_grid_mode_matrix", "coarse"), ("fade_windows", "wins"),
                           ("dyn_win_pow_table", "dyn_pow"), ("dyn_win_func", "dyn_win"), ("pow2_table", "pow2")):
            a = np.ascontiguousarray(tables[key], dtype=np.float32)
            keep.append(a)
            setattr(d, field, a.ctypes.data)
        h = ctypes.c_void_p()
        self.check(self.c.mpeghb200_hoa_spatial_create(self.ctx, ctypes.byref(d), ctypes.byref(h)))
        return h

    def hoa_spatial_destroy(self, h):
        self.c.mpeghb200_hoa_spatial_destroy(h)

    def hoa_spatial_state_bytes(self, h):
        return int(self.c.mpeghb200_hoa_spatial_state_bytes(h))

    def hoa_spatial_num_coeffs(self, h):
        return int(self.c.mpeghb200_hoa_spatial_num_coeffs(h))

    def hoa_spatial_state_init_dev(self, h, d_state, n_streams, stream=0):
        self.check(self.c.mpeghb200_hoa_spatial_state_init_dev(self.ctx, h, d_state, n_streams, stream))

    def hoa_spatial_dev(self, h, d_side, d_in, d_out, d_state, n_streams, stream=0):
        self.check(self.c.mpeghb200_hoa_spatial_dev(self.ctx, h, d_side, d_in, d_out, d_state, n_streams, stream))

    # ---- binaural renderer -----------------------------------------------------------------------
    def binaural_create(self, taps_direct, taps_diffuse, direct_fc, diffuse_fc, weight):
        """taps_direct [n_sets, 2, n_in, len], taps_diffuse [n_sets, n_blocks, 2, len], direct_fc [n_sets, 2, n_in],
        diffuse_fc [n_sets, 2, n_blocks], weight [n_sets, n_in] (a missing leading n_sets axis means one set)."""
        td = np.ascontiguousarray(taps_direct, dtype=np.float32)
        if td.ndim == 3:
            td = td[None]
        n_sets, _, n_in, length = td.shape
        tf = np.ascontiguousarray(taps_diffuse, dtype=np.float32).reshape(n_sets, -1, 2, length)
        nb = tf.shape[1]
        dfc = np.ascontiguousarray(direct_fc, dtype=np.float32).reshape(n_sets, 2, n_in)
        ffc = np.ascontiguousarray(diffuse_fc, dtype=np.float32).reshape(n_sets, 2, nb)
        w = np.ascontiguousarray(weight, dtype=np.float32).reshape(n_sets, n_in)
        h = ctypes.c_void_p()
        self.check(self.c.mpeghb200_binaural_create(self.ctx, n_in, length, nb, n_sets, td.ctypes.data,
                                                    tf.ctypes.data if nb else None, dfc.ctypes.data,
                                                    ffc.ctypes.data if nb else None, w.ctypes.data, ctypes.byref(h)))
        return h

    def binaural_destroy(self, h):
        self.c.mpeghb200_binaural_destroy(h)

    def binaural_block_len(self, h):
        return int(self.c.mpeghb200_binaural_block_len(h))

    def binaural_state_floats(self, h):
        return int(self.c.mpeghb200_binaural_state_floats(h))

    def binaural_spectra(self, h, n_sets, n_in, n_blocks):
        n = 2 * self.binaural_block_len(h)
        d = np.zeros((n_sets, 2, n_in, n), np.float32)
        f = np.zeros((n_sets, max(n_blocks, 1), 2, n), np.float32)
        self.check(self.c.mpeghb200_binaural_spectra(self.ctx, h, d.ctypes.data, f.ctypes.data if n_blocks else None))
        return d, f

    def binaural_block_dev(self, h, d_in, d_out, d_state, n_streams, layout=0, in_scale=1.0, out_scale=1.0,
                           d_set=None, stream=0):
        self.check(self.c.mpeghb200_binaural_block_dev(self.ctx, h, d_in, layout, in_scale, out_scale, d_out, d_state,
                                                       d_set, n_streams, stream))

    # ---- object renderer ------------------------------------------------------------------------
    def obj_layout_create(self, arrays):
        desc = ObjLayoutDesc.from_arrays(arrays)
        h = ctypes.c_void_p()
        self.check(self.c.mpeghb200_obj_layout_create(self.ctx, ctypes.byref(desc), ctypes.byref(h)))
        return h

    def obj_layout_destroy(self, h):
        self.c.mpeghb200_obj_layout_destroy(h)

    def obj_side_from_polar(self, params):
        """params [..., 7] float32 -> structured array [...] of mpeghb200_obj_side (host trig, no GPU)."""
        params = np.ascontiguousarray(params, dtype=np.float32)
        side = np.zeros(params.shape[:-1], OBJ_SIDE_DTYPE)
        self.c.mpeghb200_obj_side_from_polar(params.ctypes.data, int(side.size), side.ctypes.data)
        return side

    def obj_state_floats(self, layout, n_obj):
        return int(self.c.mpeghb200_obj_state_floats(layout, n_obj))

    def obj_scratch_floats(self, layout, n_obj):
        return int(self.c.mpeghb200_obj_scratch_floats(layout, n_obj))

    def obj_state_init_dev(self, layout, n_obj, d_state, n_streams, d_first=None, stream=0):
        self.check(self.c.mpeghb200_obj_state_init_dev(self.ctx, layout, n_obj, d_first, d_state, n_streams, stream))

    def obj_render_dev(self, layout, n_obj, d_side, d_in, d_out, d_state, d_scratch, n_streams, stream=0):
        self.check(self.c.mpeghb200_obj_render_dev(self.ctx, layout, n_obj, d_side, d_in, d_out, d_state, d_scratch,
                                                   n_streams, stream))

    def obj_render_sub_dev(self, layout, n_obj, d_side, d_in, d_out, d_state, d_scratch, n_streams, frame_len,
                           sample_offset, d_metadata_set=None, stream=0):
        self.check(self.c.mpeghb200_obj_render_sub_dev(self.ctx, layout, n_obj, d_side, d_in, d_out, d_state, d_scratch,
                                                       n_streams, frame_len, sample_offset, d_metadata_set, stream))

    # ---- tail of the chain ----------------------------------------------------------------------
    def limiter_params(self, n_ch, sample_rate=48000, limiter_on=1):
        p = LimiterParams()
        st = self.c.mpeghb200_limiter_params_init(ctypes.byref(p), n_ch, sample_rate, limiter_on)
        if st != OK:
            raise MpeghError(st, "limiter_params_init: unsupported channel count or sample rate")
        return p

    def limiter_state_floats(self, p):
        return int(self.c.mpeghb200_limiter_state_floats(ctypes.byref(p)))

    def limiter_state_init_dev(self, p, d_state, n_streams, stream=0):
        self.check(self.c.mpeghb200_limiter_state_init_dev(self.ctx, ctypes.byref(p), d_state, n_streams, stream))

    def limiter_dev(self, p, d_in, d_out, d_state, n_streams, d_gain=None, stream=0):
        self.check(self.c.mpeghb200_limiter_dev(self.ctx, ctypes.byref(p), d_in, d_out, d_state, d_gain, n_streams,
                                                stream))

    def chain_open(self, n_streams, **kw):
        """kw: the fields of mpeghb200_chain_desc (handles as returned by the *_create wrappers)."""
        return Chain(self, n_streams, **kw)

    def limiter_pack_scratch_floats(self, p):
        return int(self.c.mpeghb200_limiter_pack_scratch_floats(ctypes.byref(p)))

    def limiter_pack_dev(self, p, d_in, d_state, d_scratch, bits, d_pcm, n_streams, d_gain=None, ch_map=None,
                         d_delay=None, d_pcm_bytes=None, stream=0):
        m = None if ch_map is None else np.ascontiguousarray(ch_map, dtype=np.int32)
        self.check(self.c.mpeghb200_limiter_pack_dev(self.ctx, ctypes.byref(p), d_in, d_state, d_scratch, d_gain, bits,
                                                     None if m is None else m.ctypes.data, d_delay, d_pcm, d_pcm_bytes,
                                                     n_streams, stream))

    def pcm_pack_block_dev(self, d_in, n_ch, n_samples, ch_stride, bits, d_out, n_streams, ch_map=None, d_delay=None,
                           d_out_bytes=None, stream=0):
        m = None if ch_map is None else np.ascontiguousarray(ch_map, dtype=np.int32)
        self.check(self.c.mpeghb200_pcm_pack_block_dev(self.ctx, d_in, n_ch, n_samples, ch_stride, bits,
                                                       None if m is None else m.ctypes.data, d_delay, d_out,
                                                       d_out_bytes, n_streams, stream))

    def pcm_pack_dev(self, d_in, n_ch, bits, d_out, n_streams, ch_map=None, d_delay=None, d_out_bytes=None, stream=0):
        m = None if ch_map is None else np.ascontiguousarray(ch_map, dtype=np.int32)
        self.check(self.c.mpeghb200_pcm_pack_dev(self.ctx, d_in, n_ch, bits, None if m is None else m.ctypes.data,
                                                 d_delay, d_out, d_out_bytes, n_streams, stream))


class FdSession:
    """Host-buffer FD synthesis session (mpeghb200_fd_open/execute/close)."""

    def __init__(self, lib, n_streams, n_ch):
        self.lib, self.n_streams, self.n_ch = lib, n_streams, n_ch
        h = ctypes.c_void_p()
        lib.check(lib.c.mpeghb200_fd_open(lib.ctx, n_streams, n_ch, ctypes.byref(h)))
        self.h = h

    def execute(self, coef, side, out=None):
        coef = np.ascontiguousarray(coef, dtype=np.float32)
        side = np.ascontiguousarray(side, dtype=np.uint32)
        assert coef.shape == (self.n_streams, self.n_ch, FRAME_LEN) and side.shape == (self.n_streams, self.n_ch)
        if out is None:
            out = np.empty_like(coef)
        # the library writes n_units * 4096 bytes through this pointer: refuse anything that is not exactly that
        assert isinstance(out, np.ndarray) and out.dtype == np.float32 and out.flags.c_contiguous and \
            out.flags.writeable and out.shape == coef.shape, "out must be a writable C-contiguous float32 array of coef's shape"
        self.lib.check(self.lib.c.mpeghb200_fd_execute(self.h, coef.ctypes.data, side.ctypes.data, out.ctypes.data))
        return out

    def submit(self, coef, side, out):
        """Asynchronous frame: the three arrays must stay alive and untouched until the matching wait()."""
        assert coef.dtype == np.float32 and side.dtype == np.uint32 and out.dtype == np.float32
        assert coef.flags.c_contiguous and side.flags.c_contiguous and out.flags.c_contiguous
        assert coef.shape == (self.n_streams, self.n_ch, FRAME_LEN) and side.shape == (self.n_streams, self.n_ch)
        self.lib.check(self.lib.c.mpeghb200_fd_submit(self.h, coef.ctypes.data, side.ctypes.data, out.ctypes.data))

    def wait(self):
        self.lib.check(self.lib.c.mpeghb200_fd_wait(self.h))

    def reset_stream(self, idx):
        self.lib.check(self.lib.c.mpeghb200_fd_reset_stream(self.h, idx))

    def close(self):
        if self.h:
            self.lib.c.mpeghb200_fd_close(self.h)
            self.h = None


class Chain:
    """mpeghb200_chain_*: one configuration end to end; host buffers in, packed PCM out."""

    def __init__(self, lib, n_streams, core_is_spectra=0, n_bed_ch=0, n_obj=0, n_hoa_transport=0, pcm_bits=24,
                 limiter=1, start_delay=0, loudness_gain=1.0, obj_layout=None, hoa_spatial=None, hoa_renderer=None,
                 format_conv=None, binaural=None, out_ch_map=None, binaural_set_of_stream=None,
                 max_frames_per_call=0):
        self.lib = lib
        d = ChainDesc()
        d.n_streams, d.core_is_spectra, d.n_bed_ch, d.n_obj, d.n_hoa_transport = n_streams, core_is_spectra, n_bed_ch, \
            n_obj, n_hoa_transport
        d.pcm_bits, d.limiter, d.start_delay, d.loudness_gain = pcm_bits, limiter, start_delay, loudness_gain
        d.max_frames_per_call = max_frames_per_call
        for name, h in (("obj_layout", obj_layout), ("hoa_spatial", hoa_spatial), ("hoa_renderer", hoa_renderer),
                        ("format_conv", format_conv), ("binaural", binaural)):
            setattr(d, name, h.value if isinstance(h, ctypes.c_void_p) else h)
        keep = []
        for name, a in (("out_ch_map", out_ch_map), ("binaural_set_of_stream", binaural_set_of_stream)):
            if a is not None:
                a = np.ascontiguousarray(a, np.int32)
                keep.append(a)
                setattr(d, name, a.ctypes.data)
        h = ctypes.c_void_p()
        lib.check(lib.c.mpeghb200_chain_open(lib.ctx, ctypes.byref(d), ctypes.byref(h)))
        self.h, self.desc = h, d
        self.n_streams = n_streams
        self.n_core = n_bed_ch + n_obj + n_hoa_transport
        self.n_out = lib.c.mpeghb200_chain_out_channels(h)
        self.frames_per_record = lib.c.mpeghb200_chain_frames_per_record(h)
        self.core_floats = int(lib.c.mpeghb200_chain_core_floats(h))
        self.record_bytes = int(lib.c.mpeghb200_chain_pcm_record_bytes(h))

    @staticmethod
    def _ptr(a):
        if a is None:
            return None
        if isinstance(a, np.ndarray):
            assert a.flags.c_contiguous
            return a.ctypes.data
        return int(a)  # raw pointer (pinned torch tensor .data_ptr(), device pointer)

    def frame(self, core, fd_side=None, obj_side=None, hoa_side=None, pcm=None, pcm_bytes=None):
        f = ChainFrame()
        f.core, f.fd_side, f.obj_side, f.hoa_side = self._ptr(core), self._ptr(fd_side), self._ptr(obj_side), \
            self._ptr(hoa_side)
        f.pcm, f.pcm_bytes = self._ptr(pcm), self._ptr(pcm_bytes)
        return f

    def alloc_out(self, n_frames=1):
        n_rec = n_frames // self.frames_per_record
        return (np.zeros((n_rec, self.n_streams, self.record_bytes), np.uint8),
                np.zeros((n_rec, self.n_streams), np.int32))

    def execute(self, n_frames=1, **kw):
        """numpy host arrays (keyword names as in frame()); returns (pcm [records, streams, record_bytes], bytes)."""
        if kw.get("pcm") is None:
            kw["pcm"], kw["pcm_bytes"] = self.alloc_out(n_frames)
        f = self.frame(**kw)
        self.lib.check(self.lib.c.mpeghb200_chain_execute(self.h, ctypes.byref(f), n_frames))
        return kw["pcm"], kw["pcm_bytes"]

    def submit(self, frame, n_frames=1):
        self.lib.check(self.lib.c.mpeghb200_chain_submit(self.h, ctypes.byref(frame), n_frames))

    def wait(self):
        self.lib.check(self.lib.c.mpeghb200_chain_wait(self.h))

    def run_dev(self, frame, n_frames=1, stream=0):
        self.lib.check(self.lib.c.mpeghb200_chain_run_dev(self.h, ctypes.byref(frame), n_frames, stream))

    def close(self):
        if self.h:
            self.lib.c.mpeghb200_chain_close(self.h)
            self.h = None


_lib = None


def load():
    global _lib
    if _lib is None:
        _lib = Lib()
    return _lib
