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
    _ref = lib
    return lib


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
    lib.ref_samples_sat.argtypes = [_f32p, ctypes.c_int, ctypes.c_int, ctypes.c_int, _i32p, _i32p, _u8p]
    lib.ref_samples_sat.restype = ctypes.c_int
    _ref_ns = lib
    return lib


def bits_equal(a, b):
    """Bit-for-bit equality of two float32 arrays (distinguishes -0.0 from +0.0, NaN payloads)."""
    a = np.ascontiguousarray(a, dtype=np.float32)
    b = np.ascontiguousarray(b, dtype=np.float32)
    return a.shape == b.shape and bool(np.array_equal(a.view(np.uint32), b.view(np.uint32)))
