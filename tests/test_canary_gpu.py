"""Out-of-bounds writes: every output / state buffer of the stages below sits between two guard regions filled with a
sentinel, which must be intact after the launch.  (compute-sanitizer is closed on this GPU pool,
profiles/r2_sanitizer_unavailable.txt; wrong READS show up in the bit-exact parity tests instead.)"""
import os

import numpy as np
import pytest

import fd_cases
import hoa_cases
import obj_cases
from oracle import loader

pytestmark = pytest.mark.gpu
G = os.path.join(os.path.dirname(__file__), "golden")
PAD = 4096  # elements on each side
SENT = 0x7FC0DEAD  # a NaN payload no kernel produces


class Guarded:
    """A device buffer of n elements between two sentinel regions; .ptr is the address of element 0."""

    def __init__(self, n, dtype="float32", fill=None):
        import torch
        self.t = torch
        self.itemsize = {"float32": 4, "int32": 4, "uint8": 1}[dtype]
        self.n = n
        words = (n * self.itemsize + 3) // 4
        self.buf = torch.full((2 * PAD + words,), SENT, dtype=torch.int32, device="cuda")
        self.words = words
        body = self.buf[PAD:PAD + words]
        if fill is not None:
            body.view(torch.uint8)[:n * self.itemsize] = torch.from_numpy(
                np.ascontiguousarray(fill).view(np.uint8).reshape(-1)).cuda()
        else:
            body.zero_()
        self.ptr = self.buf.data_ptr() + 4 * PAD

    def check(self, name):
        b = self.buf.cpu().numpy()
        assert (b[:PAD] == SENT).all(), f"{name}: write below the buffer"
        tail = b[PAD + self.words:]
        assert (tail == SENT).all(), f"{name}: write above the buffer"


def test_fd_synth_ragged_counts(gpu_lib):
    import torch
    rng = np.random.default_rng(1)
    for U in (1, 5, 6, 7, 148 * 3 + 1, 1000):
        side = fd_cases.pack_side(fd_cases.random_side(rng, 1, U, kernel_switch_prob=0.2, legal=False)[0]).astype(np.int32)
        coef = Guarded(U * 1024, fill=fd_cases.random_coef(rng, (U, 1024)))
        ov, out = Guarded(U * 1024), Guarded(U * 1024)
        d_side = Guarded(U, "int32", fill=side)
        gpu_lib.fd_synth_dev(coef.ptr, d_side.ptr, ov.ptr, out.ptr, U, torch.cuda.current_stream().cuda_stream)
        torch.cuda.synchronize()
        for b, n in ((coef, "coef"), (ov, "overlap"), (out, "out"), (d_side, "side")):
            b.check(f"fd_synth U={U} {n}")


@pytest.mark.parametrize("n_ch,bits", [(1, 16), (12, 24), (24, 24), (5, 32)])
def test_tail_kernels(gpu_lib, n_ch, bits):
    import torch
    rng = np.random.default_rng(n_ch)
    S = 7
    st = torch.cuda.current_stream().cuda_stream
    p = gpu_lib.limiter_params(n_ch)
    x = rng.normal(0, 9000, (S, n_ch, 1024)).astype(np.float32)
    for fused in (False, True):
        d_in = Guarded(x.size, fill=x)
        d_lim = Guarded(x.size)
        state = Guarded(S * gpu_lib.limiter_state_floats(p))
        gpu_lib.limiter_state_init_dev(p, state.ptr, S, st)
        pcm = Guarded(S * 1024 * n_ch * (bits // 8), "uint8")
        delay = Guarded(S, "int32", fill=np.array([0, 1, 240, 1023, 1024, 2000, 7], np.int32))
        nb = Guarded(S, "int32")
        if fused:
            scr = Guarded(S * gpu_lib.limiter_pack_scratch_floats(p))
            gpu_lib.limiter_pack_dev(p, d_in.ptr, state.ptr, scr.ptr, bits, pcm.ptr, S, None, None, delay.ptr, nb.ptr, st)
            torch.cuda.synchronize()
            scr.check(f"tail fused {n_ch}ch {bits}bit scratch")
        else:
            gpu_lib.limiter_dev(p, d_in.ptr, d_lim.ptr, state.ptr, S, None, st)
            gpu_lib.pcm_pack_dev(d_lim.ptr, n_ch, bits, pcm.ptr, S, None, delay.ptr, nb.ptr, st)
        torch.cuda.synchronize()
        for b, n in ((d_in, "in"), (d_lim, "limited"), (state, "state"), (pcm, "pcm"), (delay, "delay"), (nb, "bytes")):
            b.check(f"tail fused={fused} {n_ch}ch {bits}bit {n}")


@pytest.mark.parametrize("n_obj,frame_len", [(1, 1024), (9, 256), (32, 1024), (17, 512)])
def test_obj_render(gpu_lib, n_obj, frame_len):
    import torch
    rng = np.random.default_rng(n_obj)
    S = 5
    lay = gpu_lib.obj_layout_create(loader.obj_layout_golden(19))
    st = torch.cuda.current_stream().cuda_stream
    side = gpu_lib.obj_side_from_polar(obj_cases.random_params(rng, 1, S * n_obj)[0].reshape(S, n_obj, 7))
    d_side = Guarded(side.nbytes, "uint8", fill=side)
    d_in = Guarded(S * n_obj * 1024, fill=obj_cases.object_audio(rng, 1, S * n_obj)[0])
    d_out = Guarded(S * 12 * 1024)
    state = Guarded(S * gpu_lib.obj_state_floats(lay, n_obj))
    scratch = Guarded(S * gpu_lib.obj_scratch_floats(lay, n_obj))
    gpu_lib.obj_state_init_dev(lay, n_obj, state.ptr, S, None, st)
    for off in range(0, 1024, frame_len):
        gpu_lib.obj_render_sub_dev(lay, n_obj, d_side.ptr, d_in.ptr, d_out.ptr, state.ptr, scratch.ptr, S, frame_len, off,
                                   None, st)
    torch.cuda.synchronize()
    for b, n in ((d_side, "side"), (d_in, "in"), (d_out, "out"), (state, "state"), (scratch, "scratch")):
        b.check(f"obj_render {n_obj} obj L={frame_len} {n}")
    gpu_lib.obj_layout_destroy(lay)


@pytest.mark.parametrize("cfg_idx,mode", [(0, 0), (1, 0), (1, 1), (1, 2), (2, 2), (4, 2), (5, 0)])
def test_hoa_render(gpu_lib, cfg_idx, mode):
    import torch
    g = np.load(os.path.join(G, "hoa_ren_golden.npz"))
    cfg = hoa_cases.CONFIGS[cfg_idx]
    k = hoa_cases.key(cfg)
    M, nfc = g[k + "_M"], (g[k + "_nfc"] if cfg[3] > 0 else None)
    S = 3
    h = gpu_lib.hoa_renderer_create(M, nfc)
    gpu_lib.hoa_renderer_set_tensor_mode(h, mode)
    rng = np.random.default_rng(cfg_idx)
    d_in = Guarded(S * M.shape[1] * 1024, fill=rng.normal(0, 2000, (S, M.shape[1], 1024)).astype(np.float32))
    d_out = Guarded(S * M.shape[0] * 1024)
    state = Guarded(max(1, S * gpu_lib.hoa_nfc_state_floats(h)))
    gpu_lib.hoa_render_dev(h, d_in.ptr, d_out.ptr, state.ptr, S, torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    for b, n in ((d_in, "in"), (d_out, "out"), (state, "nfc state")):
        b.check(f"hoa_render {k} mode {mode} {n}")
    gpu_lib.hoa_renderer_destroy(h)


@pytest.mark.parametrize("F,U", [(2, 1), (3, 7), (8, 148 * 3 + 5)])
def test_fd_synth_frames_ragged_counts(gpu_lib, F, U):
    """The multi-frame launch: coefficients / PCM [F][U][1024], overlap [U][1024], with generic-path frames inside."""
    import torch
    rng = np.random.default_rng(F * 100 + U)
    side = fd_cases.pack_side(fd_cases.random_side(rng, F, U, kernel_switch_prob=0.2)).astype(np.int32)
    coef = Guarded(F * U * 1024, fill=fd_cases.random_coef(rng, (F, U, 1024)))
    ov, out = Guarded(U * 1024), Guarded(F * U * 1024)
    d_side = Guarded(F * U, "int32", fill=side)
    gpu_lib.fd_synth_frames_dev(coef.ptr, d_side.ptr, ov.ptr, out.ptr, U, F, torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    for b, n in ((coef, "coef"), (ov, "overlap"), (out, "out"), (d_side, "side")):
        b.check(f"fd_synth_frames F={F} U={U} {n}")


@pytest.mark.parametrize("n_in,length,nb", [(3, 300, 1), (2, 40, 0), (2, 5000, 1)])
def test_binaural_generic_sizes(gpu_lib, n_in, length, nb):
    """The generic binaural path keeps its per-stream workspace behind the carried state: state and output guarded."""
    import torch
    import bin_cases
    rng = np.random.default_rng(length)
    f = bin_cases.brir_set(rng, n_in, length, nb)
    h = gpu_lib.binaural_create(f["taps_direct"][None], f["taps_diffuse"][None], f["direct_fc"][None],
                                f["diffuse_fc"][None], f["weight"][None])
    B, S = gpu_lib.binaural_block_len(h), 3
    x = rng.normal(0, 0.1, (S, n_in, B)).astype(np.float32)
    d_in = Guarded(x.size, fill=x)
    d_out = Guarded(S * 2 * B)
    state = Guarded(S * gpu_lib.binaural_state_floats(h))
    for _ in range(2):
        gpu_lib.binaural_block_dev(h, d_in.ptr, d_out.ptr, state.ptr, S, 0, 1.0, 1.0, None,
                                   torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    for b, n in ((d_in, "in"), (d_out, "out"), (state, "state")):
        b.check(f"binaural N={2 * B} {n}")
    gpu_lib.binaural_destroy(h)
