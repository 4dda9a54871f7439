"""CUDA FD synthesis (through the C ABI) against the oracle, the committed reference vectors and, when it
travelled with the snapshot, the reference build itself.  Bit-exact: float32 compared as uint32."""
import os

import numpy as np
import pytest

import fd_cases
from oracle.loader import bits_equal

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden", "fd_golden.npz")


def run_dev(lib, coef, side_packed, overlap):
    """coef [U,1024], side [U], overlap [U,1024] -> (out, new_overlap) via mpeghb200_fd_synth_dev."""
    import torch
    d_coef = torch.from_numpy(np.ascontiguousarray(coef)).cuda()
    d_side = torch.from_numpy(side_packed.astype(np.int32)).cuda()
    d_ov = torch.from_numpy(np.ascontiguousarray(overlap)).cuda()
    d_out = torch.empty_like(d_coef)
    lib.fd_synth_dev(d_coef.data_ptr(), d_side.data_ptr(), d_ov.data_ptr(), d_out.data_ptr(), coef.shape[0],
                     torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    return d_out.cpu().numpy(), d_ov.cpu().numpy()


def test_all_combinations_vs_golden(gpu_lib):
    g = np.load(GOLD)
    side = fd_cases.pack_side(g["combos_side"])
    out, ov = run_dev(gpu_lib, g["combos_coef"], side, g["combos_ov_in"])
    for i, s in enumerate(g["combos_side"]):
        assert bits_equal(out[i], g["combos_out"][i]), ("out", tuple(s), np.abs(out[i] - g["combos_out"][i]).max())
        assert bits_equal(ov[i], g["combos_ov_out"][i]), ("overlap", tuple(s))


def test_chain_vs_golden_host_session(gpu_lib):
    """The host-buffer session (H2D, kernel, D2H inside the call) over 12 frames with carried state."""
    g = np.load(GOLD)
    side, coef = g["chain_side"], g["chain_coef"]
    F, C = side.shape[:2]
    sess = gpu_lib.fd_open(1, C)
    try:
        for f in range(F):
            out = sess.execute(coef[f][None], fd_cases.pack_side(side[f])[None])
            assert bits_equal(out[0], g["chain_out"][f]), f
    finally:
        sess.close()


@pytest.mark.parametrize("kswitch", [0.0, 0.3])
def test_random_streams_vs_oracle(gpu_lib, oracle, kswitch):
    """37 streams x 5 channels (ragged vs the 4-warp CTA), 10 frames, state carried on the device."""
    rng = np.random.default_rng(1234)
    S, C, F = 37, 5, 10
    U = S * C
    side = fd_cases.random_side(rng, F, U, kernel_switch_prob=kswitch)
    coef = fd_cases.random_coef(rng, (F, U, 1024))
    ov_o = np.zeros((U, 1024), np.float32)
    out_o = np.zeros((F, U, 1024), np.float32)
    assert oracle.oracle_fd_frm_dec_batch(coef, ov_o, out_o, side, F, U) == 0
    sess = gpu_lib.fd_open(S, C)
    try:
        for f in range(F):
            out = sess.execute(coef[f].reshape(S, C, 1024), fd_cases.pack_side(side[f]).reshape(S, C))
            assert bits_equal(out.reshape(U, 1024), out_o[f]), f
    finally:
        sess.close()


def test_reset_stream(gpu_lib, oracle):
    rng = np.random.default_rng(5)
    S, C = 3, 2
    coef = fd_cases.random_coef(rng, (2, S, C, 1024))
    side = np.zeros((S, C), np.uint32)
    sess = gpu_lib.fd_open(S, C)
    try:
        first = sess.execute(coef[0], side)
        sess.execute(coef[1], side)
        sess.reset_stream(1)
        again = sess.execute(coef[0], side)
        assert bits_equal(again[1], first[1])          # stream 1 restarted from zero state
        assert not bits_equal(again[0], first[0])      # the others kept theirs
    finally:
        sess.close()


def test_pinned_and_pageable_host_buffers_agree(gpu_lib, oracle):
    """mpeghb200_fd_execute takes the DMA-direct path for page-locked buffers and the staged path otherwise;
    2600 units exercise the 8-chunk pipeline with a ragged last chunk."""
    rng = np.random.default_rng(77)
    S, C = 325, 8
    coef = fd_cases.random_coef(rng, (S, C, 1024))
    side5 = fd_cases.random_side(rng, 1, S * C, kernel_switch_prob=0.1, legal=False)[0]
    side = fd_cases.pack_side(side5).reshape(S, C)
    want = np.zeros((S * C, 1024), np.float32)
    ov = np.zeros((S * C, 1024), np.float32)
    assert oracle.oracle_fd_frm_dec_batch(coef.reshape(1, S * C, 1024), ov, want.reshape(1, S * C, 1024),
                                          side5.reshape(1, S * C, 5), 1, S * C) == 0
    p_coef = gpu_lib.host_array((S, C, 1024))
    p_out = gpu_lib.host_array((S, C, 1024))
    p_coef[...] = coef
    for pinned in (False, True):
        sess = gpu_lib.fd_open(S, C)
        try:
            if pinned:
                gpu_lib.check(gpu_lib.c.mpeghb200_fd_execute(sess.h, p_coef.ctypes.data, side.ctypes.data,
                                                             p_out.ctypes.data))
                got = p_out.copy()
            else:
                got = sess.execute(coef, side)
        finally:
            sess.close()
        assert bits_equal(got.reshape(S * C, 1024), want), pinned
    gpu_lib.host_free(p_coef)
    gpu_lib.host_free(p_out)


def test_empty_batch(gpu_lib):
    gpu_lib.fd_synth_dev(0, 0, 0, 0, 0)


def test_full_size_properties(gpu_lib, oracle):
    """BASELINE config 2 shape (256 streams x 24 channels): spot-check units against the oracle and check
    stream independence (a permutation of the units permutes the outputs) and linearity-free determinism."""
    rng = np.random.default_rng(1234)
    U = 256 * 24
    coef = fd_cases.random_coef(rng, (U, 1024))
    ov = fd_cases.random_coef(rng, (U, 1024), 300.0)
    side5 = fd_cases.random_side(rng, 1, U, kernel_switch_prob=0.05, legal=False)[0]
    side = fd_cases.pack_side(side5)
    out, ov_new = run_dev(gpu_lib, coef, side, ov)
    perm = rng.permutation(U)
    out_p, ov_p = run_dev(gpu_lib, coef[perm], side[perm], ov[perm])
    assert bits_equal(out_p, out[perm]) and bits_equal(ov_p, ov_new[perm])
    for u in rng.choice(U, 200, replace=False):
        o = np.zeros(1024, np.float32)
        v = ov[u].copy()
        oracle.oracle_fd_frm_dec(coef[u].copy(), v, o, *[int(x) for x in side5[u]])
        assert bits_equal(out[u], o) and bits_equal(ov_new[u], v), u


def test_vs_reference_build_if_present(gpu_lib):
    """If oracle/_ref travelled with the snapshot, compare directly with the unmodified reference."""
    from oracle import loader
    ref = loader.load_ref()
    if ref is None:
        pytest.skip("oracle/_ref not present")
    rng = np.random.default_rng(99)
    F, C = 6, 8
    side = fd_cases.random_side(rng, F, C, kernel_switch_prob=0.2)
    coef = fd_cases.random_coef(rng, (F, C, 1024))
    ov = np.zeros((C, 1024), np.float32)
    out_r = np.zeros((F, C, 1024), np.float32)
    h = ref.ref_fd_create()
    assert ref.ref_fd_frm_dec_batch(h, coef, ov, out_r, side, F, C) == 0
    ref.ref_fd_destroy(h)
    sess = gpu_lib.fd_open(1, C)
    try:
        for f in range(F):
            out = sess.execute(coef[f][None], fd_cases.pack_side(side[f])[None])
            assert bits_equal(out[0], out_r[f]), f
    finally:
        sess.close()


def test_session_submit_wait_pipelined(gpu_lib, oracle):
    """Two frames in flight (mpeghb200_fd_submit / _wait): results and carried overlap state equal the oracle's
    frame-by-frame run; a third submit without a wait is refused; execute() drains what is in flight first."""
    from libmpegh_b200.lib import MpeghError
    rng = np.random.default_rng(77)
    S, C, F = 40, 24, 7  # 960 units: two chunks per frame
    U = S * C
    side = fd_cases.random_side(rng, F, U, kernel_switch_prob=0.1)
    coef = fd_cases.random_coef(rng, (F, U, 1024))
    ov = np.zeros((U, 1024), np.float32)
    want = np.zeros((F, U, 1024), np.float32)
    assert oracle.oracle_fd_frm_dec_batch(coef, ov, want, side, F, U) == 0
    packed = [np.ascontiguousarray(fd_cases.pack_side(side[f]).reshape(S, C).astype(np.uint32)) for f in range(F)]
    cf = [np.ascontiguousarray(coef[f].reshape(S, C, 1024)) for f in range(F)]
    outs = [np.zeros((S, C, 1024), np.float32) for _ in range(F)]
    sess = gpu_lib.fd_open(S, C)
    try:
        sess.submit(cf[0], packed[0], outs[0])
        sess.submit(cf[1], packed[1], outs[1])
        with pytest.raises(MpeghError):
            sess.submit(cf[2], packed[2], outs[2])
        for f in range(2, F - 1):
            sess.wait()
            assert bits_equal(outs[f - 2].reshape(U, 1024), want[f - 2]), f - 2
            sess.submit(cf[f], packed[f], outs[f])
        got_last = sess.execute(cf[F - 1], packed[F - 1])  # drains frames F-3, F-2, then runs the last one
        for f in range(F - 1):
            assert bits_equal(outs[f].reshape(U, 1024), want[f]), f
        assert bits_equal(got_last.reshape(U, 1024), want[F - 1])
        sess.wait()  # nothing in flight: no-op
    finally:
        sess.close()


def run_frames_dev(lib, coef, side_packed, overlap):
    """coef [F,U,1024], side [F,U], overlap [U,1024] -> (out [F,U,1024], new_overlap) via mpeghb200_fd_synth_frames_dev."""
    import torch
    d_coef = torch.from_numpy(np.ascontiguousarray(coef)).cuda()
    d_side = torch.from_numpy(np.ascontiguousarray(side_packed.astype(np.int32))).cuda()
    d_ov = torch.from_numpy(np.ascontiguousarray(overlap)).cuda()
    d_out = torch.full_like(d_coef, float("nan"))
    lib.fd_synth_frames_dev(d_coef.data_ptr(), d_side.data_ptr(), d_ov.data_ptr(), d_out.data_ptr(), coef.shape[1],
                            coef.shape[0], torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    return d_out.cpu().numpy(), d_ov.cpu().numpy()


@pytest.mark.parametrize("F,U,kswitch,long_only", [(4, 185, 0.0, False), (7, 185, 0.3, False), (2, 1, 0.0, False),
                                                   (5, 3000, 0.0, True), (3, 3000, 0.1, False), (1, 50, 0.0, False)])
def test_frames_launch_vs_oracle(gpu_lib, oracle, F, U, kswitch, long_only):
    """F frames of every unit in one launch == the oracle walking the frames one by one, with legal window-sequence
    chains (LONG_START -> EIGHT_SHORT ... -> LONG_STOP inside the launch, so units leave the register-carried fast path
    and come back in a later launch), transform-kernel switching, ragged unit counts and two launches chained."""
    rng = np.random.default_rng(F * 1000 + U)
    side = fd_cases.random_side(rng, 2 * F, U, kernel_switch_prob=kswitch)
    if long_only:
        side[:, :, 0] = 0
    coef = fd_cases.random_coef(rng, (2 * F, U, 1024))
    ov_o = fd_cases.random_coef(rng, (U, 1024), sigma=0.3)
    ov_g = ov_o.copy()
    out_o = np.zeros((2 * F, U, 1024), np.float32)
    assert oracle.oracle_fd_frm_dec_batch(coef, ov_o, out_o, side, 2 * F, U) == 0
    for half in range(2):
        sl = slice(half * F, (half + 1) * F)
        out, ov_g = run_frames_dev(gpu_lib, coef[sl], fd_cases.pack_side(side[sl]), ov_g)
        for f in range(F):
            assert bits_equal(out[f], out_o[sl][f]), (half, f, int((out[f].view(np.uint32) != out_o[sl][f].view(np.uint32)).sum()))
    assert bits_equal(ov_g, ov_o)


def test_frames_launch_equals_single_frame_launches(gpu_lib):
    """At configs[1]'s size with ~8 % transition frames, including more generic-path units per CTA than the slow list
    holds (every unit of the first 300 goes EIGHT_SHORT in frame 1)."""
    rng = np.random.default_rng(99)
    F, U = 4, 6144
    side = fd_cases.random_side(rng, F, U)
    side[1, :300, 0] = 2
    coef = fd_cases.random_coef(rng, (F, U, 1024))
    ov0 = fd_cases.random_coef(rng, (U, 1024), sigma=0.3)
    packed = fd_cases.pack_side(side)
    out, ov = run_frames_dev(gpu_lib, coef, packed, ov0)
    ov1 = ov0
    for f in range(F):
        o1, ov1 = run_dev(gpu_lib, coef[f], packed[f], ov1)
        assert bits_equal(out[f], o1), f
    assert bits_equal(ov, ov1)


def test_frames_launch_rejects_misaligned_buffers(gpu_lib):
    """The multi-frame launch moves coefficient rows by TMA bulk copy and the rest as float4: a buffer that is not on a
    16-byte boundary is refused with MPEGHB200_ERR_BAD_ARG instead of faulting on the device."""
    import torch
    U, F = 8, 2
    coef = torch.zeros(F * U * 1024 + 4, device="cuda")
    side = torch.zeros(F * U, dtype=torch.int32, device="cuda")
    ov = torch.zeros(U * 1024, device="cuda")
    out = torch.zeros(F * U * 1024, device="cuda")
    st = torch.cuda.current_stream().cuda_stream
    gpu_lib.fd_synth_frames_dev(coef.data_ptr(), side.data_ptr(), ov.data_ptr(), out.data_ptr(), U, F, st)  # aligned: fine
    with pytest.raises(Exception):
        gpu_lib.fd_synth_frames_dev(coef.data_ptr() + 4, side.data_ptr(), ov.data_ptr(), out.data_ptr(), U, F, st)
    torch.cuda.synchronize()
