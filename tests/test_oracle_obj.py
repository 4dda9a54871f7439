"""Object renderer restatement (oracle/oracle_obj.c) against the unmodified reference functions and against
vectors generated from them."""
import ctypes
import os

import numpy as np
import pytest

import obj_cases
from oracle import loader
from oracle.loader import bits_equal

GOLD = os.path.join(os.path.dirname(__file__), "golden", "obj_golden.npz")


@pytest.mark.parametrize("cicp", obj_cases.CICPS)
@pytest.mark.parametrize("n_obj", [1, 10, 32])
def test_render_vs_reference(ref, cicp, n_obj):
    rng = np.random.default_rng(cicp * 100 + n_obj)
    err = ctypes.c_int()
    h = ref.ref_obj_create(cicp, ctypes.byref(err))
    assert err.value == 0
    arrays = loader.ref_obj_layout_arrays(ref, h)
    gold_arrays = loader.obj_layout_golden(cicp)
    for k in arrays:
        assert np.array_equal(arrays[k].view(np.uint32), gold_arrays[k].view(np.uint32)), (cicp, k)
    L = loader.obj_layout_from_arrays(arrays)
    flags = np.zeros(n_obj, np.int32)
    ref.ref_obj_first_flags(h, flags, n_obj)
    st = loader.oracle_obj_state_new(24)
    assert [st.first_frame[i] for i in range(n_obj)] == flags.tolist()
    F = 5
    params = obj_cases.random_params(rng, F, n_obj)
    x = obj_cases.object_audio(rng, F, n_obj)
    for f in range(F):
        want = np.zeros((L.n_spk, 1024), np.float32)
        gains = np.zeros((n_obj, 44), np.float32)
        assert ref.ref_obj_render(h, params[f].reshape(-1), n_obj, x[f].reshape(-1), want.reshape(-1),
                                  gains.reshape(-1)) == 0
        side = loader.oracle_obj_side(params[f])
        got = loader.oracle_obj_render(L, st, side, x[f])
        for o in range(n_obj):
            assert bits_equal(loader.oracle_obj_gains(L, side[o:o + 1]), gains[o, :L.n_non_lfe]), (cicp, f, o)
        assert bits_equal(got, want), (cicp, n_obj, f)
    ref.ref_obj_destroy(h)


def test_obj_golden(oracle):
    g = np.load(GOLD)
    for cicp in (6, 19):
        L = loader.obj_layout_from_arrays(loader.obj_layout_golden(cicp))
        st = loader.oracle_obj_state_new(24)
        params, x = g[f"cicp{cicp}_params"], g[f"cicp{cicp}_in"]
        for f in range(params.shape[0]):
            got = loader.oracle_obj_render(L, st, loader.oracle_obj_side(params[f]), x[f])
            assert bits_equal(got, g[f"cicp{cicp}_out"][f]), (cicp, f)


@pytest.mark.parametrize("frame_len", [256, 512])
def test_oam_subframes_vs_reference(ref, oracle, frame_len):
    """OAM frame length 256 / 512: the reference renders a core frame in 4 / 2 calls (decode_main.c:1383-1410); the
    oracle's sub-frame call chained with the same sub_frame_number bookkeeping must give the same frame."""
    err = ctypes.c_int()
    h = ref.ref_obj_create(19, ctypes.byref(err))
    assert err.value == 0
    L = loader.obj_layout_from_arrays(loader.ref_obj_layout_arrays(ref, h))
    n_obj, n_sub, F = 7, 1024 // frame_len, 6
    st = loader.oracle_obj_state_new(24)
    rng = np.random.default_rng(frame_len)
    for f in range(F):
        params = obj_cases.random_params(rng, n_sub, n_obj)
        present = (rng.random(n_sub) < 0.7).astype(np.int32)
        if f == 0:
            present[0] = frame_len == 512  # 256: the stream's very first call renders with set 3, not set 0
        x = obj_cases.object_audio(rng, 1, n_obj)[0]
        want = np.zeros((L.n_spk, 1024), np.float32)
        used = np.zeros(n_sub, np.int32)
        assert ref.ref_obj_render_subframes(h, params.reshape(-1), present, n_obj, frame_len, x.reshape(-1),
                                            want.reshape(-1), used) == 0
        assert list(used) == loader.subframe_sets(present, n_sub)
        got = loader.oracle_obj_render_subframes(L, st, params, present, frame_len, x)
        assert bits_equal(got, want), (frame_len, f, float(np.abs(got - want).max()))
    ref.ref_obj_destroy(h)
