"""CUDA object renderer (through the C ABI) against the oracle and the committed reference vectors.
Bit-exact (float32 compared as uint32); layout tables come from tests/golden/obj_layouts.npz (dumped from the
reference's init), the per-object trigonometry from the library's own host helper."""
import os

import numpy as np
import pytest

import obj_cases
from oracle import loader
from oracle.loader import bits_equal

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden", "obj_golden.npz")


class DevObjRenderer:
    def __init__(self, lib, cicp, n_streams, n_obj, n_first=24):
        import torch
        self.t, self.lib, self.S, self.O = torch, lib, n_streams, n_obj
        self.arrays = loader.obj_layout_golden(cicp)
        self.n_spk = int(self.arrays["counts"][0])
        self.layout = lib.obj_layout_create(self.arrays)
        self.state = torch.empty(n_streams * lib.obj_state_floats(self.layout, n_obj), device="cuda")
        self.scratch = torch.empty(n_streams * lib.obj_scratch_floats(self.layout, n_obj), device="cuda")
        first = torch.tensor([1 if o < n_first else 0 for o in range(n_obj)], dtype=torch.int32, device="cuda")
        lib.obj_state_init_dev(self.layout, n_obj, self.state.data_ptr(), n_streams, first.data_ptr())

    def run(self, params, x):
        """params [S, O, 7], x [S, O, 1024] -> [S, n_spk, 1024]"""
        t = self.t
        side = self.lib.obj_side_from_polar(params)
        d_side = t.from_numpy(side.view(np.float32).reshape(self.S, self.O, 16).copy()).cuda()
        d_in = t.from_numpy(np.ascontiguousarray(x, dtype=np.float32)).cuda()
        d_out = t.full((self.S, self.n_spk, 1024), float("nan"), device="cuda")
        self.lib.obj_render_dev(self.layout, self.O, d_side.data_ptr(), d_in.data_ptr(), d_out.data_ptr(),
                                self.state.data_ptr(), self.scratch.data_ptr(), self.S,
                                t.cuda.current_stream().cuda_stream)
        t.cuda.synchronize()
        return d_out.cpu().numpy()

    def close(self):
        self.lib.obj_layout_destroy(self.layout)


def test_host_trig_helper_matches_oracle(so_lib):
    rng = np.random.default_rng(1)
    p = obj_cases.random_params(rng, 50, 20, spread_prob=0.5).reshape(-1, 7)
    got = so_lib.obj_side_from_polar(p)
    want = loader.oracle_obj_side(p)
    assert np.array_equal(got.view(np.uint32), want.view(np.uint32))


def test_golden(gpu_lib):
    g = np.load(GOLD)
    for cicp in (6, 19):
        params, x = g[f"cicp{cicp}_params"], g[f"cicp{cicp}_in"]
        r = DevObjRenderer(gpu_lib, cicp, 1, params.shape[1])
        for f in range(params.shape[0]):
            assert bits_equal(r.run(params[f][None], x[f][None])[0], g[f"cicp{cicp}_out"][f]), (cicp, f)
        r.close()


@pytest.mark.parametrize("cicp,n_obj", [(2, 3), (6, 10), (13, 24), (14, 1), (19, 32)])
def test_streams_vs_oracle(gpu_lib, oracle, cicp, n_obj):
    """13 independent streams (ragged vs the CTA's stream group), 5 frames, moving objects, ~30 % with spread;
    objects >= 24 start with first_frame_flag = 0 like the reference (MAX_NUM_OAM_OBJS = 24)."""
    rng = np.random.default_rng(cicp * 1000 + n_obj)
    S, F = 13, 5
    L = loader.obj_layout_from_arrays(loader.obj_layout_golden(cicp))
    sts = [loader.oracle_obj_state_new(24) for _ in range(S)]
    params = np.stack([obj_cases.random_params(rng, F, n_obj) for _ in range(S)], axis=1)  # [F, S, O, 7]
    x = np.stack([obj_cases.object_audio(rng, F, n_obj) for _ in range(S)], axis=1)
    r = DevObjRenderer(gpu_lib, cicp, S, n_obj)
    for f in range(F):
        got = r.run(params[f], x[f])
        for s in range(S):
            want = loader.oracle_obj_render(L, sts[s], loader.oracle_obj_side(params[f, s]), x[f, s])
            assert bits_equal(got[s], want), (cicp, f, s, float(np.abs(got[s] - want).max()))
    r.close()


def test_full_size_properties(gpu_lib, oracle):
    """BASELINE config 5 shape on one GPU (512 streams x 32 objects -> 7.1.4): stream independence under
    permutation, and spot checks against the oracle."""
    rng = np.random.default_rng(4567)
    S, O = 512, 32
    params = obj_cases.random_params(rng, 1, S * O, spread_prob=0.1)[0].reshape(S, O, 7)
    x = obj_cases.object_audio(rng, 1, S * O)[0].reshape(S, O, 1024)
    r = DevObjRenderer(gpu_lib, 19, S, O)
    out = r.run(params, x)
    r.close()
    perm = rng.permutation(S)
    r2 = DevObjRenderer(gpu_lib, 19, S, O)
    out_p = r2.run(params[perm], x[perm])
    r2.close()
    assert bits_equal(out_p, out[perm])
    L = loader.obj_layout_from_arrays(loader.obj_layout_golden(19))
    for s in rng.choice(S, 12, replace=False):
        st = loader.oracle_obj_state_new(24)
        assert bits_equal(out[s], loader.oracle_obj_render(L, st, loader.oracle_obj_side(params[s]), x[s])), s
    assert not out[:, 3].any()  # LFE row untouched (zero)


@pytest.mark.parametrize("frame_len", [256, 512])
def test_oam_subframes_vs_oracle(gpu_lib, oracle, frame_len):
    """OAM frame length 256 / 512: 4 / 2 mpeghb200_obj_render_sub_dev calls per core frame, each with the metadata set
    the reference's sub_frame_number bookkeeping selects; 6 streams x 5 frames, state carried on the device."""
    import torch
    rng = np.random.default_rng(900 + frame_len)
    S, F, O, n_sub = 6, 5, 9, 1024 // frame_len
    L = loader.obj_layout_from_arrays(loader.obj_layout_golden(19))
    sts = [loader.oracle_obj_state_new(24) for _ in range(S)]
    r = DevObjRenderer(gpu_lib, 19, S, O)
    st = torch.cuda.current_stream().cuda_stream
    for f in range(F):
        params = np.stack([obj_cases.random_params(rng, n_sub, O) for _ in range(S)])      # [S, n_sub, O, 7]
        present = (rng.random((S, n_sub)) < 0.7).astype(np.int32)
        x = np.stack([obj_cases.object_audio(rng, 1, O)[0] for _ in range(S)])             # [S, O, 1024]
        d_in = torch.from_numpy(x).cuda()
        d_out = torch.full((S, r.n_spk, 1024), float("nan"), device="cuda")
        used = np.array([loader.subframe_sets(present[s], n_sub) for s in range(S)])        # [S, n_sub]
        for i in range(n_sub):
            side = gpu_lib.obj_side_from_polar(np.stack([params[s, used[s, i]] for s in range(S)]))
            d_side = torch.from_numpy(side.view(np.float32).reshape(S, O, 16).copy()).cuda()
            d_set = torch.from_numpy(np.ascontiguousarray(used[:, i], np.int32)).cuda()
            gpu_lib.obj_render_sub_dev(r.layout, O, d_side.data_ptr(), d_in.data_ptr(), d_out.data_ptr(),
                                       r.state.data_ptr(), r.scratch.data_ptr(), S, frame_len, i * frame_len,
                                       d_set.data_ptr(), st)
        torch.cuda.synchronize()
        got = d_out.cpu().numpy()
        for s in range(S):
            want = loader.oracle_obj_render_subframes(L, sts[s], params[s], present[s], frame_len, x[s])
            assert bits_equal(got[s], want), (frame_len, f, s, float(np.abs(got[s] - want).max()))
    r.close()
