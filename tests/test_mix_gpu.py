"""Bus add and channel mask through the C ABI.  The reference code is a single rounded float
addition per sample (decoder/ia_core_coder_decode_main.c:2185-2302) and a copy-or-zero per channel (:906-1070), so
numpy float32 arithmetic is the exact checker."""
import numpy as np
import pytest

from oracle.loader import bits_equal

pytestmark = pytest.mark.gpu


def test_bus_add(gpu_lib):
    import torch
    rng = np.random.default_rng(3)
    st = torch.cuda.current_stream().cuda_stream
    for n in (4, 1024, 24 * 1024 * 37, 512 * 12 * 1024):
        a = rng.normal(0, 3000, n).astype(np.float32)
        b = rng.normal(0, 3000, n).astype(np.float32)
        a[::17], b[::17] = -0.0, 0.0
        a[5 % n:6 % n + 1] = np.float32(1e-40)  # denormals are kept (no flush to zero on this path)
        da, db = torch.from_numpy(a).cuda(), torch.from_numpy(b).cuda()
        out = torch.empty_like(da)
        gpu_lib.bus_add_dev(da.data_ptr(), db.data_ptr(), out.data_ptr(), n, st)
        gpu_lib.bus_add_dev(da.data_ptr(), db.data_ptr(), da.data_ptr(), n, st)  # in place on the first bus
        torch.cuda.synchronize()
        assert bits_equal(out.cpu().numpy(), a + b)
        assert bits_equal(da.cpu().numpy(), a + b)
    gpu_lib.bus_add_dev(0, 0, 0, 0, st)
    from libmpegh_b200.lib import MpeghError
    with pytest.raises(MpeghError):
        gpu_lib.bus_add_dev(da.data_ptr(), db.data_ptr(), da.data_ptr(), 6, st)


def test_channel_mask(gpu_lib):
    import torch
    rng = np.random.default_rng(4)
    C = 4096 * 3
    x = rng.normal(0, 3000, (C, 1024)).astype(np.float32)
    on = (rng.random(C) < 0.7).astype(np.uint8)
    d = torch.from_numpy(x.copy()).cuda()
    gpu_lib.channel_mask_dev(d.data_ptr(), torch.from_numpy(on).cuda().data_ptr(), C, torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    want = x.copy()
    want[on == 0] = 0.0
    assert bits_equal(d.cpu().numpy(), want)


def test_gain_apply(gpu_lib):
    import torch
    rng = np.random.default_rng(5)
    C, Q = 3000, 700
    x = rng.normal(0, 3000, (C, 1024)).astype(np.float32)
    g = np.exp2(rng.normal(0, 1, (Q, 1024))).astype(np.float32)
    seq = rng.integers(-1, Q, C).astype(np.int32)
    d = torch.from_numpy(x.copy()).cuda()
    gpu_lib.gain_apply_dev(d.data_ptr(), torch.from_numpy(g).cuda().data_ptr(), torch.from_numpy(seq).cuda().data_ptr(), C,
                           torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    want = x.copy()
    sel = seq >= 0
    want[sel] = x[sel] * g[seq[sel]]
    assert bits_equal(d.cpu().numpy(), want)


@pytest.mark.parametrize("n_ch,n_samples,S", [(1, 1024, 3), (2, 1024, 5), (12, 1024, 7), (24, 1024, 2), (64, 1024, 1),
                                             (5, 100, 3), (13, 4096, 2)])
def test_interleave_round_trip(gpu_lib, n_ch, n_samples, S):
    """mpeghb200_interleave_dev: [stream][sample][channel] <-> [stream][channel][sample], the copy the reference builds
    around its limiter (decoder/ia_core_coder_decode_main.c:1429-1460); pure data movement, compared with numpy."""
    import torch
    rng = np.random.default_rng(n_ch * 1000 + n_samples)
    x = rng.standard_normal((S, n_samples, n_ch)).astype(np.float32)
    x[0, 0, 0] = -0.0
    d_x = torch.from_numpy(x).cuda()
    d_p = torch.full((S, n_ch, n_samples), float("nan"), device="cuda")
    st = torch.cuda.current_stream().cuda_stream
    gpu_lib.interleave_dev(d_x.data_ptr(), d_p.data_ptr(), n_ch, n_samples, True, S, st)
    torch.cuda.synchronize()
    assert np.array_equal(d_p.cpu().numpy().view(np.uint32), np.ascontiguousarray(x.transpose(0, 2, 1)).view(np.uint32))
    d_y = torch.full((S, n_samples, n_ch), float("nan"), device="cuda")
    gpu_lib.interleave_dev(d_p.data_ptr(), d_y.data_ptr(), n_ch, n_samples, False, S, st)
    torch.cuda.synchronize()
    assert np.array_equal(d_y.cpu().numpy().view(np.uint32), x.view(np.uint32))
