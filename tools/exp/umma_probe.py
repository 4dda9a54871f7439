"""Decode what the tcgen05 contraction computes: identity-like matrices and index-coded inputs."""
import sys, os
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from libmpegh_b200.lib import Lib
lib = Lib().create(0)
st = torch.cuda.current_stream().cuda_stream
NC, NO = 25, 24
def run(M, x):
    h = lib.hoa_renderer_create(M.astype(np.float32), None, clip=0)
    lib.hoa_renderer_set_tensor_mode(h, 2)
    d_in = torch.from_numpy(x.astype(np.float32)).cuda()
    d_out = torch.full((x.shape[0], NO, 1024), float("nan"), device="cuda")
    dummy = torch.zeros(4, device="cuda")
    lib.hoa_render_dev(h, d_in.data_ptr(), d_out.data_ptr(), dummy.data_ptr(), x.shape[0], st)
    torch.cuda.synchronize()
    lib.hoa_renderer_destroy(h)
    return d_out.cpu().numpy()
# x[k][m] = 1000 k + (m % 128) + 0.25 * (m // 128)
k = np.arange(NC)[:, None]; m = np.arange(1024)[None, :]
x = (1000.0 * k + (m % 128) + 0.25 * (m // 128))[None]
M = np.zeros((NO, NC)); M[np.arange(NO), np.arange(NO)] = 1.0
o = run(M, x)[0]
print("identity: out[n][m] for n=0..3, m=0..9"); print(o[:4, :10])
print("identity: out[n][m] n=5, m=120..135"); print(o[5, 120:136])
print("identity: out[23][:6]", o[23, :6], " max err", np.abs(o - x[0, :NO]).max())
bad = np.argwhere(np.abs(o - x[0, :NO]) > 0.01)
print("n bad", len(bad), "first", bad[:10].tolist())
M2 = np.zeros((NO, NC)); M2[:, 3] = np.arange(1, NO + 1)
o2 = run(M2, x)[0]
print("col3 scaled by n+1: out[:, 7] / x[3][7]:", o2[:, 7] / x[0, 3, 7])
rng = np.random.default_rng(0)
Mr = rng.normal(0, 0.3, (NO, NC)); xr = rng.normal(0, 2000, (2, NC, 1024))
o3 = run(Mr, xr)
ref = np.einsum("nk,skm->snm", Mr.astype(np.float32).astype(np.float64), xr.astype(np.float32).astype(np.float64))
print("random: max err", np.abs(o3 - ref).max(), "rel", np.abs(o3 - ref).max() / np.abs(ref).max())
