// ctx.cu -- context life cycle, constant-table upload, device memory helpers of the C ABI.
#include <cstdlib>
#include <cstring>
#include <mutex>

#include "common.h"

namespace mpeghb200 {

bool g_pdl_enabled = true;

static std::mutex g_err_mu;
static std::string g_create_error;

mpeghb200_status fail(mpeghb200_ctx* ctx, mpeghb200_status code, const std::string& msg) {
  if (ctx) {
    ctx->last_error = msg;
  } else {
    std::lock_guard<std::mutex> lk(g_err_mu);
    g_create_error = msg;
  }
  return code;
}

mpeghb200_status cuda_fail(mpeghb200_ctx* ctx, cudaError_t e, const char* what) {
  std::string msg = std::string("CUDA error ") + cudaGetErrorName(e) + " (" + cudaGetErrorString(e) +
                    ") in " + what;
  // Clear the sticky-free error so later calls report their own failures.
  cudaGetLastError();
  return fail(ctx, (e == cudaErrorNoDevice || e == cudaErrorInsufficientDriver) ? MPEGHB200_ERR_NO_DEVICE
                                                                               : MPEGHB200_ERR_CUDA,
              msg);
}

template <typename T>
static mpeghb200_status upload(mpeghb200_ctx* ctx, const std::vector<T>& host, const T** dev_out) {
  void* d = nullptr;
  MB_CUDA(ctx, cudaMalloc(&d, host.size() * sizeof(T)));
  ctx->owned.push_back(d);
  MB_CUDA(ctx, cudaMemcpy(d, host.data(), host.size() * sizeof(T), cudaMemcpyHostToDevice));
  *dev_out = static_cast<const T*>(d);
  return MPEGHB200_OK;
}

static std::vector<float2> interleave(const std::vector<float>& c, const std::vector<float>& s) {
  std::vector<float2> v(c.size());
  for (size_t i = 0; i < c.size(); ++i) v[i] = make_float2(c[i], s[i]);
  return v;
}

// Effective radix-4 twiddles.  The reference (decoder/impeghd_ifft.c:450-773) reads W, W^2, W^3 from a
// quarter-wave table w[0..256] = cos, w[257..513] = -sin and, once an index passes 256, switches to
// folded entries with rearranged products (four regions of j).  Every region is bit-identical to the
// plain form  (xr*c + xi*s, -(xr*s) + xi*c)  with a suitably signed/swapped pair (c, s), because
// negation commutes exactly with IEEE multiplication and addition; region 4's negated imaginary part is
// undone by the sign-swapped butterfly lines that follow it (:742-744).  So one pair per (j, power).
static std::vector<float> effective_twiddles(const std::vector<float>& w) {
  std::vector<float> e(256 * 6);
  for (int j = 0; j < 256; ++j) {
    float* o = &e[j * 6];
    o[0] = w[j];
    o[1] = w[j + 257];
    if (j <= 128) {
      o[2] = w[2 * j];
      o[3] = w[2 * j + 257];
    } else {
      o[2] = w[2 * j + 1];
      o[3] = -w[2 * j - 256];
    }
    if (j <= 85) {
      o[4] = w[3 * j];
      o[5] = w[3 * j + 257];
    } else if (j <= 170) {
      o[4] = w[3 * j + 1];
      o[5] = -w[3 * j - 256];
    } else {
      o[4] = -w[3 * j - 512];
      o[5] = -w[3 * j - 512 + 257];
    }
  }
  return e;
}

// Final radix-2 pass of the 512-point transform (decoder/impeghd_ifft.c:779-837): first half plain,
// second half restarts the table and multiplies by j*W instead.
static std::vector<float2> radix2_twiddles_512(const std::vector<float>& w) {
  std::vector<float2> r(256);
  for (int p = 0; p < 256; ++p) {
    const int k = (p < 128) ? p : p - 128;
    const float wc = w[2 * k], ws = w[2 * k + 257];
    r[p] = (p < 128) ? make_float2(wc, ws) : make_float2(ws, -wc);
  }
  return r;
}

static mpeghb200_status upload_tables(mpeghb200_ctx* ctx) {
  DeviceTables& t = ctx->tables;
  mpeghb200_status st;
  const auto& w = rom_table(ROM_FFT_TWIDDLE);
  if ((st = upload(ctx, interleave(rom_table(ROM_TWID_COS_512), rom_table(ROM_TWID_SIN_512)), &t.tw512)))
    return st;
  if ((st = upload(ctx, interleave(rom_table(ROM_TWID_COS_64), rom_table(ROM_TWID_SIN_64)), &t.tw64)))
    return st;
  if ((st = upload(ctx, effective_twiddles(w), &t.eff))) return st;
  if ((st = upload(ctx, radix2_twiddles_512(w), &t.rad2_512))) return st;
  if ((st = upload(ctx, rom_table(ROM_SINE_1024), &t.win_long[0]))) return st;
  if ((st = upload(ctx, rom_table(ROM_KBD_1024), &t.win_long[1]))) return st;
  if ((st = upload(ctx, rom_table(ROM_SINE_128), &t.win_short[0]))) return st;
  if ((st = upload(ctx, rom_table(ROM_KBD_128), &t.win_short[1]))) return st;
  if ((st = upload(ctx, rom_table(ROM_DCTDST_COS_1024), &t.dct_cos))) return st;
  if ((st = upload(ctx, rom_table(ROM_DCTDST_SIN_1024), &t.dct_sin))) return st;
  // lane-ordered packings for the fast path
  {
    const auto& tc = rom_table(ROM_TWID_COS_512);
    const auto& ts = rom_table(ROM_TWID_SIN_512);
    const std::vector<float> eff = effective_twiddles(w);
    const std::vector<float2> r2 = radix2_twiddles_512(w);
    std::vector<float4> pre(8 * 32), post(8 * 32), rad(8 * 32), e8(256 * 2);
    for (int pair = 0; pair < 8; ++pair)
      for (int lane = 0; lane < 32; ++lane) {
        const int c0 = lane + 32 * (2 * pair), c1 = lane + 32 * (2 * pair + 1);
        pre[pair * 32 + lane] = make_float4(tc[c0], ts[c0], tc[c1], ts[c1]);
        const int H = lane >> 4, m16 = lane & 15;
        const int p0 = m16 + 16 * (2 * pair), p1 = m16 + 16 * (2 * pair + 1);
        const int k0 = 256 * H + p0, k1 = 256 * H + p1;
        post[pair * 32 + lane] = make_float4(tc[k0], ts[k0], tc[k1], ts[k1]);
        rad[pair * 32 + lane] = make_float4(r2[p0].x, r2[p0].y, r2[p1].x, r2[p1].y);
      }
    for (int j = 0; j < 256; ++j) {
      e8[2 * j] = make_float4(eff[6 * j], eff[6 * j + 1], eff[6 * j + 2], eff[6 * j + 3]);
      e8[2 * j + 1] = make_float4(eff[6 * j + 4], eff[6 * j + 5], 0.f, 0.f);
    }
    for (int m = 1; m < 4; ++m)
      for (int q = 0; q < 6; ++q) t.pass1[m - 1][q] = eff[6 * 64 * m + q];
    if ((st = upload(ctx, pre, &t.pre4))) return st;
    if ((st = upload(ctx, post, &t.post4))) return st;
    if ((st = upload(ctx, rad, &t.rad2p4))) return st;
    if ((st = upload(ctx, e8, &t.eff8))) return st;
  }
  return MPEGHB200_OK;
}

}  // namespace mpeghb200

using namespace mpeghb200;

extern "C" {

const char* mpeghb200_version(void) { return "mpeghb200 0.1 (sm_100a)"; }

mpeghb200_status mpeghb200_create(mpeghb200_ctx** out, int device_ordinal) {
  if (!out) return MPEGHB200_ERR_BAD_ARG;
  *out = nullptr;
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n == 0) {
    cudaGetLastError();
    return fail(nullptr, MPEGHB200_ERR_NO_DEVICE,
                "no CUDA device visible: libmpeghb200 has no CPU fallback by design");
  }
  if (device_ordinal < 0 || device_ordinal >= n)
    return fail(nullptr, MPEGHB200_ERR_BAD_ARG, "device ordinal out of range");
  mpeghb200_ctx* ctx = new (std::nothrow) mpeghb200_ctx();
  if (!ctx) return MPEGHB200_ERR_NO_MEM;
  ctx->device = device_ordinal;
  if (const char* v = std::getenv("MPEGHB200_FD_VARIANT")) ctx->fd_variant = std::atoi(v);
  if (const char* v = std::getenv("MPEGHB200_LIM_VARIANT")) ctx->lim_variant = std::atoi(v);
  if (const char* v = std::getenv("MPEGHB200_NO_PDL")) mpeghb200::g_pdl_enabled = std::atoi(v) == 0;
  if (const char* v = std::getenv("MPEGHB200_PACK_GENERIC")) ctx->pack_generic = std::atoi(v) != 0;
  e = cudaSetDevice(device_ordinal);
  if (e == cudaSuccess) e = cudaDeviceGetAttribute(&ctx->sm_count, cudaDevAttrMultiProcessorCount, device_ordinal);
  if (e != cudaSuccess) {
    mpeghb200_status st = cuda_fail(nullptr, e, "cudaSetDevice");
    delete ctx;
    return st;
  }
  mpeghb200_status st = upload_tables(ctx);
  if (st != MPEGHB200_OK) {
    fail(nullptr, st, ctx->last_error);
    mpeghb200_destroy(ctx);
    return st;
  }
  *out = ctx;
  return MPEGHB200_OK;
}

void mpeghb200_destroy(mpeghb200_ctx* ctx) {
  if (!ctx) return;
  cudaSetDevice(ctx->device);
  for (void* p : ctx->owned) cudaFree(p);
  delete ctx;
}

const char* mpeghb200_last_error(const mpeghb200_ctx* ctx) {
  if (ctx) return ctx->last_error.c_str();
  std::lock_guard<std::mutex> lk(g_err_mu);
  static thread_local std::string copy;
  copy = g_create_error;
  return copy.c_str();
}

uint64_t mpeghb200_launch_count(const mpeghb200_ctx* ctx) { return ctx ? ctx->launches.load() : 0; }

const float* mpeghb200_rom_table(int id, int patched, int* n_out) {
  if (id < 0 || id >= ROM_COUNT) {
    if (n_out) *n_out = 0;
    return nullptr;
  }
  const std::vector<float>& t = rom_table(RomId(id), patched != 0);
  if (n_out) *n_out = int(t.size());
  return t.data();
}
int mpeghb200_rom_count(void) { return ROM_COUNT; }
const char* mpeghb200_rom_name(int id) { return rom_name(RomId(id)); }

mpeghb200_status mpeghb200_dev_alloc(mpeghb200_ctx* ctx, size_t bytes, void** dptr) {
  if (!ctx || !dptr) return MPEGHB200_ERR_BAD_ARG;
  MB_CUDA(ctx, cudaSetDevice(ctx->device));
  MB_CUDA(ctx, cudaMalloc(dptr, bytes));
  return MPEGHB200_OK;
}
mpeghb200_status mpeghb200_dev_free(mpeghb200_ctx* ctx, void* dptr) {
  if (!ctx) return MPEGHB200_ERR_BAD_ARG;
  MB_CUDA(ctx, cudaSetDevice(ctx->device));
  MB_CUDA(ctx, cudaFree(dptr));
  return MPEGHB200_OK;
}
mpeghb200_status mpeghb200_dev_memset(mpeghb200_ctx* ctx, void* dptr, int value, size_t bytes) {
  if (!ctx) return MPEGHB200_ERR_BAD_ARG;
  MB_CUDA(ctx, cudaSetDevice(ctx->device));
  MB_CUDA(ctx, cudaMemset(dptr, value, bytes));
  return MPEGHB200_OK;
}
mpeghb200_status mpeghb200_h2d(mpeghb200_ctx* ctx, void* dptr, const void* hptr, size_t bytes) {
  if (!ctx) return MPEGHB200_ERR_BAD_ARG;
  MB_CUDA(ctx, cudaSetDevice(ctx->device));
  MB_CUDA(ctx, cudaMemcpy(dptr, hptr, bytes, cudaMemcpyHostToDevice));
  return MPEGHB200_OK;
}
mpeghb200_status mpeghb200_d2h(mpeghb200_ctx* ctx, void* hptr, const void* dptr, size_t bytes) {
  if (!ctx) return MPEGHB200_ERR_BAD_ARG;
  MB_CUDA(ctx, cudaSetDevice(ctx->device));
  MB_CUDA(ctx, cudaMemcpy(hptr, dptr, bytes, cudaMemcpyDeviceToHost));
  return MPEGHB200_OK;
}
mpeghb200_status mpeghb200_host_alloc(mpeghb200_ctx* ctx, size_t bytes, void** hptr) {
  if (!ctx || !hptr) return MPEGHB200_ERR_BAD_ARG;
  MB_CUDA(ctx, cudaSetDevice(ctx->device));
  MB_CUDA(ctx, cudaHostAlloc(hptr, bytes, cudaHostAllocDefault));
  return MPEGHB200_OK;
}

mpeghb200_status mpeghb200_host_free(mpeghb200_ctx* ctx, void* hptr) {
  if (!ctx) return MPEGHB200_ERR_BAD_ARG;
  MB_CUDA(ctx, cudaFreeHost(hptr));
  return MPEGHB200_OK;
}

mpeghb200_status mpeghb200_sync(mpeghb200_ctx* ctx) {
  if (!ctx) return MPEGHB200_ERR_BAD_ARG;
  MB_CUDA(ctx, cudaSetDevice(ctx->device));
  MB_CUDA(ctx, cudaDeviceSynchronize());
  return MPEGHB200_OK;
}

}  // extern "C"
