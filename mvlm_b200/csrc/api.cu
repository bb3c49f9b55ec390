// C-ABI glue (include/mvlm_b200.h): error state, launch counter, thin wrappers.
#include <stdarg.h>

#include <atomic>

#include "../../include/mvlm_b200.h"
#include "common.cuh"
#include "conv_umma.cuh"

namespace mvlm {

static thread_local char g_err[1024] = "";
static std::atomic<long long> g_launches{0};

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

__global__ void pack_conv_weight_kernel(const float* __restrict__ w, int cout, int cin, int kh, int kw,
                                        int cout_pad, int cin_pad, __nv_bfloat16* __restrict__ out) {
  const long long total = 1ll * cout_pad * kw * kh * cin_pad;
  for (long long i = blockIdx.x * 1ll * blockDim.x + threadIdx.x; i < total;
       i += 1ll * gridDim.x * blockDim.x) {
    const int ci = static_cast<int>(i % cin_pad);
    long long r = i / cin_pad;
    const int ky = static_cast<int>(r % kh);
    r /= kh;
    const int kx = static_cast<int>(r % kw);
    const int co = static_cast<int>(r / kw);
    float v = 0.f;
    if (co < cout && ci < cin) v = w[((1ll * co * cin + ci) * kh + ky) * kw + kx];
    out[i] = __float2bfloat16_rn(v);
  }
}

}  // namespace mvlm

using namespace mvlm;

extern "C" {

const char* mvlm_last_error(void) { return g_err; }

int mvlm_version(void) { return 1; }

long long mvlm_launch_count(int reset) {
  return reset ? g_launches.exchange(0) : g_launches.load();
}

int mvlm_conv2d_bf16(const mvlm_conv_args* a, void* stream) {
  MVLM_REQUIRE(a != nullptr, "mvlm_conv2d_bf16: null args");
  ConvShape s;
  s.in = static_cast<const __nv_bfloat16*>(a->in);
  s.n = a->n; s.h = a->h; s.w = a->w; s.cin = a->cin; s.in_cs = a->in_cs;
  s.wpacked = static_cast<const __nv_bfloat16*>(a->wpacked);
  s.cout_pad = a->cout_pad; s.n_tile = a->n_tile; s.kh = a->kh; s.kw = a->kw;
  s.y_off0 = a->y_off0; s.x_off0 = a->x_off0;
  ConvEpilogue e;
  e.bias = a->bias;
  e.pre_scale = a->pre_scale; e.pre_shift = a->pre_shift;
  e.out_pre = static_cast<__nv_bfloat16*>(a->out_pre); e.pre_cs = a->pre_cs; e.pre_co = a->pre_co;
  e.res1 = static_cast<const __nv_bfloat16*>(a->res1); e.res1_cs = a->res1_cs; e.res1_co = a->res1_co;
  e.res2 = static_cast<const __nv_bfloat16*>(a->res2); e.res2_cs = a->res2_cs; e.res2_co = a->res2_co;
  e.out_raw = static_cast<__nv_bfloat16*>(a->out_raw); e.raw_cs = a->raw_cs; e.raw_co = a->raw_co;
  e.post_scale = a->post_scale; e.post_shift = a->post_shift;
  e.out_post = static_cast<__nv_bfloat16*>(a->out_post); e.post_cs = a->post_cs; e.post_co = a->post_co;
  e.out_f32 = a->out_f32; e.argmax_keys = a->argmax_keys; e.cout_real = a->cout_real;
  e.up_sy = a->up_sy > 0 ? a->up_sy : 1; e.up_sx = a->up_sx > 0 ? a->up_sx : 1;
  e.up_py = a->up_py; e.up_px = a->up_px;
  ConvParams p;
  int rc = conv_plan(s, e, &p);
  if (rc != MVLM_OK) return rc;
  return conv_launch(p, static_cast<cudaStream_t>(stream));
}

int mvlm_pack_conv_weight(const float* w_oihw, int cout, int cin, int kh, int kw, int cout_pad, int cin_pad,
                          void* out_bf16, void* stream) {
  MVLM_REQUIRE(w_oihw && out_bf16, "mvlm_pack_conv_weight: null pointer");
  MVLM_REQUIRE(cout_pad >= cout && cin_pad >= cin, "mvlm_pack_conv_weight: pads smaller than dims");
  const long long total = 1ll * cout_pad * kw * kh * cin_pad;
  const int block = 256;
  const int grid = static_cast<int>((total + block - 1) / block > 4096 ? 4096 : (total + block - 1) / block);
  pack_conv_weight_kernel<<<grid, block, 0, static_cast<cudaStream_t>(stream)>>>(
      w_oihw, cout, cin, kh, kw, cout_pad, cin_pad, static_cast<__nv_bfloat16*>(out_bf16));
  count_launch();
  MVLM_CHECK_CUDA(cudaGetLastError());
  return MVLM_OK;
}

}  // extern "C"
