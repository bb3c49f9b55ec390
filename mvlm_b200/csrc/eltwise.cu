// Memory-bound glue between convolutions (NHWC bf16, 16-byte vector accesses, 8 channels/thread):
//   pool2_act  : F.max_pool2d(x, 2)            (+ consumer BN+ReLU copy)   paulsenpredictor.py:309,314,319,324,329,411
//   bn_relu    : F.relu(bn(x)) for a second consumer of a stored tensor     paulsenpredictor.py:263-265
#include "common.cuh"
#include "stages.cuh"

namespace mvlm {

namespace {

__device__ __forceinline__ void unpack8(const uint4& q, float (&f)[8]) {
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&q);
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const float2 t = __bfloat1622float2(h[j]);
    f[2 * j] = t.x;
    f[2 * j + 1] = t.y;
  }
}
__device__ __forceinline__ uint4 pack8(const float (&f)[8]) {
  uint4 r;
  __nv_bfloat162 a = __floats2bfloat162_rn(f[0], f[1]);
  __nv_bfloat162 b = __floats2bfloat162_rn(f[2], f[3]);
  __nv_bfloat162 c = __floats2bfloat162_rn(f[4], f[5]);
  __nv_bfloat162 d = __floats2bfloat162_rn(f[6], f[7]);
  r.x = *reinterpret_cast<uint32_t*>(&a);
  r.y = *reinterpret_cast<uint32_t*>(&b);
  r.z = *reinterpret_cast<uint32_t*>(&c);
  r.w = *reinterpret_cast<uint32_t*>(&d);
  return r;
}
__device__ __forceinline__ void act8(const float (&v)[8], const float* __restrict__ scale,
                                     const float* __restrict__ shift, int c0, float (&o)[8]) {
  const float4 s0 = __ldg(reinterpret_cast<const float4*>(scale + c0));
  const float4 s1 = __ldg(reinterpret_cast<const float4*>(scale + c0 + 4));
  const float4 t0 = __ldg(reinterpret_cast<const float4*>(shift + c0));
  const float4 t1 = __ldg(reinterpret_cast<const float4*>(shift + c0 + 4));
  const float s[8] = {s0.x, s0.y, s0.z, s0.w, s1.x, s1.y, s1.z, s1.w};
  const float t[8] = {t0.x, t0.y, t0.z, t0.w, t1.x, t1.y, t1.z, t1.w};
#pragma unroll
  for (int j = 0; j < 8; ++j) o[j] = fmaxf(fmaf(v[j], s[j], t[j]), 0.f);
}

__global__ void __launch_bounds__(256) pool2_act_kernel(const uint4* __restrict__ in, int n, int h, int w, int c8,
                                                        uint4* __restrict__ out_raw, const float* __restrict__ scale,
                                                        const float* __restrict__ shift, uint4* __restrict__ out_act) {
  const int ho = h >> 1, wo = w >> 1;
  const size_t total = static_cast<size_t>(n) * ho * wo * c8;
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const int cv = static_cast<int>(i % c8);
    size_t r = i / c8;
    const int x = static_cast<int>(r % wo);
    r /= wo;
    const int y = static_cast<int>(r % ho);
    const int img = static_cast<int>(r / ho);
    const size_t base = ((static_cast<size_t>(img) * h + 2 * y) * w + 2 * x) * c8 + cv;
    float a[8], b[8], m[8];
    unpack8(__ldg(in + base), m);
    unpack8(__ldg(in + base + c8), a);
#pragma unroll
    for (int j = 0; j < 8; ++j) m[j] = fmaxf(m[j], a[j]);
    unpack8(__ldg(in + base + static_cast<size_t>(w) * c8), a);
    unpack8(__ldg(in + base + static_cast<size_t>(w) * c8 + c8), b);
#pragma unroll
    for (int j = 0; j < 8; ++j) m[j] = fmaxf(m[j], fmaxf(a[j], b[j]));
    if (out_raw) out_raw[i] = pack8(m);
    if (out_act) {
      float o[8];
      act8(m, scale, shift, cv * 8, o);
      out_act[i] = pack8(o);
    }
  }
}

__global__ void __launch_bounds__(256) bn_relu_kernel(const uint4* __restrict__ in, size_t total, int c8,
                                                      const float* __restrict__ scale, const float* __restrict__ shift,
                                                      uint4* __restrict__ out) {
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    float a[8], o[8];
    unpack8(__ldg(in + i), a);
    act8(a, scale, shift, static_cast<int>(i % c8) * 8, o);
    out[i] = pack8(o);
  }
}

inline int grid_for(size_t total) {
  const size_t b = (total + 255) / 256;
  const size_t cap = static_cast<size_t>(sm_count()) * 16;
  return static_cast<int>(b < cap ? (b ? b : 1) : cap);
}

}  // namespace

int pool2_act(const __nv_bfloat16* in, int n, int h, int w, int c, __nv_bfloat16* out_raw, const float* scale,
              const float* shift, __nv_bfloat16* out_act, cudaStream_t s) {
  MVLM_REQUIRE(in && (out_raw || out_act), "pool2_act: null pointer");
  MVLM_REQUIRE(c % 8 == 0 && h % 2 == 0 && w % 2 == 0, "pool2_act: bad shape %dx%dx%d", h, w, c);
  const size_t total = static_cast<size_t>(n) * (h / 2) * (w / 2) * (c / 8);
  pool2_act_kernel<<<grid_for(total), 256, 0, s>>>(reinterpret_cast<const uint4*>(in), n, h, w, c / 8,
                                                   reinterpret_cast<uint4*>(out_raw), scale, shift,
                                                   reinterpret_cast<uint4*>(out_act));
  count_launch();
  MVLM_CHECK_CUDA(cudaGetLastError());
  return MVLM_OK;
}

int bn_relu(const __nv_bfloat16* in, size_t npix, int c, const float* scale, const float* shift,
            __nv_bfloat16* out_act, cudaStream_t s) {
  MVLM_REQUIRE(in && out_act && scale && shift, "bn_relu: null pointer");
  MVLM_REQUIRE(c % 8 == 0, "bn_relu: channels %d not a multiple of 8", c);
  const size_t total = npix * (c / 8);
  bn_relu_kernel<<<grid_for(total), 256, 0, s>>>(reinterpret_cast<const uint4*>(in), total, c / 8, scale, shift,
                                                 reinterpret_cast<uint4*>(out_act));
  count_launch();
  MVLM_CHECK_CUDA(cudaGetLastError());
  return MVLM_OK;
}

}  // namespace mvlm
