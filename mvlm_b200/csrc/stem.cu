// CNN stem input staging.  conv1 (3x3, Cin<=4 -> 64, bias) + bn1 + ReLU (reference
// src/mvlm/prediction/paulsenpredictor.py:405-407) runs on the tcgen05 conv kernel like every other layer
// (epilogue: bias -> bn1+ReLU -> the two BatchNorm+ReLU copies its consumers conv2.bn1 / conv2.resample.0
// need, :263-267).  The tensor cores take bf16 operands, but the image must not be quantised (the reference
// feeds float32 in [0,1]); so each channel value v is split into two bf16 terms
//     hi = bf16(v),  lo = bf16(v - hi)          (v - (hi + lo) <= 2^-17 |v|)
// stored in channels 0..3 (hi) and 4..7 (lo) of a 16-channel bf16 NHWC image (one 32-byte TMA row per
// pixel; channels 8..15 zero) and conv1's weights are repeated for both halves.  The rasteriser's u8 image
// (v = byte/255 in fp32, exactly the reference's stack) and an fp32 stack go through the same arithmetic.
#include "common.cuh"
#include "stages.cuh"

namespace mvlm {

namespace {

__global__ void __launch_bounds__(256) image_to_hilo16_kernel(const unsigned char* __restrict__ u8,
                                                              const float* __restrict__ f32, int cin, size_t npix,
                                                              uint4* __restrict__ out) {
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < npix;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    float v[4] = {0.f, 0.f, 0.f, 0.f};
    if (u8) {
      const uchar4 q = reinterpret_cast<const uchar4*>(u8)[i];
      v[0] = q.x / 255.0f; v[1] = q.y / 255.0f; v[2] = q.z / 255.0f; v[3] = q.w / 255.0f;
    } else {
      const float* p = f32 + i * cin;
      for (int c = 0; c < cin; ++c) v[c] = p[c];
    }
    __nv_bfloat16 h[8];
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      const float x = c < cin ? v[c] : 0.f;
      h[c] = __float2bfloat16_rn(x);
      h[4 + c] = __float2bfloat16_rn(x - __bfloat162float(h[c]));
    }
    out[2 * i] = *reinterpret_cast<uint4*>(h);
    out[2 * i + 1] = make_uint4(0, 0, 0, 0);
  }
}

}  // namespace

int image_to_hilo16(const unsigned char* u8, const float* f32, int cin, size_t npix, __nv_bfloat16* out,
                    cudaStream_t s) {
  MVLM_REQUIRE((u8 != nullptr) != (f32 != nullptr), "stem: exactly one of img_u8 / img_f32");
  MVLM_REQUIRE(cin >= 1 && cin <= 4 && out, "stem: bad arguments");
  const size_t blocks = (npix + 255) / 256;
  const int grid = static_cast<int>(blocks < static_cast<size_t>(sm_count()) * 16 ? blocks : static_cast<size_t>(sm_count()) * 16);
  image_to_hilo16_kernel<<<grid, 256, 0, s>>>(u8, f32, cin, npix, reinterpret_cast<uint4*>(out));
  count_launch();
  MVLM_CHECK_CUDA(cudaGetLastError());
  return MVLM_OK;
}

}  // namespace mvlm
