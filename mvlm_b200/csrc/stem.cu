// CNN stem: conv1 3x3 (Cin<=4 -> 64, bias) + bn1 + ReLU, fused with the two BatchNorm+ReLU copies
// its consumers need (conv2.bn1 and conv2.resample.0) -- reference
// src/mvlm/prediction/paulsenpredictor.py:405-407 and :263-267.
//
// K = 9*Cin <= 36 is far too small for the tensor cores (arithmetic intensity 17 FLOP/B, SURVEY
// appendix A), so this runs on the fp32 pipes: one thread per output pixel, the 3x3xCin patch in
// registers, weights broadcast from shared memory as float4.  The image is read as the u8 NHWC4
// tensor the rasteriser wrote (value/255 in fp32 = exactly the reference's float32 stack) or as an
// fp32 NHWC stack (the reference interface `predict_landmarks_from_images(image_stack)`).
// Writes 2 x (N,H,W,64) bf16; the stem activation itself is never stored.
#include "common.cuh"
#include "stages.cuh"

namespace mvlm {

namespace {

constexpr int kCo = 64;
constexpr int kK = 36;  // 9 taps x 4 (zero padded) channels

__global__ void __launch_bounds__(256) stem_kernel(StemArgs g) {
  __shared__ __align__(16) float sw[kCo * kK];  // [co][tap][ci(4)]
  __shared__ float sb[kCo], s0[kCo], t0[kCo], sa[kCo], ta[kCo], sbb[kCo], tbb[kCo];
  __shared__ float tile[18 * 18 * 4];
  const int tid = threadIdx.x;
  for (int i = tid; i < kCo * kK; i += 256) {
    const int co = i / kK, k = i % kK, tap = k >> 2, ci = k & 3;
    sw[i] = ci < g.cin ? g.w_oihw[(co * g.cin + ci) * 9 + tap] : 0.f;
  }
  if (tid < kCo) {
    sb[tid] = g.bias[tid];
    s0[tid] = g.s0[tid]; t0[tid] = g.t0[tid];
    sa[tid] = g.sa[tid]; ta[tid] = g.ta[tid];
    sbb[tid] = g.sb[tid]; tbb[tid] = g.tb[tid];
  }
  const int tiles_x = (g.w + 15) >> 4, tiles_y = (g.h + 15) >> 4;
  const int total = g.n * tiles_x * tiles_y;
  for (int t = blockIdx.x; t < total; t += gridDim.x) {
    const int tx = t % tiles_x, ty = (t / tiles_x) % tiles_y, img = t / (tiles_x * tiles_y);
    __syncthreads();
    for (int i = tid; i < 18 * 18; i += 256) {
      const int yy = ty * 16 + i / 18 - 1, xx = tx * 16 + i % 18 - 1;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (yy >= 0 && yy < g.h && xx >= 0 && xx < g.w) {
        const size_t pix = (static_cast<size_t>(img) * g.h + yy) * g.w + xx;
        if (g.img_u8) {
          const uchar4 q = reinterpret_cast<const uchar4*>(g.img_u8)[pix];
          v = make_float4(q.x / 255.0f, q.y / 255.0f, q.z / 255.0f, q.w / 255.0f);
        } else {
          const float* p = g.img_f32 + pix * g.cin;
          v.x = p[0];
          if (g.cin > 1) v.y = p[1];
          if (g.cin > 2) v.z = p[2];
          if (g.cin > 3) v.w = p[3];
        }
      }
      reinterpret_cast<float4*>(tile)[i] = v;
    }
    __syncthreads();
    const int ly = tid >> 4, lx = tid & 15;
    const int y = ty * 16 + ly, x = tx * 16 + lx;
    float in[kK];
#pragma unroll
    for (int tap = 0; tap < 9; ++tap) {
      const float4 v = reinterpret_cast<const float4*>(tile)[(ly + tap / 3) * 18 + lx + tap % 3];
      in[4 * tap] = v.x; in[4 * tap + 1] = v.y; in[4 * tap + 2] = v.z; in[4 * tap + 3] = v.w;
    }
    if (y < g.h && x < g.w) {
      const size_t pix = (static_cast<size_t>(img) * g.h + y) * g.w + x;
#pragma unroll 1
      for (int c0 = 0; c0 < kCo; c0 += 8) {
        float oa[8], ob[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const int co = c0 + j;
          float acc = sb[co];
          const float4* wr = reinterpret_cast<const float4*>(sw + co * kK);
#pragma unroll
          for (int k4 = 0; k4 < 9; ++k4) {
            const float4 w4 = wr[k4];
            acc = fmaf(in[4 * k4], w4.x, acc);
            acc = fmaf(in[4 * k4 + 1], w4.y, acc);
            acc = fmaf(in[4 * k4 + 2], w4.z, acc);
            acc = fmaf(in[4 * k4 + 3], w4.w, acc);
          }
          const float x0 = fmaxf(fmaf(acc, s0[co], t0[co]), 0.f);
          oa[j] = fmaxf(fmaf(x0, sa[co], ta[co]), 0.f);
          ob[j] = fmaxf(fmaf(x0, sbb[co], tbb[co]), 0.f);
        }
        __nv_bfloat162 pa[4], pb[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          pa[j] = __floats2bfloat162_rn(oa[2 * j], oa[2 * j + 1]);
          pb[j] = __floats2bfloat162_rn(ob[2 * j], ob[2 * j + 1]);
        }
        *reinterpret_cast<uint4*>(g.out_a + pix * kCo + c0) = *reinterpret_cast<uint4*>(pa);
        *reinterpret_cast<uint4*>(g.out_b + pix * kCo + c0) = *reinterpret_cast<uint4*>(pb);
      }
    }
  }
}

}  // namespace

int stem_launch(const StemArgs& a, cudaStream_t s) {
  MVLM_REQUIRE((a.img_u8 != nullptr) != (a.img_f32 != nullptr), "stem: exactly one of img_u8 / img_f32");
  MVLM_REQUIRE(a.cin >= 1 && a.cin <= 4, "stem: cin=%d unsupported", a.cin);
  MVLM_REQUIRE(a.w_oihw && a.bias && a.s0 && a.t0 && a.sa && a.ta && a.sb && a.tb && a.out_a && a.out_b,
               "stem: null pointer");
  const int tiles = a.n * ceil_div(a.w, 16) * ceil_div(a.h, 16);
  const int grid = tiles < kNumSMs * 4 ? tiles : kNumSMs * 4;
  stem_kernel<<<grid, 256, 0, s>>>(a);
  count_launch();
  MVLM_CHECK_CUDA(cudaGetLastError());
  return MVLM_OK;
}

}  // namespace mvlm
