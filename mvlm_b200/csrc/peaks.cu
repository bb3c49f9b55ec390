// Heat-map peak extraction with warp-shuffle reductions.
//
// Replaces PaulsenModel.find_heat_map_maxima / find_maxima_in_batch_of_heatmaps
// (reference src/mvlm/prediction/paulsenpredictor.py:112-158, :160-165):
//   simple : np.argmax over the flattened (H,W) map = FIRST maximum in row-major order, a NaN
//            wins as the first NaN (:123); output (row-1, col-0.5, value) (:127)
//   moment : additionally centre-of-mass of the raw 31x31 window around the peak when
//            16 <= row,col <= H-16 (:141-156); accumulated in fp64.
// Index selection is bit-exact; HBM-bound (each heat-map element is read once, 16-byte loads).
#include "common.cuh"
#include "stages.cuh"

namespace mvlm {

namespace {

struct Best {
  float v;
  int i;
};

// numpy argmax order: NaN beats everything, then larger value, then lower index.
__device__ __forceinline__ bool better(float av, int ai, float bv, int bi) {
  const bool an = av != av, bn = bv != bv;
  if (an || bn) {
    if (an && bn) return ai < bi;
    return an;
  }
  if (av > bv) return true;
  if (av < bv) return false;
  return ai < bi;
}

__device__ __forceinline__ Best warp_best(Best b) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float ov = __shfl_xor_sync(0xffffffffu, b.v, o);
    const int oi = __shfl_xor_sync(0xffffffffu, b.i, o);
    if (better(ov, oi, b.v, b.i)) {
      b.v = ov;
      b.i = oi;
    }
  }
  return b;
}

__global__ void __launch_bounds__(256) peaks_kernel(const float* __restrict__ hm, int V, int L, int H, int W,
                                                    int method, float* __restrict__ peaks) {
  const int vl = blockIdx.x;  // v * L + l
  const int v = vl / L, l = vl % L;
  const float* map = hm + static_cast<size_t>(vl) * H * W;
  const int n = H * W;
  Best b{-INFINITY, 0x7fffffff};
  if ((n & 3) == 0) {
    const float4* m4 = reinterpret_cast<const float4*>(map);
    for (int i = threadIdx.x; i < (n >> 2); i += blockDim.x) {
      const float4 q = __ldg(m4 + i);
      const float e[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
      for (int k = 0; k < 4; ++k)
        if (better(e[k], 4 * i + k, b.v, b.i)) { b.v = e[k]; b.i = 4 * i + k; }
    }
  } else {
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
      const float e = __ldg(map + i);
      if (better(e, i, b.v, b.i)) { b.v = e; b.i = i; }
    }
  }
  __shared__ float sv[8];
  __shared__ int si[8];
  __shared__ double acc[3];
  b = warp_best(b);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) { sv[warp] = b.v; si[warp] = b.i; }
  if (threadIdx.x < 3) acc[threadIdx.x] = 0.0;
  __syncthreads();
  Best r{sv[0], si[0]};
  for (int k = 1; k < static_cast<int>(blockDim.x >> 5); ++k)
    if (better(sv[k], si[k], r.v, r.i)) { r.v = sv[k]; r.i = si[k]; }
  const int row = r.i / W, col = r.i % W;
  double frow = row, fcol = col;
  if (method == 1) {
    const int sz = 15;
    // reference window test uses the map height for both axes (square maps), :141
    if (row > sz && H - row > sz && col > sz && H - col > sz) {
      double s_tot = 0.0, s_row = 0.0, s_col = 0.0;
      for (int i = threadIdx.x; i < 31 * 31; i += blockDim.x) {
        const int dr = i / 31, dc = i % 31;
        const double e = static_cast<double>(map[(row - sz + dr) * W + (col - sz + dc)]);
        s_tot += e;
        s_row += e * dr;
        s_col += e * dc;
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        s_tot += __shfl_xor_sync(0xffffffffu, s_tot, o);
        s_row += __shfl_xor_sync(0xffffffffu, s_row, o);
        s_col += __shfl_xor_sync(0xffffffffu, s_col, o);
      }
      if (lane == 0) {
        atomicAdd(&acc[0], s_tot);
        atomicAdd(&acc[1], s_row);
        atomicAdd(&acc[2], s_col);
      }
      __syncthreads();
      frow = row + (acc[1] / acc[0] - sz);
      fcol = col + (acc[2] / acc[0] - sz);
    }
  }
  if (threadIdx.x == 0) {
    float* o = peaks + (static_cast<size_t>(l) * V + v) * 3;
    o[0] = static_cast<float>(frow - 1.0);
    o[1] = static_cast<float>(fcol - 0.5);
    o[2] = r.v;
  }
}

__device__ __forceinline__ float unorder_f32(unsigned int o) {
  const unsigned int b = (o & 0x80000000u) ? (o & 0x7FFFFFFFu) : ~o;
  return __uint_as_float(b);
}

__global__ void peaks_from_keys_kernel(const unsigned long long* __restrict__ keys, int V, int L, int W,
                                       float* __restrict__ peaks) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;  // v * L + l
  if (i >= V * L) return;
  const int v = i / L, l = i % L;
  const unsigned long long k = keys[i];
  const unsigned int idx = 0xFFFFFFFFu - static_cast<unsigned int>(k & 0xFFFFFFFFu);
  const float val = unorder_f32(static_cast<unsigned int>(k >> 32));
  const int row = idx / W, col = idx % W;
  float* o = peaks + (static_cast<size_t>(l) * V + v) * 3;
  o[0] = static_cast<float>(row - 1);
  o[1] = static_cast<float>(col) - 0.5f;
  o[2] = val;
}

// "moment" selection on the fused path (paulsenpredictor.py:129-156): the 31x31 window of heat-map values around the
// fused arg-max is RE-EVALUATED here from the last layer's input instead of being read from materialised heat maps
// (1.9 GB of fp32 written and read back at 100 views otherwise): conv11 on the nearest-x2 up-sampled conv10 output, as
// the same four 2x2 phase kernels the tensor-core path uses (bf16 operands, fp32 accumulation, + bias), 961 pixels x
// 4 taps x Lp channels per (view, landmark).  One block per (view, landmark); sums in fp64 like peaks_kernel.
struct PhaseWeights {
  const __nv_bfloat16* w[4];  // phase (a, b) -> [cout_pad][kx(2)][ky(2)][cin_pad], index 2 * a + b
};

__global__ void __launch_bounds__(256) moment_from_keys_kernel(const unsigned long long* __restrict__ keys,
                                                               const __nv_bfloat16* __restrict__ x, int cs, int cin,
                                                               PhaseWeights pw, const float* __restrict__ bias, int V,
                                                               int L, int H, int W, float* __restrict__ peaks) {
  extern __shared__ float wl[];  // [4 phases][kx][ky][cin] of output channel l
  const int vl = blockIdx.x;
  const int v = vl / L, l = vl % L;
  const unsigned long long k = keys[vl];
  const unsigned int idx = 0xFFFFFFFFu - static_cast<unsigned int>(k & 0xFFFFFFFFu);
  const float val = unorder_f32(static_cast<unsigned int>(k >> 32));
  const int row = idx / W, col = idx % W;
  double frow = row, fcol = col;
  const int sz = 15;
  __shared__ double acc[3];
  // reference window test uses the map height for both axes (square maps), :141
  if (row > sz && H - row > sz && col > sz && H - col > sz) {
    for (int i = threadIdx.x; i < 16 * cin; i += blockDim.x) {
      const int ph = i / (4 * cin), r = i % (4 * cin);
      wl[i] = __bfloat162float(pw.w[ph][static_cast<size_t>(l) * 4 * cin + r]);
    }
    if (threadIdx.x < 3) acc[threadIdx.x] = 0.0;
    __syncthreads();
    const int h2 = H >> 1, w2 = W >> 1;
    const float b = bias[l];
    double s_tot = 0.0, s_row = 0.0, s_col = 0.0;
    for (int i = threadIdx.x; i < 31 * 31; i += blockDim.x) {
      const int dr = i / 31, dc = i % 31;
      const int y = row - sz + dr, xx = col - sz + dc;
      const int pa = y & 1, pb = xx & 1;
      const float* wp = wl + (2 * pa + pb) * 4 * cin;
      float a = 0.f;
#pragma unroll
      for (int kx = 0; kx < 2; ++kx) {
#pragma unroll
        for (int ky = 0; ky < 2; ++ky) {
          const int yy = (y >> 1) + ky + pa - 1, xl = (xx >> 1) + kx + pb - 1;
          if (yy < 0 || yy >= h2 || xl < 0 || xl >= w2) continue;  // zero padding of the convolution
          const uint4* px = reinterpret_cast<const uint4*>(x + ((static_cast<size_t>(v) * h2 + yy) * w2 + xl) * cs);
          const float* wt = wp + (kx * 2 + ky) * cin;
          for (int c8 = 0; c8 < (cin >> 3); ++c8) {
            const uint4 q = __ldg(px + c8);
            const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&q);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const float2 t = __bfloat1622float2(h[j]);
              a = fmaf(t.x, wt[8 * c8 + 2 * j], a);
              a = fmaf(t.y, wt[8 * c8 + 2 * j + 1], a);
            }
          }
        }
      }
      const double e = static_cast<double>(a + b);
      s_tot += e;
      s_row += e * dr;
      s_col += e * dc;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      s_tot += __shfl_xor_sync(0xffffffffu, s_tot, o);
      s_row += __shfl_xor_sync(0xffffffffu, s_row, o);
      s_col += __shfl_xor_sync(0xffffffffu, s_col, o);
    }
    if ((threadIdx.x & 31) == 0) {
      atomicAdd(&acc[0], s_tot);
      atomicAdd(&acc[1], s_row);
      atomicAdd(&acc[2], s_col);
    }
    __syncthreads();
    frow = row + (acc[1] / acc[0] - sz);
    fcol = col + (acc[2] / acc[0] - sz);
  }
  if (threadIdx.x == 0) {
    float* o = peaks + (static_cast<size_t>(l) * V + v) * 3;
    o[0] = static_cast<float>(frow - 1.0);
    o[1] = static_cast<float>(fcol - 0.5);
    o[2] = val;
  }
}

// keys gathered from `world` ranks, slot r = (slot_views x L) keys of rank r's contiguous view block (the first
// V % world ranks hold one view more, sharding.split_views); unused slot rows are never read
__global__ void peaks_from_gathered_keys_kernel(const unsigned long long* __restrict__ keys, int V, int L, int W, int world,
                                                int slot_views, float* __restrict__ peaks) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;  // v * L + l
  if (i >= V * L) return;
  const int v = i / L, l = i % L;
  const int base = V / world, rem = V % world;
  // rank r starts at r * base + min(r, rem)
  int r = (v < rem * (base + 1)) ? v / (base + 1) : rem + (v - rem * (base + 1)) / max(base, 1);
  const int start = r * base + min(r, rem);
  const unsigned long long k = keys[(static_cast<size_t>(r) * slot_views + (v - start)) * L + l];
  const unsigned int idx = 0xFFFFFFFFu - static_cast<unsigned int>(k & 0xFFFFFFFFu);
  const float val = unorder_f32(static_cast<unsigned int>(k >> 32));
  const int row = idx / W, col = idx % W;
  float* o = peaks + (static_cast<size_t>(l) * V + v) * 3;
  o[0] = static_cast<float>(row - 1);
  o[1] = static_cast<float>(col) - 0.5f;
  o[2] = val;
}

}  // namespace

int peaks_from_gathered_keys(const unsigned long long* keys, int v, int l, int w, int world, int slot_views, float* peaks,
                             cudaStream_t s) {
  MVLM_REQUIRE(keys && peaks, "peaks_from_gathered_keys: null pointer");
  MVLM_REQUIRE(v > 0 && l > 0 && w > 0 && world > 0 && world <= v && slot_views >= ceil_div(v, world),
               "peaks_from_gathered_keys: bad layout (%d views, %d ranks, %d views per slot)", v, world, slot_views);
  peaks_from_gathered_keys_kernel<<<ceil_div(v * l, 256), 256, 0, s>>>(keys, v, l, w, world, slot_views, peaks);
  count_launch();
  MVLM_CHECK_CUDA(cudaGetLastError());
  return MVLM_OK;
}

int peaks_moment_from_keys(const unsigned long long* keys, const __nv_bfloat16* x, int x_cs, int cin,
                           const __nv_bfloat16* const* phase_w, const float* bias, int v, int l, int h, int w, float* peaks,
                           cudaStream_t s) {
  MVLM_REQUIRE(keys && x && phase_w && bias && peaks, "peaks_moment_from_keys: null pointer");
  MVLM_REQUIRE(cin % 8 == 0 && x_cs % 8 == 0 && h % 2 == 0 && w % 2 == 0, "peaks_moment_from_keys: bad shape");
  PhaseWeights pw;
  for (int i = 0; i < 4; ++i) pw.w[i] = phase_w[i];
  moment_from_keys_kernel<<<v * l, 256, 16 * cin * sizeof(float), s>>>(keys, x, x_cs, cin, pw, bias, v, l, h, w, peaks);
  count_launch();
  MVLM_CHECK_CUDA(cudaGetLastError());
  return MVLM_OK;
}

int peaks_from_heatmaps(const float* hm, int v, int l, int h, int w, int method, float* peaks, cudaStream_t s) {
  MVLM_REQUIRE(hm && peaks, "peaks: null pointer");
  MVLM_REQUIRE(v > 0 && l > 0 && h > 0 && w > 0, "peaks: bad sizes");
  MVLM_REQUIRE(method == 0 || method == 1, "peaks: unknown selection method %d", method);
  peaks_kernel<<<v * l, 256, 0, s>>>(hm, v, l, h, w, method, peaks);
  count_launch();
  MVLM_CHECK_CUDA(cudaGetLastError());
  return MVLM_OK;
}

int peaks_from_keys(const unsigned long long* keys, int v, int l, int h, int w, float* peaks, cudaStream_t s) {
  MVLM_REQUIRE(keys && peaks, "peaks_from_keys: null pointer");
  (void)h;
  peaks_from_keys_kernel<<<ceil_div(v * l, 256), 256, 0, s>>>(keys, v, l, w, peaks);
  count_launch();
  MVLM_CHECK_CUDA(cudaGetLastError());
  return MVLM_OK;
}

}  // namespace mvlm
