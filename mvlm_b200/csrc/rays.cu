// 2D peak -> 3D view ray per (landmark, view).
//
// Replaces Estimator3D.estimate_landmark_lines (reference src/mvlm/utils/estimator3d.py:31-90):
//   y,x = peak row/col (float32 scalars, :61-62), scaled by hm_size/img_size (both = S, :66-67),
//   camera point (x/S*300-150, (S-1-y)/S*300-150, +-500) evaluated IN FLOAT32 (numpy scalar
//   arithmetic on np.float32 values, :73-79), widened to float64 and rotated by R^T (:83).
// The rotation matrices R = Ry*Rx*Rz are computed on the host in the dtype flow of the reference
// (mvlm_b200/utils/estimator3d.py) and passed as (V,9) float64.
// Compiled with --fmad=false so the float32 chain rounds exactly like numpy's.
#include "common.cuh"
#include "stages.cuh"

namespace mvlm {

namespace {

__global__ void rays_kernel(const float* __restrict__ peaks, const double* __restrict__ rot, int L, int V,
                            float S, double* __restrict__ starts, double* __restrict__ ends) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= L * V) return;
  const int v = i % V;
  float y = peaks[3 * i], x = peaks[3 * i + 1];
  y = (y / S) * S;
  x = (x / S) * S;
  const float xc = (x / S) * 300.0f + (-150.0f);
  const float yc = (((S - 1.0f) - y) / S) * 300.0f + (-150.0f);
  const double px = xc, py = yc;
  const double* R = rot + 9 * v;
  // world_j = sum_i R[i][j] * p_i   (t.T @ p)
#pragma unroll
  for (int j = 0; j < 3; ++j) {
    const double base = R[j] * px + R[3 + j] * py;
    starts[3 * i + j] = base + R[6 + j] * 500.0;
    ends[3 * i + j] = base + R[6 + j] * -500.0;
  }
}

}  // namespace

int rays_from_peaks(const float* peaks, const double* rot, int l, int v, int image_size, double* starts,
                    double* ends, cudaStream_t s) {
  MVLM_REQUIRE(peaks && rot && starts && ends, "rays: null pointer");
  MVLM_REQUIRE(l > 0 && v > 0 && image_size > 0, "rays: bad sizes");
  rays_kernel<<<ceil_div(l * v, 128), 128, 0, s>>>(peaks, rot, l, v, static_cast<float>(image_size), starts, ends);
  count_launch();
  MVLM_CHECK_CUDA(cudaGetLastError());
  return MVLM_OK;
}

}  // namespace mvlm
