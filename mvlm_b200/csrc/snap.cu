// Snap landmarks to the closest point of the triangle mesh (exact, brute force over all triangles).
//
// Replaces Estimator3D.project_landmarks_to_surface (reference src/mvlm/utils/estimator3d.py:252-285):
// vtkCellLocator.FindClosestPoint returns the exact closest point on the mesh; the locator is only
// an accelerator, so a full scan gives the same answer.  Ties -> lowest triangle id.
// grid (L, splits): each block scans a slice of the triangles (vertex/index arrays are L2-resident:
// 1.8 MB for 100k triangles), block-reduces (dist^2, tri) and a finalize kernel picks the winner.
#include "common.cuh"
#include "stages.cuh"

namespace mvlm {

namespace {

constexpr int kSnapPart = 5;  // d2, tri, p[3]

__device__ void closest_on_tri(const double* p, const double* a, const double* b, const double* c, double* out) {
  double ab[3], ac[3], ap[3], bp[3], cp[3];
#pragma unroll
  for (int i = 0; i < 3; ++i) { ab[i] = b[i] - a[i]; ac[i] = c[i] - a[i]; ap[i] = p[i] - a[i]; }
  const double d1 = ab[0] * ap[0] + ab[1] * ap[1] + ab[2] * ap[2];
  const double d2 = ac[0] * ap[0] + ac[1] * ap[1] + ac[2] * ap[2];
  if (d1 <= 0.0 && d2 <= 0.0) { out[0] = a[0]; out[1] = a[1]; out[2] = a[2]; return; }
#pragma unroll
  for (int i = 0; i < 3; ++i) bp[i] = p[i] - b[i];
  const double d3 = ab[0] * bp[0] + ab[1] * bp[1] + ab[2] * bp[2];
  const double d4 = ac[0] * bp[0] + ac[1] * bp[1] + ac[2] * bp[2];
  if (d3 >= 0.0 && d4 <= d3) { out[0] = b[0]; out[1] = b[1]; out[2] = b[2]; return; }
  const double vc = d1 * d4 - d3 * d2;
  if (vc <= 0.0 && d1 >= 0.0 && d3 <= 0.0) {
    const double v = d1 / (d1 - d3);
#pragma unroll
    for (int i = 0; i < 3; ++i) out[i] = a[i] + v * ab[i];
    return;
  }
#pragma unroll
  for (int i = 0; i < 3; ++i) cp[i] = p[i] - c[i];
  const double d5 = ab[0] * cp[0] + ab[1] * cp[1] + ab[2] * cp[2];
  const double d6 = ac[0] * cp[0] + ac[1] * cp[1] + ac[2] * cp[2];
  if (d6 >= 0.0 && d5 <= d6) { out[0] = c[0]; out[1] = c[1]; out[2] = c[2]; return; }
  const double vb = d5 * d2 - d1 * d6;
  if (vb <= 0.0 && d2 >= 0.0 && d6 <= 0.0) {
    const double w = d2 / (d2 - d6);
#pragma unroll
    for (int i = 0; i < 3; ++i) out[i] = a[i] + w * ac[i];
    return;
  }
  const double va = d3 * d6 - d5 * d4;
  if (va <= 0.0 && (d4 - d3) >= 0.0 && (d5 - d6) >= 0.0) {
    const double w = (d4 - d3) / ((d4 - d3) + (d5 - d6));
#pragma unroll
    for (int i = 0; i < 3; ++i) out[i] = b[i] + w * (c[i] - b[i]);
    return;
  }
  const double denom = 1.0 / (va + vb + vc);
  const double v = vb * denom, w = vc * denom;
#pragma unroll
  for (int i = 0; i < 3; ++i) out[i] = a[i] + ab[i] * v + ac[i] * w;
}

__global__ void __launch_bounds__(256) snap_scan_kernel(const float* __restrict__ verts, const int* __restrict__ tris,
                                                        int nt, const double* __restrict__ lm, int splits,
                                                        double* __restrict__ part) {
  const int l = blockIdx.x, split = blockIdx.y;
  const double p[3] = {lm[3 * l], lm[3 * l + 1], lm[3 * l + 2]};
  const int chunk = (nt + splits - 1) / splits;
  const int t0 = split * chunk, t1 = min(nt, t0 + chunk);
  double best = INFINITY;
  int bt = 0x7fffffff;
  for (int t = t0 + threadIdx.x; t < t1; t += blockDim.x) {
    const int i0 = __ldg(tris + 3 * t), i1 = __ldg(tris + 3 * t + 1), i2 = __ldg(tris + 3 * t + 2);
    const double a[3] = {verts[3 * i0], verts[3 * i0 + 1], verts[3 * i0 + 2]};
    const double b[3] = {verts[3 * i1], verts[3 * i1 + 1], verts[3 * i1 + 2]};
    const double c[3] = {verts[3 * i2], verts[3 * i2 + 1], verts[3 * i2 + 2]};
    double q[3];
    closest_on_tri(p, a, b, c, q);
    const double dx = q[0] - p[0], dy = q[1] - p[1], dz = q[2] - p[2];
    const double d = dx * dx + dy * dy + dz * dz;
    if (d < best) { best = d; bt = t; }  // increasing t per thread: first minimum kept
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const double od = __shfl_xor_sync(0xffffffffu, best, o);
    const int ot = __shfl_xor_sync(0xffffffffu, bt, o);
    if (od < best || (od == best && ot < bt)) { best = od; bt = ot; }
  }
  __shared__ double sd[8];
  __shared__ int st[8];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) { sd[warp] = best; st[warp] = bt; }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < 8; ++w)
      if (sd[w] < best || (sd[w] == best && st[w] < bt)) { best = sd[w]; bt = st[w]; }
    double* o = part + (static_cast<size_t>(l) * splits + split) * kSnapPart;
    o[0] = best;
    o[1] = static_cast<double>(bt);
    if (bt != 0x7fffffff) {
      const int i0 = tris[3 * bt], i1 = tris[3 * bt + 1], i2 = tris[3 * bt + 2];
      const double a[3] = {verts[3 * i0], verts[3 * i0 + 1], verts[3 * i0 + 2]};
      const double b[3] = {verts[3 * i1], verts[3 * i1 + 1], verts[3 * i1 + 2]};
      const double c[3] = {verts[3 * i2], verts[3 * i2 + 1], verts[3 * i2 + 2]};
      closest_on_tri(p, a, b, c, o + 2);
    } else {
      o[2] = p[0]; o[3] = p[1]; o[4] = p[2];
    }
  }
}

__global__ void snap_finalize_kernel(const double* __restrict__ part, int L, int splits, double* __restrict__ out,
                                     int* __restrict__ out_tri) {
  const int l = blockIdx.x * blockDim.x + threadIdx.x;
  if (l >= L) return;
  const double* q = part + static_cast<size_t>(l) * splits * kSnapPart;
  int bs = 0;
  for (int s = 1; s < splits; ++s)
    if (q[s * kSnapPart] < q[bs * kSnapPart] ||
        (q[s * kSnapPart] == q[bs * kSnapPart] && q[s * kSnapPart + 1] < q[bs * kSnapPart + 1]))
      bs = s;
  out[3 * l] = q[bs * kSnapPart + 2];
  out[3 * l + 1] = q[bs * kSnapPart + 3];
  out[3 * l + 2] = q[bs * kSnapPart + 4];
  if (out_tri) out_tri[l] = static_cast<int>(q[bs * kSnapPart + 1]);
}

int snap_splits(int l, int nt) {
  int s = ceil_div(4 * kNumSMs, l);
  const int max_s = ceil_div(nt, 256);
  if (s > max_s) s = max_s;
  return s < 1 ? 1 : s;
}

}  // namespace

size_t snap_workspace_bytes(int l, int nt) {
  return static_cast<size_t>(l) * snap_splits(l, nt) * kSnapPart * sizeof(double) + 64;
}

int snap_launch(const float* verts, const int* tris, int nt, const double* lm, int l, void* workspace,
                size_t workspace_bytes, double* out, int* out_tri, cudaStream_t s) {
  MVLM_REQUIRE(verts && tris && lm && out && workspace, "snap: null pointer");
  MVLM_REQUIRE(nt > 0 && l > 0, "snap: bad sizes");
  MVLM_REQUIRE(workspace_bytes >= snap_workspace_bytes(l, nt), "snap: workspace too small");
  const int splits = snap_splits(l, nt);
  snap_scan_kernel<<<dim3(l, splits), 256, 0, s>>>(verts, tris, nt, lm, splits, static_cast<double*>(workspace));
  snap_finalize_kernel<<<ceil_div(l, 128), 128, 0, s>>>(static_cast<double*>(workspace), l, splits, out, out_tri);
  count_launch(2);
  MVLM_CHECK_CUDA(cudaGetLastError());
  return MVLM_OK;
}

}  // namespace mvlm
