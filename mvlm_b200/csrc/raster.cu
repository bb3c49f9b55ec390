// Batched multi-view orthographic rasteriser: all V views of one mesh in one launch sequence.
//
// Replaces ObjVTKRenderer3D.render_3d_multi_rgb_geometry_depth (reference
// src/mvlm/utils/render3d.py:114-177; camera :53-59,:136,:150-152; depth byte encoder
// :73-77,:166-170; row flip :177; /255 :191) and obj_to_actor's material
// (src/mvlm/utils/utils3d.py:26-64: nearest-neighbour texture, ambient 1 / diffuse 0).
//
//   raster_xform   one thread per (view, vertex): rotate the vertex (double, like
//                  vtkTransformPolyDataFilter) and map it to the window in fp32 ONCE; the (x, y, z-buffer value)
//                  triple is kept as a float4 per (view, vertex) for the two kernels below.  (Round 1 re-did this
//                  per incident triangle and again per covered pixel: ~6x + 3x redundant fp64 work, and the
//                  triangle kernel was instruction-bound on it.)
//   raster_tris    one thread per (view, triangle): three 16-byte gathers, edge functions, and an atomicMin of
//                  the packed (depth bits << 32 | triangle id) key on every covered pixel centre.  The meshes are
//                  micro-polygon (~0.5 px/triangle at 256^2, 0.1 at config 4), so the work is triangle setup plus
//                  < 1 atomic per triangle; a per-view z-buffer is 512 KB and stays in L2.  Binning triangles into
//                  screen tiles with a shared-memory z-tile (the textbook design for LARGE triangles) costs two more
//                  passes over the (view, triangle) pairs to save less than one global atomic each: DESIGN.md 7d.
//   raster_resolve one thread per pixel: winner triangle -> barycentric uv -> nearest texel (one 4-byte load from
//                  the RGBA texture), depth byte, optional geometry shade; writes the u8 NHWC4 image the CNN stem
//                  consumes and, on request, the fp32 (V,H,W,C) stack / triangle-id / z maps.
//
// This file is compiled with --fmad=false: every fp32 operation rounds separately, in the same
// order as oracle/csrc/oracle_native.c (built with -ffp-contract=off), so triangle-id maps and
// depth bytes are bit-identical to the CPU oracle.
#include <algorithm>

#include "common.cuh"
#include "stages.cuh"

namespace mvlm {

namespace {

constexpr unsigned long long kBgKey = 0xFFFFFFFFFFFFFFFFull;

__device__ __forceinline__ float edge_fn(float ax, float ay, float bx, float by, float cx, float cy) {
  const float d1 = bx - ax, d2 = cy - ay, d3 = by - ay, d4 = cx - ax;
  const float p = d1 * d2;
  const float q = d3 * d4;
  return p - q;
}

struct SV {
  float sx, sy, zb;
};

__device__ __forceinline__ SV xform_vertex(const float* __restrict__ v, const double* R, float kx, float ky) {
  const double x = v[0], y = v[1], z = v[2];
  const float xr = static_cast<float>((R[0] * x + R[1] * y) + R[2] * z);
  const float yr = static_cast<float>((R[3] * x + R[4] * y) + R[5] * z);
  const float zr = static_cast<float>((R[6] * x + R[7] * y) + R[8] * z);
  SV o;
  o.sx = (xr + 150.0f) * kx;
  o.sy = (150.0f - yr) * ky;
  o.zb = (500.0f - zr) * (1.0f / 1500.0f);
  return o;
}

// (view, vertex) -> window position + z-buffer value, computed once (sv[view * nv + vertex])
__global__ void __launch_bounds__(256) raster_xform_kernel(const float* __restrict__ verts, int nv,
                                                           const double* __restrict__ rot, int H, int W,
                                                           float4* __restrict__ sv) {
  __shared__ double R[9];
  const int view = blockIdx.y;
  if (threadIdx.x < 9) R[threadIdx.x] = rot[view * 9 + threadIdx.x];
  __syncthreads();
  const float kx = static_cast<float>(static_cast<double>(W) / 300.0);
  const float ky = static_cast<float>(static_cast<double>(H) / 300.0);
  float4* out = sv + static_cast<size_t>(view) * nv;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < nv; i += gridDim.x * blockDim.x) {
    const SV a = xform_vertex(verts + 3 * i, R, kx, ky);
    out[i] = make_float4(a.sx, a.sy, a.zb, 0.f);
  }
}

__device__ __forceinline__ SV load_sv(const float4* __restrict__ sv, int i) {
  const float4 q = __ldg(sv + i);
  SV o;
  o.sx = q.x; o.sy = q.y; o.zb = q.z;
  return o;
}

__global__ void __launch_bounds__(256) raster_tris_kernel(const float4* __restrict__ sv, int nv,
                                                          const int* __restrict__ tris, int nt, int H, int W,
                                                          unsigned long long* __restrict__ zbuf) {
  const int view = blockIdx.y;
  const float4* svv = sv + static_cast<size_t>(view) * nv;
  unsigned long long* zb = zbuf + static_cast<size_t>(view) * H * W;
  for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < nt; t += gridDim.x * blockDim.x) {
    const int i0 = __ldg(tris + 3 * t), i1 = __ldg(tris + 3 * t + 1), i2 = __ldg(tris + 3 * t + 2);
    const SV a = load_sv(svv, i0), b = load_sv(svv, i1), c = load_sv(svv, i2);
    const float minx = fminf(a.sx, fminf(b.sx, c.sx)), maxx = fmaxf(a.sx, fmaxf(b.sx, c.sx));
    const float miny = fminf(a.sy, fminf(b.sy, c.sy)), maxy = fmaxf(a.sy, fmaxf(b.sy, c.sy));
    int x0 = static_cast<int>(ceilf(minx - 0.5f)), x1 = static_cast<int>(floorf(maxx - 0.5f));
    int y0 = static_cast<int>(ceilf(miny - 0.5f)), y1 = static_cast<int>(floorf(maxy - 0.5f));
    x0 = max(x0, 0); y0 = max(y0, 0); x1 = min(x1, W - 1); y1 = min(y1, H - 1);
    if (x0 > x1 || y0 > y1) continue;  // no pixel centre inside the bounding box: most micro-polygons end here
    const float area = edge_fn(a.sx, a.sy, b.sx, b.sy, c.sx, c.sy);
    if (area == 0.0f || area != area) continue;
    for (int py = y0; py <= y1; ++py) {
      for (int px = x0; px <= x1; ++px) {
        const float cx = static_cast<float>(px) + 0.5f, cy = static_cast<float>(py) + 0.5f;
        const float w0 = edge_fn(b.sx, b.sy, c.sx, c.sy, cx, cy);
        const float w1 = edge_fn(c.sx, c.sy, a.sx, a.sy, cx, cy);
        const float w2 = edge_fn(a.sx, a.sy, b.sx, b.sy, cx, cy);
        const bool inside = area > 0.0f ? (w0 >= 0.0f && w1 >= 0.0f && w2 >= 0.0f)
                                        : (w0 <= 0.0f && w1 <= 0.0f && w2 <= 0.0f);
        if (!inside) continue;
        const float l0 = w0 / area, l1 = w1 / area, l2 = w2 / area;
        const float z = (l0 * a.zb + l1 * b.zb) + l2 * c.zb;
        if (!(z >= 0.0f && z <= 1.0f)) continue;
        const unsigned long long key =
            (static_cast<unsigned long long>(__float_as_uint(z)) << 32) | static_cast<unsigned int>(t);
        atomicMin(zb + static_cast<size_t>(py) * W + px, key);
      }
    }
  }
}

// views view0 .. view0 + n_chunk - 1 (sv holds the transformed vertices of exactly these views)
__global__ void __launch_bounds__(256) raster_resolve_kernel(RasterArgs g, const float4* __restrict__ sv, int view0,
                                                             int n_chunk) {
  // grid = (pixel blocks of one view, views of the chunk): 32-bit index arithmetic (three 64-bit divisions per pixel
  // were a large part of this kernel's 275 instructions per pixel)
  const unsigned int vp = blockIdx.x * blockDim.x + threadIdx.x;  // pixel within the view
  if (vp >= static_cast<unsigned int>(g.h * g.w) || static_cast<int>(blockIdx.y) >= n_chunk) return;
  const int view = view0 + static_cast<int>(blockIdx.y);
  const int py = static_cast<int>(vp / static_cast<unsigned int>(g.w));
  const int px = static_cast<int>(vp - static_cast<unsigned int>(py) * static_cast<unsigned int>(g.w));
  const size_t pix = static_cast<size_t>(view) * g.h * g.w + vp;
  const unsigned long long key = g.zbuf[pix];
  float r = 1.0f, gr = 1.0f, bl = 1.0f, zval = 1.0f, geo = 1.0f;
  unsigned char r8 = 255, g8 = 255, b8 = 255, geo8 = 255;
  int tid = -1;
  const int mode = g.channel_mode;
  if (key != kBgKey) {
    tid = static_cast<int>(key & 0xFFFFFFFFu);
    zval = __uint_as_float(static_cast<unsigned int>(key >> 32));
    const int i0 = __ldg(g.tris + 3 * tid), i1 = __ldg(g.tris + 3 * tid + 1), i2 = __ldg(g.tris + 3 * tid + 2);
    const double* R = g.rot + view * 9;
    if (g.tex && g.uvs && (mode == 0 || mode == 2)) {
      const float4* svv = sv + static_cast<size_t>(view - view0) * g.nv;
      const SV a = load_sv(svv, i0), b = load_sv(svv, i1), c = load_sv(svv, i2);
      const float cx = static_cast<float>(px) + 0.5f, cy = static_cast<float>(py) + 0.5f;
      const float area = edge_fn(a.sx, a.sy, b.sx, b.sy, c.sx, c.sy);
      const float l0 = edge_fn(b.sx, b.sy, c.sx, c.sy, cx, cy) / area;
      const float l1 = edge_fn(c.sx, c.sy, a.sx, a.sy, cx, cy) / area;
      const float l2 = edge_fn(a.sx, a.sy, b.sx, b.sy, cx, cy) / area;
      const float u = (l0 * g.uvs[2 * i0] + l1 * g.uvs[2 * i1]) + l2 * g.uvs[2 * i2];
      const float v = (l0 * g.uvs[2 * i0 + 1] + l1 * g.uvs[2 * i1 + 1]) + l2 * g.uvs[2 * i2 + 1];
      int tx = static_cast<int>(floorf(u * static_cast<float>(g.tw)));
      int ty = static_cast<int>(floorf(v * static_cast<float>(g.th)));
      // repeat wrap; texture coordinates inside [0, 1) (the usual case) skip the two integer divisions
      if (tx < 0 || tx >= g.tw) { tx %= g.tw; if (tx < 0) tx += g.tw; }
      if (ty < 0 || ty >= g.th) { ty %= g.th; if (ty < 0) ty += g.th; }
      const size_t ti = static_cast<size_t>(g.th - 1 - ty) * g.tw + tx;
      if (g.tex_c == 4) {  // RGBA texture: one 4-byte load per texel
        const uchar4 q4 = __ldg(reinterpret_cast<const uchar4*>(g.tex) + ti);
        r8 = q4.x; g8 = q4.y; b8 = q4.z;
      } else {
        const unsigned char* texel = g.tex + ti * 3;
        r8 = texel[0]; g8 = texel[1]; b8 = texel[2];
      }

    }
    if (mode == 1 || mode == 4) {
      double p[3][3];
      const int idx[3] = {i0, i1, i2};
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        const float* v3 = g.verts + 3 * idx[k];
#pragma unroll
        for (int rr = 0; rr < 3; ++rr)
          p[k][rr] = (R[3 * rr] * v3[0] + R[3 * rr + 1] * v3[1]) + R[3 * rr + 2] * v3[2];
      }
      const double e1x = p[1][0] - p[0][0], e1y = p[1][1] - p[0][1], e1z = p[1][2] - p[0][2];
      const double e2x = p[2][0] - p[0][0], e2y = p[2][1] - p[0][1], e2z = p[2][2] - p[0][2];
      const double nx = e1y * e2z - e1z * e2y, ny = e1z * e2x - e1x * e2z, nz = e1x * e2y - e1y * e2x;
      const double nn = sqrt((nx * nx + ny * ny) + nz * nz);
      const double s = nn > 0.0 ? fabs(nz) / nn : 0.0;
      geo8 = static_cast<unsigned char>(static_cast<int>(s * 255.0 + 0.5));

    }
  }
  const int di = static_cast<int>(-255.0f * zval);
  const unsigned char d8 = static_cast<unsigned char>(di & 0xFF);
  // the fp32 stack (reference layout, /255 as render3d.py:191) is only computed when it is asked for: four IEEE
  // divisions per pixel that the fused path (u8 image for the CNN stem) never needs
  float depth = 0.f;
  if (g.out_f32) {
    r = static_cast<float>(r8) / 255.0f; gr = static_cast<float>(g8) / 255.0f; bl = static_cast<float>(b8) / 255.0f;
    geo = static_cast<float>(geo8) / 255.0f;
    depth = static_cast<float>(d8) / 255.0f;
  }
  float o[4] = {0.f, 0.f, 0.f, 0.f};
  uchar4 q = make_uchar4(0, 0, 0, 0);
  int C = 4;
  switch (mode) {
    case 0: o[0] = r; o[1] = gr; o[2] = bl; o[3] = depth; q = make_uchar4(r8, g8, b8, d8); C = 4; break;
    case 1: o[0] = geo; o[1] = depth; q = make_uchar4(geo8, d8, 0, 0); C = 2; break;
    case 2: o[0] = r; o[1] = gr; o[2] = bl; q = make_uchar4(r8, g8, b8, 0); C = 3; break;
    case 3: o[0] = depth; q = make_uchar4(d8, 0, 0, 0); C = 1; break;
    default: o[0] = geo; q = make_uchar4(geo8, 0, 0, 0); C = 1; break;
  }
  if (g.out_u8) reinterpret_cast<uchar4*>(g.out_u8)[pix] = q;
  if (g.out_f32) {
    float* dst = g.out_f32 + pix * C;
    for (int k = 0; k < C; ++k) dst[k] = o[k];
  }
  if (g.out_tri) g.out_tri[pix] = tid;
  if (g.out_z) g.out_z[pix] = zval;
}

}  // namespace

int raster_channels(int mode) {
  switch (mode) {
    case 0: return 4;
    case 1: return 2;
    case 2: return 3;
    case 3: return 1;
    case 4: return 1;
  }
  return -1;
}

// transformed vertices are kept for at most this many bytes at a time (views are processed in chunks beyond it):
// 100 views of a 50k-vertex scan are 80 MB (one chunk), config 4 (1M vertices) runs 16 views per chunk
constexpr size_t kSvBudgetBytes = 256u << 20;

static int sv_chunk_views(int n_views, int nv) {
  const size_t per_view = static_cast<size_t>(nv) * sizeof(float4);
  const size_t fit = per_view ? kSvBudgetBytes / per_view : static_cast<size_t>(n_views);
  return static_cast<int>(std::max<size_t>(1, std::min<size_t>(static_cast<size_t>(n_views), fit)));
}

size_t raster_workspace_bytes(int n_views, int h, int w, int n_verts) {
  const size_t z = (static_cast<size_t>(n_views) * h * w * sizeof(unsigned long long) + 255) & ~static_cast<size_t>(255);
  return z + static_cast<size_t>(sv_chunk_views(n_views, n_verts)) * n_verts * sizeof(float4);
}

int raster_launch(const RasterArgs& g, cudaStream_t stream) {
  MVLM_REQUIRE(g.verts && g.tris && g.rot && g.zbuf, "raster: null pointer");
  MVLM_REQUIRE(g.nt > 0 && g.nv > 0 && g.n_views > 0 && g.h > 0 && g.w > 0, "raster: bad sizes");
  MVLM_REQUIRE(raster_channels(g.channel_mode) > 0, "raster: unknown channel_mode %d", g.channel_mode);
  MVLM_REQUIRE(!g.tex || g.tex_c == 3 || g.tex_c == 4, "raster: texture must have 3 or 4 channels (got %d)", g.tex_c);
  MVLM_REQUIRE(g.workspace_bytes >= raster_workspace_bytes(g.n_views, g.h, g.w, g.nv),
               "raster: workspace too small (%zu bytes given, %zu needed)", g.workspace_bytes,
               raster_workspace_bytes(g.n_views, g.h, g.w, g.nv));
  const size_t npix = static_cast<size_t>(g.n_views) * g.h * g.w;
  const size_t zbytes = (npix * sizeof(unsigned long long) + 255) & ~static_cast<size_t>(255);
  float4* sv = reinterpret_cast<float4*>(reinterpret_cast<unsigned char*>(g.zbuf) + zbytes);
  MVLM_CHECK_CUDA(cudaMemsetAsync(g.zbuf, 0xFF, npix * sizeof(unsigned long long), stream));
  const int chunk = sv_chunk_views(g.n_views, g.nv);
  const size_t view_pix = static_cast<size_t>(g.h) * g.w;
  for (int v0 = 0; v0 < g.n_views; v0 += chunk) {
    const int nc = std::min(chunk, g.n_views - v0);
    dim3 gx(std::min(ceil_div(g.nv, 256), 1024), nc);
    raster_xform_kernel<<<gx, 256, 0, stream>>>(g.verts, g.nv, g.rot + static_cast<size_t>(v0) * 9, g.h, g.w, sv);
    dim3 gt(std::min(ceil_div(g.nt, 256), 1024), nc);
    raster_tris_kernel<<<gt, 256, 0, stream>>>(sv, g.nv, g.tris, g.nt, g.h, g.w, g.zbuf + static_cast<size_t>(v0) * view_pix);
    raster_resolve_kernel<<<dim3(static_cast<unsigned>((view_pix + 255) / 256), nc), 256, 0, stream>>>(g, sv, v0, nc);
    count_launch(3);
  }
  MVLM_CHECK_CUDA(cudaGetLastError());
  return MVLM_OK;
}

}  // namespace mvlm
