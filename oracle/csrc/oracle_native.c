/* CPU oracle (TEST INFRASTRUCTURE ONLY, see oracle/__init__.py): plain-C restatement of the
 * two reference stages that run inside VTK and therefore cannot be executed here.
 *
 *  oracle_raster_multiview : ObjVTKRenderer3D.render_3d_multi_rgb_geometry_depth
 *                            (src/mvlm/utils/render3d.py:114-177; camera :53-59,:136,:150-152;
 *                             depth encoder :73-77,:166-170; row flip :177; /255 :191) and
 *                            obj_to_actor's material (src/mvlm/utils/utils3d.py:26-64:
 *                            nearest texture, ambient 1 / diffuse 0 = unlit).
 *  oracle_snap_to_mesh     : Estimator3D.project_landmarks_to_surface
 *                            (src/mvlm/utils/estimator3d.py:252-285): exact closest point on
 *                            the triangle mesh (vtkCellLocator is only an accelerator).
 *
 * PARITY UNPINNED against VTK itself (not installable offline); the frozen rules are in
 * DESIGN.md "Renderer rules".  Compile with -ffp-contract=off: the CUDA rasteriser uses
 * explicitly rounded fp32 operations in the same order, so both are bit-identical.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define BG_KEY 0xFFFFFFFFFFFFFFFFull

static inline float edge_fn(float ax, float ay, float bx, float by, float cx, float cy) {
  /* fl(fl(fl(bx-ax)*fl(cy-ay)) - fl(fl(by-ay)*fl(cx-ax))) */
  float d1 = bx - ax, d2 = cy - ay, d3 = by - ay, d4 = cx - ax;
  float p = d1 * d2;
  float q = d3 * d4;
  return p - q;
}

/* view transform of one vertex: double rotation (vtkTransformPolyDataFilter works in double and
 * stores float points, render3d.py:140-145), then fp32 window mapping. */
static inline void xform_vertex(const float* v, const double* R, int W, int H, float* sx, float* sy,
                                float* zb, float* zc) {
  double x = v[0], y = v[1], z = v[2];
  float xr = (float)((R[0] * x + R[1] * y) + R[2] * z);
  float yr = (float)((R[3] * x + R[4] * y) + R[5] * z);
  float zr = (float)((R[6] * x + R[7] * y) + R[8] * z);
  const float kx = (float)((double)W / 300.0);
  const float ky = (float)((double)H / 300.0);
  *sx = (xr + 150.0f) * kx;             /* pixel units, column axis */
  *sy = (150.0f - yr) * ky;             /* pixel units, row axis, top-down */
  *zb = (500.0f - zr) * (1.0f / 1500.0f); /* linear ortho depth, near 0 / far 1500 */
  *zc = zr;
}

/* channel_mode: 0 = RGB+depth (4 ch), 1 = geometry+depth (2 ch, extension), 2 = RGB (3 ch),
 *               3 = depth (1 ch), 4 = geometry (1 ch, extension) */
int oracle_channels(int mode) {
  switch (mode) { case 0: return 4; case 1: return 2; case 2: return 3; case 3: return 1; case 4: return 1; }
  return -1;
}

int oracle_raster_multiview(const float* verts, const float* uvs, int nv, const int32_t* tris, int nt,
                            const uint8_t* tex, int th, int tw, const double* rot /* V x 9 */, int nviews,
                            int H, int W, int channel_mode, float* out_img /* V,H,W,C */,
                            int32_t* out_tri /* V,H,W or NULL */, float* out_z /* V,H,W or NULL */) {
  const int C = oracle_channels(channel_mode);
  if (C < 0) return -1;
  (void)nv;
  uint64_t* zbuf = (uint64_t*)malloc(sizeof(uint64_t) * (size_t)H * W);
  float* sv = (float*)malloc(sizeof(float) * 4 * (size_t)nv);
  if (!zbuf || !sv) return -2;
  for (int view = 0; view < nviews; ++view) {
    const double* R = rot + 9 * view;
    for (size_t i = 0; i < (size_t)H * W; ++i) zbuf[i] = BG_KEY;
    for (int i = 0; i < nv; ++i) xform_vertex(verts + 3 * i, R, W, H, sv + 4 * i, sv + 4 * i + 1, sv + 4 * i + 2, sv + 4 * i + 3);
    for (int t = 0; t < nt; ++t) {
      const float* a = sv + 4 * tris[3 * t], *b = sv + 4 * tris[3 * t + 1], *c = sv + 4 * tris[3 * t + 2];
      float area = edge_fn(a[0], a[1], b[0], b[1], c[0], c[1]);
      if (area == 0.0f || area != area) continue;
      float minx = fminf(a[0], fminf(b[0], c[0])), maxx = fmaxf(a[0], fmaxf(b[0], c[0]));
      float miny = fminf(a[1], fminf(b[1], c[1])), maxy = fmaxf(a[1], fmaxf(b[1], c[1]));
      /* pixel centres i+0.5 inside [min,max] */
      int x0 = (int)ceilf(minx - 0.5f), x1 = (int)floorf(maxx - 0.5f);
      int y0 = (int)ceilf(miny - 0.5f), y1 = (int)floorf(maxy - 0.5f);
      if (x0 < 0) x0 = 0; if (y0 < 0) y0 = 0; if (x1 > W - 1) x1 = W - 1; if (y1 > H - 1) y1 = H - 1;
      for (int py = y0; py <= y1; ++py) {
        for (int px = x0; px <= x1; ++px) {
          float cx = (float)px + 0.5f, cy = (float)py + 0.5f;
          float w0 = edge_fn(b[0], b[1], c[0], c[1], cx, cy);
          float w1 = edge_fn(c[0], c[1], a[0], a[1], cx, cy);
          float w2 = edge_fn(a[0], a[1], b[0], b[1], cx, cy);
          int inside = area > 0.0f ? (w0 >= 0.0f && w1 >= 0.0f && w2 >= 0.0f) : (w0 <= 0.0f && w1 <= 0.0f && w2 <= 0.0f);
          if (!inside) continue;
          float l0 = w0 / area, l1 = w1 / area, l2 = w2 / area;
          float z = (l0 * a[2] + l1 * b[2]) + l2 * c[2];
          if (!(z >= 0.0f && z <= 1.0f)) continue; /* clipped by near/far */
          uint32_t zbits;
          memcpy(&zbits, &z, 4);
          uint64_t key = ((uint64_t)zbits << 32) | (uint32_t)t;
          if (key < zbuf[(size_t)py * W + px]) zbuf[(size_t)py * W + px] = key;
        }
      }
    }
    /* resolve */
    for (int py = 0; py < H; ++py) {
      for (int px = 0; px < W; ++px) {
        size_t pix = ((size_t)view * H + py) * W + px;
        uint64_t key = zbuf[(size_t)py * W + px];
        float r = 1.0f, g = 1.0f, bl = 1.0f, zval = 1.0f, geo = 1.0f;
        int32_t tid = -1;
        if (key != BG_KEY) {
          tid = (int32_t)(key & 0xFFFFFFFFu);
          uint32_t zbits = (uint32_t)(key >> 32);
          memcpy(&zval, &zbits, 4);
          const int i0 = tris[3 * tid], i1 = tris[3 * tid + 1], i2 = tris[3 * tid + 2];
          const float* a = sv + 4 * i0, *b = sv + 4 * i1, *c = sv + 4 * i2;
          if (tex && uvs && (channel_mode == 0 || channel_mode == 2)) {
            float cx = (float)px + 0.5f, cy = (float)py + 0.5f;
            float area = edge_fn(a[0], a[1], b[0], b[1], c[0], c[1]);
            float l0 = edge_fn(b[0], b[1], c[0], c[1], cx, cy) / area;
            float l1 = edge_fn(c[0], c[1], a[0], a[1], cx, cy) / area;
            float l2 = edge_fn(a[0], a[1], b[0], b[1], cx, cy) / area;
            float u = (l0 * uvs[2 * i0] + l1 * uvs[2 * i1]) + l2 * uvs[2 * i2];
            float v = (l0 * uvs[2 * i0 + 1] + l1 * uvs[2 * i1 + 1]) + l2 * uvs[2 * i2 + 1];
            int tx = (int)floorf(u * (float)tw), ty = (int)floorf(v * (float)th);
            tx %= tw; if (tx < 0) tx += tw;
            ty %= th; if (ty < 0) ty += th;
            const uint8_t* texel = tex + ((size_t)(th - 1 - ty) * tw + tx) * 3; /* v=0 is the bottom image row */
            r = (float)texel[0] / 255.0f; g = (float)texel[1] / 255.0f; bl = (float)texel[2] / 255.0f;
          }
          if (channel_mode == 1 || channel_mode == 4) {
            /* extension: unlit-white replaced by head-light Lambert |n_z| of the rotated face normal */
            const double* R = rot + 9 * view;
            double p[3][3];
            const int idx[3] = {i0, i1, i2};
            for (int k = 0; k < 3; ++k) {
              const float* v3 = verts + 3 * idx[k];
              for (int rr = 0; rr < 3; ++rr) p[k][rr] = (R[3 * rr] * v3[0] + R[3 * rr + 1] * v3[1]) + R[3 * rr + 2] * v3[2];
            }
            double e1[3] = {p[1][0] - p[0][0], p[1][1] - p[0][1], p[1][2] - p[0][2]};
            double e2[3] = {p[2][0] - p[0][0], p[2][1] - p[0][1], p[2][2] - p[0][2]};
            double nx = e1[1] * e2[2] - e1[2] * e2[1], ny = e1[2] * e2[0] - e1[0] * e2[2], nz = e1[0] * e2[1] - e1[1] * e2[0];
            double nn = sqrt((nx * nx + ny * ny) + nz * nz);
            double s = nn > 0.0 ? fabs(nz) / nn : 0.0;
            int gi = (int)(s * 255.0 + 0.5);
            geo = (float)gi / 255.0f;
          }
        }
        /* depth byte: (unsigned char)(int)(-255*z), wraps mod 256 (vtkImageShiftScale, clamp off) */
        int di = (int)(-255.0f * zval);
        float depth = (float)(uint8_t)(di & 0xFF) / 255.0f;
        float* o = out_img + pix * C;
        switch (channel_mode) {
          case 0: o[0] = r; o[1] = g; o[2] = bl; o[3] = depth; break;
          case 1: o[0] = geo; o[1] = depth; break;
          case 2: o[0] = r; o[1] = g; o[2] = bl; break;
          case 3: o[0] = depth; break;
          case 4: o[0] = geo; break;
        }
        if (out_tri) out_tri[pix] = tid;
        if (out_z) out_z[pix] = zval;
      }
    }
  }
  free(zbuf);
  free(sv);
  return 0;
}

/* Closest point on triangle (a,b,c) to p, all double (Ericson, Real-Time Collision Detection 5.1.5). */
static void closest_on_tri(const double* p, const double* a, const double* b, const double* c, double* out) {
  double ab[3], ac[3], ap[3], bp[3], cp[3];
  for (int i = 0; i < 3; ++i) { ab[i] = b[i] - a[i]; ac[i] = c[i] - a[i]; ap[i] = p[i] - a[i]; }
  double d1 = ab[0] * ap[0] + ab[1] * ap[1] + ab[2] * ap[2];
  double d2 = ac[0] * ap[0] + ac[1] * ap[1] + ac[2] * ap[2];
  if (d1 <= 0.0 && d2 <= 0.0) { memcpy(out, a, 24); return; }
  for (int i = 0; i < 3; ++i) bp[i] = p[i] - b[i];
  double d3 = ab[0] * bp[0] + ab[1] * bp[1] + ab[2] * bp[2];
  double d4 = ac[0] * bp[0] + ac[1] * bp[1] + ac[2] * bp[2];
  if (d3 >= 0.0 && d4 <= d3) { memcpy(out, b, 24); return; }
  double vc = d1 * d4 - d3 * d2;
  if (vc <= 0.0 && d1 >= 0.0 && d3 <= 0.0) {
    double v = d1 / (d1 - d3);
    for (int i = 0; i < 3; ++i) out[i] = a[i] + v * ab[i];
    return;
  }
  for (int i = 0; i < 3; ++i) cp[i] = p[i] - c[i];
  double d5 = ab[0] * cp[0] + ab[1] * cp[1] + ab[2] * cp[2];
  double d6 = ac[0] * cp[0] + ac[1] * cp[1] + ac[2] * cp[2];
  if (d6 >= 0.0 && d5 <= d6) { memcpy(out, c, 24); return; }
  double vb = d5 * d2 - d1 * d6;
  if (vb <= 0.0 && d2 >= 0.0 && d6 <= 0.0) {
    double w = d2 / (d2 - d6);
    for (int i = 0; i < 3; ++i) out[i] = a[i] + w * ac[i];
    return;
  }
  double va = d3 * d6 - d5 * d4;
  if (va <= 0.0 && (d4 - d3) >= 0.0 && (d5 - d6) >= 0.0) {
    double w = (d4 - d3) / ((d4 - d3) + (d5 - d6));
    for (int i = 0; i < 3; ++i) out[i] = b[i] + w * (c[i] - b[i]);
    return;
  }
  double denom = 1.0 / (va + vb + vc);
  double v = vb * denom, w = vc * denom;
  for (int i = 0; i < 3; ++i) out[i] = a[i] + ab[i] * v + ac[i] * w;
}

int oracle_snap_to_mesh(const float* verts, const int32_t* tris, int nt, const double* lm, int nl, double* out,
                        int32_t* out_tri /* or NULL */) {
  for (int l = 0; l < nl; ++l) {
    const double* p = lm + 3 * l;
    double best = INFINITY, bp[3] = {p[0], p[1], p[2]};
    int32_t bt = -1;
    for (int t = 0; t < nt; ++t) {
      double a[3], b[3], c[3], q[3];
      for (int i = 0; i < 3; ++i) {
        a[i] = verts[3 * tris[3 * t] + i];
        b[i] = verts[3 * tris[3 * t + 1] + i];
        c[i] = verts[3 * tris[3 * t + 2] + i];
      }
      closest_on_tri(p, a, b, c, q);
      double dx = q[0] - p[0], dy = q[1] - p[1], dz = q[2] - p[2];
      double d = dx * dx + dy * dy + dz * dz;
      if (d < best) { best = d; bt = t; bp[0] = q[0]; bp[1] = q[1]; bp[2] = q[2]; }
    }
    out[3 * l] = bp[0]; out[3 * l + 1] = bp[1]; out[3 * l + 2] = bp[2];
    if (out_tri) out_tri[l] = bt;
  }
  return 0;
}
