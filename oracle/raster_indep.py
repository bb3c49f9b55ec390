"""An independently formulated rasteriser (test infrastructure, like everything under oracle/).

The C oracle (oracle/csrc/oracle_native.c) and the CUDA rasteriser (mvlm_b200/csrc/raster.cu) are two
transcriptions of ONE frozen specification (edge functions in fp32, inclusive edges, barycentric depth).
Agreement between them proves the transcription, not the specification.  This file restates the
reference's render step (src/mvlm/utils/render3d.py:53-59 camera, :136-152 per-view transform,
:166-177 depth read-back and row flip) with a DIFFERENT algorithm and arithmetic:

  * fp64 throughout (vertices are rotated and projected in double, never rounded to float32);
  * scan-line conversion: for every image row the span [x_left, x_right) of a triangle is found by
    intersecting the row's centre line with the triangle's edges (no edge functions, no barycentric
    coordinates), half-open on the right / bottom like the OpenGL top-left fill rule;
  * depth from the triangle's plane equation z(x, y), nearest wins, ties to the lower triangle id.

The two formulations may only disagree where a pixel centre lies on (or within rounding of) a triangle
edge; north_star words the bar exactly so: tri-ID agreement >= 99.9 %, differences only on edges.
"""
from __future__ import annotations

import numpy as np


def project(verts: np.ndarray, rot: np.ndarray, h: int, w: int):
    """Window coordinates (x right, y DOWN, in pixels) and z-buffer value of every vertex for one view.
    Camera: orthographic, ParallelScale 150 (window +-150 mm), at z = +500 looking down -z, clipping
    range giving z_buf = (500 - z_cam) / 1500 (render3d.py:53-59); vertices <- R p (:140-145)."""
    p = verts.astype(np.float64) @ rot.astype(np.float64).T
    sx = (p[:, 0] + 150.0) * (w / 300.0)
    sy = (150.0 - p[:, 1]) * (h / 300.0)
    zb = (500.0 - p[:, 2]) / 1500.0
    return sx, sy, zb


def raster_view(verts: np.ndarray, tris: np.ndarray, rot: np.ndarray, h: int, w: int):
    """-> (tri_id (H,W) int32 with -1 = background, z (H,W) float64 with 1.0 = background)."""
    sx, sy, zb = project(verts, rot, h, w)
    tri_id = np.full((h, w), -1, np.int32)
    zbuf = np.full((h, w), np.inf)
    for t, (i0, i1, i2) in enumerate(tris):
        xs = np.array([sx[i0], sx[i1], sx[i2]])
        ys = np.array([sy[i0], sy[i1], sy[i2]])
        zs = np.array([zb[i0], zb[i1], zb[i2]])
        # plane z = z0 + gx (x - x0) + gy (y - y0)
        d = (xs[1] - xs[0]) * (ys[2] - ys[0]) - (xs[2] - xs[0]) * (ys[1] - ys[0])
        if d == 0.0 or not np.isfinite(d):
            continue
        gx = ((zs[1] - zs[0]) * (ys[2] - ys[0]) - (zs[2] - zs[0]) * (ys[1] - ys[0])) / d
        gy = ((xs[1] - xs[0]) * (zs[2] - zs[0]) - (xs[2] - xs[0]) * (zs[1] - zs[0])) / d
        r0 = max(int(np.ceil(ys.min() - 0.5)), 0)
        r1 = min(int(np.ceil(ys.max() - 0.5)) - 1, h - 1)  # half-open at the bottom
        for r in range(r0, r1 + 1):
            yc = r + 0.5
            # crossings of the row's centre line with the three edges
            cross = []
            for a, b in ((0, 1), (1, 2), (2, 0)):
                ya, yb = ys[a], ys[b]
                if ya == yb:
                    continue
                lo, hi = (ya, yb) if ya < yb else (yb, ya)
                if lo <= yc < hi:  # half-open: a vertex on the line belongs to one edge only
                    cross.append(xs[a] + (yc - ya) * (xs[b] - xs[a]) / (yb - ya))
            if len(cross) != 2:
                continue
            xl, xr = min(cross), max(cross)
            c0 = max(int(np.ceil(xl - 0.5)), 0)
            c1 = min(int(np.ceil(xr - 0.5)) - 1, w - 1)  # half-open at the right
            if c1 < c0:
                continue
            cols = np.arange(c0, c1 + 1)
            z = zs[0] + gx * (cols + 0.5 - xs[0]) + gy * (yc - ys[0])
            ok = (z >= 0.0) & (z <= 1.0) & (z < zbuf[r, cols])  # strict: ties keep the lower id (visited first)
            tri_id[r, cols[ok]] = t
            zbuf[r, cols[ok]] = z[ok]
    z_out = np.where(tri_id >= 0, zbuf, 1.0)
    return tri_id, z_out


def edge_distance_px(verts, tris, rot, h, w, rows, cols, tri_ids):
    """Distance (pixels, window space) from the centres of pixels (rows, cols) to the nearest EDGE of the
    triangles tri_ids (one per pixel)."""
    sx, sy, _ = project(verts, rot, h, w)
    px = np.asarray(cols, np.float64) + 0.5
    py = np.asarray(rows, np.float64) + 0.5
    t = tris[np.asarray(tri_ids)]
    best = np.full(len(px), np.inf)
    for a, b in ((0, 1), (1, 2), (2, 0)):
        ax, ay, bx, by = sx[t[:, a]], sy[t[:, a]], sx[t[:, b]], sy[t[:, b]]
        ex, ey = bx - ax, by - ay
        ll = ex * ex + ey * ey
        s = np.where(ll > 0, ((px - ax) * ex + (py - ay) * ey) / np.where(ll > 0, ll, 1.0), 0.0)
        s = np.clip(s, 0.0, 1.0)
        best = np.minimum(best, np.hypot(px - (ax + s * ex), py - (ay + s * ey)))
    return best


def compare_tri_maps(verts, tris, rot, got, ref):
    """(agreement fraction, largest distance in pixels from a differing pixel's centre to the nearest edge of
    the two candidate triangles).  `got`, `ref`: (H,W) triangle-id maps of one view."""
    h, w = ref.shape
    diff = np.argwhere(got != ref)
    if len(diff) == 0:
        return 1.0, 0.0
    worst = 0.0
    rows, cols = diff[:, 0], diff[:, 1]
    d = np.full(len(diff), np.inf)
    for ids in (got[rows, cols], ref[rows, cols]):
        m = ids >= 0
        if m.any():
            d[m] = np.minimum(d[m], edge_distance_px(verts, tris, rot, h, w, rows[m], cols[m], ids[m]))
    worst = float(d.max())
    return 1.0 - len(diff) / float(h * w), worst


def line_triangle_distance(p0, p1, a, b, c):
    """Distance between the infinite lines through (p0[i], p1[i]) and the triangles (a[i], b[i], c[i]) in 3D, fp64.
    0 when the line pierces the triangle, else the smallest line-to-edge-segment distance."""
    p0, p1, a, b, c = (np.asarray(x, np.float64) for x in (p0, p1, a, b, c))
    d = p1 - p0
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    # pierce test (Moeller-Trumbore without the t >= 0 condition: the line is infinite)
    e1, e2 = b - a, c - a
    pv = np.cross(d, e2)
    det = np.einsum("ij,ij->i", e1, pv)
    ok = np.abs(det) > 1e-300
    inv = np.where(ok, 1.0 / np.where(ok, det, 1.0), 0.0)
    tv = p0 - a
    u = np.einsum("ij,ij->i", tv, pv) * inv
    qv = np.cross(tv, e1)
    v = np.einsum("ij,ij->i", d, qv) * inv
    inside = ok & (u >= 0) & (v >= 0) & (u + v <= 1)

    def seg(q0, q1):
        # closest distance between the line (p0, d) and the segment q0..q1
        e = q1 - q0
        w0 = p0 - q0
        aa = np.ones(len(d))
        bb = np.einsum("ij,ij->i", d, e)
        cc = np.einsum("ij,ij->i", e, e)
        dd = np.einsum("ij,ij->i", d, w0)
        ee = np.einsum("ij,ij->i", e, w0)
        den = aa * cc - bb * bb
        s = np.where(den > 1e-300, (aa * ee - bb * dd) / np.where(den > 1e-300, den, 1.0), 0.0)
        s = np.clip(s, 0.0, 1.0)
        q = q0 + s[:, None] * e
        wq = q - p0
        perp = wq - np.einsum("ij,ij->i", wq, d)[:, None] * d
        return np.linalg.norm(perp, axis=1)

    dist = np.minimum(np.minimum(seg(a, b), seg(b, c)), seg(c, a))
    return np.where(inside, 0.0, dist)
