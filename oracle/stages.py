"""numpy restatements of the reference's numpy stages (oracle; see oracle/__init__.py).

Pinned against the reference code executed verbatim: tests/golden/stages_*.npz made by
tools/make_golden.py (and re-checked live by tests/test_oracle_vs_reference.py whenever
/root/reference is present).
"""
from __future__ import annotations

import numpy as np


# ---------------------------------------------------------------------------------------
# R5  find_heat_map_maxima / find_maxima_in_batch_of_heatmaps
#     src/mvlm/prediction/paulsenpredictor.py:112-158, :160-165
# ---------------------------------------------------------------------------------------
def heatmap_peaks(heatmaps: np.ndarray, selection_method: str = "simple") -> np.ndarray:
    """heatmaps (V,L,H,W) float32 -> (L,V,3) float32 = (row-1, col-0.5, value) (:127)."""
    v, l, h, w = heatmaps.shape
    out = np.zeros((l, v, 3), dtype=np.float32)
    flat = heatmaps.reshape(v, l, h * w)
    idx = np.argmax(flat, axis=2)  # first maximum in row-major order; first NaN wins (:123)
    rows, cols = idx // w, idx % w
    vals = np.take_along_axis(flat, idx[..., None], axis=2)[..., 0]
    if selection_method == "simple":
        out[:, :, 0] = (rows - 1).T
        out[:, :, 1] = (cols - 0.5).T
        out[:, :, 2] = vals.T
        return out
    if selection_method != "moment":
        # the reference silently returns zeros for unknown methods (:118,:129)
        return out
    sz = 15
    ar = np.arange(2 * sz + 1)
    for vi in range(v):
        for k in range(l):
            hm = heatmaps[vi, k]
            px, py = int(rows[vi, k]), int(cols[vi, k])
            fx, fy = float(px), float(py)
            # :141  (hm_size is heatmaps.shape[1] of the per-view (L,H,W) array = H; square maps)
            if px > sz and h - px > sz and py > sz and h - py > sz:
                slc = hm[px - sz:px + sz + 1, py - sz:py + sz + 1]
                sum_x = np.sum(slc, axis=1)
                fx = px + (np.sum(np.multiply(ar, sum_x)) / np.sum(sum_x) - sz)
                sum_y = np.sum(slc, axis=0)
                fy = py + (np.sum(np.multiply(ar, sum_y)) / np.sum(sum_y) - sz)
            out[k, vi] = (fx - 1, fy - 0.5, np.max(hm))
    return out


# ---------------------------------------------------------------------------------------
# R6  Estimator3D.estimate_landmark_lines   src/mvlm/utils/estimator3d.py:31-90 (+ :8-15)
# ---------------------------------------------------------------------------------------
def rotation_matrices(transform_stack: np.ndarray) -> np.ndarray:
    """(V,>=3) [rx, ry, rz] degrees -> (V,3,3) float64  R = Ry @ Rx @ Rz  (:57).

    The angle keeps the dtype of transform_stack exactly as in the reference (float32 for the
    fixed 8-view preset, float64 for random views), because np.deg2rad/np.cos/np.sin run in that
    dtype before the values are widened by np.array([...]) of mixed ints/floats.
    """
    out = np.empty((transform_stack.shape[0], 3, 3), dtype=np.float64)
    for i in range(transform_stack.shape[0]):
        rx, ry, rz = transform_stack[i, :3]
        ax, ay, az = np.deg2rad(rx), np.deg2rad(ry), np.deg2rad(rz)
        mx = np.array([[1, 0, 0], [0, np.cos(ax), -np.sin(ax)], [0, np.sin(ax), np.cos(ax)]])
        my = np.array([[np.cos(ay), 0, np.sin(ay)], [0, 1, 0], [-np.sin(ay), 0, np.cos(ay)]])
        mz = np.array([[np.cos(az), -np.sin(az), 0], [np.sin(az), np.cos(az), 0], [0, 0, 1]])
        out[i] = (my @ mx) @ mz
    return out


def landmark_lines(image_size: int, landmarks_stack: np.ndarray, transform_stack: np.ndarray):
    """landmarks (L,V,3) float32 [row, col, value] -> (starts, ends) (L,V,3) float64.

    The pixel->camera mapping is evaluated in the dtype of `landmarks_stack` (float32 in the
    pipeline), exactly like the scalar arithmetic at :61-79; the rotation is float64 (:83).
    """
    l, v = landmarks_stack.shape[:2]
    s = image_size  # img_size == hm_size == image_stack.shape[1]  (:43-44)
    rot = rotation_matrices(transform_stack)
    y = landmarks_stack[:, :, 0]
    x = landmarks_stack[:, :, 1]
    dt = landmarks_stack.dtype.type
    y = y / dt(s) * dt(s)
    x = x / dt(s) * dt(s)
    xc = (x / dt(s)) * dt(300) + dt(-150)
    yc = ((dt(s - 1) - y) / dt(s)) * dt(300) + dt(-150)
    cam_s = np.stack([xc, yc, np.full_like(xc, 500)], axis=-1).astype(np.float64)
    cam_e = np.stack([xc, yc, np.full_like(xc, -500)], axis=-1).astype(np.float64)
    # p_world = t.T @ p  (:83)  ->  p_world_j = sum_i R[i,j] * p_i
    starts = np.einsum("vij,lvi->lvj", rot, cam_s)
    ends = np.einsum("vij,lvi->lvj", rot, cam_e)
    return starts, ends


# ---------------------------------------------------------------------------------------
# R8  compute_intersection_between_lines   src/mvlm/utils/utils3d.py:99-124
# ---------------------------------------------------------------------------------------
def lsq_intersection(pa: np.ndarray, pb: np.ndarray) -> np.ndarray:
    si = pb - pa
    ni = si / np.sqrt(np.sum(si ** 2, axis=1))[:, None]
    nx, ny, nz = ni[:, 0], ni[:, 1], ni[:, 2]
    sxx, syy, szz = np.sum(nx ** 2 - 1), np.sum(ny ** 2 - 1), np.sum(nz ** 2 - 1)
    sxy, sxz, syz = np.sum(nx * ny), np.sum(nx * nz), np.sum(ny * nz)
    s = np.array([[sxx, sxy, sxz], [sxy, syy, syz], [sxz, syz, szz]])
    cx = np.sum(pa[:, 0] * (nx ** 2 - 1) + pa[:, 1] * (nx * ny) + pa[:, 2] * (nx * nz))
    cy = np.sum(pa[:, 0] * (nx * ny) + pa[:, 1] * (ny ** 2 - 1) + pa[:, 2] * (ny * nz))
    cz = np.sum(pa[:, 0] * (nx * nz) + pa[:, 1] * (ny * nz) + pa[:, 2] * (nz ** 2 - 1))
    c = np.array([[cx], [cy], [cz]])
    return np.matmul(np.linalg.pinv(s), c)[:, 0]


def point_line_sqdist(p: np.ndarray, pa: np.ndarray, pb: np.ndarray) -> np.ndarray:
    """(|(p-a) x (p-b)| / |b-a|)^2   (estimator3d.py:109-111)."""
    top = np.cross(p[None, :] - pa, p[None, :] - pb)
    bottom = pb - pa
    return (np.linalg.norm(top, axis=1) / np.linalg.norm(bottom, axis=1)) ** 2


# ---------------------------------------------------------------------------------------
# R9  compute_intersection_between_lines_ransac   src/mvlm/utils/estimator3d.py:92-137
#     generalised to an explicit hypothesis list (the reference draws exactly one,
#     :105; its `for i in range(iterations)` loop is commented out, :103).
# ---------------------------------------------------------------------------------------
def ransac_intersection(pa: np.ndarray, pb: np.ndarray, hypotheses: np.ndarray, dist_thres: float = 100.0):
    """hypotheses (H,8) integer indices into the n lines (already reduced mod n).
    Returns (point (3,), best_error).  H=1 with the reference's own draw == reference."""
    best_error = 100000000
    best_p = None
    n_lines = len(pa)
    d = n_lines / 3
    for h in range(hypotheses.shape[0]):
        ran = hypotheses[h]
        p_est = lsq_intersection(pa[ran], pb[ran])
        dist = point_line_sqdist(p_est, pa, pb)
        inl = dist < dist_thres
        n_in = int(np.sum(inl))
        if n_in > d:
            p_est = lsq_intersection(pa[inl], pb[inl])
            dist2 = point_line_sqdist(p_est, pa[inl], pb[inl])
            err = np.sum(dist2) / n_in
            if err < best_error:
                best_error = err
                best_p = p_est
    if best_p is None:
        best_p = lsq_intersection(pa, pb)
    return best_p, best_error


# ---------------------------------------------------------------------------------------
# R7 + R10  filters and estimate_landmarks_from_lines   estimator3d.py:140-183
# ---------------------------------------------------------------------------------------
def line_filter_mask(values: np.ndarray, mode: str, threshold_quantile: float, threshold_absolute: float):
    if mode == "absolute":
        return values > threshold_absolute
    if mode == "quantile":
        return values > np.quantile(values, threshold_quantile)
    raise ValueError(f"Unknown mode for line matching in Estimator: {mode}")


def hypotheses_for(raw_draws: np.ndarray, n_lines: int) -> np.ndarray:
    """Shared seeded hypothesis table: raw uint32 draws (H,8) mapped to line indices `r mod n`."""
    return (raw_draws.astype(np.uint64) % np.uint64(max(n_lines, 1))).astype(np.int64)


def landmarks_from_lines(landmark_stack, lines_s, lines_e, raw_draws, mode="quantile",
                         threshold_quantile=0.5, threshold_absolute=0.5):
    """raw_draws (L,H,8) uint32.  Returns ((L,3) float64, mean error, per-landmark error (L,))."""
    n_landmarks = lines_s.shape[0]
    landmarks = np.empty((n_landmarks, 3))
    errs = np.zeros(n_landmarks)
    for lm in range(n_landmarks):
        mask = line_filter_mask(landmark_stack[lm, :, 2], mode, threshold_quantile, threshold_absolute)
        pa, pb = lines_s[lm][mask], lines_e[lm][mask]
        if len(pa) < 3:
            landmarks[lm] = lsq_intersection(pa, pb) if len(pa) > 0 else 0.0
            # :176 len(pa)==0 -> pinv of zeros times zeros = 0
        else:
            p, e = ransac_intersection(pa, pb, hypotheses_for(raw_draws[lm], len(pa)))
            landmarks[lm] = p
            errs[lm] = e
    return landmarks, float(np.sum(errs) / n_landmarks), errs


# ---------------------------------------------------------------------------------------
# pre-align   configs/*.json:59-84, commented apply_pre_transformation  src/mvlm/utils/estimator3d.py:186-224
#             + transform_landmarks_to_original_space :226-248
#   t = vtkTransform(); t.Scale(s, s, s); t.RotateY(ry); t.RotateX(rx); t.RotateZ(rz); t.Translate(-cm)
#   vtkTransform is in PreMultiply mode by default: every call sets  M <- M @ op  (4x4, column vectors), so the
#   calls compose left to right and a point goes through the LAST call first.
# ---------------------------------------------------------------------------------------
def pre_align_matrix(verts: np.ndarray, cfg: dict) -> np.ndarray:
    """4x4 float64 homogeneous matrix of the legacy pre-alignment, built call by call like vtkTransform."""
    def rot(axis, deg):
        a = np.deg2rad(deg)
        c, s = np.cos(a), np.sin(a)
        m = np.eye(4)
        i, j = {"x": (1, 2), "y": (2, 0), "z": (0, 1)}[axis]
        m[i, i], m[i, j], m[j, i], m[j, j] = c, -s, s, c
        return m

    m = np.eye(4)                                      # t.Identity()
    sc = float(cfg.get("scale", 1))
    m = m @ np.diag([sc, sc, sc, 1.0])                 # t.Scale(s, s, s)
    m = m @ rot("y", float(cfg.get("rot_y", 0)))       # t.RotateY(ry)
    m = m @ rot("x", float(cfg.get("rot_x", 0)))       # t.RotateX(rx)
    m = m @ rot("z", float(cfg.get("rot_z", 0)))       # t.RotateZ(rz)
    tr = np.eye(4)
    if cfg.get("align_center_of_mass", False):         # vtkCenterOfMass, UseScalarsAsWeights(False): mean of the points
        tr[:3, 3] = -np.asarray(verts, np.float64).mean(axis=0)
    return m @ tr                                      # t.Translate(translation)
