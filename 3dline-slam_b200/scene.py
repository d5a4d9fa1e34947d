"""Synthetic scenes and the NVM reader for the Line3D++ matching path.

The reference never ships runnable inputs for this path (no images, testdata4/vsfm_result.nvm is
empty), so scenes are synthesised: world 3-D segments are projected into pinhole cameras, noised,
clipped, padded with clutter up to exactly N segments per view and sorted by length (the order
`Line3D::detectLineSegments` produces, reference src/line3D.cc:342-384).  Everything is numpy on
the host; nothing here touches the GPU or the oracle.
"""
from __future__ import annotations

import dataclasses
import math
from typing import Dict, List, Optional

import numpy as np

BASE_SEED = 20261018


@dataclasses.dataclass
class SceneView:
    cam_id: int
    K: np.ndarray          # (3,3) float64
    R: np.ndarray          # (3,3) float64   x_cam = R X + t
    t: np.ndarray          # (3,)  float64
    width: int
    height: int
    median_depth: float
    segs: np.ndarray       # (N,4) float32  x1 y1 x2 y2, sorted by length descending
    neighbors: List[int]   # explicit visual neighbours (camera ids)
    worldpoints: Optional[List[int]] = None  # world-point ids (NVM input)


@dataclasses.dataclass
class Scene:
    name: str
    views: List[SceneView]
    max_image_width: int
    params: Dict[str, float]
    neighbors_by_worldpoints: bool = False

    @property
    def num_views(self) -> int:
        return len(self.views)

    def total_segments(self) -> int:
        return int(sum(v.segs.shape[0] for v in self.views))


DEFAULT_PARAMS = dict(sigma_p=5.0, sigma_a=10.0, num_neighbors=10, epipolar_overlap=0.25, knn=10,
                      const_reg_depth=-1.0)  # reference include/L3DPPing.h:77-91


def look_at(center: np.ndarray, target: np.ndarray, up=(0.0, 0.0, 1.0)) -> np.ndarray:
    """World->camera rotation with +z forward, +x right, +y down."""
    f = target - center
    f = f / np.linalg.norm(f)
    upv = np.asarray(up, dtype=np.float64)
    r = np.cross(f, upv)
    if np.linalg.norm(r) < 1e-9:
        r = np.cross(f, np.array([0.0, 1.0, 0.0]))
    r = r / np.linalg.norm(r)
    d = np.cross(f, r)
    return np.stack([r, d, f], axis=0)


def rotation_from_q(qw, qx, qy, qz) -> np.ndarray:
    """Quaternion -> rotation, same convention as reference src/line3D.cc:3221-3245."""
    n = qw * qw + qx * qx + qy * qy + qz * qz
    s = 0.0 if abs(n) < 1e-12 else 2.0 / n
    wx, wy, wz = s * qw * qx, s * qw * qy, s * qw * qz
    xx, xy, xz = s * qx * qx, s * qx * qy, s * qx * qz
    yy, yz, zz = s * qy * qy, s * qy * qz, s * qz * qz
    return np.array([[1.0 - (yy + zz), xy - wz, xz + wy],
                     [xy + wz, 1.0 - (xx + zz), yz - wx],
                     [xz - wy, yz + wx, 1.0 - (xx + yy)]], dtype=np.float64)


def _clip_segment_to_image(p, q, w, h):
    """Liang-Barsky clip of a 2-D segment to [0,w]x[0,h]; returns None if outside."""
    d = q - p
    t0, t1 = 0.0, 1.0
    for pk, qk in ((-d[0], p[0]), (d[0], w - p[0]), (-d[1], p[1]), (d[1], h - p[1])):
        if abs(pk) < 1e-12:
            if qk < 0:
                return None
        else:
            r = qk / pk
            if pk < 0:
                if r > t1:
                    return None
                t0 = max(t0, r)
            else:
                if r < t0:
                    return None
                t1 = min(t1, r)
    if t1 - t0 <= 0:
        return None
    return p + t0 * d, p + t1 * d


def _project_world_segments(P1, P2, K, R, t, w, h, near=0.2, max_depth=None):
    """Project world segments, clip against the near plane and the image rectangle.  A vectorised, conservative
    pre-cull (both end points behind the near plane, or the bounding box of the projected end points outside the
    image) removes what the per-segment loop below would drop anyway, so the result is unchanged; max_depth
    (city-scale scenes only) additionally ignores segments farther away than that."""
    X1 = P1 @ R.T + t
    X2 = P2 @ R.T + t
    z1, z2 = X1[:, 2], X2[:, 2]
    cand = ~((z1 < near) & (z2 < near))
    both = (z1 >= near) & (z2 >= near)
    if both.any():
        fx, fy, cx, cy = K[0, 0], K[1, 1], K[0, 2], K[1, 2]
        skew = K[0, 1]
        with np.errstate(divide="ignore", invalid="ignore"):
            u1 = (fx * X1[:, 0] + skew * X1[:, 1]) / z1 + cx
            v1 = fy * X1[:, 1] / z1 + cy
            u2 = (fx * X2[:, 0] + skew * X2[:, 1]) / z2 + cx
            v2 = fy * X2[:, 1] / z2 + cy
        m = 1.0  # one pixel of slack: the loop below decides exactly
        outside = (np.maximum(u1, u2) < -m) | (np.minimum(u1, u2) > w + m) | (np.maximum(v1, v2) < -m) | \
                  (np.minimum(v1, v2) > h + m)
        cand &= ~(both & outside)
        if max_depth is not None:
            # city scale: a segment that projects shorter than 8 px cannot reach the 15 px minimum with 0.5 px noise
            cand &= ~(both & (np.hypot(u1 - u2, v1 - v2) < 8.0))
    if max_depth is not None:
        cand &= np.minimum(np.linalg.norm(X1, axis=1), np.linalg.norm(X2, axis=1)) < max_depth
    idx = np.nonzero(cand)[0]
    if max_depth is not None:
        return _project_clip_vectorised(X1[idx], X2[idx], K, w, h, near)
    out = []
    depths = []
    for a, b in zip(X1[idx], X2[idx]):
        if a[2] < near and b[2] < near:
            continue
        if a[2] < near:
            s = (near - a[2]) / (b[2] - a[2])
            a = a + s * (b - a)
        elif b[2] < near:
            s = (near - b[2]) / (a[2] - b[2])
            b = b + s * (a - b)
        pa = K @ a
        pb = K @ b
        pa = pa[:2] / pa[2]
        pb = pb[:2] / pb[2]
        c = _clip_segment_to_image(pa, pb, float(w), float(h))
        if c is None:
            continue
        out.append((c[0][0], c[0][1], c[1][0], c[1][1]))
        depths.append(0.5 * (np.linalg.norm(a) + np.linalg.norm(b)))
    return np.asarray(out, dtype=np.float64).reshape(-1, 4), np.asarray(depths)


def _project_clip_vectorised(A, B, K, w, h, near):
    """The loop of _project_world_segments in array form (city-scale scenes: thousands of visible segments per view):
    near-plane clip, pinhole projection, Liang-Barsky clip against the image rectangle."""
    A, B = A.copy(), B.copy()
    za, zb = A[:, 2], B[:, 2]
    keep = ~((za < near) & (zb < near))
    A, B, za, zb = A[keep], B[keep], za[keep], zb[keep]
    ca = za < near
    if ca.any():
        sa = ((near - za[ca]) / (zb[ca] - za[ca]))[:, None]
        A[ca] = A[ca] + sa * (B[ca] - A[ca])
    cb = B[:, 2] < near
    if cb.any():
        sb = ((near - B[cb, 2]) / (A[cb, 2] - B[cb, 2]))[:, None]
        B[cb] = B[cb] + sb * (A[cb] - B[cb])
    pa = A @ K.T
    pb = B @ K.T
    p = pa[:, :2] / pa[:, 2:3]
    q = pb[:, :2] / pb[:, 2:3]
    d = q - p
    t0 = np.zeros(len(p))
    t1 = np.ones(len(p))
    ok = np.ones(len(p), dtype=bool)
    for pk, qk in ((-d[:, 0], p[:, 0]), (d[:, 0], w - p[:, 0]), (-d[:, 1], p[:, 1]), (d[:, 1], h - p[:, 1])):
        par = np.abs(pk) < 1e-12
        ok &= ~(par & (qk < 0))
        with np.errstate(divide="ignore", invalid="ignore"):
            r = np.where(par, 0.0, qk / np.where(par, 1.0, pk))
        neg = (~par) & (pk < 0)
        pos = (~par) & (pk > 0)
        ok &= ~(neg & (r > t1))
        t0 = np.where(neg, np.maximum(t0, r), t0)
        ok &= ~(pos & (r < t0))
        t1 = np.where(pos, np.minimum(t1, r), t1)
    ok &= (t1 - t0) > 0
    c0 = p + t0[:, None] * d
    c1 = p + t1[:, None] * d
    out = np.concatenate([c0, c1], axis=1)[ok]
    depths = 0.5 * (np.linalg.norm(A, axis=1) + np.linalg.norm(B, axis=1))[ok]
    return out.reshape(-1, 4), depths


def _make_view_segments(rng, P1, P2, K, R, t, w, h, n_seg, noise_px, min_len, clutter_med, max_depth=None):
    real, depths = _project_world_segments(P1, P2, K, R, t, w, h, max_depth=max_depth)
    if real.shape[0]:
        real = real + rng.normal(0.0, noise_px, size=real.shape)
        real[:, 0::2] = np.clip(real[:, 0::2], 0.0, float(w))
        real[:, 1::2] = np.clip(real[:, 1::2], 0.0, float(h))
        ln = np.hypot(real[:, 0] - real[:, 2], real[:, 1] - real[:, 3])
        keep = ln >= min_len
        real, depths = real[keep], depths[keep]
    n_real = min(real.shape[0], n_seg)
    if real.shape[0] > n_seg:
        ln = np.hypot(real[:, 0] - real[:, 2], real[:, 1] - real[:, 3])
        order = np.argsort(-ln, kind="stable")[:n_seg]
        real = real[order]
    n_cl = n_seg - n_real
    segs = [real]
    if n_cl > 0:
        cx = rng.uniform(0, w, size=n_cl)
        cy = rng.uniform(0, h, size=n_cl)
        ln = np.maximum(min_len, rng.lognormal(math.log(clutter_med), 0.6, size=n_cl))
        ang = rng.uniform(0, math.pi, size=n_cl)
        dx, dy = 0.5 * ln * np.cos(ang), 0.5 * ln * np.sin(ang)
        cl = np.stack([cx - dx, cy - dy, cx + dx, cy + dy], axis=1)
        cl[:, 0::2] = np.clip(cl[:, 0::2], 0.0, float(w))
        cl[:, 1::2] = np.clip(cl[:, 1::2], 0.0, float(h))
        segs.append(cl)
    segs = np.concatenate(segs, axis=0).astype(np.float32)
    ln = np.hypot(segs[:, 0] - segs[:, 2], segs[:, 1] - segs[:, 3])
    segs = segs[np.argsort(-ln, kind="stable")]
    med = float(np.median(depths)) if depths.size else 1.0
    return np.ascontiguousarray(segs), med


def _nearest_neighbors(centers: np.ndarray, axes: np.ndarray, nbrs: int) -> List[List[int]]:
    out = []
    n = centers.shape[0]
    for i in range(n):
        d = np.linalg.norm(centers - centers[i], axis=1)
        ang_ok = (axes @ axes[i]) > 0.0  # optical-axis angle < pi/2
        if n > 2000:
            # large scenes: the nbrs nearest admissible cameras are among the 4*nbrs+8 nearest ones unless many are
            # inadmissible; fall back to the full sort then (same result, stable order by distance then index)
            k = min(n - 1, 4 * nbrs + 8)
            part = np.argpartition(d, k)[:k + 1]
            part = part[np.lexsort((part, d[part]))]
            cand = [int(j) for j in part if j != i and ang_ok[j]]
            if len(cand) >= nbrs and (k + 1 >= n or d[part[-1]] > d[cand[nbrs - 1]]):
                out.append(sorted(cand[:nbrs]))
                continue
        cand = [j for j in np.argsort(d, kind="stable") if j != i and ang_ok[j]]
        out.append(sorted(int(j) for j in cand[:nbrs]))
    return out


def _world_segments(rng, n, box_lo, box_hi, med_len):
    c = rng.uniform(box_lo, box_hi, size=(n, 3))
    ln = rng.lognormal(math.log(med_len), 0.6, size=n)
    d = rng.normal(size=(n, 3))
    # man-made bias: most segments axis-aligned
    axis = rng.integers(0, 3, size=n)
    aligned = rng.uniform(size=n) < 0.7
    d[aligned] = np.eye(3)[axis[aligned]]
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    return c - 0.5 * ln[:, None] * d, c + 0.5 * ln[:, None] * d


PRESET_SHAPES = {  # views, segments per view, neighbours, image -- the BASELINE.json configs
    "tiny": (8, 160, 4, "640x480"), "c2": (50, 1000, 10, "640x480"), "c4": (500, 3000, 20, "1920x1080"),
    "c5": (5000, 5000, 20, "1920x1080"),
}


def make_scene(kind: str = "c2", n_views: Optional[int] = None, n_seg: Optional[int] = None,
               nbrs: Optional[int] = None, seed: Optional[int] = None,
               n_world: Optional[int] = None) -> Scene:
    """Synthetic scenes of the BASELINE.json shapes.

    kind: "tiny" (CPU tests), "c2" (50x1000x10, 640x480), "c4" (500x3000x20, 1920x1080),
          "c5" (5000x5000x20 city grid).  Sizes can be overridden for scaled-down variants.
    """
    presets = {
        #        views seg  nbrs world  w     h     f       box                       layout
        "tiny": (8, 160, 4, 120, 640, 480, 517.0, ((-5, -5, 0), (5, 5, 4)), "circle"),
        "c2": (50, 1000, 10, 600, 640, 480, 517.0, ((-5, -5, 0), (5, 5, 4)), "circle"),
        "c4": (500, 3000, 20, 4000, 1920, 1080, 1500.0, ((-30, -30, 0), (30, 30, 10)), "lawn"),
        "c5": (5000, 5000, 20, 200000, 1920, 1080, 1500.0, ((-500, -500, 0), (500, 500, 30)), "grid"),
    }
    idx = {"tiny": 9, "c2": 1, "c4": 3, "c5": 4}[kind]
    V, N, NB, L, w, h, f, box, layout = presets[kind]
    V = n_views or V
    N = n_seg or N
    NB = nbrs or NB
    L = n_world or L
    rng = np.random.Generator(np.random.PCG64((seed if seed is not None else BASE_SEED + idx)))
    lo, hi = np.asarray(box[0], float), np.asarray(box[1], float)
    P1, P2 = _world_segments(rng, L, lo, hi, 1.0 if kind != "c5" else 3.0)

    K = np.array([[f, 0, w / 2.0], [0, f, h / 2.0], [0, 0, 1.0]], dtype=np.float64)
    centers, Rs = [], []
    if layout == "circle":
        for i in range(V):
            a = 2 * math.pi * i / V
            c = np.array([3.0 * math.cos(a), 3.0 * math.sin(a), 1.6 + 0.2 * math.sin(3 * a)])
            tgt = np.array([-1.0 * math.cos(a), -1.0 * math.sin(a), 1.8])
            centers.append(c)
            Rs.append(look_at(c, tgt))
    elif layout == "lawn":
        per_row = max(2, int(round(math.sqrt(V * (hi[0] - lo[0]) / (hi[1] - lo[1])))))
        for i in range(V):
            row, col = divmod(i, per_row)
            x = lo[0] + 5 + (col if row % 2 == 0 else per_row - 1 - col) * 1.0
            y = lo[1] + 5 + row * 1.0
            c = np.array([x, y, 1.6])
            yaw = 0.35 * math.sin(0.37 * i)
            tgt = c + np.array([math.sin(yaw), math.cos(yaw), 0.05])
            centers.append(c)
            Rs.append(look_at(c, tgt))
    else:  # street grid, 2 m spacing
        per_row = max(2, int(round(math.sqrt(V))))
        for i in range(V):
            row, col = divmod(i, per_row)
            c = np.array([lo[0] + 20 + 2.0 * col, lo[1] + 20 + 14.0 * row, 1.7])
            yaw = 0.5 * math.sin(0.11 * i)
            tgt = c + np.array([math.cos(yaw), math.sin(yaw), 0.1])
            centers.append(c)
            Rs.append(look_at(c, tgt))
    centers = np.asarray(centers)
    axes = np.asarray([R[2] for R in Rs])
    nbr_lists = _nearest_neighbors(centers, axes, NB)

    views = []
    for i in range(V):
        R = Rs[i]
        t = -R @ centers[i]
        segs, med = _make_view_segments(rng, P1, P2, K, R, t, w, h, N, 0.5, 15.0, 40.0)
        views.append(SceneView(i, K.copy(), R, t, w, h, med, segs, nbr_lists[i]))
    params = dict(DEFAULT_PARAMS)
    params["num_neighbors"] = NB
    return Scene(kind, views, max(w, h), params, False)


def make_city(n_views: int = 5000, n_seg: int = 5000, nbrs: int = 20, n_world: int = 200000, seed: Optional[int] = None,
              view_range=None, max_depth: float = 300.0):
    """BASELINE config 5 (city scale): the "c5" preset's street-grid cameras (2 m spacing) and world segments, with
    ONE random stream per view so that the views can be generated independently -- every rank of a sharded run
    generates a slice and the segment tables are all-gathered (the scene tables are replicated).
    Returns (Scene with empty segment arrays outside view_range, segs (hi-lo, n_seg, 4) float32, median depths).
    World segments farther than max_depth from the camera are ignored (they would project below the 15 px
    minimum length anyway: 3 m at 300 m and f = 1500 is 15 px)."""
    V, N, NB, L = n_views, n_seg, nbrs, n_world
    w, h, f = 1920, 1080, 1500.0
    lo_b, hi_b = np.array([-500.0, -500.0, 0.0]), np.array([500.0, 500.0, 30.0])
    base = (seed if seed is not None else BASE_SEED + 4)
    rng = np.random.Generator(np.random.PCG64(base))
    P1, P2 = _world_segments(rng, L, lo_b, hi_b, 3.0)
    K = np.array([[f, 0, w / 2.0], [0, f, h / 2.0], [0, 0, 1.0]], dtype=np.float64)
    per_row = max(2, int(round(math.sqrt(V))))
    centers, Rs = [], []
    for i in range(V):
        row, col = divmod(i, per_row)
        c = np.array([lo_b[0] + 20 + 2.0 * col, lo_b[1] + 20 + 14.0 * row, 1.7])
        yaw = 0.5 * math.sin(0.11 * i)
        tgt = c + np.array([math.cos(yaw), math.sin(yaw), 0.1])
        centers.append(c)
        Rs.append(look_at(c, tgt))
    centers = np.asarray(centers)
    axes = np.asarray([R[2] for R in Rs])
    nbr_lists = _nearest_neighbors(centers, axes, NB)
    lo, hi = (0, V) if view_range is None else view_range
    segs = np.zeros((hi - lo, N, 4), dtype=np.float32)
    meds = np.zeros(hi - lo, dtype=np.float32)
    for i in range(lo, hi):
        vr = np.random.Generator(np.random.PCG64([base, 1000003 + i]))
        R = Rs[i]
        t = -R @ centers[i]
        sg, med = _make_view_segments(vr, P1, P2, K, R, t, w, h, N, 0.5, 15.0, 40.0, max_depth=max_depth)
        segs[i - lo] = sg
        meds[i - lo] = med
    views = []
    empty = np.zeros((0, 4), dtype=np.float32)
    for i in range(V):
        R = Rs[i]
        own = lo <= i < hi
        views.append(SceneView(i, K.copy(), R, -R @ centers[i], w, h, float(meds[i - lo]) if own else 1.0,
                               segs[i - lo] if own else empty, nbr_lists[i]))
    params = dict(DEFAULT_PARAMS)
    params["num_neighbors"] = NB
    return Scene("c5", views, max(w, h), params, False), segs, meds


# ----------------------------------------------------------------------------------------------
# Key-frame stream (BASELINE config 3): what L3DPPing::Run feeds its Line3D object, cycle by cycle
# (reference src/L3DPPing.cpp:98-236): delete the culled key frames, add the new ones, re-pose
# every current key frame (bundle adjustment moved it), match, reconstruct.
# ----------------------------------------------------------------------------------------------
@dataclasses.dataclass
class StreamCycle:
    deletes: List[int]                 # camera ids removed before this cycle's matching
    adds: List[SceneView]              # new key frames (ascending camera id)
    updates: List[tuple]               # (cam_id, R, t, median_depth, wps_or_neighbors) of every current key frame


@dataclasses.dataclass
class Stream:
    name: str
    cycles: List[StreamCycle]
    max_image_width: int
    params: Dict[str, float]
    neighbors_by_worldpoints: bool


def _small_rotation(rng, sigma_rad):
    w = rng.normal(0.0, sigma_rad, size=3)
    th = float(np.linalg.norm(w))
    if th < 1e-15:
        return np.eye(3)
    k = w / th
    Kx = np.array([[0, -k[2], k[1]], [k[2], 0, -k[0]], [-k[1], k[0], 0]])
    return np.eye(3) + math.sin(th) * Kx + (1 - math.cos(th)) * (Kx @ Kx)


def make_stream(n_keyframes: int = 300, n_seg: int = 1000, window: int = 20, nbrs: int = 10, init: int = 5,
                by_worldpoints: bool = True, cull_every: int = 7, jitter: float = 1.0,
                seed: Optional[int] = None, n_world: int = 600, n_points: int = 2500) -> Stream:
    """Synthetic 640x480 key-frame stream: a C2-like room, the camera circles it (150 key frames per lap),
    one new key frame per cycle after `init` initial ones, at most `window` current key frames (the oldest
    is dropped), every `cull_every`-th cycle additionally culls a key frame from the middle of the window,
    and every current key frame is re-posed each cycle by a small, shrinking perturbation (`jitter` scales
    it; 0 = poses never move).  Neighbours come from shared world points (the mode L3DPPing uses) or from
    explicit lists."""
    rng = np.random.Generator(np.random.PCG64(seed if seed is not None else BASE_SEED + 2))
    w, h, f = 640, 480, 517.0
    lo, hi = np.array([-5.0, -5.0, 0.0]), np.array([5.0, 5.0, 4.0])
    P1, P2 = _world_segments(rng, n_world, lo, hi, 1.0)
    pts = rng.uniform(lo, hi, size=(n_points, 3))
    K = np.array([[f, 0, w / 2.0], [0, f, h / 2.0], [0, 0, 1.0]], dtype=np.float64)
    base = []
    for i in range(n_keyframes):
        a = 2 * math.pi * i / 150.0
        c = np.array([3.0 * math.cos(a), 3.0 * math.sin(a), 1.6 + 0.2 * math.sin(3 * a)])
        tgt = np.array([-1.0 * math.cos(a), -1.0 * math.sin(a), 1.8])
        base.append((c, look_at(c, tgt)))

    def visible_points(R, t):
        X = pts @ R.T + t
        z = X[:, 2]
        ok = z > 0.3
        u = f * X[:, 0] / np.where(ok, z, 1.0) + w / 2.0
        v = f * X[:, 1] / np.where(ok, z, 1.0) + h / 2.0
        ok &= (u >= 0) & (u < w) & (v >= 0) & (v < h)
        ids = np.nonzero(ok)[0]
        return ids, z[ids]

    def posed(i, age):
        """Key frame i as bundle adjustment sees it `age` cycles after it was added."""
        c, R = base[i]
        if jitter <= 0.0:
            return R, -R @ c
        r = np.random.Generator(np.random.PCG64([BASE_SEED, i, age]))
        s = jitter / (1.0 + age)
        Rj = _small_rotation(r, 2.0e-3 * s) @ R
        cj = c + r.normal(0.0, 5.0e-3 * s, size=3)
        return Rj, -Rj @ cj

    cycles = []
    current: List[int] = []
    born = {}
    nxt = 0
    cyc = 0
    while nxt < n_keyframes:
        n_add = init if cyc == 0 else 1
        deletes = []
        if cyc > 0:
            while len(current) + n_add > window:
                deletes.append(current.pop(0))
            if cull_every and cyc % cull_every == 0 and len(current) > 4:
                deletes.append(current.pop(len(current) // 2))
        adds = []
        for _ in range(n_add):
            if nxt >= n_keyframes:
                break
            i = nxt
            nxt += 1
            R, t = posed(i, 0)
            segs, med = _make_view_segments(rng, P1, P2, K, R, t, w, h, n_seg, 0.5, 15.0, 40.0)
            ids, z = visible_points(R, t)
            med_pts = float(np.sort(z)[len(z) // 2]) if len(z) else med
            lst = ids.tolist() if by_worldpoints else [max(0, i - 1)]
            adds.append(SceneView(i, K.copy(), R, t, w, h, med_pts, segs, [], lst if by_worldpoints else None))
            if not by_worldpoints:
                adds[-1].neighbors = lst
            current.append(i)
            born[i] = cyc
        updates = []
        for i in current:
            R, t = posed(i, cyc - born[i])
            ids, z = visible_points(R, t)
            med_pts = float(np.sort(z)[len(z) // 2]) if len(z) else 1.0
            if by_worldpoints:
                lst = ids.tolist()
            else:  # the nbrs nearest current key frames (ids may point at frames added later or culled)
                d = sorted((abs(j - i), j) for j in current if j != i)
                lst = sorted(j for _, j in d[:nbrs])
            updates.append((i, R, t, med_pts, lst))
        cycles.append(StreamCycle(sorted(deletes), adds, updates))
        cyc += 1
    params = dict(DEFAULT_PARAMS)
    params["num_neighbors"] = nbrs
    return Stream("c3", cycles, max(w, h), params, by_worldpoints)


def drive_stream(stream: Stream, begin_cycle, delete, add, update, match, reconstruct, on_cycle=None,
                 max_cycles: Optional[int] = None):
    """Replay a stream through the six calls L3DPPing::Run makes (adapters of the object under test)."""
    for ci, cy in enumerate(stream.cycles[:max_cycles]):
        begin_cycle()
        for cam in cy.deletes:
            delete(cam)
        for v in cy.adds:
            add(v, v.worldpoints if stream.neighbors_by_worldpoints else v.neighbors)
        for cam, R, t, md, lst in cy.updates:
            update(cam, R, t, md, lst)
        match(stream.params)
        reconstruct()
        if on_cycle is not None:
            on_cycle(ci, cy)


# ----------------------------------------------------------------------------------------------
# NVM (the reference's own dump format: src/System.cc:459-535, header "NVM_BTREE_Test_v1")
# ----------------------------------------------------------------------------------------------
def read_nvm(path: str):
    """Returns (cams, points): cams = list of dict(name,f,q,C,dist); points = list of
    dict(X, obs=[(cam_idx, feat_idx, x, y), ...]).  Parsed line by line: the observation count the
    writer stores (`pMP->Observations()`) can exceed the observations it actually writes (bad key
    frames are skipped, src/System.cc:519-522), so the line length is authoritative."""
    with open(path, "r") as fh:
        lines = [ln.strip() for ln in fh.read().splitlines()]
    lines = [ln for ln in lines if ln]
    if not lines or not lines[0].startswith("NVM_"):
        raise ValueError("not an NVM file: %r" % (lines[0] if lines else ""))
    ncam = int(lines[1])
    cams = []
    for ln in lines[2:2 + ncam]:
        tk = ln.split()
        vals = [float(x) for x in tk[1:10]]   # f qw qx qy qz Cx Cy Cz dist
        cams.append(dict(name=tk[0], f=vals[0], q=vals[1:5], C=np.array(vals[5:8]), dist=vals[8]))
    npts = int(lines[2 + ncam])
    pts = []
    for ln in lines[3 + ncam:3 + ncam + npts]:
        tk = ln.split()
        X = np.array([float(tk[0]), float(tk[1]), float(tk[2])])
        obs = []
        for o in range(7, len(tk) - 3, 4):
            obs.append((int(tk[o]), int(tk[o + 1]), float(tk[o + 2]), float(tk[o + 3])))
        pts.append(dict(X=X, obs=obs))
    return cams, pts


def scene_from_nvm(path: str, n_seg: int = 400, width: int = 640, height: int = 480,
                   seed: Optional[int] = None, n_world: int = 300) -> Scene:
    """BASELINE config 1: cameras + world-point visibility from the reference's NVM dump, 2-D
    segments synthesised (the images the file names refer to are not shipped)."""
    cams, pts = read_nvm(path)
    rng = np.random.Generator(np.random.PCG64(seed if seed is not None else BASE_SEED + 0))
    X = np.asarray([p["X"] for p in pts])
    lo, hi = np.percentile(X, 5, axis=0), np.percentile(X, 95, axis=0)
    P1, P2 = _world_segments(rng, n_world, lo, hi, 0.25 * float(np.median(hi - lo)))
    wps: Dict[int, List[int]] = {i: [] for i in range(len(cams))}
    for pid, p in enumerate(pts):
        for (ci, _fi, _x, _y) in p["obs"]:
            if 0 <= ci < len(cams):
                wps[ci].append(pid)
    views = []
    for i, c in enumerate(cams):
        R = rotation_from_q(*c["q"])
        t = -R @ c["C"]
        K = np.array([[c["f"], 0, width / 2.0], [0, c["f"], height / 2.0], [0, 0, 1.0]])
        segs, _ = _make_view_segments(rng, P1, P2, K, R, t, width, height, n_seg, 0.5, 15.0, 40.0)
        d = [np.linalg.norm(X[pid] - c["C"]) for pid in wps[i]]
        med = float(sorted(d)[len(d) // 2]) if d else 1.0
        views.append(SceneView(i, K, R, t, width, height, med, segs, [], list(wps[i])))
    params = dict(DEFAULT_PARAMS)
    return Scene("c1_nvm", views, max(width, height), params, True)
