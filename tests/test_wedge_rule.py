"""CPU: the skip rule of K1's wedge test (csrc/k1_pairtest.cu: k1_rowsort_kernel's keys, hull_skips) restated in
numpy and checked against the oracle -- no GPU involved.

For groups of 32 source rows (sorted the way the kernel sorts them, and also in their natural order, which must be
just as sound) the rule names the targets the warp would skip.  Checked: (1) none of them is a match of
Line3D::matchingCPU as the oracle computes it with kNN off; (2) for every skipped (row, target) the two epipolar
intersection parameters lie on the same side of the target segment, at least 0.4 px away from it (the lemma of
DESIGN.md 4.1, with the kernel's 0.5 px margin less a float rounding allowance)."""
import math

import numpy as np
import pytest


def _lines(F, pts):
    l = (F @ np.c_[pts, np.ones(len(pts))].T).T
    return l / np.linalg.norm(l[:, :2], axis=1)[:, None]


def _rule(F, S, T, xb, order, dtype):
    """-> skip[N, M] (bool) for warps of 32 consecutive rows of `order`; the kernel's arithmetic in `dtype`."""
    N, M = len(S), len(T)
    c0, c1, c2 = F[:, 0], F[:, 1], F[:, 2]
    cands = [np.cross(c0, c1), np.cross(c1, c2), np.cross(c2, c0)]
    e = max(cands, key=lambda v: float(v @ v))
    cx = 0.25 * xb
    d = np.array([cx * e[2] - e[0], cx * e[2] - e[1]])
    d = d / np.linalg.norm(d)
    g = np.array([-d[1], d[0]])
    with np.errstate(divide="ignore"):
        R = np.linalg.norm([cx * e[2] - e[0], cx * e[2] - e[1]]) / abs(e[2])
    far = not (R <= 4.0 * xb)
    l1, l2 = _lines(F, S[:, 0:2]), _lines(F, S[:, 2:4])
    l1 = l1 * np.where(l1[:, :2] @ g < 0, -1.0, 1.0)[:, None]
    l2 = l2 * np.where(l2[:, :2] @ g < 0, -1.0, 1.0)[:, None]
    if far:    # distance of the middle of the image to the line: orders parallel lines too
        k1 = (l1[:, 0] * cx + l1[:, 1] * cx + l1[:, 2]).astype(np.float32)
        k2 = (l2[:, 0] * cx + l2[:, 1] * cx + l2[:, 2]).astype(np.float32)
    else:      # sine of the angle to g
        k1 = (l1[:, 1] * g[0] - l1[:, 0] * g[1]).astype(np.float32)
        k2 = (l2[:, 1] * g[0] - l2[:, 0] * g[1]).astype(np.float32)
    lo_line = np.where((k1 <= k2)[:, None], l1, l2).astype(dtype)
    hi_line = np.where((k1 <= k2)[:, None], l2, l1).astype(dtype)
    klo, khi = np.minimum(k1, k2), np.maximum(k1, k2)
    u = T[:, 2:4] - T[:, 0:2]
    L = np.linalg.norm(u, axis=1)
    u = u / L[:, None]
    q1 = T[:, 0:2].astype(dtype)
    u_, L_ = u.astype(dtype), L.astype(dtype)
    skip = np.zeros((N, M), dtype=bool)
    for w in range(0, N, 32):
        idx = order[w:w + 32]
        a, b = idx[np.argmin(klo[idx])], idx[np.argmax(khi[idx])]
        Hl, Hh = lo_line[a], hi_line[b]
        Nl = Hl[0] * q1[:, 0] + Hl[1] * q1[:, 1] + Hl[2]
        Dl = Hl[0] * u_[:, 0] + Hl[1] * u_[:, 1]
        Nh = Hh[0] * q1[:, 0] + Hh[1] * q1[:, 1] + Hh[2]
        Dh = Hh[0] * u_[:, 0] + Hh[1] * u_[:, 1]
        Nl2, Nh2 = Nl + L_ * Dl, Nh + L_ * Dh
        mn = np.minimum(np.minimum(Nl, Nl2), np.minimum(Nh, Nh2))
        mx = np.maximum(np.maximum(Nl, Nl2), np.maximum(Nh, Nh2))
        s = ((mn > 0.5) | (mx < -0.5)) & (Dl * Dh > 0) & (np.minimum(np.abs(Dl), np.abs(Dh)) > 1e-5)
        skip[np.ix_(idx, np.nonzero(s)[0])] = True
    return skip, klo, khi


def _sorted_order(klo, khi):
    c, w = 0.5 * (klo + khi), khi - klo
    cls = np.clip(np.floor(np.log2(np.maximum(w, 1e-30) / w.mean())) + 3, 0, 7).astype(int)
    q = (c - c.min()) / max(c.max() - c.min(), 1e-30)
    q = np.where(cls & 1, 1.0 - q, q)
    return np.lexsort((q, cls))


def _check_pair(oracle, sc, va, vb):
    o = oracle.OracleLine3D(sc.max_image_width, False)
    o.load_scene(sc)
    F, _, _, _, _ = o.match_only(va.cam_id, vb.cam_id, 0.25, -1)     # kNN off: every match of matchingCPU
    off, rec = o.lists(va.cam_id, 1)
    F = np.asarray(F, dtype=np.float64).reshape(3, 3)
    S, T = va.segs.astype(np.float64), vb.segs.astype(np.float64)
    xb = float(np.max(np.abs(T[:, [0, 2]]) + np.abs(T[:, [1, 3]])))
    rows = np.repeat(np.arange(len(off) - 1), np.diff(off.astype(np.int64)))
    match = np.zeros((len(S), len(T)), dtype=bool)
    match[rows, rec["tgt_seg"]] = True
    # exact intersection parameters (double)
    l1, l2 = _lines(F, S[:, 0:2]), _lines(F, S[:, 2:4])
    u = T[:, 2:4] - T[:, 0:2]
    L = np.linalg.norm(u, axis=1)
    u = u / L[:, None]
    q1h = np.c_[T[:, 0:2], np.ones(len(T))]
    with np.errstate(divide="ignore", invalid="ignore"):
        s1 = -(l1 @ q1h.T) / (l1[:, :2] @ u.T)
        s2 = -(l2 @ q1h.T) / (l2[:, :2] @ u.T)
    skipped_total = 0
    for dtype in (np.float64, np.float32):
        k = _rule(F, S, T, xb, np.arange(len(S)), dtype)
        for order in (np.arange(len(S)), _sorted_order(k[1], k[2])):
            skip, _, _ = _rule(F, S, T, xb, order, dtype)
            assert not (skip & match).any(), "the wedge rule skips a match of matchingCPU"
            below = (s1 < -0.4) & (s2 < -0.4)
            above = (s1 > L[None, :] + 0.4) & (s2 > L[None, :] + 0.4)
            assert (below | above)[skip].all(), "a skipped pair has an intersection within 0.4 px of the target segment"
            skipped_total += int(skip.sum())
    o.close() if hasattr(o, "close") else None
    return skipped_total, match.sum()


def test_wedge_rule_never_skips_a_match_side_by_side(oracle, scene_mod):
    sc = scene_mod.make_scene("tiny", n_views=6, n_seg=320, nbrs=3)
    tot = 0
    for a, b in ((0, 1), (2, 1), (3, 5)):
        n, m = _check_pair(oracle, sc, sc.views[a], sc.views[b])
        assert m > 0
        tot += n
    assert tot > 0      # the rule does skip something (sorted order: most of the matrix)


def test_wedge_rule_never_skips_a_match_c4_shape(oracle, scene_mod):
    sc = scene_mod.make_scene("c4", n_views=4, n_seg=640, nbrs=3)
    n, m = _check_pair(oracle, sc, sc.views[0], sc.views[1])
    assert n > 0.5 * 4 * 640 * 640 * 0.5 and m > 0


def rectified_scene(scene_mod, n_views=4, n_seg=400):
    """cameras side by side with one common rotation: every pair is a rectified stereo pair, its epipole a point
    at infinity, its epipolar lines parallel"""
    sc = scene_mod.make_scene("c2", n_views=n_views, n_seg=n_seg, nbrs=2)
    rng = np.random.Generator(np.random.PCG64(77))
    P1, P2 = scene_mod._world_segments(rng, 600, np.array([-5.0, -5.0, 0.0]), np.array([5.0, 5.0, 4.0]), 1.0)
    for i, v in enumerate(sc.views):
        c = np.array([-1.0 + 0.25 * i, -6.0, 1.6])
        R = scene_mod.look_at(c, c + np.array([0.0, 1.0, 0.0]))
        t = -R @ c
        segs, med = scene_mod._make_view_segments(rng, P1, P2, v.K, R, t, v.width, v.height, n_seg, 0.5, 15.0, 40.0)
        v.R, v.t, v.segs, v.median_depth = R, t, segs, med
    return sc


def test_wedge_rule_rectified_stereo(oracle, scene_mod):
    """parallel epipolar lines: the angle key cannot tell them apart, the distance key must"""
    sc = rectified_scene(scene_mod)
    tot = matches = 0
    for a, b in ((0, 1), (2, 1), (0, 3)):
        n, m = _check_pair(oracle, sc, sc.views[a], sc.views[b])
        tot += n
        matches += m
    assert tot > 0 and matches > 100


def test_wedge_rule_forward_motion(oracle, scene_mod):
    """epipole inside both images: lines in every direction, wide wedges, the key wraps"""
    sc = scene_mod.make_scene("c2", n_views=4, n_seg=320, nbrs=2)
    for i, v in enumerate(sc.views):
        c = np.array([0.4 * i - 6.0, 0.03 * i, 1.7])
        R = scene_mod.look_at(c, c + np.array([1.0, 0.0, 0.02]))
        v.R, v.t = R, -R @ c
    for a, b in ((0, 1), (2, 1)):
        _check_pair(oracle, sc, sc.views[a], sc.views[b])


def random_two_view_scene(scene_mod, seed, n_seg=300):
    """Two views of one random 3-D scene; the relative pose cycles through the cases that stress the sort key:
    translation along the optical axis (epipole at the image centre), sideways (far away or at infinity), vertical,
    oblique, with small and large rotations."""
    rng = np.random.Generator(np.random.PCG64(1000 + seed))
    sc = scene_mod.make_scene("c2", n_views=2, n_seg=n_seg, nbrs=1)
    P1, P2 = scene_mod._world_segments(rng, 500, np.array([-4.0, -2.0, 0.0]), np.array([4.0, 6.0, 3.5]), 1.0)
    direction = [np.array([0.0, 1.0, 0.0]), np.array([1.0, 0.0, 0.0]), np.array([0.0, 0.0, 1.0]),
                 rng.standard_normal(3)][seed % 4]
    direction = direction / np.linalg.norm(direction)
    base = np.array([0.0, -7.0, 1.6])
    yaw = [0.0, 0.0, 0.02, 0.3][(seed // 2) % 4]
    for i, v in enumerate(sc.views):
        c = base + i * (0.3 + 0.4 * rng.random()) * direction
        tgt = c + np.array([math.sin(i * yaw), math.cos(i * yaw), 0.03 * i * (seed % 3)])
        R = scene_mod.look_at(c, tgt)
        t = -R @ c
        segs, med = scene_mod._make_view_segments(rng, P1, P2, v.K, R, t, v.width, v.height, n_seg, 0.5, 15.0, 40.0)
        v.R, v.t, v.segs, v.median_depth = R, t, segs, med
        v.neighbors = [sc.views[1 - i].cam_id]
    return sc


@pytest.mark.parametrize("seed", range(8))
def test_wedge_rule_random_two_view_geometries(oracle, scene_mod, seed):
    sc = random_two_view_scene(scene_mod, seed)
    n, m = _check_pair(oracle, sc, sc.views[0], sc.views[1])
    n2, m2 = _check_pair(oracle, sc, sc.views[1], sc.views[0])
    assert m + m2 > 0
