"""CPU: a second, independent restatement of Line3D::matchingCPU (src/line3D.cc:1097-1212) with
mutualOverlap (:1283-1362), pointOnSegment (:1274-1280), triangulationDepths (:1365-1390) and
View::getNormalizedRay (src/view.cc:346-350), written from the reference in vectorised numpy float64 /
float32 with the reference's operation order -- and compared bit for bit with the oracle's C++ restatement
(kNN off, so no priority queue is involved).  Two readings of the same source agreeing is not the reference
itself, but it catches slips of either reading."""
import numpy as np

EPS = 1e-12


def _dot3(a, b):
    return a[..., 0] * b[..., 0] + a[..., 1] * b[..., 1] + a[..., 2] * b[..., 2]


def _cross(a, b):
    return np.stack([a[..., 1] * b[..., 2] - a[..., 2] * b[..., 1],
                     a[..., 2] * b[..., 0] - a[..., 0] * b[..., 2],
                     a[..., 0] * b[..., 1] - a[..., 1] * b[..., 0]], axis=-1)


def _matvec(M, v):
    return np.stack([M[i, 0] * v[..., 0] + M[i, 1] * v[..., 1] + M[i, 2] * v[..., 2] for i in range(3)], axis=-1)


def _normalized(v):
    return v / np.sqrt(_dot3(v, v))[..., None]


def _match_pair_numpy(ls, lt, F, Ms, Mt, Cs, Ct, W, thr):
    ns, nt = len(ls), len(lt)
    one = np.ones(ns)
    p1 = np.stack([ls[:, 0].astype(np.float64), ls[:, 1].astype(np.float64), one], axis=1)
    p2 = np.stack([ls[:, 2].astype(np.float64), ls[:, 3].astype(np.float64), one], axis=1)
    one = np.ones(nt)
    q1 = np.stack([lt[:, 0].astype(np.float64), lt[:, 1].astype(np.float64), one], axis=1)
    q2 = np.stack([lt[:, 2].astype(np.float64), lt[:, 3].astype(np.float64), one], axis=1)
    e1, e2 = _matvec(F, p1)[:, None, :], _matvec(F, p2)[:, None, :]
    l2 = _cross(q1, q2)[None, :, :]
    a, b = _cross(l2, e1), _cross(l2, e2)                      # (ns, nt, 3)
    ok = (np.abs(a[..., 2]) > EPS) & (np.abs(b[..., 2]) > EPS)
    with np.errstate(divide="ignore", invalid="ignore"):
        a = a / a[..., 2:3]
        b = b / b[..., 2:3]
    Wd = float(W)
    inb = ~((a[..., 0] < 0) | (a[..., 0] > Wd) | (a[..., 1] < 0) | (a[..., 1] > Wd) |
            (b[..., 0] < 0) | (b[..., 0] > Wd) | (b[..., 1] < 0) | (b[..., 1] > Wd))
    ok &= inb
    Q1 = np.broadcast_to(q1[None, :, :], a.shape)
    Q2 = np.broadcast_to(q2[None, :, :], a.shape)
    pts = [a, b, Q1, Q2]

    def on_seg(x, s1, s2):
        return ((s1[..., 0] - x[..., 0]) * (s2[..., 0] - x[..., 0]) + (s1[..., 1] - x[..., 1]) * (s2[..., 1] - x[..., 1])) < EPS
    touch = on_seg(a, Q1, Q2) | on_seg(b, Q1, Q2) | on_seg(Q1, a, b) | on_seg(Q2, a, b)

    def dist(u, v):
        d = u - v
        return np.sqrt(d[..., 0] * d[..., 0] + d[..., 1] * d[..., 1] + d[..., 2] * d[..., 2])
    max_dist = np.zeros(a.shape[:2], dtype=np.float32)
    o1 = np.zeros(a.shape[:2], dtype=np.int64)
    o2 = np.full(a.shape[:2], 3, dtype=np.int64)
    for i in range(3):
        for j in range(i + 1, 4):
            with np.errstate(invalid="ignore"):
                dij = dist(pts[i], pts[j]).astype(np.float32)
                better = dij > max_dist
            max_dist = np.where(better, dij, max_dist)
            o1 = np.where(better, i, o1)
            o2 = np.where(better, j, o2)
    inner = {(0, 1): (2, 3), (0, 2): (1, 3), (0, 3): (1, 2), (1, 2): (0, 3), (1, 3): (0, 2), (2, 3): (0, 1)}
    overlap = np.zeros(a.shape[:2], dtype=np.float32)
    for (x, y), (i1, i2) in inner.items():
        sel = (o1 == x) & (o2 == y)
        with np.errstate(divide="ignore", invalid="ignore"):
            ov = (dist(pts[i1], pts[i2]) / max_dist.astype(np.float64)).astype(np.float32)
        overlap = np.where(sel, ov, overlap)
    overlap = np.where(touch & ~(max_dist < np.float32(1.0)), overlap, np.float32(0.0))
    ok &= overlap > np.float32(thr)
    # triangulationDepths, both ways
    rp1, rp2 = _normalized(_matvec(Ms, p1)), _normalized(_matvec(Ms, p2))      # (ns,3)
    rq1, rq2 = _normalized(_matvec(Mt, q1)), _normalized(_matvec(Mt, q2))      # (nt,3)
    n_t = _normalized(_cross(rq1, rq2))                                       # plane through the tgt segment
    n_s = _normalized(_cross(rp1, rp2))

    def depths(Ca, ra1, ra2, Cb, nb):      # a: (A,3) rays of the view being measured, nb: (B,3) normals of the other
        d1 = ra1 @ nb.T if False else (ra1[:, None, 0] * nb[None, :, 0] + ra1[:, None, 1] * nb[None, :, 1] + ra1[:, None, 2] * nb[None, :, 2])
        d2 = (ra2[:, None, 0] * nb[None, :, 0] + ra2[:, None, 1] * nb[None, :, 1] + ra2[:, None, 2] * nb[None, :, 2])
        num = (_dot3(np.broadcast_to(Cb, nb.shape), nb) - _dot3(nb, np.broadcast_to(Ca, nb.shape)))[None, :]
        bad = (np.abs(d1) < EPS) | (np.abs(d2) < EPS)
        with np.errstate(divide="ignore", invalid="ignore"):
            return np.where(bad, -1.0, num / d1), np.where(bad, -1.0, num / d2)
    ds1, ds2 = depths(Cs, rp1, rp2, Ct, n_t)                   # (ns, nt)
    dt1, dt2 = depths(Ct, rq1, rq2, Cs, n_s)                   # (nt, ns)
    dt1, dt2 = dt1.T, dt2.T
    ok &= (ds1 > EPS) & (ds2 > EPS) & (dt1 > EPS) & (dt2 > EPS)
    rows, cols = np.nonzero(ok)                                # row-major = (r asc, c asc): push order with kNN off
    return rows, cols, overlap[rows, cols], ds1[rows, cols].astype(np.float32), ds2[rows, cols].astype(np.float32), \
        dt1[rows, cols].astype(np.float32), dt2[rows, cols].astype(np.float32)


def test_matching_second_reading_equals_the_oracle(oracle, scene_mod):
    sc = scene_mod.make_scene("tiny", seed=3, n_views=4, n_seg=220, nbrs=3)
    n_total = 0
    for ia, ib in ((1, 2), (2, 0), (0, 3)):
        va, vb = sc.views[ia], sc.views[ib]
        o = oracle.OracleLine3D(sc.max_image_width, False)
        o.load_scene(sc)
        F, Ms, Mt, Cs, Ct = o.match_only(va.cam_id, vb.cam_id, 0.25, -1)
        off, rec = o.lists(va.cam_id, 1)
        o.close()
        rows, cols, ov, d1, d2, d3, d4 = _match_pair_numpy(va.segs, vb.segs, np.asarray(F).reshape(3, 3),
                                                           np.asarray(Ms).reshape(3, 3), np.asarray(Mt).reshape(3, 3),
                                                           np.asarray(Cs), np.asarray(Ct), sc.max_image_width, 0.25)
        want_rows = np.repeat(np.arange(len(off) - 1), np.diff(off.astype(np.int64)))
        assert len(rows) == len(rec) > 200
        assert (rows == want_rows).all() and (cols == rec["tgt_seg"]).all()
        for got, name in ((ov, "overlap"), (d1, "d_p1"), (d2, "d_p2"), (d3, "d_q1"), (d4, "d_q2")):
            assert (got.view(np.uint32) == rec[name].view(np.uint32)).all(), (ia, ib, name)
        n_total += len(rec)
    assert n_total > 1000
