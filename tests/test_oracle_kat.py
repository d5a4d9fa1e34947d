"""CPU: hand-derivable known-answer tests that pin the oracle's restatement of the reference
primitives (the reference ships no tests or golden vectors for this path: SURVEY.md section 8c)."""
import ctypes as C
import math

import numpy as np
import pytest


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def overlap(L, pts):
    a = np.ascontiguousarray(np.asarray(pts, dtype=np.float64).reshape(12))
    return L.orc_kat_mutual_overlap(_p(a))


def test_mutual_overlap_known_answers(oracle):
    """Line3D::mutualOverlap (src/line3D.cc:1283-1362) on collinear points of the x axis."""
    L = oracle.lib()
    # [0,10] vs [5,20]: inner 5, outer 20
    assert overlap(L, [[0, 0, 1], [10, 0, 1], [5, 0, 1], [20, 0, 1]]) == pytest.approx(0.25, abs=0)
    # identical segments: overlap 1
    assert overlap(L, [[3, 4, 1], [9, 12, 1], [3, 4, 1], [9, 12, 1]]) == 1.0
    # containment: [0,40] contains [10,20] -> 10/40
    assert overlap(L, [[0, 0, 1], [40, 0, 1], [10, 0, 1], [20, 0, 1]]) == 0.25
    # disjoint segments: no endpoint inside the other -> 0
    assert overlap(L, [[0, 0, 1], [10, 0, 1], [11, 0, 1], [20, 0, 1]]) == 0.0
    # outer distance below one pixel -> 0 (src/line3D.cc:1319)
    assert overlap(L, [[0, 0, 1], [0.5, 0, 1], [0.1, 0, 1], [0.6, 0, 1]]) == 0.0
    # order of the end points does not matter
    assert overlap(L, [[10, 0, 1], [0, 0, 1], [20, 0, 1], [5, 0, 1]]) == 0.25
    # touching intervals count as "on segment" (dot < eps) but overlap is 0/len
    assert overlap(L, [[0, 0, 1], [10, 0, 1], [10, 0, 1], [20, 0, 1]]) == 0.0


def test_fundamental_matrix_pure_translation(oracle):
    """Two fronto-parallel cameras, pure x translation: F ~ [t]_x K-scaled, horizontal epipolar
    lines (src/line3D.cc:1058-1094)."""
    L = oracle.lib()
    f = 500.0
    K = np.array([[f, 0, 320], [0, f, 240], [0, 0, 1.0]])
    R = np.eye(3)
    t1 = np.zeros(3)
    t2 = np.array([-1.0, 0, 0])  # camera 2 centre at x=+1
    F = np.zeros(9)
    L.orc_kat_fundamental(_p(K.ravel().copy()), _p(R.ravel().copy()), _p(t1), _p(K.ravel().copy()),
                          _p(R.ravel().copy()), _p(t2), _p(F))
    F = F.reshape(3, 3)
    x = np.array([100.0, 77.0, 1.0])
    l = F @ x
    # epipolar line of (100,77) is y = 77: a = 0, -c/b = 77
    assert abs(l[0]) < 1e-15
    assert -l[2] / l[1] == pytest.approx(77.0, rel=1e-12)
    # epipolar constraint for a true correspondence (depth 5 -> disparity f*B/Z = 100)
    xp = np.array([0.0, 77.0, 1.0])
    assert abs(xp @ F @ x) < 1e-12


def test_inverse3_matches_numpy(oracle):
    L = oracle.lib()
    rng = np.random.default_rng(0)
    for _ in range(20):
        A = rng.normal(size=(3, 3))
        out = np.zeros(9)
        L.orc_kat_inverse3(_p(A.ravel().copy()), _p(out))
        assert np.allclose(out.reshape(3, 3), np.linalg.inv(A), rtol=1e-10, atol=1e-12)


def test_angle_between_segments(oracle):
    """Line3D::angleBetweenSeg3D (src/line3D.cc:1841-1853): undirected angle in degrees."""
    L = oracle.lib()
    x = np.array([1.0, 0, 0]); y = np.array([0, 1.0, 0])
    assert L.orc_kat_angle(_p(x), _p(y)) == pytest.approx(90.0, abs=1e-5)
    assert L.orc_kat_angle(_p(x), _p(x)) == pytest.approx(0.0, abs=1e-5)
    assert L.orc_kat_angle(_p(x), _p(-x)) == pytest.approx(0.0, abs=1e-5)      # 180 -> 0 (undirected)
    d = np.array([math.cos(math.radians(30)), math.sin(math.radians(30)), 0])
    assert L.orc_kat_angle(_p(x), _p(d)) == pytest.approx(30.0, abs=1e-4)
    d2 = np.array([math.cos(math.radians(150)), math.sin(math.radians(150)), 0])
    assert L.orc_kat_angle(_p(x), _p(d2)) == pytest.approx(30.0, abs=1e-4)
    # orthogonal directions: sim_a = exp(-8100/200)
    assert L.orc_kat_expf(-90.0 * 90.0 / 200.0) == pytest.approx(math.exp(-40.5), rel=1e-6)


def test_distance_point_to_line(oracle):
    """Segment3D::distance_Point2Line (include/segment3D.h:80-84)."""
    L = oracle.lib()
    P1 = np.array([0.0, 0, 0]); P2 = np.array([10.0, 0, 0])
    assert L.orc_kat_dist_point_line(_p(P1), _p(P2), _p(np.array([3.0, 4.0, 0]))) == pytest.approx(4.0, abs=1e-6)
    assert L.orc_kat_dist_point_line(_p(P1), _p(P2), _p(np.array([-5.0, 0, 12.0]))) == pytest.approx(12.0, abs=1e-6)
    assert L.orc_kat_dist_point_line(_p(P1), _p(P2), _p(np.array([7.0, 0, 0]))) == 0.0


def test_clustering_known_answer(oracle):
    """Felzenszwalb-Huttenlocher union-find as src/clustering.cc:7-48 with c=3: with thresholds
    starting at c=3 > any affinity (<=1) every edge joins, so connected components are clusters."""
    L = oracle.lib()
    ij = np.array([[0, 1], [1, 0], [1, 2], [2, 1], [3, 4], [4, 3]], dtype=np.int32)
    w = np.array([0.9, 0.9, 0.6, 0.6, 0.8, 0.8], dtype=np.float32)
    out = np.zeros(6, dtype=np.int32)
    n = L.orc_kat_cluster(_p(ij), _p(w), len(w), 6, _p(out))
    assert n == 6
    assert out[0] == out[1] == out[2]
    assert out[3] == out[4] and out[3] != out[0]
    assert out[5] == 5  # isolated node keeps its own id
    # rank-union tie: the second argument becomes the root (include/universe.h:95-105)
    assert out[3] == 4


def test_two_view_triangulation_depth(oracle, scene_mod):
    """Two fronto-parallel cameras with baseline B: a vertical 3-D segment at depth Z gets
    triangulated depths |X - C| for both views (src/line3D.cc:1365-1390), i.e. Z/cos of the ray."""
    f, B, Z = 517.0, 0.5, 4.0
    K = np.array([[f, 0, 320], [0, f, 240], [0, 0, 1.0]])
    R = np.eye(3)
    C1, C2 = np.array([0.0, 0, 0]), np.array([B, 0, 0])
    P = [np.array([0.3, -0.5, Z]), np.array([0.3, 0.6, Z])]

    def proj(C, X):
        x = K @ (R @ (X - C))
        return x[:2] / x[2]
    s1 = np.array([[*proj(C1, P[0]), *proj(C1, P[1])]], dtype=np.float32)
    s2 = np.array([[*proj(C2, P[0]), *proj(C2, P[1])]], dtype=np.float32)
    # pad with far-away clutter so that every view has a few segments
    o = oracle.OracleLine3D(640, False, threads=1)
    o.add_image(0, K, R, -R @ C1, 640, 480, Z, [1], s1)
    o.add_image(1, K, R, -R @ C2, 640, 480, Z, [0], s2)
    o.update_image(0, R, -R @ C1, Z, [1])
    o.update_image(1, R, -R @ C2, Z, [0])
    o.match_images(5.0, 10.0, 10, 0.25, 10, -1.0)
    off, rec = o.lists(0, 0)
    assert len(rec) == 1 and rec[0]["tgt_cam"] == 1 and rec[0]["tgt_seg"] == 0
    assert rec[0]["overlap"] == pytest.approx(1.0, abs=1e-4)
    # the oracle shifts the scene by the median camera centre; depths are distances, unaffected
    for d, X in ((rec[0]["d_p1"], P[0]), (rec[0]["d_p2"], P[1])):
        assert d == pytest.approx(np.linalg.norm(X - C1), rel=1e-4)
    for d, X in ((rec[0]["d_q1"], P[0]), (rec[0]["d_q2"], P[1])):
        assert d == pytest.approx(np.linalg.norm(X - C2), rel=1e-4)


def test_serial_and_parallel_oracle_identical(oracle, scene_mod):
    """OpenMP is only used where results are order-free: 1 thread == all threads, bit for bit."""
    sc = scene_mod.make_scene("tiny", seed=5)
    a = oracle.run_scene(sc, threads=1)
    b = oracle.run_scene(sc, threads=0)
    assert (a.entries().tobytes() == b.entries().tobytes())
    ea, wa = a.edges(); eb, wb = b.edges()
    assert (ea == eb).all() and wa.tobytes() == wb.tobytes()
    assert (a.cluster_ids() == b.cluster_ids()).all()
    for v in sc.views:
        assert a.lists(v.cam_id, 0)[1].tobytes() == b.lists(v.cam_id, 0)[1].tobytes()


def test_sparse_matrix_known_answer(oracle):
    """SparseMatrix::SparseMatrix (src/sparsematrix.cc:8-61) on a hand-made A_: 4 nodes, node 3 isolated.
    Column sort = (j, i) ascending; start index of a column = position of its first entry, -1 if empty."""
    ij = np.array([[0, 1], [1, 0], [2, 0], [0, 2], [2, 1], [1, 2]], dtype=np.int32)
    w = np.array([0.9, 0.9, 0.6, 0.6, 0.75, 0.75], dtype=np.float32)
    ent, st = oracle.sparse_matrix(ij, w, 4)
    assert ent[:, :2].tolist() == [[1, 0], [2, 0], [0, 1], [2, 1], [0, 2], [1, 2]]
    assert ent[:, 2].tolist() == [np.float32(0.9), np.float32(0.6), np.float32(0.9), np.float32(0.75),
                                  np.float32(0.6), np.float32(0.75)]
    assert (ent[:, 3] == 0).all() and st.tolist() == [0, 2, 4, -1]
    ent, st = oracle.sparse_matrix(ij, w, 4, norm=3.0, sort_by_row=True)
    assert ent[:, :2].tolist() == [[0, 1], [0, 2], [1, 0], [1, 2], [2, 0], [2, 1]]
    assert ent[0, 2] == np.float32(0.9) / np.float32(3.0) and st.tolist() == [0, 2, 4, -1]


def test_find_collinear_known_answer(oracle):
    """View::findCollinCPU (src/view.cc:238-293) on hand-made segments:
    0: (0,0)-(10,0); 1: (20,0)-(30,0): same line, gap -> collinear with 0;
    2: (5,0)-(25,0): an endpoint of it lies between the endpoints of 0 and of 1 (dot test <= 0) -> overlap,
       never collinear;
    3: (20,1.5)-(30,1.5): 1.5 px beside 1 -- the overlap test is a 2-D dot product, (p1-x).(p2-x) = 2.25 > 0
       for both of its endpoints, so it does NOT count as overlapping 1, and all distances are 1.5;
    4: (20,0)-(30,2): shares an endpoint with 1 (dot = 0 -> overlap), crosses 3 (dot = -0.75), and the
       endpoints of 0 are 3.9 px / 1.96 px from its line."""
    L = np.array([[0, 0, 10, 0], [20, 0, 30, 0], [5, 0, 25, 0], [20, 1.5, 30, 1.5], [20, 0, 30, 2]], dtype=np.float32)
    t2 = oracle.find_collinear(L, 2.0)
    assert t2.tolist() == [[0, 1, 0, 1, 0], [1, 0, 0, 1, 0], [0, 0, 0, 0, 0], [1, 1, 0, 0, 0], [0, 0, 0, 0, 0]]
    assert (t2 == t2.T).all()                       # max(d1, d2) is symmetric in the pair
    t1 = oracle.find_collinear(L, 1.0)              # the 1.5 px neighbours drop out
    assert t1.tolist() == [[0, 1, 0, 0, 0], [1, 0, 0, 0, 0], [0] * 5, [0] * 5, [0] * 5]
    assert (oracle.find_collinear(L, 4.0)[0] == [0, 1, 0, 1, 1]).all()   # 3.92 px < 4 px: segment 4 joins


def test_sparse_matrix_against_an_independent_sort(oracle):
    """The oracle's SparseMatrix restatement against numpy's lexsort on random A_-shaped edge lists
    (every directed pair at most once, as Line3D::unused guarantees)."""
    rng = np.random.default_rng(7)
    for n, m in ((5, 6), (40, 300), (500, 4000)):
        pairs = set()
        while len(pairs) < m:
            i, j = rng.integers(0, n, size=2)
            if i != j:
                pairs.add((int(i), int(j)))
        und = sorted({tuple(sorted(p)) for p in pairs})
        rng.shuffle(und)
        ij = np.array([q for (a, b) in und for q in ((a, b), (b, a))], dtype=np.int32)
        w = np.repeat(rng.uniform(0.5, 1.0, size=len(und)).astype(np.float32), 2)
        for by_row in (False, True):
            ent, st = oracle.sparse_matrix(ij, w, n, 1.0, by_row)
            prim, sec = (ij[:, 0], ij[:, 1]) if by_row else (ij[:, 1], ij[:, 0])
            order = np.lexsort((sec, prim))
            assert (ent[:, 0] == ij[order, 0]).all() and (ent[:, 1] == ij[order, 1]).all()
            assert (ent[:, 2].view(np.uint32) == w[order].view(np.uint32)).all()
            want = np.full(n, -1, dtype=np.int64)
            ps = prim[order]
            first = np.r_[True, ps[1:] != ps[:-1]]
            want[ps[first]] = np.nonzero(first)[0]
            assert (st == want).all()


def test_find_collinear_against_numpy(oracle, scene_mod):
    """View::findCollinCPU restated a second time, vectorised in numpy with the same operation widths
    (double cross products and dot tests, sqrtf of the float-rounded squared norm, float maxima)."""
    base = scene_mod.make_scene("tiny").views[2].segs.astype(np.float32)
    a, b = base[:40, :2], base[:40, 2:]
    segs = np.ascontiguousarray(np.concatenate([np.concatenate([a, a + 0.4 * (b - a)], axis=1),
                                                np.concatenate([a + 0.62 * (b - a), b], axis=1), base[40:]]).astype(np.float32))
    n = len(segs)
    P0 = np.stack([segs[:, 0], segs[:, 1], np.ones(n)], axis=1).astype(np.float64)
    P1 = np.stack([segs[:, 2], segs[:, 3], np.ones(n)], axis=1).astype(np.float64)
    L = np.cross(P0, P1)                                            # line of every segment

    def on_seg(p1, p2, x):                                         # (n,1,3),(n,1,3),(1,n,3) -> (n,n)
        return ((p1[..., 0] - x[..., 0]) * (p2[..., 0] - x[..., 0]) + (p1[..., 1] - x[..., 1]) * (p2[..., 1] - x[..., 1])) < 1e-12

    def dist(l, p):
        num = l[..., 0] * p[..., 0] + l[..., 1] * p[..., 1] + l[..., 2]
        den = np.sqrt((l[..., 0] * l[..., 0] + l[..., 1] * l[..., 1]).astype(np.float32))   # sqrtf(float)
        return np.abs(num / den.astype(np.float64)).astype(np.float32)
    r0, r1, c0, c1 = P0[:, None, :], P1[:, None, :], P0[None, :, :], P1[None, :, :]
    overlap = on_seg(r0, r1, c0) | on_seg(r0, r1, c1) | on_seg(c0, c1, r0) | on_seg(c0, c1, r1)
    d1 = np.maximum(dist(L[:, None, :], c0), dist(L[:, None, :], c1))
    d2 = np.maximum(dist(L[None, :, :], r0), dist(L[None, :, :], r1))
    for t in (0.75, 2.0, 6.0):
        want = (~overlap) & (np.maximum(d1, d2) < np.float32(t))
        np.fill_diagonal(want, False)
        got = oracle.find_collinear(segs, t)
        assert (got.astype(bool) == want).all() and want.sum() >= 80
