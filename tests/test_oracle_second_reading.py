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


def _score_row_numpy(seg, entries, cam, cams, k_of, two_sigA_sqr, expf, acosf):
    """Line3D::scoringCPU, new-match branch (src/line3D.cc:1513-1547) for ONE row, read a second time:
    similarityForScoring (:1685-1716), angleBetweenSeg3D (:1841-1853), View::unprojectSegment
    (src/view.cc:385-400), Segment3D ctor (include/segment3D.h:58-77), regularizerFrom3Dpoint (src/view.cc:474-477).
    `entries` = the row's matches in list order (tgt_cam, d_p1, d_p2); returns their score3D_."""
    M, C = cams[cam]
    p1 = np.array([float(seg[0]), float(seg[1]), 1.0])
    p2 = np.array([float(seg[2]), float(seg[3]), 1.0])
    r1, r2 = _normalized(_matvec(M, p1)), _normalized(_matvec(M, p2))
    k = np.float32(k_of[cam])
    geo = []
    for tgt_cam, d1, d2 in entries:
        P1 = C + r1 * float(d1)
        P2 = C + r2 * float(d2)
        d = P1 - P2
        length = np.float32(np.sqrt(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]))
        if length > 1e-12:
            dirv = _normalized(P2 - P1)
        else:
            P1 = P2 = dirv = np.zeros(3)
            length = np.float32(0.0)
        sig1, sig2 = np.float32(d1) * k, np.float32(d2) * k
        reg1 = np.float32(2.0) * sig1 * sig1
        reg2 = np.float32(2.0) * sig2 * sig2
        Ct, kt = cams[tgt_cam][1], np.float32(k_of[tgt_cam])
        e1, e2 = P1 - Ct, P2 - Ct
        s1t = np.float32(np.sqrt(e1[0] * e1[0] + e1[1] * e1[1] + e1[2] * e1[2]) * float(kt))
        s2t = np.float32(np.sqrt(e2[0] * e2[0] + e2[1] * e2[1] + e2[2] * e2[2]) * float(kt))
        reg1 = np.float32(0.5) * (reg1 + np.float32(2.0) * s1t * s1t)
        reg2 = np.float32(0.5) * (reg2 + np.float32(2.0) * s2t * s2t)
        geo.append((length, dirv, reg1, reg2))
    out = []
    for a, (cam_a, da1, da2) in enumerate(entries):
        la, dir_a, reg1, reg2 = geo[a]
        score = np.float32(0.0)
        per_cam = {}
        for b, (cam_b, db1, db2) in enumerate(entries):
            if cam_b == cam_a:
                continue
            lb, dir_b, _, _ = geo[b]
            if la < 1e-12 or lb < 1e-12:
                sim = np.float32(0.0)
            else:
                d1 = np.float32(da1) - np.float32(db1)
                d2 = np.float32(da2) - np.float32(db2)
                sim_p = min(expf(-d1 * d1 / reg1), expf(-d2 * d2 / reg2))
                dot_p = np.float32(dir_a[0] * dir_b[0] + dir_a[1] * dir_b[1] + dir_a[2] * dir_b[2])
                ang = np.float32(float(acosf(np.float32(max(min(dot_p, np.float32(1.0)), np.float32(-1.0))))) / np.pi * 180.0)
                if ang > np.float32(90.0):
                    ang = np.float32(180.0) - ang
                sim_a = expf(-ang * ang / np.float32(two_sigA_sqr))
                sim = min(sim_a, sim_p)
                if not sim > np.float32(0.5):
                    sim = np.float32(0.0)
            if cam_b in per_cam:
                if sim > per_cam[cam_b]:
                    score = score - per_cam[cam_b]
                    score = score + sim
                    per_cam[cam_b] = sim
            else:
                score = score + sim
                per_cam[cam_b] = sim
        out.append(score)
    return np.array(out, dtype=np.float32)


def test_scoring_second_reading_equals_the_oracle(oracle, scene_mod):
    import ctypes as C
    L = oracle.lib()
    L.orc_kat_expf.restype = C.c_float
    L.orc_kat_expf.argtypes = [C.c_float]
    L.orc_kat_acosf.restype = C.c_float
    L.orc_kat_acosf.argtypes = [C.c_float]

    def expf(x):
        return np.float32(L.orc_kat_expf(float(np.float32(x))))

    def acosf(x):
        return np.float32(L.orc_kat_acosf(float(np.float32(x))))
    sc = scene_mod.make_scene("tiny")
    o = oracle.run_scene(sc)
    cams = {v.cam_id: o.match_camera(v.cam_id) for v in sc.views}
    k_of = {v.cam_id: o.view_info(v.cam_id)["k"] for v in sc.views}
    checked = positive = 0
    with np.errstate(over="ignore", under="ignore"):
        for v in (sc.views[0], sc.views[3], sc.views[7]):      # first view: forward matches only; others: inverse too
            off, rec = o.lists(v.cam_id, 0)                    # lists as scored (before filterMatches)
            cand = [r for r in range(len(off) - 1) if 2 <= off[r + 1] - off[r] <= 40]
            hot = [r for r in cand if (rec["score"][off[r]:off[r + 1]] > 0).any()]     # rows where something scores
            rows = hot[:25] + [r for r in cand if r not in hot][:10]
            for r in rows:
                e = rec[off[r]:off[r + 1]]
                got = _score_row_numpy(v.segs[r], list(zip(e["tgt_cam"].tolist(), e["d_p1"], e["d_p2"])), v.cam_id, cams,
                                       k_of, 200.0, expf, acosf)
                assert (got.view(np.uint32) == e["score"].view(np.uint32)).all(), (v.cam_id, r, got, e["score"])
                checked += len(e)
                positive += int((got > 0).sum())
    assert checked > 1000 and positive > 100
    o.close()


def _det_funcs(oracle):
    import ctypes as C
    L = oracle.lib()
    L.orc_kat_expf.restype = C.c_float
    L.orc_kat_expf.argtypes = [C.c_float]
    L.orc_kat_acosf.restype = C.c_float
    L.orc_kat_acosf.argtypes = [C.c_float]
    f32 = np.float32
    return (lambda x: f32(L.orc_kat_expf(float(f32(x))))), (lambda x: f32(L.orc_kat_acosf(float(f32(x)))))


def _affinity_sim_numpy(a, b, va, vb, msdl, expf, acosf):
    """Line3D::similarity(Segment3D, Match, Segment2D, false) (src/line3D.cc:1737-1823) for two rows of
    estimated_position3D_ (a, b) and their views' (k, median depth); Segment3D::distance_Point2Line
    (include/segment3D.h:80-84) evaluated as (dir * diff^T) * dir."""
    f32 = np.float32

    def d_p2l(P1, dirv, P):
        d = P - P1
        h = np.array([(dirv[i] * d[0]) * dirv[0] + (dirv[i] * d[1]) * dirv[1] + (dirv[i] * d[2]) * dirv[2] for i in range(3)])
        e = (P1 + h) - P
        return f32(np.sqrt(e[0] * e[0] + e[1] * e[1] + e[2] * e[2]))
    if a["length"] < 1e-12 or b["length"] < 1e-12:
        return f32(0.0)
    dot_p = f32(a["dir"][0] * b["dir"][0] + a["dir"][1] * b["dir"][1] + a["dir"][2] * b["dir"][2])
    ang = f32(float(acosf(f32(max(min(dot_p, f32(1.0)), f32(-1.0))))) / np.pi * 180.0)
    if ang > f32(90.0):
        ang = f32(180.0) - ang
    sim_a = expf(-ang * ang / f32(200.0))
    c1, c2 = f32(va["median_depth"]), f32(vb["median_depth"])
    if msdl > 1e-12:
        c1, c2 = min(c1, msdl), min(c2, msdl)
    d11, d12 = d_p2l(b["P1"], b["dir"], a["P1"]), d_p2l(b["P1"], b["dir"], a["P2"])
    d21, d22 = d_p2l(a["P1"], a["dir"], b["P1"]), d_p2l(a["P1"], a["dir"], b["P2"])
    ka, kb = f32(va["k"]), f32(vb["k"])
    s11 = (c1 if a["d_p1"] > c1 else f32(a["d_p1"])) * ka
    s12 = (c1 if a["d_p2"] > c1 else f32(a["d_p2"])) * ka
    s21 = (c2 if b["d_p1"] > c2 else f32(b["d_p1"])) * kb
    s22 = (c2 if b["d_p2"] > c2 else f32(b["d_p2"])) * kb
    r11, r12, r21, r22 = f32(2.0) * s11 * s11, f32(2.0) * s12 * s12, f32(2.0) * s21 * s21, f32(2.0) * s22 * s22
    sp1 = min(expf(-d11 * d11 / r11), expf(-d12 * d12 / r12))
    sp2 = min(expf(-d21 * d21 / r21), expf(-d22 * d22 / r22))
    return f32(min(sim_a, min(sp1, sp2)))


def test_affinity_matrix_second_reading_equals_the_oracle(oracle, scene_mod):
    """Line3D::computingAffinityMatrix (src/line3D.cc:2275-2402, collinearity off) with Line3D::unused
    (:2405-2425) and getLocalID (:2428-2446) read a second time: walking estimated_position3D_ and the filtered
    lists reproduces A_ -- every (i, j, w), (j, i, w) pair in order -- and local2global_."""
    expf, acosf = _det_funcs(oracle)
    sc = scene_mod.make_scene("tiny")
    o = oracle.run_scene(sc)
    entries = o.entries()
    ent = {(int(e["src_cam"]), int(e["src_seg"])): e for e in entries}
    info = {v.cam_id: o.view_info(v.cam_id) for v in sc.views}
    lists = {v.cam_id: o.lists(v.cam_id, 1) for v in sc.views}
    msdl = np.float32(o.med_scene_depth_lines())
    A, ids, used = [], {}, set()

    def local_id(seg):
        if seg not in ids:
            ids[seg] = len(ids)
        return ids[seg]
    with np.errstate(over="ignore", under="ignore"):
        for e in entries:                                              # traversal order = estimated_position3D_ order
            seg = (int(e["src_cam"]), int(e["src_seg"]))
            off, rec = lists[seg[0]]
            id1 = -1
            for m2 in rec[off[seg[1]]:off[seg[1] + 1]]:
                seg2 = (int(m2["tgt_cam"]), int(m2["tgt_seg"]))
                if seg2 not in ent:
                    continue                                           # similarity() returns 0 without an entry
                sim = _affinity_sim_numpy(e, ent[seg2], info[seg[0]], info[seg2[0]], msdl, expf, acosf)
                if sim > np.float32(0.5) and frozenset((seg, seg2)) not in used:
                    used.add(frozenset((seg, seg2)))
                    if id1 < 0:
                        id1 = local_id(seg)
                    id2 = local_id(seg2)
                    A += [(id1, id2, sim), (id2, id1, sim)]
    ij, w = o.edges()
    assert len(A) == len(w) > 150
    assert [(a, b) for a, b, _ in A] == [tuple(x) for x in ij.tolist()]
    assert np.array([x for _, _, x in A], np.float32).tobytes() == w.tobytes()
    l2g = [tuple(x) for x in o.local2global().tolist()]
    assert [s for s, _ in sorted(ids.items(), key=lambda kv: kv[1])] == l2g
    o.close()


def test_orientation_filter_second_reading_equals_the_oracle(oracle, scene_mod):
    """Line3D::checkMatchOrientation (src/line3D.cc:962-1014) with View::segmentQualityAngle (src/view.cc:495-513):
    with kNN off, the list of the FIRST view at scoring time is exactly its forward matches (targets ascending) that
    pass the orientation test -- recomputed here from the numpy matching above plus a numpy orientation predicate,
    with the cameras and fundamental matrices matchImages used."""
    import ctypes as C
    L = oracle.lib()
    L.orc_kat_acos.restype = C.c_double
    L.orc_kat_acos.argtypes = [C.c_double]
    sc = scene_mod.make_scene("tiny", n_seg=120)
    sc.params = dict(sc.params, knn=-1)
    o = oracle.run_scene(sc)
    v = sc.views[0]
    off, rec = o.lists(v.cam_id, 0)
    Ms, Cs = o.match_camera(v.cam_id)
    per_row = [[] for _ in range(len(v.segs))]
    rejected = []
    for tgt in sorted(o.neighbors(v.cam_id)):
        F = o.fundamental(v.cam_id, tgt)
        assert F is not None                       # view 0 is the source of all its pairs
        vt = next(x for x in sc.views if x.cam_id == tgt)
        Mt, Ct = o.match_camera(tgt)
        rows, cols, ov, d1, d2, d3, d4 = _match_pair_numpy(v.segs, vt.segs, F, Ms, Mt, Cs, Ct, sc.max_image_width, 0.25)
        for r, c, a, x1, x2, x3, x4 in zip(rows, cols, ov, d1, d2, d3, d4):
            s = v.segs[r]
            p1 = np.array([float(s[0]), float(s[1]), 1.0])
            p2 = np.array([float(s[2]), float(s[3]), 1.0])
            P1 = Cs + _normalized(_matvec(Ms, p1)) * float(x1)
            P2 = Cs + _normalized(_matvec(Ms, p2)) * float(x2)
            d = P1 - P2
            length = np.float32(np.sqrt(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]))
            dirv = _normalized(P2 - P1) if length > 1e-12 else np.zeros(3)
            mid = np.array([0.5 * (float(s[0]) + float(s[2])), 0.5 * (float(s[1]) + float(s[3])), 1.0])
            rm = _normalized(_matvec(Ms, mid))
            dotp = rm[0] * dirv[0] + rm[1] * dirv[1] + rm[2] * dirv[2]
            ang = L.orc_kat_acos(min(max(float(dotp), -1.0), 1.0))
            if ang > float(np.float32(0.098174771)) and ang < float(np.float32(3.043417886)):
                per_row[r].append((tgt, int(c), a, x1, x2, x3, x4))
            else:
                rejected.append((r, tgt, int(c)))
    n = 0
    for r in range(len(v.segs)):
        e = rec[off[r]:off[r + 1]]
        assert len(e) == len(per_row[r]), r
        for x, y in zip(e, per_row[r]):
            assert int(x["tgt_cam"]) == y[0] and int(x["tgt_seg"]) == y[1]
            for name, val in (("overlap", y[2]), ("d_p1", y[3]), ("d_p2", y[4]), ("d_q1", y[5]), ("d_q2", y[6])):
                assert np.float32(x[name]).tobytes() == np.float32(val).tobytes(), (r, name)
        n += len(e)
    assert n > 300 and len(rejected) > 0          # the filter did reject something
    o.close()


def test_filter_matches_second_reading_equals_the_oracle(oracle, scene_mod):
    """Line3D::filterMatches (src/line3D.cc:1911-1983) read a second time: from the lists as scored, the kept
    entries (score > 0 and > 10 % of the view's maximum), the first strict maximum as best match, the
    hypothesis of every row whose best match scores > 0.75 (unprojectMatch, :1826-1838) and the view's median
    depth (View::update_median_depth, include/view.h:122-135)."""
    sc = scene_mod.make_scene("tiny")
    o = oracle.run_scene(sc)
    ent = {(int(e["src_cam"]), int(e["src_seg"])): e for e in o.entries()}
    n_ent = 0
    for v in sc.views:
        off0, rec0 = o.lists(v.cam_id, 0)
        off1, rec1 = o.lists(v.cam_id, 1)
        M, Cm = o.match_camera(v.cam_id)
        max_score = np.float32(0.0)
        for s in rec0["score"]:
            max_score = max(max_score, np.float32(s))
        lim = np.float32(0.1) * max_score
        depths = []
        for r in range(len(v.segs)):
            e = rec0[off0[r]:off0[r + 1]]
            kept = [x for x in e if x["score"] > np.float32(0.0) and x["score"] > lim]
            got = rec1[off1[r]:off1[r + 1]]
            assert len(kept) == len(got) and all(a.tobytes() == b.tobytes() for a, b in zip(kept, got)), (v.cam_id, r)
            best, best_score = None, np.float32(0.0)
            for x in kept:
                if x["score"] > best_score:
                    best, best_score = x, x["score"]
            if best is not None and best_score > np.float32(0.75):
                E = ent[(v.cam_id, r)]
                s = v.segs[r]
                P1 = Cm + _normalized(_matvec(M, np.array([float(s[0]), float(s[1]), 1.0]))) * float(best["d_p1"])
                P2 = Cm + _normalized(_matvec(M, np.array([float(s[2]), float(s[3]), 1.0]))) * float(best["d_p2"])
                assert (E["P1"] == P1).all() and (E["P2"] == P2).all() and (E["dir"] == _normalized(P2 - P1)).all()
                assert int(E["tgt_cam"]) == int(best["tgt_cam"]) and int(E["tgt_seg"]) == int(best["tgt_seg"])
                assert E["score"] == best["score"] and E["d_p1"] == best["d_p1"] and E["d_q2"] == best["d_q2"]
                depths += [np.float32(best["d_p1"]), np.float32(best["d_p2"])]
                n_ent += 1
            else:
                assert (v.cam_id, r) not in ent
        med = np.float32(1e-12) if not depths else sorted(depths)[len(depths) // 2]
        info = o.view_info(v.cam_id)
        assert np.float32(info["median_depth"]).tobytes() == np.float32(med).tobytes()
        assert np.float32(info["median_sigma"]).tobytes() == (np.float32(info["k"]) * np.float32(med)).tobytes()
    assert n_ent == len(ent) > 80
    o.close()


def test_inverse_matches_second_reading_equals_the_oracle(oracle, scene_mod):
    """Line3D::storeInverseMatches (src/line3D.cc:1986-2015): every match that scored > 0 in a view processed
    earlier appears, with views and depths swapped, score 0 and the orientation flag set, in the list of its
    target segment -- pushed in the order the sources were processed (camera id, row, list position), ahead of
    the target view's own forward matches."""
    sc = scene_mod.make_scene("tiny")
    o = oracle.run_scene(sc)
    lists = {v.cam_id: o.lists(v.cam_id, 0) for v in sc.views}       # as scored
    expect = {v.cam_id: [[] for _ in range(len(v.segs))] for v in sc.views}
    for v in sc.views:                                                # ascending camera id = processing order
        off, rec = lists[v.cam_id]
        for r in range(len(v.segs)):
            for x in rec[off[r]:off[r + 1]]:
                if x["score"] > np.float32(0.0) and int(x["tgt_cam"]) > v.cam_id:      # target not processed yet
                    expect[int(x["tgt_cam"])][int(x["tgt_seg"])].append(
                        (v.cam_id, r, x["overlap"], x["d_q1"], x["d_q2"], x["d_p1"], x["d_p2"]))
    n_inv = 0
    for v in sc.views:
        off, rec = lists[v.cam_id]
        for r in range(len(v.segs)):
            e = rec[off[r]:off[r + 1]]
            inv = [x for x in e if x["flags"] == 1]
            assert len(inv) == len(expect[v.cam_id][r]), (v.cam_id, r)
            assert list(e["flags"]) == [1] * len(inv) + [0] * (len(e) - len(inv))      # inverse first, then forward
            for x, y in zip(inv, expect[v.cam_id][r]):
                assert (int(x["tgt_cam"]), int(x["tgt_seg"])) == (y[0], y[1])
                assert (x["overlap"], x["d_p1"], x["d_p2"], x["d_q1"], x["d_q2"]) == y[2:]
            n_inv += len(inv)
    assert n_inv > 200
    o.close()


def test_incremental_score_deltas_second_reading_equals_the_oracle(oracle, scene_mod):
    """The incremental branch of Line3D::scoringCPU (src/line3D.cc:1439-1512) read a second time on a key-frame
    stream with additions, deletions and pose changes: a match scored in an earlier cycle keeps its score, plus
    the per-camera maximum similarity to every camera added this cycle, minus that to every camera deleted this
    cycle (both summed in ascending camera id), all evaluated with the CURRENT poses; a match that is new this
    cycle is scored from scratch against the siblings that are not being deleted.  The forward matches that fail
    this cycle's orientation re-test (src/line3D.cc:962-1014) are gone before scoring."""
    import ctypes as C
    import stream_utils
    expf, acosf = _det_funcs(oracle)
    L = oracle.lib()
    L.orc_kat_acos.restype = C.c_double
    L.orc_kat_acos.argtypes = [C.c_double]
    f32 = np.float32
    st = scene_mod.make_stream(n_keyframes=10, n_seg=220, window=6, nbrs=4, jitter=0.25, n_world=700, cull_every=3)
    o, calls = stream_utils.oracle_driver(oracle, st)
    segs = {}
    prev = {}                     # cam -> (off, rec) filtered lists after the previous cycle
    stats = dict(delta=0, fresh=0, adds=0, dels=0, dropped=0)

    def geometry(cam, seg, d1, d2, tgt_cam, cams, kk):
        M, Cc = cams[cam]
        s = segs[cam][seg]
        r1 = _normalized(_matvec(M, np.array([float(s[0]), float(s[1]), 1.0])))
        r2 = _normalized(_matvec(M, np.array([float(s[2]), float(s[3]), 1.0])))
        P1, P2 = Cc + r1 * float(d1), Cc + r2 * float(d2)
        d = P1 - P2
        length = f32(np.sqrt(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]))
        if length > 1e-12:
            dirv = _normalized(P2 - P1)
        else:
            P1 = P2 = dirv = np.zeros(3)
            length = f32(0.0)
        k = f32(kk[cam])
        sig1, sig2 = f32(d1) * k, f32(d2) * k
        reg1, reg2 = f32(2.0) * sig1 * sig1, f32(2.0) * sig2 * sig2
        Ct, kt = cams[tgt_cam][1], f32(kk[tgt_cam])
        e1, e2 = P1 - Ct, P2 - Ct
        s1t = f32(np.sqrt(e1[0] * e1[0] + e1[1] * e1[1] + e1[2] * e1[2]) * float(kt))
        s2t = f32(np.sqrt(e2[0] * e2[0] + e2[1] * e2[1] + e2[2] * e2[2]) * float(kt))
        reg1 = f32(0.5) * (reg1 + f32(2.0) * s1t * s1t)
        reg2 = f32(0.5) * (reg2 + f32(2.0) * s2t * s2t)
        mid = _normalized(_matvec(M, np.array([0.5 * (float(s[0]) + float(s[2])), 0.5 * (float(s[1]) + float(s[3])), 1.0])))
        ang = L.orc_kat_acos(min(max(float(mid[0] * dirv[0] + mid[1] * dirv[1] + mid[2] * dirv[2]), -1.0), 1.0))
        oriented = float(f32(0.098174771)) < ang < float(f32(3.043417886))
        return length, dirv, reg1, reg2, oriented

    def sim(gm, m, g2, m2):
        if gm[0] < 1e-12 or g2[0] < 1e-12:
            return f32(0.0)
        d1, d2 = f32(m["d_p1"]) - f32(m2["d_p1"]), f32(m["d_p2"]) - f32(m2["d_p2"])
        sim_p = min(expf(-d1 * d1 / gm[2]), expf(-d2 * d2 / gm[3]))
        dot_p = f32(gm[1][0] * g2[1][0] + gm[1][1] * g2[1][1] + gm[1][2] * g2[1][2])
        ang = f32(float(acosf(f32(max(min(dot_p, f32(1.0)), f32(-1.0))))) / np.pi * 180.0)
        if ang > f32(90.0):
            ang = f32(180.0) - ang
        s = min(expf(-ang * ang / f32(200.0)), sim_p)
        return s if s > f32(0.5) else f32(0.0)

    with np.errstate(over="ignore", under="ignore"):
        for ci, cy in enumerate(st.cycles):
            calls["begin_cycle"]()
            for cam in cy.deletes:
                calls["delete"](cam)
            for v in cy.adds:
                calls["add"](v, v.worldpoints)
                segs[v.cam_id] = v.segs
            for cam, R, t, md, lst in cy.updates:
                calls["update"](cam, R, t, md, lst)
            calls["match"](st.params)
            calls["reconstruct"]()
            current = [u[0] for u in cy.updates]
            add_cams, del_cams = {v.cam_id for v in cy.adds}, set(cy.deletes)
            cams = {c: o.match_camera(c) for c in current}
            for c in segs:                                     # deleted views: last pose, never translated again
                if c not in cams:
                    cams[c] = (o.match_camera(c)[0], o.view_info(c)["C"])
            kk = {c: o.view_info(c)["k"] for c in segs}
            for cam in current:
                if cam not in prev:
                    continue                                   # a new view: only fresh matches (checked elsewhere)
                off, rec = o.lists(cam, 0)                     # as scored, entries to deleted cameras removed
                poff, prec = prev[cam]
                for r in range(len(segs[cam])):
                    S = rec[off[r]:off[r + 1]]
                    P = prec[poff[r]:poff[r + 1]]
                    old = {(int(x["tgt_cam"]), int(x["tgt_seg"])): x for x in P}
                    if len(S) == 0 and len(P) == 0:
                        continue
                    # the list at scoring time: the snapshot plus the persisted entries to cameras deleted now
                    T = [(x, geometry(cam, r, x["d_p1"], x["d_p2"], int(x["tgt_cam"]), cams, kk)) for x in S]
                    for x in P:
                        if int(x["tgt_cam"]) in del_cams:
                            g = geometry(cam, r, x["d_p1"], x["d_p2"], int(x["tgt_cam"]), cams, kk)
                            if x["flags"] == 1 or g[4]:
                                T.append((x, g))
                            else:
                                stats["dropped"] += 1
                    for x, g in T[:len(S)]:
                        key = (int(x["tgt_cam"]), int(x["tgt_seg"]))
                        if key in old:                         # scored before: add / delete deltas
                            score = f32(old[key]["score"])
                            for bit, cset in (("adds", add_cams), ("dels", del_cams)):
                                for c2 in sorted(cset):
                                    if c2 == key[0]:
                                        continue
                                    sims = [sim(g, x, g2, x2) for x2, g2 in T if int(x2["tgt_cam"]) == c2]
                                    if sims:
                                        score = score + max(sims) if bit == "adds" else score - max(sims)
                                        stats[bit] += 1
                            stats["delta"] += 1
                        else:                                  # new this cycle: from scratch, deleted siblings excluded
                            score, per_cam = f32(0.0), {}
                            for x2, g2 in T[:len(S)]:
                                c2 = int(x2["tgt_cam"])
                                if c2 == key[0]:
                                    continue
                                sv = sim(g, x, g2, x2)
                                if c2 in per_cam:
                                    if sv > per_cam[c2]:
                                        score = score - per_cam[c2]
                                        score = score + sv
                                        per_cam[c2] = sv
                                else:
                                    score = score + sv
                                    per_cam[c2] = sv
                            stats["fresh"] += 1
                        assert f32(score).tobytes() == f32(x["score"]).tobytes(), (ci, cam, r, key, score, x["score"])
            prev = {cam: o.lists(cam, 1) for cam in current}
    assert stats["delta"] > 500 and stats["fresh"] > 500 and stats["adds"] > 100 and stats["dels"] > 20, stats
    o.close()


def test_retriangulated_hypotheses_second_reading_equals_the_oracle(oracle, scene_mod):
    """Line3D::update_Matches_and_Estimated_position3D (src/line3D.cc:1857-1908) on a key-frame stream whose poses
    move: every row of estimated_position3D_ carries the depths of its best match triangulated AGAIN with the
    cycle's poses (triangulationDepths both ways), and the 3-D segment unprojected from them; rows whose new depths
    are not all positive are gone."""
    import stream_utils
    st = scene_mod.make_stream(n_keyframes=9, n_seg=220, window=6, nbrs=4, jitter=0.6, n_world=700, cull_every=0)
    o, calls = stream_utils.oracle_driver(oracle, st)
    segs = {}
    checked = changed = 0

    def rays(cam, seg, cams):
        M = cams[cam][0]
        s = segs[cam][seg]
        return (_normalized(_matvec(M, np.array([float(s[0]), float(s[1]), 1.0]))),
                _normalized(_matvec(M, np.array([float(s[2]), float(s[3]), 1.0]))))

    def tri(Ca, ra1, ra2, Cb, rb1, rb2):
        n = _normalized(_cross(rb1, rb2))
        a, b = _dot3(ra1, n), _dot3(ra2, n)
        if abs(a) < EPS or abs(b) < EPS:
            return -1.0, -1.0
        num = _dot3(Cb, n) - _dot3(n, Ca)
        return num / a, num / b
    for ci, cy in enumerate(st.cycles):
        calls["begin_cycle"]()
        for cam in cy.deletes:
            calls["delete"](cam)
        for v in cy.adds:
            calls["add"](v, v.worldpoints)
            segs[v.cam_id] = v.segs
        for cam, R, t, md, lst in cy.updates:
            calls["update"](cam, R, t, md, lst)
        calls["match"](st.params)
        cams = {u[0]: o.match_camera(u[0]) for u in cy.updates}
        filt = {u[0]: o.lists(u[0], 1) for u in cy.updates}
        for E in o.entries():
            cam, seg, tc, ts = int(E["src_cam"]), int(E["src_seg"]), int(E["tgt_cam"]), int(E["tgt_seg"])
            rs1, rs2 = rays(cam, seg, cams)
            rt1, rt2 = rays(tc, ts, cams)
            ds1, ds2 = tri(cams[cam][1], rs1, rs2, cams[tc][1], rt1, rt2)
            dt1, dt2 = tri(cams[tc][1], rt1, rt2, cams[cam][1], rs1, rs2)
            assert ds1 > EPS and ds2 > EPS and dt1 > EPS and dt2 > EPS
            for name, val in (("d_p1", ds1), ("d_p2", ds2), ("d_q1", dt1), ("d_q2", dt2)):
                assert np.float32(E[name]).tobytes() == np.float32(val).tobytes(), (ci, cam, seg, name)
            P1 = cams[cam][1] + rs1 * float(np.float32(ds1))
            P2 = cams[cam][1] + rs2 * float(np.float32(ds2))
            assert (E["P1"] == P1).all() and (E["P2"] == P2).all() and (E["dir"] == _normalized(P2 - P1)).all()
            off, rec = filt[cam]                                   # the match itself keeps the depths it was made with
            m = [x for x in rec[off[seg]:off[seg + 1]] if int(x["tgt_cam"]) == tc and int(x["tgt_seg"]) == ts]
            assert len(m) == 1 and m[0]["score"] == E["score"]
            changed += int(m[0]["d_p1"] != E["d_p1"])
            checked += 1
    assert checked > 400 and changed > 100            # later cycles: old matches, new poses -> new depths
    o.close()


def test_spatial_regulariser_second_reading_equals_the_oracle(oracle, scene_mod):
    """View::computeSpatialRegularizer (src/view.cc:330-343): k = sin(angle between the rays through the principal
    point and through the point sigma_p pixels beside it)."""
    import ctypes as C
    L = oracle.lib()
    L.orc_kat_acos.restype = C.c_double
    L.orc_kat_acos.argtypes = [C.c_double]
    L.orc_kat_sin.restype = C.c_double
    L.orc_kat_sin.argtypes = [C.c_double]
    for name, sig in (("tiny", 5.0), ("tiny", 2.5)):
        sc = scene_mod.make_scene(name)
        sc.params = dict(sc.params, sigma_p=sig)
        o = oracle.run_scene(sc)
        for v in sc.views:
            M, _ = o.match_camera(v.cam_id)
            pp = np.array([v.K[0, 2], v.K[1, 2], 1.0])
            pps = np.array([pp[0] + float(np.float32(sig)), pp[1] + 0.0, pp[2] + 0.0])
            a, b = _normalized(_matvec(M, pp)), _normalized(_matvec(M, pps))
            alpha = L.orc_kat_acos(min(max(float(_dot3(a, b)), -1.0), 1.0))
            k = np.float32(L.orc_kat_sin(alpha))
            assert k.tobytes() == np.float32(o.view_info(v.cam_id)["k"]).tobytes() and 0.001 < k < 0.02
        o.close()
