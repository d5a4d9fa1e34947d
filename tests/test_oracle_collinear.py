"""CPU: the oracle's restatement of the collinearity-extended affinity stage (src/line3D.cc:2328-2396,
src/view.cc:180-318).  The reference runs with collinearity = -1 (include/L3DPPing.h:82) and the CUDA
product builds only that case; this is the checker the extended stage will be built against."""
import numpy as np


def _split_scene(scene_mod):
    """The tiny scene with the 60 longest segments of every view split into two collinear fragments."""
    sc = scene_mod.make_scene("tiny", n_seg=200)
    for v in sc.views:
        s = v.segs
        a, b = s[:60, :2], s[:60, 2:]
        f1 = np.concatenate([a, a + 0.45 * (b - a)], axis=1)
        f2 = np.concatenate([a + 0.55 * (b - a), b], axis=1)
        v.segs = np.ascontiguousarray(np.concatenate([f1, f2, s[60:]]).astype(np.float32))
    return sc


def _pairs(o):
    ij, w = o.edges()
    g = [tuple(x) for x in o.local2global().tolist()]
    assert len(w) % 2 == 0 and (ij[0::2, 0] == ij[1::2, 1]).all() and (ij[0::2, 1] == ij[1::2, 0]).all()
    assert (w[0::2].view(np.uint32) == w[1::2].view(np.uint32)).all() and (w > np.float32(0.5)).all()
    return {frozenset((g[i], g[j])): float(x) for (i, j), x in zip(ij[0::2].tolist(), w[0::2].tolist())}


def test_collinear_links_extend_the_affinity_matrix(oracle, scene_mod):
    sc = _split_scene(scene_mod)
    o = oracle.run_scene(sc)                       # collinearity off
    off = _pairs(o)
    lists = {v.cam_id: o.lists(v.cam_id, 1) for v in sc.views}
    o.reconstruct(2.0)                             # same matches, links to collinear segments added
    on = _pairs(o)
    assert set(off) <= set(on) and len(on) > len(off) + 50
    assert all(on[p] == off[p] for p in off)       # the similarity of a pair does not depend on the route
    coll = {v.cam_id: oracle.find_collinear(v.segs, 2.0) for v in sc.views}

    def targets(seg):
        off_, rec = lists[seg[0]]
        r = rec[off_[seg[1]]:off_[seg[1] + 1]]
        return {(int(c), int(s)) for c, s in zip(r["tgt_cam"], r["tgt_seg"])}
    for p in set(on) - set(off):
        a, b = tuple(p)
        ok = False
        for x, y in ((a, b), (b, a)):
            if x[0] == y[0] and coll[x[0]][x[1], y[1]]:                    # collinear to the entry itself
                ok = True
            if any(t[0] == y[0] and coll[t[0]][t[1], y[1]] for t in targets(x)):   # collinear to a direct target
                ok = True
        assert ok, (a, b)
    n_off = len(set(o.cluster_ids().tolist()))
    o.reconstruct(-1.0)                            # back to the reference configuration: identical to the first run
    assert _pairs(o) == off
    assert len(set(o.cluster_ids().tolist())) >= n_off
    o.close()


def test_collinear_table_is_kept_until_the_threshold_changes(oracle, scene_mod):
    """View::findCollinearSegments returns early when the threshold is unchanged (src/view.cc:182-186), and
    Line3D::reconstruct3Dlines only calls it when the threshold changed (src/line3D.cc:2068-2072)."""
    sc = _split_scene(scene_mod)
    o = oracle.run_scene(sc)
    o.reconstruct(2.0)
    a = _pairs(o)
    o.reconstruct(2.0)
    assert _pairs(o) == a
    o.reconstruct(0.5)                             # tighter threshold: fewer collinear links
    assert len(_pairs(o)) <= len(a)
    o.close()
