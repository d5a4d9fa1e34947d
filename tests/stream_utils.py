"""Drivers that replay a key-frame stream (scene.make_stream) through the oracle and through the
CUDA path, and the per-cycle comparison of everything the path produces."""
import importlib

import numpy as np

scene_mod = importlib.import_module("3dline-slam_b200.scene")


def oracle_driver(oracle, stream, threads=0):
    o = oracle.OracleLine3D(stream.max_image_width, stream.neighbors_by_worldpoints, threads)

    def match(p):
        o.match_images(p["sigma_p"], p["sigma_a"], p["num_neighbors"], p["epipolar_overlap"], p["knn"],
                       p["const_reg_depth"])

    calls = dict(begin_cycle=o.begin_cycle, delete=o.delete_image,
                 add=lambda v, lst: o.add_image(v.cam_id, v.K, v.R, v.t, v.width, v.height, v.median_depth, lst, v.segs),
                 update=o.update_image, match=match, reconstruct=o.reconstruct)
    return o, calls


def cuda_driver(api, stream):
    l3 = api.Line3DStream("", False, stream.max_image_width, 3000, stream.neighbors_by_worldpoints, True)

    def match(p):
        l3.matchImages(p["sigma_p"], p["sigma_a"], p["num_neighbors"], p["epipolar_overlap"], p["knn"],
                       p["const_reg_depth"])

    calls = dict(begin_cycle=l3.beginCycle, delete=l3.deleteImage,
                 add=lambda v, lst: l3.addImage(v.cam_id, (v.width, v.height), v.K, v.R, v.t, v.median_depth, lst, v.segs),
                 update=l3.UpdataImage, match=match, reconstruct=l3.reconstruct3Dlines)
    return l3, calls


def snapshot(obj, cams, is_oracle):
    """Everything a cycle produced, as plain arrays (same layout for both implementations)."""
    out = {}
    for cam in cams:
        for which in (0, 1):
            off, rec = obj.lists(cam, which)
            out["off_%d_%d" % (cam, which)] = np.asarray(off).copy()
            out["rec_%d_%d" % (cam, which)] = np.asarray(rec).copy()
        out["nb_%d" % cam] = np.asarray(sorted(obj.neighbors(cam)), dtype=np.int64)
        vi = obj.view_info(cam)
        out["vi_%d" % cam] = np.array([vi["k"], vi["median_depth"]], dtype=np.float32)
    out["entries"] = obj.entries().copy()
    ij, w = obj.edges()
    out["edges_ij"], out["edges_w"] = np.asarray(ij).copy(), np.asarray(w).copy()
    out["l2g"] = np.asarray(obj.local2global()).copy()
    out["clusters"] = np.asarray(obj.cluster_ids()).copy()
    return out


def assert_same(a, b, where=""):
    assert a.keys() == b.keys()
    for k in a:
        x, y = a[k], b[k]
        assert x.shape == y.shape, (where, k, x.shape, y.shape)
        if x.dtype.names:
            for f in x.dtype.names:
                if f == "pad":
                    continue
                xv, yv = x[f], y[f]
                same = (xv.view(np.uint8) == yv.view(np.uint8)).all() if xv.size else True
                assert same, (where, k, f, np.nonzero(xv != yv)[0][:5])
        else:
            assert (x.view(np.uint8) == y.view(np.uint8)).all(), (where, k)


def run_lockstep(api, oracle, stream, max_cycles=None, check_scored=True):
    """Replay the stream through both implementations; after EVERY cycle compare, bit for bit, the
    cycle's new pairs, the neighbour sets, every current view's scored and filtered lists, k and
    median depth, estimated_position3D_, A_, the local ids and the cluster ids.  Returns totals."""
    from parity_utils import assert_struct_equal
    o, oc = oracle_driver(oracle, stream)
    l3, gc = cuda_driver(api, stream)
    tot = dict(cycles=0, pairs=0, scored=0, filtered=0, entries=0, edges=0, deleted=0, tests=0)
    seen_pairs = 0
    for ci, cy in enumerate(stream.cycles[:max_cycles]):
        for calls in (oc, gc):
            calls["begin_cycle"]()
            for cam in cy.deletes:
                calls["delete"](cam)
            for v in cy.adds:
                calls["add"](v, v.worldpoints if stream.neighbors_by_worldpoints else v.neighbors)
            for cam, R, t, md, lst in cy.updates:
                calls["update"](cam, R, t, md, lst)
            calls["match"](stream.params)
            calls["reconstruct"]()
        where = "cycle %d" % ci
        op = o.pairs()[seen_pairs:]
        gp = l3.pairs()
        assert gp.shape == op.shape and (gp == op).all(), where + ": new pair list differs"
        seen_pairs += len(op)
        for cam, *_ in cy.updates:
            assert l3.neighbors(cam) == sorted(o.neighbors(cam)), (where, cam)
            for which in ((0, 1) if check_scored else (1,)):
                go, gr = l3.lists(cam, which)
                oo, orr = o.lists(cam, which)
                assert (go == oo).all(), "%s view %d: row offsets of list %d differ" % (where, cam, which)
                assert_struct_equal(gr, orr, "%s view %d list %d" % (where, cam, which))
                tot["scored" if which == 0 else "filtered"] += len(gr)
            gi, oi = l3.view_info(cam), o.view_info(cam)
            assert np.float32(gi["k"]).tobytes() == np.float32(oi["k"]).tobytes(), (where, cam, "k")
            assert np.float32(gi["median_depth"]).tobytes() == np.float32(oi["median_depth"]).tobytes(), \
                (where, cam, gi["median_depth"], oi["median_depth"])
            assert (gi["C"] == oi["C"]).all(), (where, cam, "C")
        assert_struct_equal(l3.entries(), o.entries(), where + " estimated_position3D_")
        gij, gw = l3.edges()
        oij, ow = o.edges()
        assert gij.shape == oij.shape, "%s: edge count %s vs %s" % (where, gij.shape, oij.shape)
        assert (gij == oij).all() and (gw.view(np.uint32) == ow.view(np.uint32)).all(), where + ": A_ differs"
        assert (l3.local2global() == o.local2global()).all(), where + ": local2global differs"
        g_ids, o_ids = l3.cluster_ids(), o.cluster_ids()
        assert g_ids.shape == o_ids.shape and (g_ids == o_ids).all(), where + ": cluster ids differ"
        tot["cycles"] += 1
        tot["pairs"] += len(gp)
        tot["entries"] += len(l3.entries())
        tot["edges"] += len(gw)
        tot["deleted"] += len(cy.deletes)
        tot["tests"] += l3.counts()["pair_tests"]
    o.close()
    return tot


GOLDEN_STREAM = dict(n_keyframes=14, n_seg=600, window=10, nbrs=6, jitter=0.2, n_world=1200, cull_every=3)


def cycle_digest(obj, cams):
    """sha256 over everything a cycle produced (canonical field order, independent of who ran it)."""
    import hashlib
    import oracle_py
    h = hashlib.sha256()

    def canon(a, dt):
        c = np.zeros(len(a), dtype=dt)
        for f in dt.names:
            c[f] = a[f]
        return c.tobytes()
    for cam in cams:
        off, rec = obj.lists(cam, 1)
        h.update(np.asarray(off, dtype=np.uint32).tobytes())
        h.update(canon(rec, oracle_py.REC_DTYPE))
        vi = obj.view_info(cam)
        h.update(np.float32(vi["k"]).tobytes() + np.float32(vi["median_depth"]).tobytes())
        h.update(np.asarray(sorted(obj.neighbors(cam)), dtype=np.uint32).tobytes())
    h.update(canon(obj.entries(), oracle_py.ENTRY_DTYPE))
    ij, w = obj.edges()
    h.update(np.asarray(ij, dtype=np.int32).tobytes() + np.asarray(w, dtype=np.float32).tobytes())
    h.update(np.asarray(obj.local2global(), dtype=np.uint32).tobytes())
    h.update(np.asarray(obj.cluster_ids(), dtype=np.int32).tobytes())
    return np.frombuffer(h.digest(), dtype=np.uint8).copy()


def stream_digests(stream, obj, calls):
    out, sizes = [], []

    def on_cycle(ci, cy):
        out.append(cycle_digest(obj, [u[0] for u in cy.updates]))
        sizes.append((len(obj.entries()), len(obj.edges()[1]), len(obj.cluster_ids())))
    scene_mod.drive_stream(stream, on_cycle=on_cycle, **calls)
    return np.stack(out), np.asarray(sizes, dtype=np.int64)
