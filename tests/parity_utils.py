"""Shared comparison helpers: product (CUDA, through the C ABI) versus the CPU oracle."""
import numpy as np


def assert_struct_equal(a, b, what):
    assert a.shape == b.shape, "%s: %s vs %s records" % (what, a.shape, b.shape)
    for name in a.dtype.names:
        if name == "pad":
            continue
        x, y = a[name], b[name]
        if x.dtype.kind == "f":
            same = (x.view(np.uint32 if x.dtype.itemsize == 4 else np.uint64) ==
                    y.view(np.uint32 if y.dtype.itemsize == 4 else np.uint64))
        else:
            same = x == y
        if not np.all(same):
            idx = np.argwhere(~same)[0]
            raise AssertionError("%s: field %s differs at %s: %r vs %r (%d of %d differ)" %
                                 (what, name, idx, x[tuple(idx)], y[tuple(idx)], int((~same).sum()), same.size))


def compare_full(l3, orc, scene, check_scored=True):
    """Bit-exact comparison of every stage output. Returns a dict of sizes."""
    gp, op = l3.pairs(), orc.pairs()
    assert gp.shape == op.shape and (gp == op).all(), "matched pair list differs"
    n_scored = n_filt = 0
    for v in scene.views:
        if check_scored:
            go, gr = l3.lists(v.cam_id, 0)
            oo, orr = orc.lists(v.cam_id, 0)
            assert (go == oo).all(), "view %d: scored row offsets differ" % v.cam_id
            assert_struct_equal(gr, orr, "view %d scored lists" % v.cam_id)
            n_scored += len(gr)
        go, gr = l3.lists(v.cam_id, 1)
        oo, orr = orc.lists(v.cam_id, 1)
        assert (go == oo).all(), "view %d: filtered row offsets differ" % v.cam_id
        assert_struct_equal(gr, orr, "view %d filtered lists" % v.cam_id)
        n_filt += len(gr)
        gi, oi = l3.view_info(v.cam_id), orc.view_info(v.cam_id)
        assert np.float32(gi["k"]).tobytes() == np.float32(oi["k"]).tobytes(), "view %d: k differs" % v.cam_id
        assert np.float32(gi["median_depth"]).tobytes() == np.float32(oi["median_depth"]).tobytes(), \
            "view %d: median depth %r vs %r" % (v.cam_id, gi["median_depth"], oi["median_depth"])
        assert (gi["C"] == oi["C"]).all(), "view %d: camera centre differs" % v.cam_id
    assert_struct_equal(l3.entries(), orc.entries(), "estimated_position3D_")
    assert np.float32(l3.med_scene_depth_lines()).tobytes() == np.float32(orc.med_scene_depth_lines()).tobytes()
    gij, gw = l3.edges()
    oij, ow = orc.edges()
    assert gij.shape == oij.shape, "edge count %s vs %s" % (gij.shape, oij.shape)
    assert (gij == oij).all(), "edge ids differ"
    assert (gw.view(np.uint32) == ow.view(np.uint32)).all(), "edge weights differ"
    assert (l3.local2global() == orc.local2global()).all(), "local2global differs"
    gc, oc = l3.cluster_ids(), orc.cluster_ids()
    assert gc.shape == oc.shape and (gc == oc).all(), "cluster ids differ"
    return dict(pairs=len(gp), scored=n_scored, filtered=n_filt, entries=len(l3.entries()), edges=len(gw),
                local=len(gc), clusters=len(set(gc.tolist())))
