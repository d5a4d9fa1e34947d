"""The CPU restatement (oracle/l3d_oracle.cpp) against the REFERENCE'S OWN Line3D++ SOURCES -- src/line3D.cc,
src/view.cc, src/clustering.cc and their headers, compiled unmodified from /root/reference into
oracle/_ref/libref_line3d_{det,libm}.so (oracle/Makefile target `ref`; oracle/ref_line3d_wrap.cpp;
oracle/standin/ supplies the part of Eigen / Boost / OpenCV they touch, none of which is installed here).

  * det build (expf / acos / sin routed to oracle/detmath.h): every result of the path is compared BIT FOR BIT --
    filtered match lists of every view (targets, overlaps, depths, scores, order), estimated_position3D_, A_
    (pairs, weights, order), local2global_, cluster roots, k, median depths -- on batch scenes with explicit
    neighbour lists, on the world-point (.nvm) path, and cycle by cycle on a key-frame stream with deleted views
    (the incremental add / delete scoring branches, Line3D::scoringCPU src/line3D.cc:1439-1512).
  * libm build (glibc, as the reference links it): identical match sets and cluster roots, scores and 3-D
    endpoints within the north star's 1e-4 relative.

This is what pins the oracle: what runs on the other side is the reference's control flow, thresholds,
containers and list handling, not a reading of them."""
import importlib

import numpy as np
import pytest

import stream_utils


def _need(oracle, kind):
    if oracle.ref_lib(kind) is None:
        pytest.skip("oracle/_ref/libref_line3d_%s.so is not built (needs /root/reference)" % kind)


def _struct_equal(a, b, what):
    assert a.shape == b.shape, (what, a.shape, b.shape)
    for f in a.dtype.names:
        if f == "pad":
            continue
        assert (np.ascontiguousarray(a[f]).view(np.uint8) == np.ascontiguousarray(b[f]).view(np.uint8)).all(), (what, f)


def _compare_exact(o, r, cams, where=""):
    n = 0
    for cam in cams:
        oo, orec = o.lists(cam, 1)
        ro, rrec = r.lists(cam, 1)
        assert (oo == ro).all(), (where, cam, "row offsets")
        _struct_equal(orec, rrec, "%s view %d filtered list" % (where, cam))
        n += len(orec)
        a, b = o.view_info(cam), r.view_info(cam)
        for k in ("k", "median_depth", "median_sigma"):
            assert np.float32(a[k]).tobytes() == np.float32(b[k]).tobytes(), (where, cam, k)
        assert (a["C"] == b["C"]).all(), (where, cam, "C")
        assert sorted(o.neighbors(cam)) == sorted(r.neighbors(cam)), (where, cam, "neighbours")
    _struct_equal(o.entries(), r.entries(), where + " estimated_position3D_")
    ij1, w1 = o.edges()
    ij2, w2 = r.edges()
    assert ij1.shape == ij2.shape and (ij1 == ij2).all(), where + " A_ pairs / order"
    assert w1.tobytes() == w2.tobytes(), where + " A_ weights"
    assert o.local2global().tobytes() == r.local2global().tobytes(), where + " local2global_"
    assert (o.cluster_ids() == r.cluster_ids()).all(), where + " cluster roots"
    assert np.float32(o.med_scene_depth_lines()).tobytes() == np.float32(r.med_scene_depth_lines()).tobytes()
    assert o.pair_tests() == r.pair_tests()
    assert o.pairs().shape == r.pairs().shape and (o.pairs() == r.pairs()).all(), where + " matched pairs / order"
    return dict(filtered=n, entries=len(o.entries()), edges=len(w1), local=len(o.local2global()))


@pytest.mark.parametrize("kind,kw", [("tiny", {}), ("c2", dict(n_views=14, n_seg=400)), ("c4", dict(n_views=6, n_seg=900))])
def test_restatement_equals_the_compiled_reference_bit_for_bit(oracle, scene_mod, kind, kw):
    _need(oracle, "det")
    scene = scene_mod.make_scene(kind, **kw)
    o = oracle.run_scene(scene)
    r = oracle.run_scene_ref(scene, "det")
    sizes = _compare_exact(o, r, [v.cam_id for v in scene.views], kind)
    assert sizes["entries"] > 50 and sizes["edges"] > 50 and sizes["filtered"] > 200
    o.close()
    r.close()


def test_world_point_neighbours_path_equals_the_compiled_reference(oracle):
    """neighbors_by_worldpoints = true (the .nvm input): Line3D::findVisualNeighborsFromWPs (src/line3D.cc:723-843)
    on the reference's own NVM dump (tests/golden/c1_nvm_wps.npz), then the whole path."""
    _need(oracle, "det")
    import golden_utils
    import os
    scene = golden_utils.load_scene("c1_nvm_scene.npz")
    wps = np.load(os.path.join(golden_utils.HERE, "c1_nvm_wps.npz"))
    for i, v in enumerate(scene.views):
        v.worldpoints = wps["wps_%d" % i].tolist() if "wps_%d" % i in wps else wps["wps_%d" % v.cam_id].tolist()
    scene.neighbors_by_worldpoints = True
    o = oracle.run_scene(scene)
    r = oracle.run_scene_ref(scene, "det")
    sizes = _compare_exact(o, r, [v.cam_id for v in scene.views], "c1 nvm")
    assert sizes["entries"] > 100
    o.close()
    r.close()


def test_key_frame_stream_equals_the_compiled_reference_cycle_by_cycle(oracle, scene_mod):
    """What L3DPPing::Run does (src/L3DPPing.cpp:98-236): delete culled key frames, add new ones, re-pose every current
    one, matchImages, reconstruct3Dlines -- 14 cycles with a sliding window, so both scoringCPU branches, updateMatch
    and update_Matches_and_Estimated_position3D run in the reference's own code."""
    _need(oracle, "det")
    st = scene_mod.make_stream(n_keyframes=18, n_seg=220, window=6, nbrs=4, jitter=0.3, n_world=700)
    o, oc = stream_utils.oracle_driver(oracle, st)
    r = oracle.RefLine3D(st.max_image_width, st.neighbors_by_worldpoints, "det")
    rc = dict(begin_cycle=r.begin_cycle, delete=r.delete_image,
              add=lambda v, lst: r.add_image(v.cam_id, v.K, v.R, v.t, v.width, v.height, v.median_depth, lst, v.segs),
              update=r.update_image,
              match=lambda p: r.match_images(p["sigma_p"], p["sigma_a"], p["num_neighbors"], p["epipolar_overlap"], p["knn"],
                                             p["const_reg_depth"]),
              reconstruct=r.reconstruct)
    deleted = 0
    for ci, cy in enumerate(st.cycles):
        for calls in (oc, rc):
            calls["begin_cycle"]()
            for cam in cy.deletes:
                calls["delete"](cam)
            for v in cy.adds:
                calls["add"](v, v.worldpoints if st.neighbors_by_worldpoints else v.neighbors)
            for cam, R, t, md, lst in cy.updates:
                calls["update"](cam, R, t, md, lst)
            calls["match"](st.params)
            calls["reconstruct"]()
        deleted += len(cy.deletes)
        _compare_exact(o, r, [u[0] for u in cy.updates], "cycle %d" % ci)
    assert deleted >= 8 and len(st.cycles) >= 12
    o.close()
    r.close()


def test_glibc_build_same_match_sets_scores_within_1e4(oracle, scene_mod):
    """The reference linked against glibc's libm, as its own build would: the deterministic replacements change
    no decision on this scene (same matches, same hypotheses, same cluster partition) and the values agree to the
    north star's tolerance."""
    _need(oracle, "libm")
    scene = scene_mod.make_scene("c2", n_views=14, n_seg=400)
    o = oracle.run_scene(scene)
    r = oracle.run_scene_ref(scene, "libm")
    for v in scene.views:
        oo, orec = o.lists(v.cam_id, 1)
        ro, rrec = r.lists(v.cam_id, 1)
        assert (oo == ro).all()
        for f in ("tgt_cam", "tgt_seg", "flags"):
            assert (orec[f] == rrec[f]).all(), (v.cam_id, f)
        for f in ("overlap", "d_p1", "d_p2", "d_q1", "d_q2"):     # no libm call on these
            assert (orec[f].view(np.uint32) == rrec[f].view(np.uint32)).all(), (v.cam_id, f)
        assert np.allclose(orec["score"], rrec["score"], rtol=1e-4, atol=0)
    eo, er = o.entries(), r.entries()
    for f in ("src_cam", "src_seg", "tgt_cam", "tgt_seg"):
        assert (eo[f] == er[f]).all()
    assert np.allclose(eo["score"], er["score"], rtol=1e-4) and np.allclose(eo["P1"], er["P1"], rtol=1e-4, atol=1e-9)
    ij1, w1 = o.edges()
    ij2, w2 = r.edges()
    assert (ij1 == ij2).all() and np.allclose(w1, w2, rtol=1e-4)
    assert (o.local2global() == r.local2global()).all() and (o.cluster_ids() == r.cluster_ids()).all()
    o.close()
    r.close()
