"""GPU parity: the CUDA path (through the C ABI) against the CPU oracle on the same seeded scenes.
Bit-exact on every stage: match set, depths, scores, hypotheses, affinity edges, cluster IDs."""
import numpy as np
import pytest

from parity_utils import assert_struct_equal, compare_full

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("kind,kw", [("tiny", {}), ("tiny", dict(seed=77, n_views=6, n_seg=97, nbrs=3)),
                                     ("c2", dict(n_views=12, n_seg=300, nbrs=6, n_world=250))])
def test_full_pipeline_bit_exact(api, oracle, scene_mod, kind, kw):
    sc = scene_mod.make_scene(kind, **kw)
    orc = oracle.run_scene(sc)
    l3 = api.run_scene(sc, filter_mode=0, keep_scored=True)
    sizes = compare_full(l3, orc, sc)
    assert sizes["scored"] > 0 and sizes["entries"] > 0 and sizes["edges"] > 0 and sizes["clusters"] > 1, sizes
    c = l3.counts()
    assert c["pair_tests"] == orc.pair_tests()
    assert c["candidates"] < 0.2 * c["pair_tests"], "the FP32 pre-filter is not selective: %r" % c
    assert c["gpu_launches"] > 0


def test_prefilter_is_conservative(api, oracle, scene_mod):
    """filter_mode=1 sends every pair through the exact kernel: same results as with the filter."""
    sc = scene_mod.make_scene("tiny", seed=1234)
    orc = oracle.run_scene(sc)
    a = api.run_scene(sc, filter_mode=1, keep_scored=True)
    compare_full(a, orc, sc)
    assert a.counts()["candidates"] == a.counts()["pair_tests"]


def test_long_rows_use_the_global_staging_path(api, oracle, scene_mod, monkeypatch):
    """Rows whose potential list does not fit the shared-memory staging (forced here) give the same result."""
    monkeypatch.setenv("L3D_K3_MAXM", "6")
    sc = scene_mod.make_scene("tiny", seed=5, n_views=6, n_seg=150, nbrs=4)
    orc = oracle.run_scene(sc)
    l3 = api.run_scene(sc, keep_scored=True)
    compare_full(l3, orc, sc)


def test_knn_variants(api, oracle, scene_mod):
    for knn in (1, 3, -1):
        sc = scene_mod.make_scene("tiny", seed=4242, n_views=5, n_seg=120, nbrs=3)
        sc.params["knn"] = knn
        orc = oracle.run_scene(sc)
        l3 = api.run_scene(sc, keep_scored=True)
        compare_full(l3, orc, sc)


def test_rematch_with_other_parameters_on_the_same_context(api, oracle, scene_mod):
    """A context that is matched again with parameters that make the lists much longer (kNN 10 -> all matches):
    the cached longest-list size of the first run must not size the second one (it did: illegal memory access,
    found by tools/sanitize_run.py)."""
    sc = scene_mod.make_scene("tiny")
    l3 = api.run_scene(sc, keep_scored=True)
    for knn in (0, 3):
        sc.params["knn"] = knn
        p = sc.params
        l3.matchImages(p["sigma_p"], p["sigma_a"], p["num_neighbors"], p["epipolar_overlap"], p["knn"], p["const_reg_depth"])
        l3.reconstruct3Dlines()
        orc = oracle.run_scene(sc)
        compare_full(l3, orc, sc)
        orc.close()


@pytest.mark.parametrize("knn", [-1, 40, 10])
def test_dense_rows_take_the_rare_kernel_paths(api, oracle, scene_mod, knn):
    """Without the pre-filter every target is a candidate: more than 256 candidates per mask chunk
    (enumerated 8 words at a time), rows with more matches than the shared-memory selection holds
    (heap replay from the global staging), kNN above the warp width and kNN <= 0 (all matches)."""
    sc = scene_mod.make_scene("tiny", seed=808, n_views=5, n_seg=640, nbrs=3)
    sc.params["knn"] = knn
    orc = oracle.run_scene(sc)
    l3 = api.run_scene(sc, filter_mode=1, keep_scored=True)
    compare_full(l3, orc, sc)
    assert l3.counts()["candidates"] == l3.counts()["pair_tests"]


@pytest.mark.parametrize("variant", ["0", "1", "2"])
def test_large_target_views_in_both_k2_kernels(api, oracle, scene_mod, monkeypatch, variant):
    """More than 1024 segments per view: several mask chunks per row.  The row kernel is the default
    for such pairs; both kernels (forced through the tuning hook) must give the same result, and so
    must the row kernel on small views."""
    monkeypatch.setenv("L3D_K2_VARIANT", variant)
    sc = scene_mod.make_scene("tiny", seed=909, n_views=4, n_seg=1300, nbrs=3)
    orc = oracle.run_scene(sc)
    l3 = api.run_scene(sc, keep_scored=True)
    compare_full(l3, orc, sc)
    sc.params["knn"] = -1
    orc = oracle.run_scene(sc)
    l3 = api.run_scene(sc, filter_mode=1, keep_scored=True)
    compare_full(l3, orc, sc)


def test_ties_follow_priority_queue_order(api, oracle, scene_mod):
    """Duplicated target segments give exactly equal overlaps: the kNN pop order must be the
    std::priority_queue's (include/commons.h:233-244)."""
    sc = scene_mod.make_scene("tiny", seed=99, n_views=4, n_seg=80, nbrs=3)
    for v in sc.views:
        s = v.segs.copy()
        s[40:80] = s[0:40]  # every segment twice
        v.segs = s
    sc.params["knn"] = 4
    orc = oracle.run_scene(sc)
    l3 = api.run_scene(sc, keep_scored=True)
    compare_full(l3, orc, sc)


def test_device_math_bit_exact(api, oracle):
    ctx = api.Context()
    rng = np.random.default_rng(3)
    x = np.concatenate([rng.uniform(-20, 0.5, 200000), rng.uniform(-104, -80, 2000), [0.0, -0.0, -0.6931472, 1.0, 88.0,
                        -150.5, np.inf, -np.inf]]).astype(np.float32)
    got = ctx.test_expf(x)
    L = oracle.lib()
    exp = np.array([L.orc_kat_expf(float(v)) for v in x], dtype=np.float32)
    assert (got.view(np.uint32) == exp.view(np.uint32)).all()
    xd = np.concatenate([rng.uniform(-1, 1, 100000), [1.0, -1.0, 0.0, 0.5, -0.5, 0.9999999999, 1e-300]])
    gotd = ctx.test_acos(xd)
    expd = np.array([L.orc_kat_acos(float(v)) for v in xd])
    assert (gotd.view(np.uint64) == expd.view(np.uint64)).all()


@pytest.mark.parametrize("world", [2, 3])
def test_sharded_run_matches_the_oracle(api, oracle, scene_mod, world):
    """`world` shard contexts on one GPU (contiguous view slices): matching, scoring rows, hypotheses
    and edges are computed per slice and exchanged (host buffers); every rank ends with the result
    of the unsharded run = the oracle's, bit for bit."""
    import importlib
    shd = importlib.import_module("3dline-slam_b200.sharding")
    sc = scene_mod.make_scene("tiny", seed=31, n_views=9, n_seg=140, nbrs=4)
    shards = []
    for r in range(world):
        l3 = api.Line3D("", False, sc.max_image_width)
        l3.keep_scored = False
        l3.shard = (r, world)
        l3.load_scene(sc)
        shards.append(l3)
    import torch
    grp = shd.LocalGroup(shards, torch, torch.device("cuda", 0))
    grp.run(sc.params)          # first pass: size exchange + host buffers
    grp.run(sc.params)          # second pass: self-describing blobs in device buffers
    assert grp.fallbacks == 0
    grp.stride[shd.X_PROGRAMS] = 64   # far too small: every rank must detect it and fall back
    grp.run(sc.params)
    assert grp.fallbacks == 1
    orc = oracle.run_scene(sc)
    assert sum(s.counts()["pair_tests"] for s in shards) == orc.pair_tests()
    assert sum(s.counts()["num_pairs_local"] for s in shards) == len(orc.pairs())
    ref = api.run_scene(sc, reconstruct=False)
    for s in shards:
        s._ck(s.L.l3d_cluster(s.h))
        sizes = compare_full(s, orc, sc, check_scored=False)
        assert sizes["edges"] > 0 and sizes["clusters"] > 1
        for k in ("forward_matches", "scored_entries", "sim_evals", "filtered_entries", "num_entries"):
            assert s.counts()[k] == ref.counts()[k], k


def test_sparse_matrix_layout_matches_the_oracle(api, oracle, scene_mod):
    """A_ in the SparseMatrix layout (src/sparsematrix.cc:8-61), column- and row-sorted, with and
    without normalisation, against the oracle's restatement on the same edge list."""
    sc = scene_mod.make_scene("tiny")
    l3 = api.run_scene(sc)
    ij, w = l3.edges()
    n = len(l3.local2global())
    assert len(w) > 100
    for by_row, norm in ((False, 1.0), (True, 1.0), (False, 2.5)):
        ge, gs = l3.sparse_matrix(by_row, norm)
        oe, os_ = oracle.sparse_matrix(ij, w, n, norm, by_row)
        assert ge.tobytes() == oe.tobytes() and (gs == os_).all()
        key = ge[:, 0 if by_row else 1]
        assert (np.diff(key) >= 0).all()


def test_metric_regulariser_with_constant_depth(api, oracle, scene_mod):
    """sigma_p < 0: the spatial regulariser is given in world units and k = |sigma_p| / const_regularization_depth
    for every view (fixed3Dregularizer_, src/line3D.cc:525-530, 576-582); a negative depth is refused."""
    sc = scene_mod.make_scene("tiny")
    sc.params = dict(sc.params, sigma_p=-0.03, const_reg_depth=3.5)
    l3 = api.run_scene(sc, keep_scored=True)
    orc = oracle.run_scene(sc)
    sizes = compare_full(l3, orc, sc)
    assert sizes["entries"] > 50 and sizes["edges"] > 50
    assert np.float32(l3.view_info(sc.views[0].cam_id)["k"]) == np.float32(0.03) / np.float32(3.5)
    orc.close()
    sc.params = dict(sc.params, const_reg_depth=-1.0)
    with pytest.raises(api.L3DError, match="const_regularization_depth"):
        api.run_scene(sc)
