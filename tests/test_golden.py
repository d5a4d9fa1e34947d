"""Golden fixtures (tests/golden): BASELINE config 1 (cameras + world-point visibility of the
reference's NVM dump, frozen neighbour lists) and the tiny synthetic scene.  CPU: the oracle
reproduces the committed expectation; GPU: so does the CUDA path, through the C ABI."""
import os

import numpy as np
import pytest

import golden_utils


def _c1_with_worldpoints():
    sc = golden_utils.load_scene("c1_nvm_scene.npz")
    wps = np.load(os.path.join(golden_utils.HERE, "c1_nvm_wps.npz"))
    for v in sc.views:
        v.worldpoints = wps["wps_%d" % v.cam_id].tolist()
    return sc


def test_world_point_neighbours_match_golden_c1(api, oracle):
    """Host-only part of the nvm input path (Line3D::findVisualNeighborsFromWPs): from the world-point
    lists of the reference's NVM dump the product chooses exactly the frozen neighbour lists, for the
    reference's num_neighbors and for others (against the oracle)."""
    sc = _c1_with_worldpoints()
    got = api.neighbors_from_worldpoints(sc, 10)
    for v in sc.views:
        assert got[v.cam_id] == list(v.neighbors), v.cam_id
    # sparse world-point ids (the inverted index switches from counting sort to key sort): same choice
    for v in sc.views:
        v.worldpoints = [(w * 2654435761 + 12345) % (1 << 32) for w in v.worldpoints]
    got = api.neighbors_from_worldpoints(sc, 10)
    for v in sc.views:
        assert got[v.cam_id] == list(v.neighbors), v.cam_id
    # a view that lists some points twice: the reference's nested list walks count the multiplicities
    # (the bit-row popcount path must step aside for the counting walk)
    sc = _c1_with_worldpoints()
    sc.neighbors_by_worldpoints = True
    for v in sc.views[::3]:
        v.worldpoints = list(v.worldpoints) + list(v.worldpoints[:40])
    p = sc.params
    o = oracle.OracleLine3D(sc.max_image_width, True)
    o.load_scene(sc)
    o.match_images(p["sigma_p"], p["sigma_a"], 6, p["epipolar_overlap"], p["knn"], p["const_reg_depth"])
    got = api.neighbors_from_worldpoints(sc, 6)
    for v in sc.views:
        assert got[v.cam_id] == list(o.neighbors(v.cam_id)), ("dup", v.cam_id)
    o.close()
    sc = _c1_with_worldpoints()
    sc.neighbors_by_worldpoints = True
    p = sc.params
    for nn in (2, 5, 16):
        o = oracle.OracleLine3D(sc.max_image_width, True)
        o.load_scene(sc)
        o.match_images(p["sigma_p"], p["sigma_a"], nn, p["epipolar_overlap"], p["knn"], p["const_reg_depth"])
        got = api.neighbors_from_worldpoints(sc, nn)
        for v in sc.views:
            assert got[v.cam_id] == list(o.neighbors(v.cam_id)), (nn, v.cam_id)
        o.close()


def test_oracle_reproduces_golden_c1(oracle):
    sc = golden_utils.load_scene("c1_nvm_scene.npz")
    assert len(sc.views) == 25
    o = oracle.run_scene(sc)
    golden_utils.check_against_golden(o, sc, "c1_nvm_expected.npz")


def test_oracle_reproduces_golden_tiny(oracle, scene_mod):
    sc = scene_mod.make_scene("tiny")
    golden_utils.check_against_golden(oracle.run_scene(sc), sc, "tiny_expected.npz")


@pytest.mark.gpu
def test_cuda_reproduces_golden_c1(api):
    sc = golden_utils.load_scene("c1_nvm_scene.npz")
    l3 = api.run_scene(sc)
    golden_utils.check_against_golden(l3, sc, "c1_nvm_expected.npz")
    c = l3.counts()
    assert c["num_clusters"] == 286 and c["num_edges"] == 31608


@pytest.mark.gpu
def test_cuda_reproduces_golden_tiny(api, scene_mod):
    sc = scene_mod.make_scene("tiny")
    golden_utils.check_against_golden(api.run_scene(sc), sc, "tiny_expected.npz")


@pytest.mark.gpu
def test_cuda_world_point_mode_reproduces_golden_c1(api):
    """neighbors_by_worldpoints=True through the C ABI: same neighbours, same everything."""
    sc = _c1_with_worldpoints()
    frozen = {v.cam_id: list(v.neighbors) for v in sc.views}
    sc.neighbors_by_worldpoints = True
    l3 = api.run_scene(sc)
    for v in sc.views:
        assert l3.neighbors(v.cam_id) == frozen[v.cam_id]
    golden_utils.check_against_golden(l3, sc, "c1_nvm_expected.npz")


def _stream_expected():
    return np.load(os.path.join(golden_utils.HERE, "stream_expected.npz"))


def test_oracle_reproduces_golden_stream(oracle, scene_mod):
    """Incremental mode (delete / add / re-pose / match / reconstruct per cycle): every cycle's outputs
    hash to the committed digests, with one OpenMP thread and with all of them."""
    import stream_utils
    g = _stream_expected()
    stream = scene_mod.make_stream(**stream_utils.GOLDEN_STREAM)
    for threads in (1, 0):
        o, calls = stream_utils.oracle_driver(oracle, stream, threads)
        dig, sizes = stream_utils.stream_digests(stream, o, calls)
        assert (sizes == g["sizes"]).all()
        assert (dig == g["digests"]).all(), np.nonzero((dig != g["digests"]).any(axis=1))[0]
        o.close()


@pytest.mark.gpu
def test_cuda_reproduces_golden_stream(api, scene_mod):
    import stream_utils
    g = _stream_expected()
    stream = scene_mod.make_stream(**stream_utils.GOLDEN_STREAM)
    l3, calls = stream_utils.cuda_driver(api, stream)
    dig, sizes = stream_utils.stream_digests(stream, l3, calls)
    assert (sizes == g["sizes"]).all()
    assert (dig == g["digests"]).all(), np.nonzero((dig != g["digests"]).any(axis=1))[0]


def test_world_point_neighbours_randomised(api, oracle, scene_mod):
    """The three counting paths of the host-side neighbour selection (bit rows, dense counting walk, sparse
    key sort) against the oracle's restatement of Line3D::findVisualNeighborsFromWPs on random world-point
    lists: few / many points, duplicates inside a list, huge ids, views without points."""
    rng = np.random.default_rng(2026)
    found = 0
    for trial in range(24):
        sc = scene_mod.make_scene("tiny", n_views=10, n_seg=40, nbrs=3, seed=100 + trial)
        n_pts = int(rng.choice([30, 200, 3000]))
        big = trial % 3 == 2                      # sparse ids
        dup = trial % 4 == 1                      # a list names points twice
        for v in sc.views:
            k = int(rng.integers(0, min(n_pts, 150)))
            ids = rng.choice(n_pts, size=k, replace=False)
            if dup and k > 4:
                ids = np.concatenate([ids, ids[:k // 3]])
            if big:
                ids = (ids.astype(np.uint64) * 2654435761 + 977) % (1 << 32)
            v.worldpoints = [int(x) for x in ids]
        if all(len(v.worldpoints) == 0 for v in sc.views):
            continue
        sc.neighbors_by_worldpoints = True
        nn = int(rng.choice([2, 3, 6]))
        o = oracle.OracleLine3D(sc.max_image_width, True)
        for v in sc.views:                        # addImage refuses an empty list; such views simply stay out
            if len(v.worldpoints):
                o.add_image(v.cam_id, v.K, v.R, v.t, v.width, v.height, v.median_depth, v.worldpoints, v.segs)
                o.update_image(v.cam_id, v.R, v.t, v.median_depth, v.worldpoints)
        p = sc.params
        o.match_images(p["sigma_p"], p["sigma_a"], nn, p["epipolar_overlap"], p["knn"], p["const_reg_depth"])
        sc.views = [v for v in sc.views if len(v.worldpoints)]
        got = api.neighbors_from_worldpoints(sc, nn)
        for v in sc.views:
            assert got[v.cam_id] == list(o.neighbors(v.cam_id)), (trial, v.cam_id, n_pts, big, dup)
            found += len(got[v.cam_id])
        o.close()
    assert found > 150
