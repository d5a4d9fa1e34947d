"""GPU: the incremental (key-frame stream) mode against the oracle, cycle by cycle -- SURVEY.md
section 8f rank 1 / BASELINE config 3.  The oracle restates what Line3D keeps between two
matchImages calls (matched_, processed_, the filtered lists and their scores, the Add / Delete
score deltas of Line3D::scoringCPU src/line3D.cc:1439-1512, the per-cycle orientation re-test of
checkMatchOrientation src/line3D.cc:962-1014 and update_Matches_and_Estimated_position3D
src/line3D.cc:1857-1908); the CUDA path must reproduce every cycle bit for bit."""
import numpy as np
import pytest

import stream_utils

pytestmark = pytest.mark.gpu


def test_stream_world_point_neighbours(api, oracle, scene_mod):
    """The mode L3DPPing uses: neighbours from shared world points, re-chosen every cycle; sliding
    window with culling from the middle; poses re-estimated every cycle."""
    st = scene_mod.make_stream(n_keyframes=22, n_seg=500, window=12, nbrs=6, jitter=0.3, n_world=1200, cull_every=3)
    tot = stream_utils.run_lockstep(api, oracle, st)
    assert tot["cycles"] == 18 and tot["deleted"] >= 8 and tot["entries"] > 3000 and tot["edges"] > 3000


def test_stream_fixed_neighbours(api, oracle, scene_mod):
    """Explicit neighbour lists (setVisualNeighbors): a view's set is filled once and may name key
    frames that were culled since."""
    st = scene_mod.make_stream(n_keyframes=18, n_seg=400, window=7, nbrs=4, jitter=0.3, by_worldpoints=False,
                               cull_every=3)
    tot = stream_utils.run_lockstep(api, oracle, st)
    assert tot["cycles"] == 14 and tot["deleted"] >= 8 and tot["entries"] > 300


def test_stream_static_poses(api, oracle, scene_mod):
    """Without pose changes the re-triangulation is the identity and no orientation test flips."""
    st = scene_mod.make_stream(n_keyframes=12, n_seg=500, window=6, nbrs=5, jitter=0.0)
    tot = stream_utils.run_lockstep(api, oracle, st)
    assert tot["entries"] > 300


def test_stream_c3_shape(api, oracle, scene_mod):
    """BASELINE config 3's shape, shortened: 640x480, 1000 segments per key frame, window of 20,
    10 neighbours, 45 key frames (41 cycles)."""
    st = scene_mod.make_stream(n_keyframes=45, n_seg=1000, window=20, nbrs=10, jitter=0.3)
    tot = stream_utils.run_lockstep(api, oracle, st, check_scored=False)
    assert tot["cycles"] == 41 and tot["entries"] > 10000


def test_stream_idle_cycle_is_a_fixed_point(api, scene_mod):
    import test_stream_oracle as props
    st = scene_mod.make_stream(n_keyframes=8, n_seg=400, window=8, nbrs=5, jitter=0.0, n_world=900, cull_every=0)
    l3, calls = stream_utils.cuda_driver(api, st)
    scene_mod.drive_stream(st, **calls)
    cams = [u[0] for u in st.cycles[-1].updates]
    before, e0, _ = props.check_idle_cycle_is_a_fixed_point(l3, calls, st, cams)
    assert len(l3.pairs()) == 0 and l3.counts()["pair_tests"] == 0   # nothing is matched twice
    for c in cams:
        off, rec = l3.lists(c, 1)
        assert (off == before[c][0]).all() and rec.tobytes() == before[c][1].tobytes()
    assert l3.entries().tobytes() == e0.tobytes() and len(e0) > 100


def test_stream_ragged_segment_counts(api, oracle, scene_mod):
    """Key frames with different numbers of segments (7 ... 400), including one with fewer segments than
    a warp and one new key frame per cycle receiving inverse matches into short rows."""
    st = scene_mod.make_stream(n_keyframes=14, n_seg=400, window=9, nbrs=5, jitter=0.2, n_world=1000, cull_every=4)
    keep = [400, 7, 333, 129, 400, 31, 257, 64, 400, 199, 33, 400, 150, 65]
    for cy in st.cycles:
        for v in cy.adds:
            v.segs = np.ascontiguousarray(v.segs[:keep[v.cam_id % len(keep)]])
    tot = stream_utils.run_lockstep(api, oracle, st)
    assert tot["cycles"] == 10 and tot["entries"] > 300


def test_stream_delete_everything_then_continue(api, oracle, scene_mod):
    """Edge cases of the cycle: a cycle that deletes every key frame (nothing to match, no hypotheses,
    empty A_), then new key frames whose explicit neighbour lists still name the deleted ones (views_
    keeps them, src/line3D.cc:396-430, so they ARE matched against, with their last pose)."""
    base = scene_mod.make_stream(n_keyframes=9, n_seg=300, window=9, nbrs=4, jitter=0.2, by_worldpoints=False,
                                 cull_every=0, init=3, n_world=900)
    c0, c1, c2, c3 = base.cycles[0], base.cycles[1], base.cycles[2], base.cycles[3]
    wipe = scene_mod.StreamCycle([0, 1, 2, 3], [], [])
    # after the wipe: key frames 4 and 5 arrive; their lists name 2 and 3 (deleted) and each other
    c2.deletes, c3.deletes = [], []
    c2.updates = [(4, c2.updates[-1][1], c2.updates[-1][2], c2.updates[-1][3], [2, 3, 5])]
    c3.updates = [(4, c3.updates[-2][1], c3.updates[-2][2], c3.updates[-2][3], [2, 3, 5]),
                  (5, c3.updates[-1][1], c3.updates[-1][2], c3.updates[-1][3], [3, 4])]
    base.cycles = [c0, c1, wipe, c2, c3]
    tot = stream_utils.run_lockstep(api, oracle, base)
    assert tot["cycles"] == 5 and tot["deleted"] == 4 and tot["pairs"] >= 8


def test_stream_errors_mirror_the_reference(api, scene_mod):
    st = scene_mod.make_stream(n_keyframes=6, n_seg=100, window=6, jitter=0.0)
    l3, calls = stream_utils.cuda_driver(api, st)
    v = st.cycles[0].adds[0]
    calls["add"](v, v.worldpoints)
    with pytest.raises(api.L3DError, match="already in use"):
        calls["add"](v, v.worldpoints)
    assert l3.deleteImage(99) is False           # "camera ID [99] non_existent!"
    assert l3.deleteImage(v.cam_id) is True
    with pytest.raises(api.L3DError, match="views_reserved_"):
        l3.UpdataImage(v.cam_id, v.R, v.t, v.median_depth, v.worldpoints)
    with pytest.raises(api.L3DError, match="has no worldpoints"):
        calls["add"](st.cycles[0].adds[1], [])
    calls["add"](st.cycles[0].adds[3], st.cycles[0].adds[3].worldpoints)
    with pytest.raises(api.L3DError, match="ascending order"):
        calls["add"](st.cycles[0].adds[2], st.cycles[0].adds[2].worldpoints)
    with pytest.raises(api.L3DError, match="was deleted"):
        calls["add"](v, v.worldpoints)
