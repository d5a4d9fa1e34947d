"""CPU: properties of the incremental mode that follow from the reference's code, checked on the
oracle (tests/test_stream_gpu.py checks the same on the CUDA path and compares the two bit for bit)."""
import numpy as np

import stream_utils


def _replay(obj, calls, stream, upto):
    scene_mod = stream_utils.scene_mod
    scene_mod.drive_stream(stream, max_cycles=upto, **calls)


def check_idle_cycle_is_a_fixed_point(obj, calls, stream, cams):
    """A cycle without adds / deletes / pose changes: matched_ holds every pair (nothing is matched
    again, src/line3D.cc:867), Add_camID_ and Delete_camID_ are empty, so every score keeps its value
    (src/line3D.cc:1439-1512) and filterMatches keeps the same entries."""
    before = {c: obj.lists(c, 1) for c in cams}
    e0 = obj.entries().copy()
    n_pairs = len(obj.pairs())
    calls["begin_cycle"]()
    calls["match"](stream.params)
    calls["reconstruct"]()
    return before, e0, n_pairs


def test_oracle_idle_cycle_keeps_everything(oracle, scene_mod):
    st = scene_mod.make_stream(n_keyframes=8, n_seg=400, window=8, nbrs=5, jitter=0.0, n_world=900, cull_every=0)
    o, calls = stream_utils.oracle_driver(oracle, st)
    _replay(o, calls, st, None)
    cams = [u[0] for u in st.cycles[-1].updates]
    before, e0, n_pairs = check_idle_cycle_is_a_fixed_point(o, calls, st, cams)
    assert len(o.pairs()) == n_pairs                       # the oracle's pair log is cumulative
    for c in cams:
        off, rec = o.lists(c, 1)
        assert (off == before[c][0]).all() and rec.tobytes() == before[c][1].tobytes()
    assert o.entries().tobytes() == e0.tobytes() and len(e0) > 100
    o.close()


def test_oracle_delete_removes_every_match_to_the_camera(oracle, scene_mod):
    st = scene_mod.make_stream(n_keyframes=7, n_seg=400, window=8, nbrs=5, jitter=0.0, n_world=900, cull_every=0)
    o, calls = stream_utils.oracle_driver(oracle, st)
    _replay(o, calls, st, None)
    cams = [u[0] for u in st.cycles[-1].updates]
    victim = cams[2]
    had = sum(int((o.lists(c, 1)[1]["tgt_cam"] == victim).sum()) for c in cams if c != victim)
    s_before = {c: o.lists(c, 1)[1]["score"].sum() for c in cams if c != victim}
    assert had > 20
    calls["begin_cycle"]()
    assert o.delete_image(victim)
    assert not o.delete_image(victim)                     # "non_existent" the second time
    calls["match"](st.params)
    calls["reconstruct"]()
    for c in cams:
        if c == victim:
            continue
        rec = o.lists(c, 1)[1]
        assert not (rec["tgt_cam"] == victim).any()
        assert rec["score"].sum() <= s_before[c]          # per-camera maxima of the victim were subtracted
    e = o.entries()
    assert not (e["src_cam"] == victim).any() and not (e["tgt_cam"] == victim).any()
    o.close()
