"""CPU: the synthetic key-frame stream generator (BASELINE config 3 inputs) is deterministic and keeps the
invariants the incremental mode relies on."""
import numpy as np


def test_stream_generator_invariants(scene_mod):
    st = scene_mod.make_stream(n_keyframes=40, n_seg=60, window=10, nbrs=5, cull_every=4)
    again = scene_mod.make_stream(n_keyframes=40, n_seg=60, window=10, nbrs=5, cull_every=4)
    current, seen, last_added = set(), set(), -1
    for a, b in zip(st.cycles, again.cycles):
        assert a.deletes == b.deletes and [v.cam_id for v in a.adds] == [v.cam_id for v in b.adds]
        for va, vb in zip(a.adds, b.adds):
            assert va.segs.tobytes() == vb.segs.tobytes() and (va.R == vb.R).all() and (va.t == vb.t).all()
        for ua, ub in zip(a.updates, b.updates):
            assert ua[0] == ub[0] and (ua[1] == ub[1]).all() and (ua[2] == ub[2]).all() and list(ua[4]) == list(ub[4])
        assert a.deletes == sorted(a.deletes) and set(a.deletes) <= current
        current -= set(a.deletes)
        for v in a.adds:
            assert v.cam_id > last_added and v.cam_id not in seen          # ascending ids, never re-used
            assert v.segs.shape == (60, 4) and v.segs.dtype == np.float32
            ln = np.hypot(v.segs[:, 0] - v.segs[:, 2], v.segs[:, 1] - v.segs[:, 3])
            assert (np.diff(ln) <= 1e-3).all()                              # sorted by length, longest first
            last_added = v.cam_id
            seen.add(v.cam_id)
            current.add(v.cam_id)
        assert len(current) <= 10
        assert [u[0] for u in a.updates] == sorted(current)                 # every current key frame is re-posed
        for cam, R, t, md, lst in a.updates:
            assert np.allclose(R @ R.T, np.eye(3), atol=1e-12) and md > 0 and len(lst) > 4
    assert sum(len(c.adds) for c in st.cycles) == 40 and sum(len(c.deletes) for c in st.cycles) >= 30
    # the mode L3DPPing uses: world points shared between neighbouring key frames
    u = st.cycles[10].updates
    assert len(set(u[0][4]) & set(u[1][4])) > 20
