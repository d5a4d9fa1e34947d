"""GPU: the CUDA path against the REFERENCE'S OWN Line3D++ SOURCES compiled into oracle/_ref (see
tests/test_ref_line3d.py and oracle/ref_line3d_wrap.cpp) -- no restatement in between.  The prebuilt library
travels to the GPU box with the repo (it is built where /root/reference exists)."""
import numpy as np
import pytest

import stream_utils
from parity_utils import compare_full

pytestmark = pytest.mark.gpu


def _need(oracle):
    if oracle.ref_lib("det") is None:
        pytest.skip("oracle/_ref/libref_line3d_det.so did not travel / is not built")


@pytest.mark.parametrize("kind,kw", [("tiny", {}), ("c2", dict(n_views=20, n_seg=500)), ("c4", dict(n_views=6, n_seg=1500))])
def test_cuda_path_equals_the_compiled_reference(api, oracle, scene_mod, kind, kw):
    _need(oracle)
    scene = scene_mod.make_scene(kind, **kw)
    l3 = api.run_scene(scene)
    ref = oracle.run_scene_ref(scene, "det")
    sizes = compare_full(l3, ref, scene, check_scored=False)   # match lists, hypotheses, A_, ids, cluster roots: bit for bit
    assert sizes["entries"] > 50 and sizes["edges"] > 50
    assert l3.counts()["pair_tests"] == ref.pair_tests()
    ref.close()


def test_cuda_stream_mode_equals_the_compiled_reference(api, oracle, scene_mod):
    """The incremental mode against the reference's own Line3D object driven like L3DPPing::Run drives it."""
    _need(oracle)
    st = scene_mod.make_stream(n_keyframes=16, n_seg=300, window=6, nbrs=4, jitter=0.3, n_world=700)

    class RefAsOracle:      # stream_utils.oracle_driver builds `oracle.OracleLine3D(width, by_wps, threads)`
        @staticmethod
        def OracleLine3D(width, by_wps, threads=0):
            return oracle.RefLine3D(width, by_wps, "det")

    tot = stream_utils.run_lockstep(api, RefAsOracle, st, check_scored=False)
    assert tot["cycles"] >= 10 and tot["deleted"] >= 6 and tot["entries"] > 0
