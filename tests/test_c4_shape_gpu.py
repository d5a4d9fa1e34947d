"""GPU, BASELINE config 4's SHAPE (1920x1080, f = 1500, 3000 segments per view, 20 neighbours) on a cut of
the scene the oracle finishes in well under a minute: bit-exact parity of everything the path produces.
The shape matters: K1's guard band scales with the image size (Xb, cN), target views have three mask
chunks per row, K2 runs the multi-chunk mapping, and the candidate rate per row is far above kNN."""
import importlib

import numpy as np
import pytest

from parity_utils import compare_full

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def c4cut(scene_mod):
    return scene_mod.make_scene("c4", n_views=40)


def test_c4_shape_bit_exact_against_the_oracle(api, oracle, c4cut):
    l3 = api.run_scene(c4cut)
    orc = oracle.run_scene(c4cut)
    sizes = compare_full(l3, orc, c4cut, check_scored=False)
    assert sizes["pairs"] > 300 and sizes["entries"] > 5000
    c = l3.counts()
    assert c["pair_tests"] == orc.pair_tests() == sizes["pairs"] * 3000 * 3000
    # the pre-filter is a filter: far fewer candidates than tests, more than the matches it must keep
    assert c["forward_matches"] < c["candidates"] < c["pair_tests"] // 20
    orc.close()


def test_c4_shape_prefilter_is_conservative(api, c4cut):
    """filter_mode = 1 sends EVERY pair through the exact kernel: same results as with the FP32 pre-filter."""
    small = importlib.import_module("3dline-slam_b200.scene").make_scene("c4", n_views=6)
    a = api.run_scene(small, filter_mode=0)
    b = api.run_scene(small, filter_mode=1)
    views = [v.cam_id for v in small.views]
    assert api.result_digest(a, views) == api.result_digest(b, views)
    assert b.counts()["candidates"] == b.counts()["pair_tests"]


def test_c4_shape_sharded_equals_unsharded(api, c4cut):
    import torch
    shd = importlib.import_module("3dline-slam_b200.sharding")
    views = [v.cam_id for v in c4cut.views]
    ref = api.result_digest(api.run_scene(c4cut), views)
    shards = []
    for r in range(3):
        l3 = api.Line3D("", False, c4cut.max_image_width)
        l3.shard = (r, 3)
        l3.load_scene(c4cut)
        shards.append(l3)
    grp = shd.LocalGroup(shards, torch, torch.device("cuda", 0))
    grp.run(c4cut.params)
    grp.run(c4cut.params)      # steady state: self-describing device blobs
    for s in shards:
        assert api.result_digest(s, views) == ref
