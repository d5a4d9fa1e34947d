"""GPU: K1's row sort + wedge test (k1_rowsort_kernel / hull_skips in csrc/k1_pairtest.cu).

The order of the source rows only selects which pair tests are SKIPPED; whatever it is, the results must be the
oracle's.  Covered here: views with more rows than one sort chunk (4096), so a pair is sorted in several chunks;
that the wedge test really skips most tests on a scene of the C4 shape and none when it is switched off; views
with fewer rows than a warp; an epipole inside the image (forward motion)."""
import importlib
import os
import subprocess
import sys

import numpy as np
import pytest

from parity_utils import compare_full

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_more_rows_than_one_sort_chunk(api, oracle, scene_mod):
    sc = scene_mod.make_scene("c4", n_views=4, n_seg=4500, nbrs=2)   # 4500 rows: chunks of 4096 + 404
    assert all(v.segs.shape[0] == 4500 for v in sc.views)
    l3 = api.run_scene(sc)
    orc = oracle.run_scene(sc)
    sizes = compare_full(l3, orc, sc, check_scored=False)
    assert sizes["pairs"] >= 4
    c = l3.counts()
    assert c["pair_tests"] == sizes["pairs"] * 4500 * 4500
    assert 0 < c["pair_tests_run"] < c["pair_tests"]
    orc.close()


def test_wedge_test_skips_most_tests_and_counts_them(api, scene_mod):
    sc = scene_mod.make_scene("c4", n_views=12)
    c = api.run_scene(sc).counts()
    assert c["pair_tests_run"] < 0.4 * c["pair_tests"]      # measured 0.16-0.18 on this shape
    assert c["pair_tests_run"] > c["candidates"]            # every candidate was evaluated


def test_fewer_rows_than_a_warp(api, oracle, scene_mod):
    sc = scene_mod.make_scene("tiny", n_views=6, n_seg=20, nbrs=3)
    l3 = api.run_scene(sc)
    orc = oracle.run_scene(sc)
    compare_full(l3, orc, sc, check_scored=False)
    orc.close()


def test_forward_motion_epipole_inside_the_image(api, oracle, scene_mod):
    """Cameras in a line along their viewing direction: the epipole of every pair lies inside both images, so
    the epipolar lines of a pair point in every direction (the sort key wraps, wide wedges)."""
    sc = scene_mod.make_scene("c2", n_views=6, n_seg=600, nbrs=3)
    look = scene_mod.look_at
    for i, v in enumerate(sc.views):
        c = np.array([0.3 * i - 6.0, 0.02 * i, 1.7])
        R = look(c, c + np.array([1.0, 0.0, 0.02]))
        v.R, v.t = R, -R @ c
    # the segments were projected with the old poses: as 2-D input they are as good as any
    l3 = api.run_scene(sc)
    orc = oracle.run_scene(sc)
    compare_full(l3, orc, sc, check_scored=False)
    orc.close()


def test_rectified_stereo_parallel_epipolar_lines(api, oracle, scene_mod):
    """Cameras side by side with one common rotation: the epipole of every pair is a point at infinity and all its
    epipolar lines are parallel -- the angle of a line no longer orders the rows, the sort key has to be the
    distance of the line from the middle of the image (an angle-only key made every warp take the wedge of its
    first row for its hull and skip the matches of the others)."""
    from test_wedge_rule import rectified_scene
    sc = rectified_scene(scene_mod, n_views=5, n_seg=700)
    l3 = api.run_scene(sc)
    orc = oracle.run_scene(sc)
    sizes = compare_full(l3, orc, sc, check_scored=False)
    c = l3.counts()
    assert c["forward_matches"] > 1000 and c["pair_tests_run"] < c["pair_tests"]
    assert sizes["pairs"] >= 5
    orc.close()


@pytest.mark.parametrize("seed", [0, 1, 2, 3, 6, 7])
def test_random_two_view_geometries(api, oracle, scene_mod, seed):
    """epipole at the image centre / far away / at infinity / oblique, small and large rotations (the scenes of
    tests/test_wedge_rule.py, 900 segments per view so that several warps per class exist)"""
    from test_wedge_rule import random_two_view_scene
    sc = random_two_view_scene(scene_mod, seed, n_seg=900)
    l3 = api.run_scene(sc)
    orc = oracle.run_scene(sc)
    compare_full(l3, orc, sc, check_scored=False)
    assert l3.counts()["forward_matches"] > 0
    orc.close()


_HOOK_SCRIPT = r"""
import importlib, sys
sys.path.insert(0, %r)
api = importlib.import_module("3dline-slam_b200.api")
scene_mod = importlib.import_module("3dline-slam_b200.scene")
sc = scene_mod.make_scene("c4", n_views=8)
l3 = api.run_scene(sc)
c = l3.counts()
print(api.result_digest(l3, [v.cam_id for v in sc.views]), c["pair_tests"], c["pair_tests_run"])
"""


def _run_with(env_extra):
    env = dict(os.environ)
    env.update(env_extra)
    out = subprocess.check_output([sys.executable, "-c", _HOOK_SCRIPT % ROOT], env=env, text=True).strip().split()
    return out[0], int(out[1]), int(out[2])


def test_results_do_not_depend_on_the_row_order():
    """L3D_K1_SORT=0 keeps the natural row order, L3D_K1_HULL=0 evaluates every test (read once per process, hence
    the subprocesses): same digest three times."""
    d0, n0, r0 = _run_with({})
    d1, n1, r1 = _run_with({"L3D_K1_SORT": "0"})
    d2, n2, r2 = _run_with({"L3D_K1_HULL": "0"})
    assert d0 == d1 == d2
    assert n0 == n1 == n2
    assert r2 == n2          # no wedge test: every pair test is evaluated
    assert r0 < r1 <= n1     # unsorted rows: the wedge of a warp covers almost everything
