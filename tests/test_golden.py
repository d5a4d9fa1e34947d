"""Golden fixtures (tests/golden): BASELINE config 1 (cameras + world-point visibility of the
reference's NVM dump, frozen neighbour lists) and the tiny synthetic scene.  CPU: the oracle
reproduces the committed expectation; GPU: so does the CUDA path, through the C ABI."""
import pytest

import golden_utils


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
