"""CPU: the one piece of the reference's path that compiles from its own sources without external
libraries -- the graph clustering (src/clustering.cc, include/clustering.h, include/universe.h) -- is
built in place into oracle/_ref/ (oracle/Makefile, target `ref`) and used to pin, against the REFERENCE
ITSELF, both the oracle's restatement and the product's host clustering (l3d_cluster_edges)."""
import numpy as np
import pytest


def _ref_or_skip(oracle):
    probe = oracle.ref_cluster(np.zeros((0, 2), np.int32), np.zeros(0, np.float32), 1)
    if probe is None:
        pytest.skip("oracle/_ref/libref_clustering.so is not built (needs /root/reference once)")


def _random_affinity(rng, n, m, levels):
    """An A_-shaped edge list: every undirected edge twice, (i,j,w) then (j,i,w), weights in (0.5, 1]
    quantised to `levels` values so that ties (and the stable sort) matter."""
    i = rng.integers(0, n, size=m)
    j = rng.integers(0, n, size=m)
    keep = i != j
    i, j = i[keep], j[keep]
    w = (0.5 + 0.5 * (rng.integers(1, levels + 1, size=len(i)) / levels)).astype(np.float32)
    ij = np.empty((2 * len(i), 2), np.int32)
    ij[0::2, 0], ij[0::2, 1], ij[1::2, 0], ij[1::2, 1] = i, j, j, i
    return ij, np.repeat(w, 2)


@pytest.mark.parametrize("n,m,levels", [(50, 120, 4), (400, 1500, 16), (3000, 9000, 1000), (64, 2000, 3)])
def test_clustering_equals_the_reference_on_random_graphs(api, oracle, n, m, levels):
    _ref_or_skip(oracle)
    rng = np.random.default_rng(n * 7 + levels)
    ij, w = _random_affinity(rng, n, m, levels)
    ref = oracle.ref_cluster(ij, w, n)
    assert (oracle.kat_cluster(ij, w, n) == ref).all()
    assert (api.cluster_edges(ij, w, n) == ref).all()
    assert len(set(ref.tolist())) < n        # something was merged


def test_clustering_equals_the_reference_on_scene_matrices(api, oracle, scene_mod):
    """A_ of the tiny scene and of the C1 (NVM) golden scene, as the oracle builds it."""
    import golden_utils
    _ref_or_skip(oracle)
    for sc in (scene_mod.make_scene("tiny"), golden_utils.load_scene("c1_nvm_scene.npz")):
        o = oracle.run_scene(sc)
        ij, w = o.edges()
        n = len(o.local2global())
        ref = oracle.ref_cluster(ij, w, n)
        assert n > 50 and (o.cluster_ids() == ref).all()
        assert (api.cluster_edges(ij, w, n) == ref).all()
        o.close()


def test_clustering_equals_the_reference_property(api, oracle):
    """Property test (hypothesis): arbitrary small multigraphs with heavily tied weights, self-consistent
    A_ layout or not -- the three implementations agree on every cluster id."""
    from hypothesis import given, settings, strategies as st
    _ref_or_skip(oracle)

    @settings(max_examples=150, deadline=None)
    @given(st.integers(2, 12).flatmap(lambda n: st.tuples(
        st.just(n),
        st.lists(st.tuples(st.integers(0, n - 1), st.integers(0, n - 1), st.sampled_from([0.51, 0.6, 0.6, 0.75, 0.9, 1.0])),
                 min_size=1, max_size=40))))
    def check(case):
        n, edges = case
        ij = np.array([(a, b) for a, b, _ in edges], dtype=np.int32)
        w = np.array([x for _, _, x in edges], dtype=np.float32)
        ref = oracle.ref_cluster(ij, w, n)
        assert (oracle.kat_cluster(ij, w, n) == ref).all()
        assert (api.cluster_edges(ij, w, n) == ref).all()
    check()
