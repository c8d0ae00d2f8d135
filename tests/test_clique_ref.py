"""Pins the oracle's clique restatement against the reference's OWN max-clique finder (FMC), compiled
from /root/reference/include/KimeraRPGO/max_clique_finder/*.cpp into oracle/_ref/libref_fmc.so
(oracle/Makefile).  The prebuilt .so travels to the GPU box; if it is absent the tests skip."""
import numpy as np
import pytest

import orc

pytestmark = pytest.mark.skipif(orc.ref_fmc() is None, reason="oracle/_ref/libref_fmc.so not built")


def rand_graph(rng, n, p):
    a = (rng.random((n, n)) < p).astype(np.uint8)
    a = np.triu(a, 1)
    return a + a.T


def test_survey_probe_graph():
    """4-vertex graph {01,02,03,13}: heuristic size 3 ids [0,1,2] (not a clique), exact [0,1,3]."""
    a = np.zeros((4, 4), dtype=np.uint8)
    for i, j in [(0, 1), (0, 2), (0, 3), (1, 3)]:
        a[i, j] = a[j, i] = 1
    k, ids = orc.ref_clique_heu(a)
    assert (k, ids.tolist()) == (3, [0, 1, 2])
    k2, ids2 = orc.clique_heu(a)
    assert (k2, ids2.tolist()) == (3, [0, 1, 2])
    k3, ids3 = orc.ref_clique_exact(a)
    assert k3 == 3 and sorted(ids3.tolist()) == [0, 1, 3]


def test_empty_graph_is_size_one():
    a = np.zeros((5, 5), dtype=np.uint8)
    assert orc.ref_clique_heu(a)[0] == 1 and orc.ref_clique_heu(a)[1].tolist() == [0]
    assert orc.clique_heu(a)[0] == 1 and orc.clique_heu(a)[1].tolist() == [0]
    a1 = np.zeros((1, 1), dtype=np.uint8)
    assert orc.ref_clique_heu(a1)[0] == orc.clique_heu(a1)[0] == 1


def test_heuristic_matches_reference_fmc():
    rng = np.random.default_rng(7)
    for t in range(300):
        n = int(rng.integers(2, 70))
        a = rand_graph(rng, n, rng.uniform(0.1, 0.95))
        kr, ir = orc.ref_clique_heu(a)
        ko, io = orc.clique_heu(a)
        assert kr == ko and ir.tolist() == io.tolist(), (t, n)


def test_incremental_matches_reference_fmc():
    rng = np.random.default_rng(8)
    for t in range(300):
        n = int(rng.integers(3, 60))
        a = rand_graph(rng, n, rng.uniform(0.1, 0.95))
        num_new = int(rng.integers(1, n))
        prev = int(rng.integers(0, 6))
        kr, ir = orc.ref_clique_heu_incremental(a, num_new, prev)
        ko, io = orc.clique_heu_incremental(a, num_new, prev)
        assert kr == ko and ir.tolist() == io.tolist(), (t, n, num_new, prev)


def test_exact_matches_reference_fmc():
    rng = np.random.default_rng(9)
    for t in range(200):
        n = int(rng.integers(2, 40))
        a = rand_graph(rng, n, rng.uniform(0.1, 0.9))
        kr, ir = orc.ref_clique_exact(a)
        ko, io = orc.clique_exact(a)
        assert kr == ko and ir.tolist() == io.tolist(), (t, n)
        # the exact finder returns a true clique
        s = ir.tolist()
        assert all(a[x, y] for x in s for y in s if x != y)


def test_incremental_on_the_1x1_matrix_of_a_disabled_loop_check():
    """With the pairwise check disabled the reference never grows its matrices (Pcm.h:484-486), so findInliersIncremental
    runs FMC on a 1x1 zero matrix with num_new >= 1: vertex 0 is selected only for num_new == 1 and prev == 0; for
    num_new > n the size_t candidate index wraps and nothing is selected (findCliqueHeu.cpp:141).  Pinned on the
    reference's own code when it is built, and on the restatement."""
    z = np.zeros((1, 1), dtype=np.uint8)
    impls = [orc.clique_heu_incremental]
    if orc.ref_fmc() is not None:
        impls.append(orc.ref_clique_heu_incremental)
    for f in impls:
        k, ids = f(z, 1, 0)
        assert k == 1 and ids.tolist() == [0]
        assert f(z, 1, 1)[0] == 0
        assert f(z, 3, 0)[0] == 0
