"""GPU candidate-reference costing of BvComp against the host front end: the cost of every
(node, reference offset) candidate record and the symbols of the chosen records must be identical."""
import numpy as np
import pytest

from conftest import random_graph

pytestmark = pytest.mark.gpu


def entropy_tables(W, off, succ, window, max_ref, min_interval):
    """Model of the Log2Estimator pass: what the second and third pass of ANSBvGraph::store cost with."""
    comps, syms = W.bvcomp_symbols(off, succ, window, max_ref, min_interval)
    mb = W.ANSModel4EncoderBuilder()
    mb.push_symbols(comps, syms)
    return mb.build()[0]


@pytest.mark.parametrize("params", [(7, 3, 4), (1, 3, 2), (16, 1 << 30, 4), (7, 3, 0), (0, 3, 4), (32, 2, 3)])
@pytest.mark.parametrize("estimator", ["log2", "entropy"])
def test_candidate_costs_and_chosen_symbols_equal_the_host(W, gpu, params, estimator):
    window, max_ref, min_interval = params
    off, succ = random_graph(np.random.default_rng(100 + window), 6000, 14)
    tables = entropy_tables(W, off, succ, *params) if estimator == "entropy" else None
    for chunk in (0, 1000):
        if window:
            g = W.bvcomp_costs(off, succ, window, min_interval, tables, chunk_nodes=chunk, use_gpu=True)
            h = W.bvcomp_costs(off, succ, window, min_interval, tables, chunk_nodes=chunk, use_gpu=False)
            assert g.shape == h.shape == (6000, window + 1)
            assert (g == h).all(), np.argwhere(g != h)[:5]
            assert (h[:, 0] != np.uint64(0xFFFFFFFFFFFFFFFF)).all()  # the reference-free candidate always exists
        c1, s1 = W.bvcomp_symbols(off, succ, window, max_ref, min_interval, tables, chunk_nodes=chunk, threads=3)
        c2, s2 = W.bvcomp_symbols(off, succ, window, max_ref, min_interval, tables, chunk_nodes=chunk, threads=3,
                                  gpu_costing=True)
        assert (c1 == c2).all() and (s1 == s2).all()


def test_gpu_costing_on_a_node_range_of_a_larger_graph(W, gpu):
    """One rank's share: nodes [first, first + n) with global successor ids, chunks aligned to the whole graph."""
    off, succ = random_graph(np.random.default_rng(5), 9000, 12)
    first, last = 3000, 7500
    r_off = (off[first:last + 1] - off[first]).astype(np.uint64)
    r_succ = succ[int(off[first]):int(off[last])]
    for chunk in (0, 1000, 2048):
        a = W.bvcomp_symbols(r_off, r_succ, 7, 3, 4, None, chunk_nodes=chunk, threads=2, first_node=first)
        b = W.bvcomp_symbols(r_off, r_succ, 7, 3, 4, None, chunk_nodes=chunk, threads=2, first_node=first, gpu_costing=True)
        assert (a[0] == b[0]).all() and (a[1] == b[1]).all()
        g = W.bvcomp_costs(r_off, r_succ, 7, 4, None, chunk_nodes=chunk, first_node=first, use_gpu=True)
        h = W.bvcomp_costs(r_off, r_succ, 7, 4, None, chunk_nodes=chunk, first_node=first, use_gpu=False)
        assert (g == h).all()


def test_gpu_costing_bench_shaped_graph_and_long_lists(W, gpu):
    off, succ = W.synth_graph("web", 120_000, 34.3, seed=3)
    tables = entropy_tables(W, off, succ, 7, 3, 4)
    a = W.bvcomp_symbols(off, succ, 7, 3, 4, tables, chunk_nodes=65536, threads=4)
    b = W.bvcomp_symbols(off, succ, 7, 3, 4, tables, chunk_nodes=65536, threads=4, gpu_costing=True)
    assert (a[0] == b[0]).all() and (a[1] == b[1]).all()
    off, succ = W.synth_graph("social", 60_000, 35.3, seed=4)
    a = W.bvcomp_symbols(off, succ, 7, 3, 4, None, chunk_nodes=0)
    b = W.bvcomp_symbols(off, succ, 7, 3, 4, None, chunk_nodes=0, gpu_costing=True)
    assert (a[0] == b[0]).all() and (a[1] == b[1]).all()
